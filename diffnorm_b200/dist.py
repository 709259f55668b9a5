"""Data-parallel gradient exchange for the denoiser training step (the ONLY collective on any path here).

One process per GPU (torch.distributed, backend "nccl" over NVLink / NVSwitch; "gloo" in the CPU tests).  Gradients
are copied into persistent flat fp32 buckets in the order the backward pass produces them; a bucket's all-reduce is
launched (async) the moment it fills, so the exchange of the transformer's gradients overlaps the WaveNet part of
the backward pass.  ``finish()`` waits, divides by the world size (DDP's mean, trainer.py:918-932) and returns views.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, bucket_bytes: int = 64 << 20, group=None, average: bool = True):
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.group, self.average = group, average
        self.buckets: List[torch.Tensor] = []
        self.reset()

    def reset(self):
        self.cur, self.off = 0, 0
        self.entries: List[Tuple[str, torch.Size, int, int]] = []   # name, shape, bucket, offset
        self.works = []
        self.launched = -1

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _bucket(self, i: int, like: torch.Tensor, need: int) -> torch.Tensor:
        while len(self.buckets) <= i:
            self.buckets.append(None)
        b = self.buckets[i]
        size = max(self.bucket_elems, need)
        if b is None or b.numel() < size or b.device != like.device:
            b = torch.empty(size, dtype=torch.float32, device=like.device)
            self.buckets[i] = b
        return b

    def _launch(self, i: int, used: int):
        if dist.is_initialized() and self.world > 1:
            self.works.append(dist.all_reduce(self.buckets[i][:used], group=self.group, async_op=True))
        self.launched = i

    def hook(self, name: str, g: torch.Tensor):
        """Called by DenoiserTrainer.step as soon as a gradient is final (in stream order)."""
        n = g.numel()
        if self.off > 0 and self.off + n > self.bucket_elems:
            self._launch(self.cur, self.off)
            self.cur, self.off = self.cur + 1, 0
        b = self._bucket(self.cur, g, n)
        b[self.off:self.off + n].copy_(g.reshape(-1))
        self.entries.append((name, g.shape, self.cur, self.off))
        self.off += n

    def finish(self) -> Dict[str, torch.Tensor]:
        if self.off > 0:
            self._launch(self.cur, self.off)
        for w in self.works:
            w.wait()
        out = {}
        scale = 1.0 / self.world if (self.average and self.world > 1) else None
        used: Dict[int, int] = {}
        for name, shape, bi, off in self.entries:
            used[bi] = max(used.get(bi, 0), off + shape.numel())
        if scale is not None:
            for bi, n in used.items():
                self.buckets[bi][:n].mul_(scale)
        for name, shape, bi, off in self.entries:
            out[name] = self.buckets[bi][off:off + shape.numel()].view(shape)
        self.reset()
        return out
