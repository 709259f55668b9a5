"""Data-parallel gradient exchange for the denoiser training step (the ONLY collective on any path here).

One process per GPU (torch.distributed, backend "nccl" over NVLink / NVSwitch; "gloo" in the CPU tests).  Gradients
are copied into persistent flat buckets in the order the backward pass produces them; a bucket's all-reduce is
launched (async, on NCCL's own stream) the moment it fills, so the exchange of the transformer's gradients overlaps the
WaveNet part of the backward pass.  ``finish()`` waits and returns per-parameter fp32 views; the result is the MEAN over
ranks (DDP's semantics, fairseq/trainer.py:918-932, models/distributed_fairseq_model.py:59-69).

Three things decide how much of the exchange is hidden (profiles/r02_*train_comm*):
  * the persistent GEMM kernels stride their tiles statically over one CTA per SM, so an NCCL kernel that takes SMs away
    mid-backward delays whole CTAs; ``sm_reserve`` caps the GEMM grids (dn_set_sm_limit) while buckets are in flight so
    the collective owns its SMs instead of stealing them;
  * ``comm_dtype=torch.bfloat16`` halves the bytes on the wire (the sum is then formed in bf16: NOT DDP's fp32
    arithmetic — opt-in, the default stays fp32);
  * the mean is taken by NCCL itself (ReduceOp.AVG) instead of a separate pass over 1 GB.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, bucket_bytes: int = 64 << 20, group=None, average: bool = True, comm_dtype: torch.dtype = torch.float32,
                 overlap: bool = True, sm_reserve: int = 0):
        self.comm_dtype = comm_dtype
        self.bucket_elems = max(1, bucket_bytes // torch.empty((), dtype=comm_dtype).element_size())
        self.group, self.average, self.overlap, self.sm_reserve = group, average, overlap, int(sm_reserve)
        self.buckets: List[Optional[torch.Tensor]] = []
        self.out32: List[Optional[torch.Tensor]] = []     # fp32 landing buffers when the wire format is 16 bit
        self._limited = False
        self.reset()

    def reset(self):
        self.cur, self.off = 0, 0
        self.entries: List[Tuple[str, torch.Size, int, int]] = []   # name, shape, bucket, offset
        self.works = []
        self.pending: List[Tuple[int, int]] = []                    # filled buckets not launched yet (overlap off)
        self.used: Dict[int, int] = {}

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _active(self) -> bool:
        return dist.is_initialized() and self.world > 1

    def _bucket(self, i: int, like: torch.Tensor, need: int) -> torch.Tensor:
        while len(self.buckets) <= i:
            self.buckets.append(None)
            self.out32.append(None)
        b = self.buckets[i]
        size = max(self.bucket_elems, need)
        if b is None or b.numel() < size or b.device != like.device:
            b = torch.empty(size, dtype=self.comm_dtype, device=like.device)
            self.buckets[i] = b
            self.out32[i] = None
        return b

    def _avg_op(self):
        if not self.average:
            return dist.ReduceOp.SUM, False
        if dist.get_backend(self.group) == "nccl":
            return dist.ReduceOp.AVG, False       # the mean is formed inside the collective
        return dist.ReduceOp.SUM, True            # gloo: sum, then scale

    def _reserve_sms(self):
        if self.sm_reserve > 0 and not self._limited and self.buckets and self.buckets[0] is not None and self.buckets[0].is_cuda:
            from . import ops
            n = torch.cuda.get_device_properties(self.buckets[0].device).multi_processor_count - self.sm_reserve
            ops.set_sm_limit(n - (n & 1))
            self._limited = True

    def _release_sms(self):
        if self._limited:
            from . import ops
            ops.set_sm_limit(0)
            self._limited = False

    def _launch(self, i: int, used: int):
        self.used[i] = used
        if not self._active():
            return
        if not self.overlap:
            self.pending.append((i, used))
            return
        self._reserve_sms()
        op, _ = self._avg_op()
        self.works.append(dist.all_reduce(self.buckets[i][:used], op=op, group=self.group, async_op=True))

    def hook(self, name: str, g: torch.Tensor):
        """Called by the trainers as soon as a gradient is final (in stream order)."""
        n = g.numel()
        if self.off > 0 and self.off + n > self.bucket_elems:
            self._launch(self.cur, self.off)
            self.cur, self.off = self.cur + 1, 0
        b = self._bucket(self.cur, g, n)
        b[self.off:self.off + n].copy_(g.reshape(-1))      # casts when the wire format is 16 bit
        self.entries.append((name, g.shape, self.cur, self.off))
        self.off += n

    def finish(self) -> Dict[str, torch.Tensor]:
        if self.off > 0:
            self._launch(self.cur, self.off)
        if self._active():
            op, scale_after = self._avg_op()
            for i, used in self.pending:
                self.works.append(dist.all_reduce(self.buckets[i][:used], op=op, group=self.group, async_op=True))
            for w in self.works:
                w.wait()
            if scale_after:
                for bi, n in self.used.items():
                    self.buckets[bi][:n].mul_(1.0 / self.world)
        self._release_sms()
        src = self.buckets
        if self.comm_dtype != torch.float32:     # hand back fp32 like the parameters: one widening pass per bucket
            for bi, n in self.used.items():
                if self.out32[bi] is None or self.out32[bi].numel() < self.buckets[bi].numel():
                    self.out32[bi] = torch.empty(self.buckets[bi].numel(), dtype=torch.float32, device=self.buckets[bi].device)
                self.out32[bi][:n].copy_(self.buckets[bi][:n])
            src = self.out32
        out = {name: src[bi][off:off + shape.numel()].view(shape) for name, shape, bi, off in self.entries}
        self.reset()
        return out
