#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[3] on the GPU(s) of one node.

  --what sweep    noise-ratio x sampler sweep: ratio {0.25, 0.5, 0.75} -> start_step {50, 100, 150} with the reference
                  sampler (DDIM eta=0, stride 1), plus at ratio 0.5 ancestral DDPM (99 calls) and strided DDIM
                  (25 calls), over length buckets T in {200, 500, 1000, 2000} with B*T ~ 64k.
  --what dataset  dataset-scale normalization of `--utts` synthetic variable-length utterances
                  (N_i ~ round(exp(N(ln 600, 0.5^2))) clipped to [200, 2000], seed 1234), length-bucketed
                  under a 64k padded-frame budget and sharded by utterance over the ranks (torchrun for > 1 GPU; no
                  data-path collective).  Reports valid (un-padded) frames/s = sum N_i / max-over-ranks device time.

Prints one JSON line per measurement (CUDA events; clocks not pinned).
"""
import argparse
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from diffnorm_b200 import data  # noqa: E402
from diffnorm_b200.plugin.latent_module import LatentDiscreteModel, SpeechVAEEncoderDecoder  # noqa: E402


def build_engine(z, dev):
    torch.manual_seed(0)
    vae = types.SimpleNamespace(encoder=SpeechVAEEncoderDecoder(768, z))
    ldm = LatentDiscreteModel(vae, 512, z, timesteps=200, multitask=False).to(dev).eval()
    return ldm._engine()


def timed(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def sweep(a, eng, dev):
    for T in (200, 500, 1000, 2000):
        B = max(1, 64000 // T)
        feat = torch.randn(B, T, 768, device=dev)
        lens = torch.full((B,), T, dtype=torch.int32, device=dev)
        cases = [("ddim", 50, None), ("ddim", 100, None), ("ddim", 150, None), ("ddpm", 100, None),
                 ("ddim_strided", 100, list(range(0, 100, 4)))]
        for sampler, start, keep in cases:
            def fn():
                return eng.normalize(feat, lens, start, sampler=sampler, timesteps=keep)
            calls = fn()["calls"]
            sec = timed(fn, 1, a.iters)
            print(json.dumps({"config": "sweep", "T": T, "B": B, "ratio": start / 200, "sampler": sampler,
                              "denoiser_calls": calls, "ms_per_pass": sec * 1e3, "frames_per_s": B * T / sec,
                              "frames_per_s_per_call_x99": B * T / sec * calls / 99}), flush=True)


def dataset(a, eng, dev, rank, world):
    import torch.distributed as dist
    rng = np.random.default_rng(1234)
    n = np.clip(np.rint(np.exp(rng.normal(np.log(600.0), 0.5, size=a.utts))), 200, 2000).astype(np.int64)
    plan = data.plan_batches(n, a.max_tokens, world_size=world, pad_multiple=a.pad_multiple)[rank]
    eng.reserve(a.max_tokens)
    # features are generated on the host per batch (pinned) and copied in the timed region: the reduced features the
    # pass consumes, N_i x 768 fp32 each
    my_frames = int(sum(int(n[idx].sum()) for idx in plan))
    padded = int(sum(len(idx) * int(n[idx].max()) for idx in plan))
    shapes = sorted({(len(idx), int(n[idx].max())) for idx in plan})
    host = torch.randn(a.max_tokens * 768, generator=torch.Generator().manual_seed(rank)).pin_memory()

    def run_all(limit=None):
        tot = 0
        for k, idx in enumerate(plan):
            if limit is not None and k >= limit:
                break
            B, T = len(idx), int(n[idx].max())
            lens = torch.from_numpy(n[idx].astype(np.int32)).to(dev, non_blocking=True)
            feat = host[: B * T * 768].view(B, T, 768).to(dev, non_blocking=True)
            out = eng.normalize(feat, lens, a.start_step)
            tot += int(out["counts"].sum().item())  # D2H of the result (forces completion, like the TSV writer)
        return tot

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_all()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    wall = time.perf_counter() - t0
    stats = torch.tensor([sec, wall, my_frames, padded, len(plan), len(shapes)], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
    else:
        gathered = [stats]
    if rank == 0:
        g = torch.stack(gathered).cpu().numpy()
        tmax = float(g[:, 0].max())
        print(json.dumps({"config": "dataset", "utterances": a.utts, "n_gpus": world, "start_step": a.start_step,
                          "valid_frames": int(g[:, 2].sum()), "padded_frames": int(g[:, 3].sum()),
                          "batches": int(g[:, 4].sum()), "distinct_shapes_per_rank": [int(x) for x in g[:, 5]],
                          "device_s_per_rank": [round(float(x), 3) for x in g[:, 0]], "max_device_s": tmax,
                          "valid_frames_per_s": float(g[:, 2].sum()) / tmax,
                          "padded_frames_per_s": float(g[:, 3].sum()) / tmax,
                          "note": "includes the pinned-host H2D of every batch's features; (B,T) shapes seen once run eager, a "
                                  "shape that comes back is captured as a CUDA graph (engine.graph_after)"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="sweep", choices=["sweep", "dataset"])
    ap.add_argument("--latent-dim", type=int, default=16)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--utts", type=int, default=20000)
    ap.add_argument("--max-tokens", type=int, default=64000)
    ap.add_argument("--pad-multiple", type=int, default=8)
    ap.add_argument("--start-step", type=int, default=100)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    eng = build_engine(a.latent_dim, dev)
    with torch.no_grad():
        if a.what == "sweep":
            if rank == 0:
                sweep(a, eng, dev)
        else:
            dataset(a, eng, dev, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
