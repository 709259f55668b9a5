#!/usr/bin/env python
"""Achieved HBM bandwidth of every HBM-bound kernel on the path at the bench shape (B 64 x T 1000 = 64,000 frames),
against the measured copy peak in MEASURED_PEAKS.json.  Algorithmic bytes per frame follow SURVEY.md §8(d) /
DESIGN.md §4 (ideal single pass: every input read once, every output written once).  Each kernel is timed alone
with CUDA events over `--iters` back-to-back launches after warm-up; the working sets (65 MB .. 390 MB per launch,
rotated over several buffers where they would fit the 126 MB L2) exceed L2.

    python tools/hbm_kernels.py [--batch 64 --frames 1000] > profiles/rNN_hbm_kernels.txt
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import ops  # noqa: E402

bf16, f32, i32, i64 = torch.bfloat16, torch.float32, torch.int32, torch.int64


def timeit(fns, iters):
    """fns: list of closures over distinct buffers (rotated so consecutive launches do not hit L2)."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--rot", type=int, default=4, help="distinct buffer sets rotated between launches")
    a = ap.parse_args()
    B, T = a.batch, a.frames
    M = B * T
    dev = "cuda"
    peaks = {"hbm_gbs": 6650.0}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    peak = peaks["hbm_gbs"]
    rows = []

    def rec(name, bytes_per_frame, sec, note=""):
        gbs = bytes_per_frame * M / sec / 1e9
        rows.append((name, bytes_per_frame, sec * 1e6, gbs, gbs / peak, gbs / 8000.0, note))

    R = a.rot
    t_idx = torch.tensor([50], dtype=i32, device=dev)
    lens = torch.full((B,), T, dtype=i32, device=dev)

    # E1 adaptive RMSNorm, C = 512 (denoiser) and 768 (VAE decoder): read fp32, write bf16
    for C in (512, 768):
        xs = [torch.randn(M, C, device=dev) for _ in range(R)]
        os_ = [torch.empty(M, C, dtype=bf16, device=dev) for _ in range(R)]
        if C == 512:
            gb = torch.randn(200, 2 * C, device=dev)
            fns = [lambda x=x, o=o: ops.adarmsnorm(x, o, B, T, None, gb.view(-1), 2 * C, t_idx, 0) for x, o in zip(xs, os_)]
        else:
            gp = torch.randn(C, device=dev)
            fns = [lambda x=x, o=o: ops.adarmsnorm(x, o, B, T, gp) for x, o in zip(xs, os_)]
        rec(f"adarmsnorm C={C}", C * 4 + C * 2, timeit(fns, a.iters), "fp32 in, bf16 out")
        del xs, os_

    # E2 stand-alone WaveNet gate, C = 512: bf16 u, res in; bf16 y out
    C = 512
    us = [torch.randn(M, C, device=dev).to(bf16) for _ in range(R)]
    rs = [torch.randn(M, C, device=dev).to(bf16) for _ in range(R)]
    ys = [torch.empty(M, C, dtype=bf16, device=dev) for _ in range(R)]
    gb = torch.randn(200, 2 * C, device=dev)
    fns = [lambda u=u, r=r, y=y: ops.wavenet_gate(u, r, y, B, T, gb.view(-1), 2 * C, t_idx, 0) for u, r, y in zip(us, rs, ys)]
    rec("wavenet_gate C=512", 3 * C * 2, timeit(fns, a.iters), "bf16 u,res in; bf16 out (engine uses the fused GEMM epilogue)")
    del us, rs, ys

    # E4/E5 sampler updates and E6 reparam at z = 128 (2 KB/frame) and z = 16 (256 B/frame)
    for z in (128, 16):
        zp = max(64, z)
        Rz = R if z == 128 else 16  # z=16: 4 MB tensors — rotate more sets, still L2 resident by nature (16 MB/launch)
        xs = [torch.randn(M, z, device=dev) for _ in range(Rz)]
        es = [torch.randn(M, z, device=dev) for _ in range(Rz)]
        ns = [torch.randn(M, z, device=dev) for _ in range(Rz)]
        xbs = [torch.zeros(M, zp, dtype=bf16, device=dev) for _ in range(Rz)]
        from diffnorm_b200.schedule import DDPMScheduler
        s = DDPMScheduler(200)
        ddim_rows = torch.from_numpy(s.ddim_rows()).to(dev)
        ddpm_rows = torch.from_numpy(s.ddpm_rows()).to(dev)
        fns = [lambda x=x, e=e, xb=xb: ops.ddim_step(x, e, ddim_rows, t_idx, 0, xb) for x, e, xb in zip(xs, es, xbs)]
        rec(f"ddim_step z={z}", (2 + 1) * z * 4 + zp * 2, timeit(fns, a.iters), "x, eps_hat in; x + bf16 staging out")
        fns = [lambda x=x, e=e, n=n, xb=xb: ops.ddpm_step(x, e, n, ddpm_rows, t_idx, xb) for x, e, n, xb in zip(xs, es, ns, xbs)]
        rec(f"ddpm_step z={z}", (3 + 1) * z * 4 + zp * 2, timeit(fns, a.iters), "x, eps_hat, noise in; x + bf16 staging out")
        fns = [lambda x=x, e=e, n=n, xb=xb: ops.q_sample(e, n, 0.7, 0.7, x, xb) for x, e, n, xb in zip(xs, es, ns, xbs)]
        rec(f"q_sample z={z}", (2 + 1) * z * 4 + zp * 2, timeit(fns, a.iters), "z, eps in; x + bf16 staging out")
        ps = [torch.randn(B, T, 2 * z, device=dev) for _ in range(Rz)]
        eps = [torch.randn(B, z, T, device=dev) for _ in range(Rz)]
        outs = [torch.empty(B, T, z, device=dev) for _ in range(Rz)]
        fns = [lambda p_=p_, e=e, o=o: ops.vae_reparam(p_, e, z, True, o) for p_, e, o in zip(ps, eps, outs)]
        rec(f"vae_reparam z={z}", 2 * z * 4 + z * 4 + z * 4, timeit(fns, a.iters), "params + channel-first eps in; z out")
        del xs, es, ns, xbs, ps, eps, outs

    # A1 argmax over 1004 (ld 1008) fp32 logits
    ls = [torch.randn(M, 1008, device=dev) for _ in range(2)]
    uo = [torch.empty(M, dtype=i64, device=dev) for _ in range(2)]
    fns = [lambda l=l, o=o: ops.argmax_units(l, 1004, 4, o) for l, o in zip(ls, uo)]
    rec("argmax_units 1004 fp32", 1004 * 4 + 8, timeit(fns, a.iters), "fp32 logits in, int64 unit out")
    del ls

    # R1 run-length reduce (int64 units in; int64 dedup/duration/index out, geometric run lengths p = 0.6)
    g = torch.Generator(device="cpu").manual_seed(0)
    runs = torch.randint(0, 1000, (B, T), generator=g)
    keep = torch.rand(B, T, generator=g) < 0.6
    idx = torch.cumsum(keep.long(), 1).clamp_(max=T - 1)
    units = torch.gather(runs, 1, idx).to(dev).contiguous()
    fns = [lambda: ops.reduce_tgt(units, lens)]
    sec = timeit(fns, a.iters)
    rec("reduce_tgt", 8 + 24 * 0.6, sec, "latency-bound at this size (0.5 MB in): one block per utterance")

    # P1 gather + pad 768-d fp32 rows
    src = torch.randn(int(M * 1.67), 768, device=dev)
    row0 = (torch.arange(B, dtype=i64) * int(T * 1.67)).to(dev)
    ik = torch.sort(torch.randperm(int(T * 1.67))[:T])[0].to(dev).repeat(B, 1).contiguous()
    cnt = torch.full((B,), T, dtype=i32, device=dev)
    fns = [lambda: ops.gather_pack(src, row0, ik, cnt, T)]
    rec("gather_pack 768 fp32->fp32", 768 * 4 * 2, timeit(fns, a.iters), "includes torch.empty of the 197 MB destination")

    # cast + pad (operand staging)
    xs = [torch.randn(M, 768, device=dev) for _ in range(2)]
    os_ = [torch.empty(M, 768, dtype=bf16, device=dev) for _ in range(2)]
    fns = [lambda x=x, o=o: ops.cast_pad_bf16(x, 768, out=o) for x, o in zip(xs, os_)]
    rec("cast_pad_bf16 C=768", 768 * 6, timeit(fns, a.iters), "fp32 in, bf16 out")

    print(f"# HBM-bound kernels at B {B} x T {T} = {M} frames; peak = {peak:.1f} GB/s measured copy "
          f"(MEASURED_PEAKS.json), spec 8000 GB/s; {a.iters} launches each, CUDA events")
    print(f"{'kernel':28s} {'B/frame':>9s} {'us/launch':>10s} {'GB/s':>9s} {'%meas':>7s} {'%spec':>7s}  note")
    for n, b, us_, gbs, f1, f2, note in rows:
        print(f"{n:28s} {b:9.0f} {us_:10.1f} {gbs:9.1f} {100 * f1:6.1f}% {100 * f2:6.1f}%  {note}")


if __name__ == "__main__":
    main()
