#!/usr/bin/env python
"""CTA-pair (cta_group::2) GEMM vs the single-CTA kernel: bit-compare outputs and time both at the bench shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import _lib, packing  # noqa: E402

DEV = "cuda"
bf16, f32 = torch.bfloat16, torch.float32


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


def run(plan, A, out_shape, dt, B, T, impl, iters=0, **kw):
    torch.manual_seed(0)
    out = torch.randn(out_shape, device=DEV).to(dt).contiguous()
    plan.run(A, out, B, T, impl=impl, **kw)
    torch.cuda.synchronize()
    ms = None
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            plan.run(A, out, B, T, impl=impl, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
    return out, ms


def compare(name, plan, A, out_shape, dt, B, T, iters=0, resid=False, **kw):
    o1, t1 = run(plan, A, out_shape, dt, B, T, _lib.GEMM_TCGEN05, iters=0 if resid else iters, **kw)
    o2, t2 = run(plan, A, out_shape, dt, B, T, _lib.GEMM_TCGEN05_2CTA, iters=0 if resid else iters, **kw)
    d = (o1.float() - o2.float()).abs().max().item()
    print(f"{name:28s} max|d| {d:.3e} equal={torch.equal(o1, o2)}" + (f"  1cta {t1*1e3:.1f} us  2cta {t2*1e3:.1f} us" if t1 else ""), flush=True)
    return d


small = os.environ.get("SMALL", "1") == "1"
# ---- small correctness cases (ragged T, partial tiles)
B, T = 3, 300
A = rnd(B * T, 192, seed=1).bfloat16()
W, b = rnd(272, 192, seed=2, scale=0.1), rnd(272, seed=3)
for epi, dt in ((_lib.EPI_BF16, bf16), (_lib.EPI_F32, f32), (_lib.EPI_RESID, f32)):
    compare(f"linear epi{epi}", packing.pack_linear(W.cpu(), b.cpu(), epi=epi).to(DEV), A, (B * T, 272), dt, B, T)
x = rnd(B * T, 128, seed=4).bfloat16()
Wc, bc = rnd(144, 128, 3, seed=5, scale=0.1), rnd(144, seed=6)
compare("conv3 dil4", packing.pack_conv3(Wc.cpu(), bc.cpu(), dilation=4).to(DEV), x, (B * T, 144), bf16, B, T)
Wg, bg = rnd(400, 128, seed=7, scale=0.1), rnd(400, seed=8)
compare("geglu", packing.pack_geglu(Wg.cpu(), bg.cpu()).to(DEV), x, (B * T, 256), bf16, B, T)
G, C = 3, 256
xw = rnd(B * T, C, seed=9).bfloat16()
lv = packing.pack_wavenet_level([rnd(C, C, 3, seed=10 + g, scale=0.05).cpu() for g in range(G)], [rnd(C, seed=20 + g).cpu() for g in range(G)],
                                [rnd(C, C, 1, seed=30 + g, scale=0.05).cpu() for g in range(G)], [rnd(C, seed=40 + g).cpu() for g in range(G)], C).to(DEV)
compare("wavenet level", lv, xw, (B * T, G * C), bf16, B, T, g_a_col=0, g_out_col=C)
if not small:
    B, T = 64, 1000
    M = B * T
    ip = 1408
    m1 = rnd(M, ip, seed=50).bfloat16()
    compare("ff.conv 1408x4224", packing.pack_conv3(rnd(1365, 1365, 3, seed=51, scale=0.02).cpu(), rnd(1365, seed=52).cpu(), cin_pad=ip, n_pad=ip).to(DEV),
            m1, (M, ip), bf16, B, T, iters=20)
    hb = rnd(M, 512, seed=53).bfloat16()
    compare("ff.geglu 512->2730", packing.pack_geglu(rnd(2730, 512, seed=54, scale=0.04).cpu(), rnd(2730, seed=55).cpu()).to(DEV), hb, (M, ip), bf16, B, T, iters=20)
    compare("qkv 512->1536", packing.pack_linear(rnd(1536, 512, seed=56, scale=0.04).cpu(), None).to(DEV), hb, (M, 1536), bf16, B, T, iters=20)
    G, C = 8, 512
    y = rnd(M, G * C, seed=57).bfloat16()
    lv = packing.pack_wavenet_level([rnd(C, C, 3, seed=60 + g, scale=0.03).cpu() for g in range(G)], [rnd(C, seed=70 + g).cpu() for g in range(G)],
                                    [rnd(C, C, 1, seed=80 + g, scale=0.03).cpu() for g in range(G)], [rnd(C, seed=90 + g).cpu() for g in range(G)], C).to(DEV)
    compare("wavenet level 8x512", lv, y, (M, G * C), bf16, B, T, iters=10, g_a_col=C, g_out_col=C)
    compare("ff.out 1408->512 resid", packing.pack_linear(rnd(512, 1365, seed=95, scale=0.03).cpu(), rnd(512, seed=96).cpu(), epi=_lib.EPI_RESID, k_pad=ip).to(DEV),
            m1, (M, 512), f32, B, T, resid=True)
if not small:
    # is the WaveNet level epilogue-bound?  same MMAs, plain +bias epilogue (the training step's un-fused form)
    from diffnorm_b200.ops import GemmPlan
    tiles = C // 128
    bi = torch.stack([lv.bias.view(G, tiles, 128), lv.bias2.view(G, tiles, 128)], dim=2).reshape(-1).contiguous()
    plain = GemmPlan(lv.W, lv.segs, 2 * C, tiles, _lib.EPI_BF16, bias=bi, groups=G, g_w_row=lv.g_w_row, g_bias=2 * C, dilation=1,
                     dilation_shl_group=1, name="wn.plain")
    compare("wavenet level, +bias only", plain, y, (M, G * 2 * C), bf16, B, T, iters=10, g_a_col=C, g_out_col=2 * C)
    nob = GemmPlan(lv.W, lv.segs, 2 * C, tiles, _lib.EPI_BF16, bias=None, groups=G, g_w_row=lv.g_w_row, g_bias=2 * C, dilation=1,
                   dilation_shl_group=1, name="wn.nobias")
    compare("wavenet level, no bias", nob, y, (M, G * 2 * C), bf16, B, T, iters=10, g_a_col=C, g_out_col=2 * C)
