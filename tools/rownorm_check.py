"""dn_gemm_resid_norm vs the un-fused pair (EPI_RESID dn_gemm + dn_adarmsnorm): values and time.
    python tools/rownorm_check.py            # correctness at small / ragged M, then timing at M = 64000
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffnorm_b200 import _lib, ops  # noqa: E402
from diffnorm_b200.packing import pack_linear  # noqa: E402

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32


def case(M, K, bias, cond, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    W = torch.randn(512, K, generator=g) / K ** 0.5
    b = torch.randn(512, generator=g) if bias else None
    plan = pack_linear(W, b, epi=_lib.EPI_RESID, name="chk").to(dev)
    Kp = plan.W.shape[1]
    A = torch.zeros(M, Kp, dtype=bf16, device=dev)
    A[:, :K] = torch.randn(M, K, generator=g).to(dev)
    x0 = (torch.randn(M, 512, generator=g) * 3).to(dev)
    table = torch.randn(7, 1024, generator=g).to(dev)
    table[:, :512] += 1.0
    t_idx = torch.tensor([3], dtype=torch.int32, device=dev)
    gamma_p = (torch.randn(512, generator=g) * 0.1 + 1).to(dev)
    kw_n = dict(gamma_p=None, gb=table, gb_t_stride=1024, t_idx=t_idx) if cond else dict(gamma_p=gamma_p)
    return plan, A, x0, kw_n


def pair(plan, A, x, hb, kw):
    plan.run(A, x, 1, x.shape[0])
    ops.adarmsnorm(x, hb, 1, x.shape[0], kw.get("gamma_p"), kw.get("gb"), kw.get("gb_t_stride", 0), kw.get("t_idx"), 0)


def fused(plan, A, x, hb, kw):
    ops.gemm_resid_norm(plan, A, x, hb, **kw)


def check():
    ok = True
    for M, K, bias, cond in [(128, 512, False, True), (1000, 512, False, True), (12000, 1365, True, True), (777, 1365, True, False),
                             (64000, 512, False, True)]:
        plan, A, x0, kw = case(M, K, bias, cond, seed=M)
        x1, x2 = x0.clone(), x0.clone()
        h1 = torch.zeros(M, 512, dtype=bf16, device=dev)
        h2 = torch.zeros_like(h1)
        pair(plan, A, x1, h1, kw)
        fused(plan, A, x2, h2, kw)
        torch.cuda.synchronize()
        xe = (x1 != x2).sum().item()
        dh = (h1.float() - h2.float()).abs()
        tol = h1.float().abs() * 2 ** -7 + 1e-6         # one bf16 ulp
        hbad = (dh > tol).sum().item()
        hdiff = (h1 != h2).float().mean().item()
        print(f"M={M} K={K} bias={bias} cond={cond}: x mismatches {xe}, hb beyond 1 ulp {hbad}, hb differing {hdiff:.2e}, "
              f"max |dh| {dh.max().item():.3e}", flush=True)
        ok &= xe == 0 and hbad == 0
    return ok


def bench():
    for K in (512, 1365):
        plan, A, x0, kw = case(64000, K, K != 512, True)
        x = x0.clone()
        hb = torch.zeros(64000, 512, dtype=bf16, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for name, fn in (("pair", pair), ("fused", fused)):
            ts = []
            for it in range(13):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(plan, A, x, hb, kw)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            print(f"K={K} {name}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us", flush=True)


if __name__ == "__main__":
    good = check()
    bench()
    print("ROWNORM_OK" if good else "ROWNORM_MISMATCH")
    sys.exit(0 if good else 1)
