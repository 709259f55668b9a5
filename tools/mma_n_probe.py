#!/usr/bin/env python
"""How long does a tcgen05 MMA with N = 128 take relative to N = 256?  L2-resident A [16384 x 2048], 8 N tiles of 256 weight
rows; the same launch with every K segment flagged n_mma = 128 multiplies only the first 128 rows of each tile."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import _lib  # noqa: E402
from diffnorm_b200.ops import GemmPlan  # noqa: E402

M, K, NT = 16384, 2048, 8
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(NT * 256, K, device="cuda") * 0.02).bfloat16()
out = torch.empty(M, NT * 256, dtype=torch.bfloat16, device="cuda")
for n_mma in (0, 128):
    plan = GemmPlan(W, [(0, 0, K // 64, 0, n_mma)], NT * 256, NT, _lib.EPI_BF16, name=f"probe{n_mma}")
    for impl, nm in ((_lib.GEMM_TCGEN05, "1cta"), (_lib.GEMM_TCGEN05_2CTA, "2cta")):
        for _ in range(3):
            plan.run(A, out, 1, M, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.run(A, out, 1, M, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        n = 128 if n_mma else 256
        print(f"N per MMA {n:3d} {nm}: {us:7.1f} us  {2.0 * M * n * NT * K / us / 1e6:7.1f} TFLOP/s useful")
