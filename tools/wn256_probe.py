#!/usr/bin/env python
"""Upper bound for a WaveNet-level kernel whose shifted taps run as N = 256 MMAs: time the conv part (3 taps, 256-row tiles)
and the 1x1 res part as two separate grouped launches of the existing kernel at the bench shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import _lib  # noqa: E402
from diffnorm_b200.ops import GemmPlan  # noqa: E402

B, T, G, C = 64, 1000, 8, 512
M = B * T
y = torch.randn(M, G * C, device="cuda").bfloat16()
Wc = (torch.randn(G * C, 3 * C, device="cuda") * 0.02).bfloat16()
Wr = (torch.randn(G * C, C, device="cuda") * 0.02).bfloat16()
kb = C // 64
conv = GemmPlan(Wc, [(0, 0, kb, 0, 0), (0, 2, kb, C, 0), (0, 1, kb, 2 * C, 0)], C, C // 256, _lib.EPI_BF16, groups=G, g_w_row=C,
                dilation=1, dilation_shl_group=1, name="conv256")
res = GemmPlan(Wr, [(0, 0, kb, 0, 0)], C, C // 256, _lib.EPI_BF16, groups=G, g_w_row=C, name="res256")
out = torch.empty(M, G * C, dtype=torch.bfloat16, device="cuda")


def t(plan, impl):
    for _ in range(2):
        plan.run(y, out, B, T, g_a_col=C, g_out_col=C, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.run(y, out, B, T, g_a_col=C, g_out_col=C, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


for impl, nm in ((_lib.GEMM_TCGEN05, "1cta"), (_lib.GEMM_TCGEN05_2CTA, "2cta")):
    a, b = t(conv, impl), t(res, impl)
    print(f"{nm}: conv (3 taps, N=256) {a:.1f} us + res (N=256) {b:.1f} us = {a + b:.1f} us   (fused level today: ~905 us)")
