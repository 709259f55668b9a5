#!/usr/bin/env python
"""Time dn_attention alone at the bench shape (B 64 x T 1000, 8 heads x 64) with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import ops  # noqa: E402

B, T, H, dh = int(os.environ.get("B", 64)), int(os.environ.get("T", 1000)), 8, int(os.environ.get("DH", 64))
DT = torch.float16 if os.environ.get("F16", "1") == "1" else torch.bfloat16
qkvs = [torch.randn(B * T, 3 * H * dh, device="cuda").to(DT) for _ in range(3)]
out = torch.empty(B * T, H * dh, dtype=DT, device="cuda")
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
for q in qkvs:
    ops.attention(q, out, lens, B, T, H, dh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for i in range(n):
    ops.attention(qkvs[i % 3], out, lens, B, T, H, dh)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
fl = 4.0 * B * H * T * T * dh
print(f"attention B={B} T={T} H={H} dh={dh} opt={os.environ.get("DN_ATTN_OPT", "1")} {DT}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")

if os.environ.get("BWD"):
    lse = torch.empty(B * H, T, device="cuda")
    bits = torch.randint(-2**31, 2**31 - 1, (B, H, T, (T + 31) // 32), dtype=torch.int32, device="cuda")
    ops.attention_train(qkvs[0], out, lse, lens, bits, 1 / 0.9, B, T, H, dh)
    dout = torch.randn(B * T, H * dh, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkvs[0])
    dws = torch.empty(B * H, T, device="cuda")
    for keep in (bits, None):
        ops.attention_bwd(qkvs[0], out, dout, lse, lens, keep, 1 / 0.9, dqkv, dws, B, T, H, dh)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            ops.attention_bwd(qkvs[0], out, dout, lse, lens, keep, 1 / 0.9, dqkv, dws, B, T, H, dh)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"attention bwd (delta + dQ + dKdV) dropout={'on' if keep is not None else 'off'}: {us:.1f} us  {2.5 * fl / us / 1e6:.0f} TFLOP/s")
