#!/usr/bin/env python
"""Second, fairer baseline: the reference algorithm as plain PyTorch fp32 eager ON THE B200 — the oracle
port (oracle/diffnorm_oracle.py, the functional restatement pinned to the live reference) with weights and inputs moved
to cuda:0.  Same workload as bench.py (config 2); the denoiser loop is timed over `--calls` calls and extrapolated to 99
(one call materialises the 64 x 8 x 1000 x 1000 attention matrix per layer, as the reference does).  Test/bench
infrastructure only: nothing in diffnorm_b200/ imports it."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import diffnorm_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--tf32", action="store_true")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    torch.backends.cudnn.allow_tf32 = a.tf32
    dev = "cuda"
    z, B, T, start = 16, a.batch, a.frames, 100
    arch = O.Arch(latent_dim=z)
    sd = {k: v.to(dev) for k, v in O.init_state_dict(arch, seed=0).items()}
    g = torch.Generator().manual_seed(1234)
    feat = torch.randn(B, T, 768, generator=g).to(dev)
    mask = torch.ones(B, T, dtype=torch.bool, device=dev)
    ev, eq = torch.randn(B, z, T, generator=g).to(dev), torch.randn(B, T, z, generator=g).to(dev)
    sch = O.Schedule(arch.timesteps)
    # the oracle builds its positional table on the CPU; keep that semantics but place tensors on the device
    orig_pe = O.pos_embed
    O.pos_embed = lambda m, dim, dtype=torch.float32: orig_pe(m.cpu(), dim, dtype).to(m.device)

    def ev_time(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return r, e0.elapsed_time(e1) * 1e-3

    with torch.no_grad():
        O.vae_encode(sd, arch, feat[:2, :64], ev[:2, :, :64])  # warm-up (cuDNN / cuBLAS init)
        zl, t_enc = ev_time(lambda: O.vae_encode(sd, arch, feat, ev))
        x = O.q_sample(sch, zl, start, eq)
        tt = torch.full((B,), start - 1, dtype=torch.long, device=dev)
        O.denoiser(sd, arch, x, tt, mask)  # warm-up
        t_calls = 0.0
        for k in range(a.calls):
            t = start - 1 - k
            tt = torch.full((B,), t, dtype=torch.long, device=dev)
            eh, dt = ev_time(lambda: O.denoiser(sd, arch, x, tt, mask))
            x = O.ddim_step(sch, x, eh, t)
            t_calls += dt
        (rec, logits), t_dec = ev_time(lambda: O.vae_decode(sd, arch, x, mask))
    per_call = t_calls / a.calls
    total = t_enc + per_call * (start - 1) + t_dec
    print(json.dumps({"baseline": "oracle port, PyTorch eager on cuda:0, " + ("tf32" if a.tf32 else "fp32"),
                      "workload": f"B {B} x T {T}, z {z}, 99 calls (timed {a.calls}, extrapolated)", "encode_s": t_enc,
                      "denoiser_call_s": per_call, "decode_s": t_dec, "pass_s": total, "frames_per_s": B * T / total,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
