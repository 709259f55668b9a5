#!/usr/bin/env python
"""Per-op device-time breakdown of ONE eager denoiser call (CUDA events around every C-ABI launch) at the bench
shape.  Diagnostic only: event bracketing serialises launches, so compare shares, not absolutes."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffnorm_b200 import ops  # noqa: E402
from diffnorm_b200.engine import DiffNormEngine  # noqa: E402
from diffnorm_b200.plugin.latent_module import LatentDiscreteModel, SpeechVAEEncoderDecoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--latent-dim", type=int, default=16)
    ap.add_argument("--what", default="denoise", choices=["denoise", "decode", "encode"])
    a = ap.parse_args()
    import types
    torch.manual_seed(0)
    vae = types.SimpleNamespace(encoder=SpeechVAEEncoderDecoder(768, a.latent_dim))
    ldm = LatentDiscreteModel(vae, 512, a.latent_dim, timesteps=200, multitask=False).cuda().eval()
    eng: DiffNormEngine = ldm._engine()
    B, T = a.batch, a.frames
    lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
    xb = eng.buf("s.xb", B * T, eng.xw)
    t_idx = torch.tensor([50], dtype=torch.int32, device="cuda")
    feat = torch.randn(B, T, 768, device="cuda")
    eps = torch.randn(B, a.latent_dim, T, device="cuda")
    fn = {"denoise": lambda: eng.denoise(xb, lens, B, T, t_idx), "decode": lambda: eng.decode(xb, lens, B, T),
          "encode": lambda: eng.encode(feat, eps)}[a.what]
    fn()
    torch.cuda.synchronize()
    events = []

    def wrap(name, f):
        def g(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*args, **kw)
            e1.record()
            events.append((name, e0, e1))
            return r
        return g

    for n in ("adarmsnorm", "attention", "ddim_step", "cast_pad_bf16", "vae_reparam", "argmax_units"):
        setattr(ops, n, wrap(n, getattr(ops, n)))
    orig_run = ops.GemmPlan.run

    def run(self, *args, **kw):
        import re
        key = "gemm:" + re.sub(r"\d+", "#", self.name)
        return wrap(key, lambda: orig_run(self, *args, **kw))()

    ops.GemmPlan.run = run
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    fn()
    w1.record()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for n, e0, e1 in events:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    print(f"{a.what} B={B} T={T}: sum of op times {tot:.3f} ms, wall (events) {w0.elapsed_time(w1):.3f} ms")
    for n, (k, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {ms:8.3f} ms {100 * ms / tot:5.1f}%  n={k:3d}  avg {1e3 * ms / k:8.1f} us  {n}")


if __name__ == "__main__":
    main()
