#!/usr/bin/env python
"""Which rounding drives the accumulated error of the 99-call loop and of the once-per-pass encode / decode?  Emulates
operand formats inside the fp32 torch oracle on the GPU: weights rounded once to bf16 / fp16 (a FIXED perturbation, the
same at every step) and / or the inputs of every linear / conv rounded on the fly (a fresh perturbation per step), then
runs the loop from the exact x_start and decodes with the exact decoder.  Design evidence for the operand formats of
diffnorm_b200 (DESIGN.md "operand formats"); measurement infrastructure only."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

import oracle_cuda as OC  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

FMT = {"f32": None, "bf16": torch.bfloat16, "f16": torch.float16}


def rnd(t, fmt):
    return t if fmt is None else t.to(fmt).float()


def round_weights(sd, fmt, pre):
    out = {}
    for k, v in sd.items():
        is_w = k.startswith(pre) and v.dim() >= 2 and "to_time_cond" not in k and "to_gamma_beta" not in k
        out[k] = rnd(v, fmt) if is_w else v
    return out


class ActRounding:
    """Patch the oracle's F.linear / F.conv1d so their activation operand is rounded to `fmt` (what a 16-bit GEMM operand
    staging does); attention operands q, k, v are linear outputs and are rounded by the einsum patch."""

    def __init__(self, fmt):
        self.fmt = fmt

    def __enter__(self):
        fmt = self.fmt
        self.F, self.einsum = O.F, torch.einsum
        if fmt is None:
            return self
        shim = types.SimpleNamespace(**{n: getattr(TF, n) for n in ("pad", "normalize", "gelu", "silu", "log_softmax")})
        shim.linear = lambda x, w, b=None: TF.linear(rnd(x, fmt) if x.dim() == 3 else x, w, b)
        shim.conv1d = lambda x, w, b=None, **kw: TF.conv1d(rnd(x, fmt), w, b, **kw)
        O.F = shim
        torch.einsum = lambda eq, a, b: self.einsum(eq, rnd(a, fmt), rnd(b, fmt))
        return self

    def __exit__(self, *a):
        O.F, torch.einsum = self.F, self.einsum


def rel(got, want):
    return float(((got - want).pow(2).mean().sqrt()) / want.std())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--start", type=int, default=100)
    a = ap.parse_args()
    dev, z = ("cuda" if torch.cuda.is_available() else "cpu"), 16
    arch = O.Arch(latent_dim=z)
    sd = {k: v.to(dev) for k, v in O.init_state_dict(arch, seed=1, gains=O.PARITY_GAINS).items()}
    c = OC.case_inputs(z, a.batch, a.frames)
    mask = c["mask"].to(dev)
    feat, ev, eq = c["feat"].to(dev), c["eps_vae"].to(dev), c["eps_q"].to(dev)
    ref = OC.oracle_pass(sd, arch, feat, mask, a.start, ev, eq)
    sch = O.Schedule(arch.timesteps)

    def flips(logits):
        u = torch.argmax(logits, -1) - O.UNIT_OFFSET
        return int((u != ref["units"])[mask].sum())

    n = int(mask.sum())
    rows = []
    with torch.no_grad(), OC.strict_fp32():
        for wf, af in (("bf16", "f32"), ("f32", "bf16"), ("bf16", "bf16"), ("f16", "bf16"), ("f16", "f16")):
            sdw = round_weights(sd, FMT[wf], "model.")
            x = ref["x_start"].clone()
            with ActRounding(FMT[af]):
                for t in range(a.start - 1, 0, -1):
                    tt = torch.full((a.batch,), t, dtype=torch.long, device=dev)
                    x = O.ddim_step(sch, x, O.denoiser(sdw, arch, x, tt, mask), t)
            _, lg = O.vae_decode(sd, arch, x, mask)
            rows.append(dict(stage="loop", weights=wf, acts=af, x0_rel_rms=rel(x[mask], ref["x0"][mask]),
                             logits_rel_rms=rel(lg[mask], ref["logits"][mask]), flips=flips(lg), frames=n))
            print(json.dumps(rows[-1]), flush=True)
        for wf, af in (("bf16", "bf16"), ("f16", "bf16"), ("f16", "f16")):
            sdw = round_weights(sd, FMT[wf], "speech_decoder.")
            with ActRounding(FMT[af]):
                _, lg = O.vae_decode(sdw, arch, ref["x0"], mask)
            rows.append(dict(stage="decode", weights=wf, acts=af, logits_rel_rms=rel(lg[mask], ref["logits"][mask]),
                             flips=flips(lg), frames=n))
            print(json.dumps(rows[-1]), flush=True)
            with ActRounding(FMT[af]):
                zz = O.vae_encode(sdw, arch, feat, ev)
            rows.append(dict(stage="encode", weights=wf, acts=af, z_rel_rms=rel(zz[mask], ref["z"][mask])))
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
