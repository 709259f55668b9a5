#!/usr/bin/env python
"""Where does the unit disagreement of a full-size pass come from?  Runs the fp32 oracle on the GPU (tests/oracle_cuda.py)
and the CUDA path on the same inputs at a BASELINE config shape and reports, per stage, the error against the oracle,
the overall / confident unit agreement, and two isolations: our decode on the ORACLE's x0 (decode error alone) and the
oracle's decode on OUR x0 (encode + 99-call loop error alone).  Test/measurement infrastructure only."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import oracle_cuda as OC  # noqa: E402
from diffnorm_b200.engine import DiffNormEngine  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402


def err(tag, got, want, rep):
    d = (got.float() - want.float()).abs()
    rep[tag] = dict(max_abs=float(d.max()), mean_abs=float(d.mean()), ref_std=float(want.std()),
                    rel_rms=float((d.pow(2).mean().sqrt()) / want.std()))
    return d


def agreement(units, ref_units, ref_logits, mask, ltol_frac=5e-2):
    top2 = ref_logits.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])[mask]
    a = (units == ref_units)[mask]
    conf = margin > 2 * ltol_frac * float(ref_logits.std())
    return dict(overall=float(a.float().mean()), flips=int((~a).sum()), frames=int(a.numel()),
                confident=float(a[conf].float().mean()), n_confident=int(conf.sum()),
                flip_margin_over_sigma_max=float((margin[~a] / ref_logits.std()).max()) if (~a).any() else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--z", type=int, default=16)
    ap.add_argument("--start", type=int, default=100)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    dev = "cuda"
    arch = O.Arch(latent_dim=a.z)
    sd = O.init_state_dict(arch, seed=a.seed, gains=O.PARITY_GAINS)
    eng = DiffNormEngine(sd, dev)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    c = OC.case_inputs(a.z, a.batch, a.frames)
    B, T, z = a.batch, a.frames, a.z
    mask = c["mask"].to(dev)
    feat, ev, eq = c["feat"].to(dev), c["eps_vae"].to(dev), c["eps_q"].to(dev)
    lens = c["lens"].to(torch.int32).to(dev)
    keep = sorted({a.start - 1, a.start // 2, 1})
    ref = OC.oracle_pass(sdg, arch, feat, mask, a.start, ev, eq, keep_steps=keep, chunk=a.chunk)
    rep = {"shape": f"B {B} x T {T}, z {z}, start_step {a.start}, valid frames {int(mask.sum())}", "mode": eng.precision_mode()
           if hasattr(eng, "precision_mode") else "bf16"}
    out = eng.normalize(feat, lens, a.start, ev, eq, collect=True, reduce=False)
    err("z", out["z"][mask], ref["z"][mask], rep)
    err("x_start", out["x_start"][mask], ref["x_start"][mask], rep)
    err("x0", out["x0"].view(B, T, z)[mask], ref["x0"][mask], rep)
    err("recon", out["recon"][mask], ref["recon"][mask], rep)
    err("logits", out["logits"][..., :arch.vocab][mask], ref["logits"][mask], rep)
    rep["agreement_pass"] = agreement(out["units"], ref["units"], ref["logits"], mask)
    # one denoiser call on the oracle's own input at three depths of the loop
    for t in keep:
        xb = eng.stage_latent(ref["x_at"][t])
        t_idx = torch.tensor([t], dtype=torch.int32, device=dev)
        eh = eng.denoise(xb, lens, B, T, t_idx).view(B, T, -1)[..., :z]
        err(f"eps_hat_t{t}", eh[mask], ref["eps_at"][t][mask], rep)
    # isolation 1: our decode on the oracle's x0
    xb = eng.stage_latent(ref["x0"])
    recon, logits = eng.decode(xb, lens, B, T)
    err("decode_only_logits", logits[..., :arch.vocab][mask], ref["logits"][mask], rep)
    units = torch.argmax(logits[..., :arch.vocab], -1) - O.UNIT_OFFSET
    rep["agreement_decode_only"] = agreement(units, ref["units"], ref["logits"], mask)
    # isolation 2: the oracle's decode on our x0
    _, lg = OC.oracle_decode(sdg, arch, out["x0"].view(B, T, z).clone(), mask, chunk=a.chunk)
    err("loop_only_logits", lg[mask], ref["logits"][mask], rep)
    rep["agreement_loop_only"] = agreement(torch.argmax(lg, -1) - O.UNIT_OFFSET, ref["units"], ref["logits"], mask)
    # isolation 3: the 99-call loop alone, started from the ORACLE's x_start, decoded by the oracle
    from diffnorm_b200 import ops
    i32 = torch.int32
    x = eng.buf("s.x", B * T, z, torch.float32)
    x.copy_(ref["x_start"].reshape(B * T, z))
    eng.stage_latent(ref["x_start"])
    t_idx = eng.buf("s.t", 1, 1, i32, frames=False).view(-1)
    eng.buf("s.len", B, 1, i32, frames=False).view(-1).copy_(lens)
    t_idx.fill_(a.start - 1)
    for _ in range(a.start - 1):
        eng._ddim_step(B, T)
    x0p = x.view(B, T, z).clone()
    err("pure_loop_x0", x0p[mask], ref["x0"][mask], rep)
    _, lg = OC.oracle_decode(sdg, arch, x0p, mask, chunk=a.chunk)
    err("pure_loop_logits", lg[mask], ref["logits"][mask], rep)
    rep["agreement_pure_loop"] = agreement(torch.argmax(lg, -1) - O.UNIT_OFFSET, ref["units"], ref["logits"], mask)
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
