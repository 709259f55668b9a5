"""End-to-end parity of the CUDA path against the oracle / golden fixtures (B200 only).

Tolerances (bf16 operands, fp32 accumulation, fp32 residual stream / latent / logits), stated per stage:
  z (VAE encode)            max-abs <= 3e-2 * (1 + |z|)      after 3 WaveNet blocks in bf16
  eps_hat (one denoiser call) max-abs <= 5e-2 * std(eps_hat) + 2e-2 (SURVEY §8c proposes 2e-2 abs at unit scale)
  logits                    max-abs <= 5e-2 * std(logits)
  units                     >= 99.5 % agreement on frames whose fp32 top-1/top-2 margin exceeds the logit tolerance;
                            overall agreement reported (random-init logits are near-ties).
Integer post-processing (reduce) is bit-exact given identical units.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200.engine import DiffNormEngine  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def setup_case(name):
    if name in _cache:
        return _cache[name]
    _cache.clear()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    z = int(g["latent_dim"])
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=int(g["weight_seed"]), gains=O.PARITY_GAINS if int(g["parity_gains"]) else None)
    eng = DiffNormEngine(sd, DEV)
    _cache[name] = (g, arch, sd, eng)
    return _cache[name]


def stats(tag, got, want):
    d = (got - want).abs()
    print(f"[parity] {tag}: max_abs {d.max():.4e} mean_abs {d.mean():.4e} ref_std {want.std():.4e} ref_absmax {want.abs().max():.4e}")
    return d


@pytest.mark.parametrize("name", ["pass_z16_parity", "pass_z16_default", "pass_z128_parity"])
def test_stages_and_pass_against_golden(name):
    g, arch, sd, eng = setup_case(name)
    z = arch.latent_dim
    t = lambda k: torch.from_numpy(g[k])
    feat, eps_vae, eps_q = t("feat").to(DEV), t("eps_vae").to(DEV), t("eps_q").to(DEV)
    B, T, _ = feat.shape
    lens = torch.from_numpy(g["lengths"]).to(torch.int32).to(DEV)
    mask_cpu = O.lengths_to_mask(torch.from_numpy(g["lengths"]), T)
    start = int(g["start_step"])

    # --- stage: encode
    zl = eng.encode(feat, eps_vae).cpu()
    d = stats("z", zl, t("z"))
    assert d.max() <= 3e-2 * (1 + t("z").abs().max())

    # --- stage: one denoiser call on the golden x_start
    xb = eng.stage_latent(t("x_start").to(DEV))
    t_idx = torch.tensor([start - 1], dtype=torch.int32, device=DEV)
    eh = eng.denoise(xb, lens, B, T, t_idx).view(B, T, -1)[..., :z].cpu()
    want = t("eps_first")
    d = stats("eps_first", eh[mask_cpu], want[mask_cpu])
    assert d.max() <= 5e-2 * want.std() + 2e-2

    # --- stage: decode on the golden latent
    xb = eng.stage_latent(t("dec_latent").to(DEV))
    recon, logits = eng.decode(xb, lens, B, T)
    want_l = t("dec_logits")
    d = stats("dec_logits", logits.cpu()[..., : arch.vocab][mask_cpu], want_l[mask_cpu])
    tol = 5e-2 * float(want_l.std())
    assert d.max() <= tol
    stats("dec_feat", recon.cpu()[mask_cpu], t("dec_feat")[mask_cpu])

    # --- the full pass, graph and eager
    ref = O.normalize_pass(sd, arch, t("feat"), mask_cpu, start, t("eps_vae"), t("eps_q"))
    for use_graph in (False, True):
        out = eng.normalize(feat, lens, start, eps_vae, eps_q, ref_units=t("ref_units").to(DEV), use_graph=use_graph, logits=True)
        # default path: the unit head's epilogue takes the argmax itself, logits never materialised -> identical units
        fused = eng.normalize(feat, lens, start, eps_vae, eps_q, use_graph=use_graph)
        assert "logits" not in fused and torch.equal(fused["units"], out["units"])
        units = out["units"].cpu()
        gold_units = t("units")
        top2 = ref["logits"].topk(2, dim=-1).values
        margin = top2[..., 0] - top2[..., 1]
        agree = (units == gold_units)[mask_cpu]
        ltol = 5e-2 * float(ref["logits"].std())
        confident = (margin > 2 * ltol)[mask_cpu]
        stats("pass_logits", out["logits"].cpu()[..., : arch.vocab][mask_cpu], ref["logits"][mask_cpu])
        stats("pass_x0", out["x0"].cpu()[mask_cpu], ref["x0"][mask_cpu])
        print(f"[parity] {name} graph={use_graph}: unit agreement overall {agree.float().mean():.4f} "
              f"({int(agree.sum())}/{agree.numel()}), on confident frames {agree[confident].float().mean():.4f} "
              f"({int(confident.sum())} frames, margin > {2 * ltol:.4f}), margin median {margin.median():.4f}")
        assert agree[confident].float().mean() >= 0.995
        assert out["acc"].cpu().tolist()[1] == int(g["total"])
        # integer tail: the device reduce equals the oracle reduce of the SAME device units (bit-exact)
        cnt = out["counts"].cpu()
        for b in range(B):
            n = int(g["lengths"][b])
            dd, du, kp = O.reduce_tgt(units[b, :n].tolist())
            r = int(cnt[b])
            assert out["dedup"][b, :r].cpu().tolist() == dd
            assert out["duration"][b, :r].cpu().tolist() == du
            assert out["index_to_keep"][b, :r].cpu().tolist() == kp


def test_fused_resid_norm_denoiser_call_matches_unfused():
    """engine.fuse_norm (dn_gemm_resid_norm in the transformer stack) against the default un-fused stack on the same
    denoiser call: eps_hat within the bf16 rounding noise of the stack (hb differs by <= 1 bf16 ulp per norm)."""
    g, arch, sd, _ = setup_case("pass_z16_parity")
    eng = DiffNormEngine(sd, DEV, wfmt="bf16")    # dn_gemm_resid_norm multiplies bf16 weights
    t = lambda k: torch.from_numpy(g[k])
    B, T, _ = t("feat").shape
    lens = torch.from_numpy(g["lengths"]).to(torch.int32).to(DEV)
    mask_cpu = O.lengths_to_mask(torch.from_numpy(g["lengths"]), T)
    xb = eng.stage_latent(t("x_start").to(DEV))
    t_idx = torch.tensor([int(g["start_step"]) - 1], dtype=torch.int32, device=DEV)
    was = eng.fuse_norm
    try:
        eng.fuse_norm = False
        a = eng.denoise(xb, lens, B, T, t_idx).clone()
        eng.fuse_norm = True
        b = eng.denoise(xb, lens, B, T, t_idx).clone()
    finally:
        eng.fuse_norm = was
    z = arch.latent_dim
    a, b = a.view(B, T, -1)[..., :z].cpu()[mask_cpu], b.view(B, T, -1)[..., :z].cpu()[mask_cpu]
    d = stats("eps fused vs unfused", b, a)
    assert d.max() <= 2e-2 * a.std() + 5e-3
    want = t("eps_first")[mask_cpu]
    assert (b - want).abs().max() <= 5e-2 * want.std() + 2e-2


def test_sampler_variants_against_reference_generic_lib():
    """Config-3 samplers: ancestral DDPM (fixed-small / fixed-large variance) and strided DDIM against outputs of the
    reference's generic diffusion lib (gaussian_diffusion.py p_sample / ddim_sample + respace.SpacedDiffusion) minted
    with the reference denoiser as model_fn (tests/golden/samplers_z16_parity.npz)."""
    from diffnorm_b200 import ops
    g, arch, sd, eng = setup_case("pass_z16_parity")
    s = np.load(os.path.join(GOLD, "samplers_z16_parity.npz"))
    t = lambda k: torch.from_numpy(s[k])
    x0 = t("x")
    B, T, z = x0.shape
    lens = torch.from_numpy(g["lengths"]).to(torch.int32).to(DEV)
    mask_cpu = O.lengths_to_mask(torch.from_numpy(g["lengths"]), T)
    t_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    for tag, large in (("small", False), ("large", True)):
        rows = torch.from_numpy(eng.sched.ddpm_rows(large)).to(DEV)
        for step in (37, 1, 0):
            xb = eng.stage_latent(x0.to(DEV))
            x = x0.to(DEV).clone().view(B * T, z)
            t_idx.fill_(step)
            eh = eng.denoise(xb, lens, B, T, t_idx)
            d = stats(f"ddpm_{tag}_t{step}_eps", eh.view(B, T, -1)[..., :z].cpu()[mask_cpu], t(f"ddpm_{tag}_t{step}_eps")[mask_cpu])
            assert d.max() <= 5e-2 * t(f"ddpm_{tag}_t{step}_eps").std() + 2e-2
            ops.ddpm_step(x, eh, t(f"ddpm_{tag}_t{step}_noise").to(DEV).view(B * T, z).contiguous(), rows, t_idx, xb)
            want = t(f"ddpm_{tag}_t{step}_sample")
            d = stats(f"ddpm_{tag}_t{step}_sample", x.view(B, T, z).cpu()[mask_cpu], want[mask_cpu])
            assert d.max() <= 3e-2 * (1 + want.abs().max())
    # strided DDIM through the engine's own sampler loop: 3 steps from the top of range(0, 40, 4)
    keep = s["strided_keep"].tolist()
    sp, tmap = eng.sched.spaced(keep)
    rows = torch.from_numpy(sp.ddim_rows()).to(DEV)
    x = x0.to(DEV).clone().view(B * T, z)
    xb = eng.stage_latent(x0.to(DEV))
    r_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    for i in range(len(tmap) - 1, len(tmap) - 4, -1):
        t_idx.fill_(tmap[i])
        r_idx.fill_(i)
        eh = eng.denoise(xb, lens, B, T, t_idx)
        ops.ddim_step(x, eh, rows, r_idx, 1, xb)
    want = t("strided_after3")
    d = stats("strided_after3", x.view(B, T, z).cpu()[mask_cpu], want[mask_cpu])
    assert d.max() <= 3e-2 * (1 + want.abs().max())


def test_pass_sampler_modes_run_and_agree_with_oracle():
    """engine.normalize with sampler='ddpm' (replayed per-step noise) and 'ddim_strided' vs the oracle pass."""
    g, arch, sd, eng = setup_case("pass_z16_parity")
    t = lambda k: torch.from_numpy(g[k])
    feat, eps_vae, eps_q = t("feat"), t("eps_vae"), t("eps_q")
    B, T, _ = feat.shape
    z = arch.latent_dim
    lens = torch.from_numpy(g["lengths"]).to(torch.int32).to(DEV)
    mask_cpu = O.lengths_to_mask(torch.from_numpy(g["lengths"]), T)
    start = 6
    gen = torch.Generator().manual_seed(5)
    noises = [torch.randn(B, T, z, generator=gen) for _ in range(start - 1)]
    ref = O.normalize_pass(sd, arch, feat, mask_cpu, start, eps_vae, eps_q, sampler="ddpm", step_noise=noises)
    out = eng.normalize(feat.to(DEV), lens, start, eps_vae.to(DEV), eps_q.to(DEV), sampler="ddpm",
                        step_noise=[n.to(DEV) for n in noises])
    d = stats("ddpm_pass_x0", out["x0"].cpu()[mask_cpu], ref["x0"][mask_cpu])
    assert out["calls"] == ref["calls"] == start - 1
    assert d.max() <= 3e-2 * (1 + ref["x0"].abs().max())
    keep = list(range(0, 12, 3))
    ref = O.normalize_pass(sd, arch, feat, mask_cpu, 12, eps_vae, eps_q, sampler="ddim_strided", timesteps=keep)
    out = eng.normalize(feat.to(DEV), lens, 12, eps_vae.to(DEV), eps_q.to(DEV), sampler="ddim_strided", timesteps=keep)
    d = stats("strided_pass_x0", out["x0"].cpu()[mask_cpu], ref["x0"][mask_cpu])
    assert out["calls"] == ref["calls"] == len(keep) - 1
    assert d.max() <= 3e-2 * (1 + ref["x0"].abs().max())
