"""Pins the oracle (oracle/diffnorm_oracle.py, oracle/reduce_tgt.c) against fixtures minted from the
live reference by oracle/make_golden.py.  CPU only."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import diffnorm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
FP32_TOL = dict(rtol=1e-4, atol=2e-5)  # fp32 restatement vs fp32 reference (different op order only)

_sd_cache = {}


def weights(z, seed, parity):
    key = (z, seed, parity)
    if key not in _sd_cache:
        _sd_cache.clear()  # 1.6 GB each
        arch = O.Arch(latent_dim=z)
        _sd_cache[key] = (arch, O.init_state_dict(arch, seed=seed, gains=O.PARITY_GAINS if parity else None))
    return _sd_cache[key]


def test_schedule_matches_reference_tables():
    g = np.load(os.path.join(GOLD, "schedule_T200.npz"))
    s = O.Schedule(200)
    for k in g.files:
        np.testing.assert_allclose(getattr(s, k), g[k], rtol=1e-12, atol=0, err_msg=k)
    assert abs(s.alphas_cumprod[99] - 0.49384) < 1e-5 and abs(s.betas[0] - 2.55e-4) < 1e-6  # SURVEY a4 probe


def _reduce_cases():
    g = np.load(os.path.join(GOLD, "reduce_tgt.npz"))
    names = sorted({f[:-3] for f in g.files if f.endswith("_in")})
    return g, names


def test_reduce_tgt_python_matches_reference():
    g, names = _reduce_cases()
    assert "empty" in names and len(names) >= 16
    for n in names:
        d, du, keep = O.reduce_tgt(g[n + "_in"].tolist())
        assert d == g[n + "_dedup"].tolist(), n
        assert du == g[n + "_dur"].tolist(), n
        assert keep == g[n + "_keep"].tolist(), n
        d2, du2, k2 = O.reduce_tgt_np(g[n + "_in"])
        assert d2.tolist() == d and du2.tolist() == du and k2.tolist() == keep, n
    assert O.reduce_tgt([]) == ([], [1], [])  # reference quirk (diff_norm_synthesis.py:45)


def test_reduce_tgt_c_matches_reference():
    so = os.path.join(ROOT, "oracle", "liboracle_reduce.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    P = ctypes.POINTER(ctypes.c_int64)
    lib.dn_oracle_reduce_tgt.restype = ctypes.c_int64
    lib.dn_oracle_reduce_tgt.argtypes = [P, ctypes.c_int64, P, P, P, P]
    g, names = _reduce_cases()
    for n in names:
        tok = np.ascontiguousarray(g[n + "_in"], dtype=np.int64)
        m = max(len(tok), 1)
        d, du, kp = (np.zeros(m, np.int64) for _ in range(3))
        nd = ctypes.c_int64(0)
        r = lib.dn_oracle_reduce_tgt(tok.ctypes.data_as(P), len(tok), d.ctypes.data_as(P), du.ctypes.data_as(P),
                                     kp.ctypes.data_as(P), ctypes.byref(nd))
        assert d[:r].tolist() == g[n + "_dedup"].tolist(), n
        assert du[: nd.value].tolist() == g[n + "_dur"].tolist(), n
        assert kp[:r].tolist() == g[n + "_keep"].tolist(), n


@pytest.mark.parametrize("name", ["pass_z16_default", "pass_z16_parity", "pass_z128_parity", "pass_z16_start1"])
def test_pass_matches_reference(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    z, parity = int(g["latent_dim"]), bool(g["parity_gains"])
    arch, sd = weights(z, int(g["weight_seed"]), parity)
    t = lambda k: torch.from_numpy(g[k])
    lens = torch.from_numpy(g["lengths"])
    mask = O.lengths_to_mask(lens, g["feat"].shape[1])
    out = O.normalize_pass(sd, arch, t("feat"), mask, int(g["start_step"]), t("eps_vae"), t("eps_q"),
                           t("ref_units"), collect=True)
    np.testing.assert_allclose(out["z"].numpy(), g["z"], **FP32_TOL)
    np.testing.assert_allclose(out["x_start"].numpy(), g["x_start"], **FP32_TOL)
    np.testing.assert_allclose(out["eps_first"].numpy(), g["eps_first"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(out["recon"].numpy(), g["recon"], rtol=1e-3, atol=2e-4)
    for i, n in enumerate(lens.tolist()):
        assert out["out_tokens"][i].tolist() == g["units"][i, :n].tolist()
    assert out["match"] == int(g["match"]) and out["total"] == int(g["total"])
    # decode stage on a seeded latent
    rec, logits = O.vae_decode(sd, arch, t("dec_latent"), mask)
    np.testing.assert_allclose(rec.numpy(), g["dec_feat"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(logits.numpy(), g["dec_logits"], rtol=1e-3, atol=2e-4)


def test_sampler_variants_match_reference_generic_lib():
    g = np.load(os.path.join(GOLD, "samplers_z16_parity.npz"))
    p = np.load(os.path.join(GOLD, "pass_z16_parity.npz"))
    arch, sd = weights(16, int(p["weight_seed"]), True)
    mask = O.lengths_to_mask(torch.from_numpy(p["lengths"]), p["feat"].shape[1])
    sch = O.Schedule(200)
    x = torch.from_numpy(g["x"])
    b = x.shape[0]
    for tag in ("small", "large"):
        for t in (37, 1, 0):
            eps = O.denoiser(sd, arch, x, torch.full((b,), t, dtype=torch.long), mask)
            np.testing.assert_allclose(eps.numpy(), g[f"ddpm_{tag}_t{t}_eps"], rtol=1e-3, atol=1e-4)
            y = O.ddpm_step(sch, x, torch.from_numpy(g[f"ddpm_{tag}_t{t}_eps"]), t,
                            torch.from_numpy(g[f"ddpm_{tag}_t{t}_noise"]), large_var=(tag == "large"))
            np.testing.assert_allclose(y.numpy(), g[f"ddpm_{tag}_t{t}_sample"], rtol=1e-4, atol=1e-5)
    keep = g["strided_keep"].tolist()
    sp, tmap = O.Schedule.spaced(sch, keep)
    xs = x.clone()
    for i in range(len(tmap) - 1, len(tmap) - 4, -1):
        eh = O.denoiser(sd, arch, xs, torch.full((b,), tmap[i], dtype=torch.long), mask)
        xs = O.ddim_generic_step(sp, xs, eh, i)
    np.testing.assert_allclose(xs.numpy(), g["strided_after3"], rtol=1e-3, atol=2e-4)


def test_inline_ddim_equals_generic_ddim():
    # SURVEY §8c probe: latent_module.py:1419-1438 == gaussian_diffusion.py ddim_sample(eta=0) to ~2e-7
    sch = O.Schedule(200)
    g = torch.Generator().manual_seed(3)
    x, e = torch.randn(2, 9, 16, generator=g), torch.randn(2, 9, 16, generator=g)
    for t in (150, 99, 1):
        a, b = O.ddim_step(sch, x, e, t), O.ddim_generic_step(sch, x, e, t)
        assert (a - b).abs().max() < 5e-5 * (1 + a.abs().max())


# ---------------------------------------------------------------------------------------------- training step (a11)
@pytest.mark.parametrize("name", ["train_z16_dropout", "train_z16_nodrop_multitask"])
def test_oracle_train_loss_and_grads_match_reference(name):
    """LatentDiscreteModel.forward + autograd of the live reference (oracle/make_golden.py make_train: every random
    draw replayed, incl. the 12 attention-dropout masks) vs the oracle's functional restatement under autograd."""
    g = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    z, T = int(g["latent_dim"]), int(g["T"])
    lengths, times = g["lengths"].tolist(), g["times"].tolist()
    drop_p, multitask = float(g["drop_p"]), bool(int(g["multitask"]))
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=int(g["weight_seed"]), gains=O.PARITY_GAINS)
    train_keys = [k for k in sd if k.startswith("model.") and sd[k].is_floating_point() and "pos_embed" not in k]
    for k in train_keys:
        sd[k].requires_grad_(True)
    audio, units, mask, eps_vae, eps0, eps, keeps = O.train_case_inputs(z, len(lengths), T, lengths, int(g["data_seed"]), drop_p)
    out = O.train_loss(sd, arch, audio, units, mask, torch.tensor(times), eps_vae, eps0, eps, keeps, drop_p, multitask)
    for k in ("total_loss", "nll_loss", "recon_mse_loss", "noise_loss", "acc"):
        assert abs(float(out[k]) - float(g[k])) <= 2e-5 * max(1.0, abs(float(g[k]))), (k, float(out[k]), float(g[k]))
    out["total_loss"].backward()
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(train_keys)
    worst = 0.0
    for n, norm, samp in zip(names, g["grad_norms"], g["grad_samples"]):
        gr = sd[n].grad.double().flatten()
        assert abs(float(gr.norm()) - norm) <= 1e-3 * norm + 1e-9, (n, float(gr.norm()), norm)
        got = gr[torch.from_numpy(O.grad_probe(n, gr.numel()))].numpy()
        err = np.abs(got - samp).max() / (np.abs(samp).max() + norm / np.sqrt(gr.numel()) + 1e-12)
        worst = max(worst, err)
        assert err <= 5e-3, (n, got, samp)
    for key in g.files:
        if key.startswith("full:"):
            np.testing.assert_allclose(sd[key[5:]].grad.numpy(), g[key], rtol=2e-3, atol=1e-6 * float(np.abs(g[key]).max()) + 1e-9)
    print(f"[parity] {name}: oracle autograd vs reference autograd, worst sampled-gradient error {worst:.2e}")


def test_oracle_vae_train_loss_and_grads_match_reference():
    """VAE training: SpeechVAEEncoderDecoder.forward + speech_vae_decoder_loss of the live reference in train
    mode (posterior draw and the 6 attention-dropout masks replayed) vs the oracle under autograd."""
    g = np.load(os.path.join(GOLD, "vae_train_z16.npz"), allow_pickle=False)
    z, T = int(g["latent_dim"]), int(g["T"])
    lengths, drop_p = g["lengths"].tolist(), float(g["drop_p"])
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=int(g["weight_seed"]), gains=O.PARITY_GAINS)
    keys = [str(n) for n in g["grad_names"]]
    for k in keys:
        sd[k].requires_grad_(True)
    audio, units, mask, eps_vae, _, _, keeps = O.train_case_inputs(z, len(lengths), T, lengths, int(g["data_seed"]), drop_p,
                                                                   depth=arch.vae_depth)
    out = O.vae_train_loss(sd, arch, audio, units, mask, eps_vae, keeps, drop_p, int(g["ntokens"]))
    for k in ("loss", "nll_loss", "mse_loss", "kl_loss", "acc"):
        assert abs(float(out[k].detach()) - float(g[k])) <= 2e-5 * max(1.0, abs(float(g[k]))), (k, float(out[k].detach()), float(g[k]))
    out["loss"].backward()
    assert sorted(keys) == sorted(k for k in sd if k.startswith("speech_decoder.") and sd[k].is_floating_point())
    for n, norm, samp in zip(keys, g["grad_norms"], g["grad_samples"]):
        gr = sd[n].grad.double().flatten()
        assert abs(float(gr.norm()) - norm) <= 1e-3 * norm + 1e-9, (n, float(gr.norm()), norm)
        got = gr[torch.from_numpy(O.grad_probe(n, gr.numel()))].numpy()
        assert np.abs(got - samp).max() / (np.abs(samp).max() + norm / np.sqrt(gr.numel()) + 1e-12) <= 5e-3, n


def test_oracle_kmeans_predict_matches_sklearn_golden():
    """k-means unit quantisation: the algorithm lives in scikit-learn (not vendored by the reference; 1.9.0
    here); the fixture holds KMeans.predict's own labels (oracle/make_golden.py make_kmeans)."""
    g = np.load(os.path.join(GOLD, "kmeans_predict.npz"))
    centers, feats = O.kmeans_case(int(g["seed"]), int(g["K"]), int(g["D"]), int(g["N"]))
    got = O.kmeans_predict(centers, feats)
    assert (got == g["labels"].astype(np.int64)).mean() == 1.0


def test_vocoder_oracle_matches_reference_golden():
    """SURVEY §8f-4 groundwork: oracle/vocoder_oracle.py (CodeHiFiGAN generator + duration predictor + the driver's unit
    handling) against the waveforms the untouched reference modules produced on the same seeded weights
    (oracle/make_golden.py --vocoder-only).  Durations are integers: bit-exact; waveform fp32 vs fp32: 1e-5."""
    from oracle import vocoder_oracle as V
    g = np.load(os.path.join(GOLD, "vocoder_code_hifigan.npz"))
    sd = V.init_state_dict(int(g["weight_seed"]))
    assert [k for k, _ in V.weight_norm_keys()] == list(sd) and len(sd) == 302
    reduced = V.process_units(g["raw_units"].tolist(), reduce=True)
    assert reduced == g["reduced_code"].tolist() and len(reduced) < len(g["raw_units"])
    assert V.process_units([3, 3, 4, 4, 4, 3], reduce=True) == [3, 4, 3] and V.process_units([3, 3], reduce=False) == [3, 3]
    for name, dp in (("dur", True), ("nodur", False), ("reduced", True)):
        wav, dur = V.code_to_waveform(sd, torch.from_numpy(g[f"{name}_code"]), dur_prediction=dp)
        assert dur.tolist() == g[f"{name}_dur"].tolist()
        assert wav.numel() == V.HOP * int(dur.sum())
        want = torch.from_numpy(g[f"{name}_wave"])
        assert float((wav - want).abs().max()) < 1e-5
        assert 0.05 < float(want.std()) < 0.5 and float(want.abs().max()) < 0.999     # audible, un-saturated signal
    assert int(g["dur_dur"].max()) > 3 and int(g["dur_dur"].min()) == 1              # the duration head is exercised
    assert (g["reduced_code"] < 0).sum() == 2                                         # invalid codes reach, and are dropped by, the model wrapper
