"""Mint golden vectors from the LIVE, UNMODIFIED reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden   (needs /root/reference; ~2 min on 8 cores)

Recipe (SURVEY.md §8c "determinism recipe"): weights come from ``diffnorm_oracle.init_state_dict(seed)``
(regenerable anywhere from the seed — 400 M parameters cannot be committed) and are *loaded into the
reference's own modules*; inputs and every noise tensor come from a seeded ``torch.Generator`` and are stored
in the fixture; the reference's ``torch.randn`` draws are replaced by replay (ref_loader.ReplayNoise).
Outputs of the reference are stored as float32/int64 arrays in ``tests/golden/*.npz``.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import diffnorm_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (latent_dim, weight seed, gains, B, T, lengths, start_step, data seed)
PASS_CASES = {
    "pass_z16_default": (16, 0, None, 2, 40, [40, 29], 6, 11),
    "pass_z16_parity": (16, 1, "parity", 2, 48, [48, 33], 8, 12),
    "pass_z128_parity": (128, 2, "parity", 1, 40, [40], 5, 13),
    # start_step = 1: timesteps = [0], the break at time == 1 never fires -> ONE call at t = 0 (LM:1402,1444)
    "pass_z16_start1": (16, 1, "parity", 1, 24, [24], 1, 14),
}


def case_inputs(z, B, T, lengths, dseed):
    g = torch.Generator().manual_seed(dseed)
    lens = torch.tensor(lengths)
    mask = O.lengths_to_mask(lens, T)
    feat = torch.randn(B, T, 768, generator=g) * mask[:, :, None]
    eps_vae = torch.randn(B, z, T, generator=g)
    eps_q = torch.randn(B, T, z, generator=g)
    ref_units = torch.randint(0, 1000, (B, T), generator=g) * mask
    return feat, mask, eps_vae, eps_q, ref_units


def gains_of(tag):
    return O.PARITY_GAINS if tag == "parity" else None


@torch.no_grad()
def make_pass(name):
    z, wseed, gtag, B, T, lengths, start, dseed = PASS_CASES[name]
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=wseed, gains=gains_of(gtag))
    ldm = ref_loader.build_reference_model(z)
    ldm.load_state_dict(sd)
    feat, mask, eps_vae, eps_q, ref_units = case_inputs(z, B, T, lengths, dseed)
    sch = ldm.scheduler
    # stage-wise outputs of the reference's own modules
    with ref_loader.ReplayNoise([eps_vae]):
        zlat = ldm.speech_decoder.encode_feature(feat).transpose(1, 2)
    x_start = (torch.tensor(np.float32(sch.sqrt_alphas_cumprod[start])) * zlat
               + torch.tensor(np.float32(sch.sqrt_one_minus_alphas_cumprod[start])) * eps_q)
    t_first = torch.full((B,), start - 1, dtype=torch.long)
    eps_first = ldm.model(x_start.clone(), t_first, input_mask=mask, cond_drop_prob=0)
    # full pass through the reference's public entry
    with ref_loader.ReplayNoise([eps_vae, eps_q]) as rn:
        toks, match, total, recon = ldm.ddim_sample(feat, input_mask=mask, ref_units=ref_units, start_step=start)
    assert rn.used == 2
    # x0 is not returned by the reference; recover it by re-running its own loop body pieces is overkill —
    # instead the decode stage is pinned separately on a seeded latent:
    g = torch.Generator().manual_seed(dseed + 100)
    lat = torch.randn(B, T, z, generator=g)
    dec_feat, dec_logits = ldm.speech_decoder.decode_feature(lat, mask)
    units_pad = np.full((B, T), -1000, dtype=np.int64)
    for i, tk in enumerate(toks):
        units_pad[i, : len(tk)] = tk.numpy()
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        latent_dim=z, weight_seed=wseed, parity_gains=int(gtag == "parity"), start_step=start, data_seed=dseed,
        lengths=np.array(lengths), feat=feat.numpy(), eps_vae=eps_vae.numpy(), eps_q=eps_q.numpy(),
        ref_units=ref_units.numpy(), z=zlat.numpy(), x_start=x_start.numpy(), eps_first=eps_first.numpy(),
        recon=recon.numpy(), units=units_pad, match=match, total=total,
        dec_latent=lat.numpy(), dec_feat=dec_feat.numpy(), dec_logits=dec_logits.numpy(),
    )
    print(name, "match/total", match, total, "units uniq", len(np.unique(units_pad)))
    return sd, arch, ldm, (feat, mask, eps_vae, eps_q)


@torch.no_grad()
def make_samplers(sd, arch, ldm, inputs):
    """DDPM ancestral and strided-DDIM steps from the reference's generic diffusion lib
    (diffusion/gaussian_diffusion.py:376-417, 513-560; respace.py:65-129), model = the reference denoiser."""
    gd = ref_loader.load().diffusion.gaussian_diffusion
    rs = ref_loader.load().diffusion.respace
    feat, mask, eps_vae, eps_q = inputs
    B, T, z = eps_q.shape
    betas = gd.get_named_beta_schedule("squaredcos_cap_v2", 200)
    assert np.allclose(betas, ldm.scheduler.betas)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, T, z, generator=g)
    out = {"x": x.numpy()}
    model_fn = lambda xx, tt, **kw: ldm.model(xx, tt, input_mask=mask, cond_drop_prob=0)
    for tag, vt in (("small", gd.ModelVarType.FIXED_SMALL), ("large", gd.ModelVarType.FIXED_LARGE)):
        diff = gd.GaussianDiffusion(betas=betas, model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=vt,
                                    loss_type=gd.LossType.MSE)
        for t in (37, 1, 0):
            noise = torch.randn(B, T, z, generator=g)
            tt = torch.full((B,), t, dtype=torch.long)
            orig = torch.randn_like
            torch.randn_like = lambda _x, **kw: noise.clone()
            try:
                r = diff.p_sample(model_fn, x, tt, clip_denoised=False)
            finally:
                torch.randn_like = orig
            out[f"ddpm_{tag}_t{t}_noise"] = noise.numpy()
            out[f"ddpm_{tag}_t{t}_sample"] = r["sample"].numpy()
            out[f"ddpm_{tag}_t{t}_eps"] = model_fn(x, tt).numpy()
    # strided DDIM: keep every 4th step of range(0, 40) -> 10 steps; run 3 of them from the top
    keep = list(range(0, 40, 4))
    sp = rs.SpacedDiffusion(use_timesteps=keep, betas=betas, model_mean_type=gd.ModelMeanType.EPSILON,
                            model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE)
    xs = x.clone()
    for i in range(len(keep) - 1, len(keep) - 4, -1):
        tt = torch.full((B,), i, dtype=torch.long)
        xs = sp.ddim_sample(model_fn, xs, tt, clip_denoised=False, eta=0.0)["sample"]
    out["strided_keep"] = np.array(keep)
    out["strided_after3"] = xs.numpy()
    np.savez_compressed(os.path.join(GOLD, "samplers_z16_parity.npz"), **out)
    print("samplers done")


TRAIN_CASES = {
    # name -> (latent_dim, weight seed, B, T, lengths, times, data seed, dropout p, multitask)
    "train_z16_dropout": (16, 3, 2, 24, [24, 17], [37, 142], 21, 0.1, False),
    "train_z16_nodrop_multitask": (16, 3, 2, 24, [24, 17], [5, 199], 22, 0.0, True),
}


train_inputs = O.train_case_inputs


class ReplayTraining:
    """Replays every random draw of LatentDiscreteModel.forward (LM:1514-1613) in call order: torch.randint (times,
    :1528), torch.randn on the CPU (posterior draw, distributions.py:38), two torch.randn_like (:1528, :1534) and the
    12 attention-dropout draws (F.dropout behind nn.Dropout, LM:338)."""

    def __init__(self, times, eps_vae, eps0, eps, keeps, drop_p):
        self.times, self.randn_q, self.like_q = times, [eps_vae], [eps0, eps]
        self.keeps, self.drop_p, self.drop_used = list(keeps or []), drop_p, 0

    def __enter__(self):
        import torch.nn.functional as F
        self._saved = (torch.randint, torch.randn, torch.randn_like, F.dropout)
        rt = self

        def randint(*a, **kw):
            return rt.times.clone()

        def randn(*shape, **kw):
            if len(shape) == 1 and not isinstance(shape[0], int):
                shape = tuple(shape[0])
            t = rt.randn_q.pop(0)
            assert tuple(t.shape) == tuple(shape)
            return t.clone()

        def randn_like(x, **kw):
            t = rt.like_q.pop(0)
            assert t.shape == x.shape
            return t.clone()

        def dropout(inp, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return inp
            assert abs(p - rt.drop_p) < 1e-12
            k = rt.keeps.pop(0)
            rt.drop_used += 1
            return inp * k.to(inp.dtype) / (1.0 - p)

        torch.randint, torch.randn, torch.randn_like, F.dropout = randint, randn, randn_like, dropout
        return self

    def __exit__(self, *a):
        import torch.nn.functional as F
        torch.randint, torch.randn, torch.randn_like, F.dropout = self._saved


grad_probe = O.grad_probe


def make_train(name):
    z, wseed, B, T, lengths, times, dseed, drop_p, multitask = TRAIN_CASES[name]
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=wseed, gains=O.PARITY_GAINS)
    ldm = ref_loader.build_reference_model(z, multitask=multitask)
    ldm.load_state_dict(sd)
    ldm.train()
    for n_, p_ in ldm.named_parameters():   # diff_discrete.py:79-81: the VAE is frozen
        p_.requires_grad_(not n_.startswith("speech_decoder."))
    if drop_p == 0.0:                        # dropout is a hyper-parameter of the reference modules (LM:668)
        for m in ldm.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    audio, units, mask, eps_vae, eps0, eps, keeps = train_inputs(z, B, T, lengths, dseed, drop_p)
    tt = torch.tensor(times, dtype=torch.long)
    with ReplayTraining(tt, eps_vae, eps0, eps, keeps, drop_p) as rt:
        out = ldm(audio.clone(), units.clone(), tgt_mask=mask)
    assert rt.drop_used == (12 if drop_p > 0 else 0) and not rt.like_q and not rt.randn_q
    out["total_loss"].backward()
    store = dict(latent_dim=z, weight_seed=wseed, lengths=np.array(lengths), times=np.array(times), data_seed=dseed,
                 drop_p=drop_p, multitask=int(multitask), T=T)
    for k in ("total_loss", "nll_loss", "recon_mse_loss", "noise_loss", "acc"):
        store[k] = np.float64(out[k].detach().double().item())
    names, norms, samples = [], [], []
    for n_, p_ in ldm.named_parameters():
        if p_.grad is None:
            continue
        gflat = p_.grad.detach().double().flatten()
        names.append(n_)
        norms.append(float(gflat.norm()))
        samples.append(gflat[torch.from_numpy(grad_probe(n_, gflat.numel()))].numpy())
    store["grad_names"] = np.array(names)
    store["grad_norms"] = np.array(norms)
    store["grad_samples"] = np.stack(samples)
    # full gradients of a few small tensors
    for n_ in ("model.to_time_cond.0.weights", "model.final_proj.bias", "model.transformer.to_pred.0.gamma",
               "model.wavenet.stacks.0.blocks.0.to_time_cond.bias", "model.transformer.layers.0.0.to_gamma_beta.bias"):
        store["full:" + n_] = dict(ldm.named_parameters())[n_].grad.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **store)
    print(name, {k: float(store[k]) for k in ("total_loss", "nll_loss", "recon_mse_loss", "noise_loss", "acc")},
          "params with grad", len(names))


VAE_TRAIN_CASES = {
    # name -> (latent_dim, weight seed, B, T, lengths, data seed, dropout p)
    "vae_train_z16": (16, 4, 2, 24, [24, 15], 31, 0.1),
}


def make_vae_train(name):
    """SpeechVAEEncoderDecoder.forward in train mode (LM:1118-1142) + the loss arithmetic of
    speech_vae_decoder_loss.py:60-82 (restated here in 6 lines: that file imports fairseq at module top), backward."""
    z, wseed, B, T, lengths, dseed, drop_p = VAE_TRAIN_CASES[name]
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=wseed, gains=O.PARITY_GAINS)
    ldm = ref_loader.build_reference_model(z)
    ldm.load_state_dict(sd)
    vae = ldm.speech_decoder
    vae.train()
    audio, units, mask, eps_vae, _, _, keeps = O.train_case_inputs(z, B, T, lengths, dseed, drop_p, depth=arch.vae_depth)
    rt = ReplayTraining(None, eps_vae, None, None, keeps, drop_p)
    rt.like_q = []
    with rt:
        mse_loss, lm_pred, kl_loss = vae(audio.clone(), units.clone(), mask)
    assert rt.drop_used == arch.vae_depth and not rt.randn_q
    lprobs = torch.log_softmax(lm_pred, dim=-1).view(-1, lm_pred.size(-1))
    target = units.view(-1)
    tmask = target.ne(0)
    acc = (lprobs.argmax(1)[tmask] == target[tmask]).sum() / tmask.sum()
    from fairseq.criterions.label_smoothed_cross_entropy import label_smoothed_nll_loss   # the stub of ref_loader (:34-51)
    ls, nll = label_smoothed_nll_loss(lprobs, target, 0.1, ignore_index=0, reduce=True)
    ntokens = int(tmask.sum())
    loss = 0.1 * (ls / ntokens) + 10 * mse_loss + 0.0001 * kl_loss
    loss.backward()
    store = dict(latent_dim=z, weight_seed=wseed, lengths=np.array(lengths), data_seed=dseed, drop_p=drop_p, T=T, ntokens=ntokens)
    for k, v in (("loss", loss), ("nll_loss", nll / ntokens), ("mse_loss", mse_loss), ("kl_loss", kl_loss), ("acc", acc)):
        store[k] = np.float64(v.detach().double().item())
    names, norms, samples = [], [], []
    for n_, p_ in vae.named_parameters():
        if p_.grad is None:
            continue
        gflat = p_.grad.detach().double().flatten()
        names.append("speech_decoder." + n_)
        norms.append(float(gflat.norm()))
        samples.append(gflat[torch.from_numpy(grad_probe("speech_decoder." + n_, gflat.numel()))].numpy())
    store["grad_names"], store["grad_norms"], store["grad_samples"] = np.array(names), np.array(norms), np.stack(samples)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **store)
    print(name, {k: float(store[k]) for k in ("loss", "nll_loss", "mse_loss", "kl_loss", "acc")}, "params with grad", len(names))


def kmeans_case(seed=11, K=1000, D=768, N=4000):
    """Seeded centroids / features of the k-means fixture (features = planted centroid + noise, plus 500 pure-noise rows)."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((K, D)).astype(np.float32) * 0.6
    lab = rng.integers(0, K, size=N)
    feats = (centers[lab] + rng.standard_normal((N, D)).astype(np.float32) * 0.9).astype(np.float32)
    feats[-500:] = rng.standard_normal((500, D)).astype(np.float32)
    return centers, feats


def make_kmeans():
    """Labels from scikit-learn's own KMeans.predict (what quantize_with_kmeans.py:115 calls on the joblib-loaded model)."""
    from sklearn.cluster import KMeans
    centers, feats = kmeans_case()
    km = KMeans(n_clusters=len(centers), n_init=1)
    km.fit(feats[: len(centers)])          # creates the fitted attributes; the centroids are then replaced
    km.cluster_centers_ = centers.copy()
    pred = km.predict(feats)
    np.savez_compressed(os.path.join(GOLD, "kmeans_predict.npz"), seed=11, K=len(centers), D=centers.shape[1], N=len(feats),
                        labels=pred.astype(np.int16))
    print("kmeans labels", len(pred), "agreement with oracle", float((pred == O.kmeans_predict(centers, feats)).mean()))


def make_schedule(ldm):
    s = ldm.scheduler
    np.savez_compressed(
        os.path.join(GOLD, "schedule_T200.npz"),
        betas=s.betas, alphas_cumprod=s.alphas_cumprod, alphas_cumprod_prev=s.alphas_cumprod_prev,
        sqrt_alphas_cumprod=s.sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod=s.sqrt_one_minus_alphas_cumprod,
        posterior_variance=s.posterior_variance, posterior_log_variance_clipped=s.posterior_log_variance_clipped,
        posterior_mean_coef1=s.posterior_mean_coef1, posterior_mean_coef2=s.posterior_mean_coef2,
    )


def reference_reduce_token():
    """Extract ONLY the `reduce_token` function object from the reference driver (the file itself cannot be
    imported: it imports fairseq at module top) by compiling that one FunctionDef node in memory."""
    path = os.path.join(ref_loader.REF_ROOT, "research", "TranSpeech", "diff_norm_synthesis.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "reduce_token"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["reduce_token"]


def make_reduce():
    red = reference_reduce_token()
    rng = np.random.default_rng(5)
    cases = {
        "empty": [], "single": [7], "all_same": [3] * 17, "all_diff": list(range(40)),
        "two_runs": [5] * 31 + [6] * 33, "warp_edge": [1] * 32 + [2] * 32 + [2] * 1 + [3] * 63,
        "negatives": [-4, -4, -1, -1, -1, 0, 0, 999, 999],
        "alt": [1, 2] * 50,
    }
    for n in (1, 31, 32, 33, 500, 1000, 2000, 4097):
        runs = rng.geometric(0.6, size=n)
        ids = rng.integers(0, 1000, size=n)
        cases[f"geom_{n}"] = np.repeat(ids, runs)[:n].tolist()
    out = {}
    for k, toks in cases.items():
        d, du, keep = red(list(toks))
        out[k + "_in"] = np.array(toks, dtype=np.int64)
        out[k + "_dedup"] = np.array(d, dtype=np.int64)
        out[k + "_dur"] = np.array(du, dtype=np.int64)
        out[k + "_keep"] = keep.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(GOLD, "reduce_tgt.npz"), **out)
    print("reduce cases", len(cases))


def make_batcher():
    """Golden vectors from the reference's OWN compiled Cython batcher (oracle/_ref/data_utils_fast*.so, built by
    `make -C oracle ref` from /root/reference/fairseq/data/data_utils_fast.pyx)."""
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import data_utils_fast as ref
    rng = np.random.default_rng(7)
    out, n_cases = {}, 0
    for trial in range(300):
        n = int(rng.integers(1, 80))
        toks = rng.integers(1, 60, size=n).astype(np.int64)
        if rng.random() < 0.6:
            toks = np.sort(toks)
        mt, ms, bm = int(rng.choice([0, 60, 64, 100, 200, 400])), int(rng.choice([0, 1, 3, 8])), int(rng.choice([1, 2, 4, 8]))
        if mt and toks.max() > mt:
            mt = int(toks.max())
        sizes = [len(b) for b in ref.batch_by_size_vec(np.arange(n, dtype=np.int64), toks, mt, ms, bm)]
        out[f"toks_{n_cases}"], out[f"args_{n_cases}"], out[f"sizes_{n_cases}"] = toks, np.array([mt, ms, bm]), np.array(sizes)
        n_cases += 1
    # dataset-like case: 20k log-normal lengths, length-sorted, --max-tokens 12000 (scripts/diffusion/train.sh:31)
    lens = np.sort(np.clip(np.round(np.exp(rng.normal(np.log(600), 0.5, size=20000))), 200, 2000).astype(np.int64))
    sizes = [len(b) for b in ref.batch_by_size_vec(np.arange(len(lens), dtype=np.int64), lens, 12000, 0, 1)]
    out[f"toks_{n_cases}"], out[f"args_{n_cases}"], out[f"sizes_{n_cases}"] = lens, np.array([12000, 0, 1]), np.array(sizes)
    n_cases += 1
    np.savez_compressed(os.path.join(GOLD, "batch_by_size.npz"), n_cases=n_cases, **out)
    print("batcher cases", n_cases)


def make_vocoder():
    """SURVEY §8f-4 groundwork: the reference's CodeGenerator (codehifigan.py / hifigan.py / fastspeech2.VariancePredictor,
    loaded untouched by ref_loader.load_vocoder) on the oracle's seeded weights -> tests/golden/vocoder_code_hifigan.npz.
    Cases: durations predicted; durations off; a unit stream that is first reduced by the driver's process_units and carries
    invalid (negative) codes, which CodeHiFiGANVocoder.forward drops (vocoder.py:231-237)."""
    from oracle import vocoder_oracle as V
    voc = ref_loader.load_vocoder()
    seed = 7
    sd = V.init_state_dict(seed)
    m = voc.CodeGenerator(dict(V.VOCODER_CFG)).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(3)
    out = {"weight_seed": seed}
    codes = {"dur": torch.randint(0, 1000, (40,), generator=g), "nodur": torch.randint(0, 1000, (25,), generator=g)}
    raw = torch.randint(0, 6, (60,), generator=g) * 150 + 3           # few distinct units -> runs of duplicates
    raw[[5, 17, 18]] = -1
    reduced = V.process_units(raw.tolist(), reduce=True)
    codes["reduced"] = torch.tensor(reduced)
    out["raw_units"] = raw.numpy()
    for name, code in codes.items():
        dp = name != "nodur"
        with torch.no_grad():
            x = {"code": code.view(1, -1).clone()}
            mask = x["code"] >= 0                                        # vocoder.py:234-235
            x["code"] = x["code"][mask].unsqueeze(0)
            wav = m(**x, dur_prediction=dp).detach().squeeze()
            emb = m.dict(x["code"])
            dur = torch.clamp(torch.round(torch.exp(m.dur_predictor(emb)) - 1).long(), min=1).view(-1) if dp \
                else torch.ones(x["code"].shape[1], dtype=torch.long)
        w2, d2 = V.code_to_waveform(sd, code, dur_prediction=dp)
        err = float((wav - w2).abs().max())
        assert torch.equal(dur, d2) and wav.shape == w2.shape and err < 5e-6, (name, err)
        assert wav.numel() == V.HOP * int(dur.sum())
        print(f"vocoder {name}: {code.numel()} codes -> {int(dur.sum())} frames -> {wav.numel()} samples, dur max {int(dur.max())}, "
              f"wave std {float(wav.std()):.3f}, oracle max err {err:.2e}")
        out[f"{name}_code"], out[f"{name}_dur"], out[f"{name}_wave"] = code.numpy(), dur.numpy(), wav.numpy()
    np.savez_compressed(os.path.join(GOLD, "vocoder_code_hifigan.npz"), **out)


def make_keys():
    """state_dict key ORDER of the reference's LatentDiscreteModel (= its parameter registration order, which is what an
    optimizer state of a resumed checkpoint is matched by) -> tests/golden/state_dict_keys_z16.txt."""
    ldm = ref_loader.build_reference_model(16)
    with open(os.path.join(GOLD, "state_dict_keys_z16.txt"), "w") as f:
        f.write("\n".join(ldm.state_dict().keys()) + "\n")


def make_dataset():
    """The reference's own ReprToReprUnitDataset / Creator (repr_to_repr_unit_dataset.py, loaded untouched with its real
    Dictionary by ref_loader.load_dataset_module) on the synthetic corpus of oracle/dataset_fixture.py ->
    tests/golden/dataset_collate.npz: the kept utterance ids, ordered_indices, and every field of two collated batches."""
    import tempfile

    from oracle.dataset_fixture import write_corpus
    ns = ref_loader.load_dataset_module()
    d = ns.Dictionary()
    for i in range(1000):
        d.add_symbol(str(i))
    out = {}
    with tempfile.TemporaryDirectory() as root:
        src_dir, tgt_dir, tsv_dir = write_corpus(root)
        ds = ns.module.ReprToReprUnitDatasetCreator.from_tsv(src_dir, tgt_dir, tsv_dir, ns.S2SDataConfig(shuffle=False), "train",
                                                             True, 1, 1, tgt_dict=d)
        out["ids"] = np.array(ds.ids)
        out["ordered_indices"] = ds.ordered_indices()
        out["sizes"] = ds.sizes
        for b, idx in enumerate(([0, 3, 5, 1], [6, 2, 4])):
            batch = ds.collater([ds[i] for i in idx])
            out[f"b{b}_idx"] = np.array(idx)
            for k in ("id", "target", "target_unit", "reduce_target", "reduce_target_unit", "target_lengths",
                      "reduce_target_lengths"):
                out[f"b{b}_{k}"] = batch[k].numpy()
            out[f"b{b}_src_tokens"] = batch["net_input"]["src_tokens"].numpy()
            out[f"b{b}_src_lengths"] = batch["net_input"]["src_lengths"].numpy()
            out[f"b{b}_ntokens"] = np.int64(batch["ntokens"])
            out[f"b{b}_nsentences"] = np.int64(batch["nsentences"])
    np.savez_compressed(os.path.join(GOLD, "dataset_collate.npz"), **out)
    print("dataset golden:", out["ids"].tolist(), out["ordered_indices"].tolist())


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    if "--dataset-only" in sys.argv:
        return make_dataset()
    if "--keys-only" in sys.argv:
        return make_keys()
    if "--pass-only" in sys.argv:
        return make_pass(sys.argv[sys.argv.index("--pass-only") + 1])
    if "--vocoder-only" in sys.argv:
        return make_vocoder()
    if "--batcher-only" in sys.argv:
        return make_batcher()
    if "--train-only" in sys.argv:
        for name in TRAIN_CASES:
            make_train(name)
        return
    if "--kmeans-only" in sys.argv:
        return make_kmeans()
    if "--vae-train-only" in sys.argv:
        for name in VAE_TRAIN_CASES:
            make_vae_train(name)
        return
    make_batcher()
    make_reduce()
    for name in PASS_CASES:
        sd, arch, ldm, inputs = make_pass(name)
        if name == "pass_z16_parity":
            make_samplers(sd, arch, ldm, inputs)
            make_schedule(ldm)
    for name in TRAIN_CASES:
        make_train(name)
    for name in VAE_TRAIN_CASES:
        make_vae_train(name)
    make_kmeans()
    make_vocoder()
    make_dataset()
    make_keys()


if __name__ == "__main__":
    main()
