"""Training step of the denoiser (BASELINE config 5) on the GPU: every new kernel against a plain
PyTorch fp32 reference of the same op, then the whole forward + backward step against the oracle's autograd
(oracle.train_loss, itself pinned to the live reference's losses and gradients in tests/test_oracle_golden.py)."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200 import ops  # noqa: E402
from diffnorm_b200.train import DenoiserTrainer, pack_keep_bits  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

DEV = "cuda"
bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("B,T,N,K,shift,dy0,x0", [(1, 640, 128, 256, 0, 0, 0), (3, 200, 512, 512, 0, 0, 0),
                                                  (2, 333, 96, 1408, 2, 0, 0), (2, 150, 512, 512, 8, 1024, 512),
                                                  (2, 77, 16, 64, 0, 0, 0), (4, 1000, 1408, 1408, 1, 0, 0)])
def test_wgrad_kernel(B, T, N, K, shift, dy0, x0):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    ldy, ldx = dy0 + rup8(N) + 64, x0 + rup8(K) + 64
    dY = (torch.randn(B * T, ldy, generator=g) * 0.5).to(bf16).to(DEV)
    X = (torch.randn(B * T, ldx, generator=g) * 0.5).to(bf16).to(DEV)
    dW = torch.zeros(N, (K + 3) // 4 * 4, device=DEV)
    ops.wgrad(dY, X, dW, B, T, N, K, dy0, x0, shift)
    ops.wgrad(dY, X, dW, B, T, N, K, dy0, x0, shift, splits=3)        # accumulates: expect 2x
    y = dY.float().view(B, T, ldy)[:, :, dy0:dy0 + N]
    x = X.float().view(B, T, ldx)[:, :, x0:x0 + K]
    xs = torch.zeros_like(x)
    if shift < T:
        xs[:, shift:] = x[:, :T - shift]
    want = 2 * torch.einsum("btn,btk->nk", y, xs)
    err = rel(dW[:, :K], want)
    print(f"[parity] wgrad B{B} T{T} N{N} K{K} shift{shift}: rel err {err:.2e}")
    assert err < 2e-3


def rup8(v):
    return (v + 7) // 8 * 8


def test_wgrad_grouped_dilated():
    """8 chains in one launch: group g reads its own column blocks and shift << g (WaveNet level weight gradients)."""
    g_ = torch.Generator().manual_seed(3)
    B, T, G, C = 2, 300, 8, 512
    dY = (torch.randn(B * T, G * 2 * C, generator=g_) * 0.5).to(bf16).to(DEV)
    X = (torch.randn(B * T, G * C, generator=g_) * 0.5).to(bf16).to(DEV)
    dW = torch.zeros(G, C, C, device=DEV)
    ops.wgrad(dY, X, dW, B, T, C, C, 0, 0, 1, groups=G, g_dy_col=2 * C, g_x_col=C, shift_shl_group=True)
    y = dY.float().view(B, T, G, 2 * C)[..., :C]
    x = X.float().view(B, T, G, C)
    for g in range(G):
        sh = 1 << g
        xs = torch.zeros(B, T, C, device=DEV)
        xs[:, sh:] = x[:, : T - sh, g]
        want = torch.einsum("btn,btk->nk", y[:, :, g], xs)
        assert rel(dW[g], want) < 2e-3, g


def test_geglu_gate_norm_colsum_backward():
    g = torch.Generator().manual_seed(5)
    B, T, C, ip = 2, 70, 512, 256
    M = B * T
    # GEGLU
    h = (torch.randn(M, 2 * ip, generator=g)).to(bf16).to(DEV)
    dm = (torch.randn(M, ip, generator=g)).to(bf16).to(DEV)
    m = ops.geglu_fwd(h, torch.empty(M, ip, dtype=bf16, device=DEV))
    dh = ops.geglu_bwd(h, dm, torch.empty(M, 2 * ip, dtype=bf16, device=DEV))
    hv = h.float().view(M, ip // 128, 2, 128).requires_grad_(True)
    ref = torch.nn.functional.gelu(hv[:, :, 1]) * hv[:, :, 0]
    assert rel(m.float(), ref.reshape(M, ip)) < 6e-3
    ref.backward(dm.float().view(M, ip // 128, 128))
    assert rel(dh.float(), hv.grad.reshape(M, 2 * ip)) < 6e-3
    # WaveNet gate, G chains, per-utterance gamma/beta rows
    G = 3
    ur = (torch.randn(M, G * 2 * C, generator=g)).to(bf16).to(DEV)
    dy = (torch.randn(M, G * C, generator=g)).to(bf16).to(DEV)
    gb = torch.randn(B, G * 2 * C, generator=g).to(DEV)
    rows = torch.arange(B, dtype=i32, device=DEV)
    y = ops.wn_gate_fwd(ur, torch.empty(M, G * C, dtype=bf16, device=DEV), B, T, C, G, gb.view(-1), G * 2 * C, 2 * C, rows, 1)
    dgb = torch.zeros(B, G * 2 * C, device=DEV)
    dur = ops.wn_gate_bwd(ur, dy, torch.empty(M, G * 2 * C, dtype=bf16, device=DEV), B, T, C, G, gb.view(-1), G * 2 * C, 2 * C,
                          rows, 1, dgb.view(-1), G * 2 * C, 2 * C)
    urv = ur.float().view(B, T, G, C // 128, 2, 128)
    u = urv[..., 0, :].reshape(B, T, G, C).clone().requires_grad_(True)
    r = urv[..., 1, :].reshape(B, T, G, C).clone().requires_grad_(True)
    gbv = gb.view(B, 1, G, 2, C).clone().requires_grad_(True)
    up = u * gbv[:, :, :, 0] + gbv[:, :, :, 1]
    yr = up.tanh() * up.sigmoid() + r
    assert rel(y.float(), yr.reshape(M, G * C)) < 6e-3
    yr.backward(dy.float().view(B, T, G, C))
    durv = dur.float().view(B, T, G, 2, C)
    assert rel(durv[:, :, :, 0], u.grad) < 6e-3 and rel(durv[:, :, :, 1], r.grad) < 1e-6
    assert rel(dgb.view(B, G, 2, C), gbv.grad.view(B, G, 2, C)) < 2e-3
    # adaptive RMSNorm backward (conditioned) and gamma-parameter form
    x = torch.randn(M, C, generator=g).to(DEV)
    dyn = torch.randn(M, C, generator=g).to(bf16).to(DEV)
    gbn = torch.randn(B, 2 * C, generator=g).to(DEV)
    dx0 = torch.randn(M, C, generator=g).to(DEV)
    dx = dx0.clone()
    dxb = torch.empty(M, C, dtype=bf16, device=DEV)
    dgbn = torch.zeros(B, 2 * C, device=DEV)
    ops.adarmsnorm_bwd(x, dyn, dx, dxb, B, T, gb=gbn.view(-1), gb_t_stride=2 * C, t_idx=rows, t_idx_stride=1, dgb=dgbn.view(-1),
                       dgb_b_stride=2 * C)
    xr = x.clone().requires_grad_(True)
    gr = gbn.clone().requires_grad_(True)
    out = torch.nn.functional.normalize(xr.view(B, T, C), dim=-1) * C ** 0.5 * gr[:, None, :C] + gr[:, None, C:]
    out.backward(dyn.float().view(B, T, C))
    assert rel(dx - dx0, xr.grad) < 1e-4 and rel(dxb.float(), dx) < 5e-3 and rel(dgbn, gr.grad) < 1e-4
    gp = torch.randn(C, generator=g).to(DEV)
    dgp = torch.zeros(C, device=DEV)
    dx = torch.zeros(M, C, device=DEV)
    ops.adarmsnorm_bwd(x, dyn, dx, None, B, T, gamma_p=gp, dgamma_p=dgp)
    xr = x.clone().requires_grad_(True)
    gpr = gp.clone().requires_grad_(True)
    (torch.nn.functional.normalize(xr, dim=-1) * C ** 0.5 * gpr).backward(dyn.float())
    assert rel(dx, xr.grad) < 1e-4 and rel(dgp, gpr.grad) < 1e-4
    # column sums
    src = torch.randn(M, 1408, generator=g).to(bf16).to(DEV)
    cs = ops.colsum(src, 128, 1280, torch.zeros(1280, device=DEV))
    assert rel(cs, src.float()[:, 128:].sum(0)) < 1e-5


@pytest.mark.parametrize("T,lens,drop,dh", [(200, [200, 131], True, 64), (128, [128, 1], False, 64), (300, [300, 257], True, 64),
                                            (200, [200, 131], False, 96), (260, [260, 3], False, 96)])
def test_attention_train_forward_backward(T, lens, drop, dh):
    g = torch.Generator().manual_seed(T)
    B, H = len(lens), 8
    M = B * T
    qkv = (torch.randn(M, 3 * H * dh, generator=g) * 1.5).to(bf16).to(DEV)
    dout = torch.randn(M, H * dh, generator=g).to(bf16)
    ln = torch.tensor(lens, dtype=i32, device=DEV)
    p = 0.1
    keep = (torch.rand(B, H, T, T, generator=g) >= p) if drop else None
    bits = pack_keep_bits(keep).to(DEV) if drop else None
    ks = 1 / (1 - p) if drop else 1.0
    out = torch.empty(M, H * dh, dtype=bf16, device=DEV)
    lse = torch.empty(B * H, T, device=DEV)
    ops.attention_train(qkv, out, lse, ln, bits, ks, B, T, H, dh)
    # fp32 reference on the same bf16 inputs
    leaf = qkv.float().cpu().requires_grad_(True)
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in leaf.chunk(3, dim=-1))
    mask = torch.arange(T)[None, :] < torch.tensor(lens)[:, None]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
    attn = sim.softmax(-1)
    if drop:
        attn = attn * keep.float() * ks
    ref = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(B, T, H * dh)
    valid_q = torch.ones(B, T, dtype=torch.bool)   # padded queries are computed too (LM:333 masks keys only)
    e_out = rel(out.float().cpu().view(B, T, -1)[valid_q], ref.detach()[valid_q])
    lse_ref = torch.logsumexp(sim, dim=-1) / np.log(2.0)
    e_lse = float((lse.cpu().view(B, H, T) - lse_ref.detach()).abs().max())
    # backward: zero the upstream gradient on padded frames (as the training loss does)
    dmask = dout.float().view(B, T, -1) * mask[:, :, None]
    ref.backward(dmask)
    dqkv = torch.empty(M, 3 * H * dh, dtype=bf16, device=DEV)
    ops.attention_bwd(qkv, out, dmask.view(M, -1).to(bf16).to(DEV).contiguous(), lse, ln, bits, ks, dqkv,
                      torch.empty(B * H, T, device=DEV), B, T, H, dh)
    got = dqkv.float().cpu().view(B, T, 3, H * dh)
    want = leaf.grad.view(B, T, 3, H * dh)
    errs = [rel(got[:, :, i], want[:, :, i]) for i in range(3)]
    print(f"[parity] attention train dh{dh} T{T} drop={drop}: out {e_out:.2e} lse {e_lse:.2e} dq/dk/dv {errs}")
    assert e_out < 1.5e-2 and e_lse < 2e-2 and max(errs) < 2.5e-2


def _build(z, wseed, multitask=False):
    from diffnorm_b200.plugin.latent_module import LatentDiscreteModel, SpeechVAEEncoderDecoder
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=wseed, gains=O.PARITY_GAINS)
    vae = types.SimpleNamespace(encoder=SpeechVAEEncoderDecoder(768, z))
    ldm = LatentDiscreteModel(vae, 512, z, timesteps=200, multitask=multitask)
    ldm.load_state_dict(sd, strict=True)
    return arch, sd, ldm.to(DEV)


@pytest.mark.parametrize("drop_p,multitask,extreme_t", [(0.1, False, False), (0.0, False, False), (0.0, True, True),
                                                        (0.1, True, True), (0.1, True, False)])
def test_train_step_against_oracle_autograd(drop_p, multitask, extreme_t):
    """Whole step (forward losses + every parameter gradient) vs the oracle's autograd on the same replayed draws;
    multitask adds (50 mse + nll) / T back-propagated through the frozen VAE decoder (LM:1572-1604)."""
    z, wseed, B, T, lengths, times, dseed = 16, 3, 2, 24, [24, 17], [37, 142], 21
    if extreme_t:
        times, dseed = [5, 199], 22          # the case pinned in tests/golden/train_z16_nodrop_multitask.npz
    arch, sd, ldm = _build(z, wseed, multitask)
    audio, units, mask, eps_vae, eps0, eps, keeps = O.train_case_inputs(z, B, T, lengths, dseed, drop_p)
    train_keys = [k for k in sd if k.startswith("model.") and sd[k].is_floating_point() and "pos_embed" not in k]
    for k in train_keys:
        sd[k].requires_grad_(True)
    ref = O.train_loss(sd, arch, audio, units, mask, torch.tensor(times), eps_vae, eps0, eps, keeps, drop_p, multitask)
    ref["total_loss"].backward()
    tr = DenoiserTrainer(ldm, drop_p=drop_p)
    bits = [pack_keep_bits(k).to(DEV) for k in keeps] if keeps is not None else None
    out, grads = tr.step(audio.to(DEV), units.to(DEV), torch.tensor(lengths, dtype=i32, device=DEV),
                         times=torch.tensor(times), noise={"vae": eps_vae, "eps0": eps0, "eps": eps}, keep_bits=bits)
    torch.cuda.synchronize()
    for k in ("noise_loss", "recon_mse_loss", "nll_loss", "total_loss"):
        a, b = float(out[k]), float(ref[k])
        print(f"[parity] train {k}: cuda {a:.6f} oracle {b:.6f}")
        assert abs(a - b) <= 2e-2 * abs(b) + 1e-4
    assert abs(float(out["acc"]) - float(ref["acc"])) <= 0.1
    e_pred = rel(out["pred_noise"].cpu()[mask], ref["pred_noise"].detach()[mask])
    assert e_pred < 3e-2, e_pred
    assert sorted("model." + k for k in grads) == sorted(train_keys)
    worst, tot_num, tot_den = ("", 0.0), 0.0, 0.0
    for k in train_keys:
        gg, gr = grads[k[6:]].cpu().reshape(sd[k].shape), sd[k].grad
        e = rel(gg, gr)
        tot_num += float((gg.double() - gr.double()).pow(2).sum())
        tot_den += float(gr.double().pow(2).sum())
        if e > worst[1]:
            worst = (k, e)
        # softmax-gradient cancellation (dP - D) makes the to_q / to_kv gradients the noisiest under bf16 operands: measured
        # worst 8.1 % (21 % at the extreme timesteps), everything else below 6 % (profiles/r02_b3_gpu_tests_85_passed.log);
        # bounds = 2x the measured worst
        assert e < (0.25 if extreme_t else (0.16 if ".1.to_" in k else 0.12)), (k, e, float(gr.norm()))
        # rounding noise is (nearly) orthogonal to the gradient, a wrong scale factor is not: the NORM of every tensor's
        # gradient must match far more tightly than its direction
        ratio = float(gg.double().norm() / gr.double().norm().clamp_min(1e-30))
        assert abs(ratio - 1.0) < (0.15 if extreme_t else 0.04), (k, ratio)    # measured worst 1.085 at the extreme timesteps
    print(f"[parity] train grads (drop_p={drop_p}, multitask={multitask}): t={times}: global rel err {np.sqrt(tot_num / tot_den):.3e}, worst {worst}, "
          f"pred_noise rel err {e_pred:.2e}")
    # extreme_t case: utterance 1 sits at t = 199 where x1_hat = (x_t - s1 pred) / sqrt(ab) is scaled by 1/2.5e-4, so the
    # decode branch runs on inputs of magnitude ~1e3 and its gradient (x -s1/sa) dominates: the worst-conditioned step there is
    assert np.sqrt(tot_num / tot_den) < (8e-2 if extreme_t else 4e-2)


def test_repack_kernel_bit_identical_to_packing_on_full_model(monkeypatch):
    """dn_pack_weights (one launch over the recorded descriptor table) against diffnorm_b200.packing on every packed tensor
    of the full-size denoiser after an in-place weight update (bit-identical); then two SGD steps with the kernel refresh
    and with the torch-indexing refresh agree (to the run-to-run noise of the step's atomically reduced gradients)."""
    from diffnorm_b200.repack import plan_tensors
    arch, sd, ldm = _build(16, 3)
    tr = DenoiserTrainer(ldm, drop_p=0.0)
    plans = tr._packed()                      # packing.* + record
    named = plan_tensors(plans)
    before = {n: t.clone() for n, t in named}
    g = torch.Generator(device=DEV).manual_seed(5)
    with torch.no_grad():
        for p in tr.P.values():
            p.add_(torch.randn(p.shape, generator=g, device=DEV) * 0.02)
    assert tr._packed() is plans              # refreshed in place by the kernel
    torch.cuda.synchronize()
    fresh = dict(plan_tensors(tr._pack()))
    assert len(named) > 150
    for n, t in named:
        assert torch.equal(t, fresh[n]), n
        assert not torch.equal(t, before[n]), n

    z, B, T, lengths, times = 16, 2, 24, [24, 17], [37, 142]
    audio, units, mask, eps_vae, eps0, eps, _ = O.train_case_inputs(z, B, T, lengths, 21, 0.0)
    res = {}
    for mode in ("kernel", "torch"):
        monkeypatch.setenv("DN_REPACK", mode)
        arch, sd, ldm = _build(16, 3)
        tr = DenoiserTrainer(ldm, drop_p=0.0)
        losses = []
        for it in range(2):
            out, grads = tr.step(audio.to(DEV), units.to(DEV), torch.tensor(lengths, dtype=i32, device=DEV),
                                 times=torch.tensor(times), noise={"vae": eps_vae, "eps0": eps0, "eps": eps})
            losses.append(float(out["total_loss"]))
            with torch.no_grad():             # plain SGD on the masters, in place
                for k, p in tr.P.items():
                    p.add_(grads[k].reshape(p.shape), alpha=-1e-3)
        res[mode] = losses
    print(f"[parity] two SGD steps, total_loss: kernel repack {res['kernel']}, torch repack {res['torch']}")
    assert abs(res["kernel"][0] - res["kernel"][1]) > 0.05          # the update did something
    for a, b in zip(res["kernel"], res["torch"]):
        assert abs(a - b) <= 5e-3 * abs(b)


def test_plugin_forward_backward_through_autograd():
    """The fairseq-facing entry: model(...) -> loss dict; loss.backward() fills .grad of every denoiser parameter with
    the CUDA step's gradients; eval mode runs without dropout and without backward (valid_step)."""
    z, B, T, lengths = 16, 2, 24, [24, 17]
    arch, sd, ldm = _build(z, 3)
    audio, units, mask, eps_vae, eps0, eps, keeps = O.train_case_inputs(z, B, T, lengths, 21, 0.1)
    for n_, p_ in ldm.named_parameters():
        p_.requires_grad_(not n_.startswith("speech_decoder."))
    ldm.train()
    rp = {"times": torch.tensor([37, 142]), "noise": {"vae": eps_vae, "eps0": eps0, "eps": eps},
          "keep_bits": [pack_keep_bits(k).to(DEV) for k in keeps]}
    out = ldm(audio.to(DEV), units.to(DEV), tgt_mask=mask.to(DEV), _replay=rp)
    assert set(out) == {"total_loss", "nll_loss", "recon_mse_loss", "noise_loss", "acc"}
    (out["total_loss"] * 2.0).backward()
    tr = ldm._trainer()
    _, grads = tr.step(audio.to(DEV), units.to(DEV), torch.tensor(lengths, dtype=i32, device=DEV), times=rp["times"],
                       noise=rp["noise"], keep_bits=rp["keep_bits"])
    n = 0
    for name, p_ in ldm.model.named_parameters():
        assert p_.grad is not None, name
        assert rel(p_.grad, 2.0 * grads[name].reshape(p_.shape)) < 1e-3, name   # atomics: not bit-reproducible
        n += 1
    assert n == 377
    assert all(p_.grad is None for n_, p_ in ldm.named_parameters() if n_.startswith("speech_decoder."))
    ldm.eval()
    with torch.no_grad():
        ev = ldm(audio.to(DEV), units.to(DEV), tgt_mask=mask.to(DEV), _replay={"times": rp["times"], "noise": rp["noise"]})
    assert abs(float(ev["total_loss"]) - 0.904014) < 2e-2 * 0.904014   # the oracle's no-dropout loss for this case


@pytest.mark.parametrize("z", [16, 128])
def test_vae_repack_kernel_bit_identical_to_packing(z):
    """The VAE training step's one-launch weight refresh against diffnorm_b200.packing on the full-size VAE (channel-padded
    WaveNet blocks, 768-wide decoder transformer, gamma copies) after an in-place weight update."""
    from diffnorm_b200.plugin.latent_module import SpeechVAEEncoderDecoder
    from diffnorm_b200.repack import plan_tensors
    from diffnorm_b200.train_vae import VaeTrainer
    torch.manual_seed(2)
    tr = VaeTrainer(SpeechVAEEncoderDecoder(768, z).to(DEV), drop_p=0.1)

    def tree(enc_blocks):
        d = tr.dec
        return plan_tensors([enc_blocks, d.blocks, d.layers, d.pred, d.pred_T, d.lm, d.lm_T, d.pred_gamma])
    named = tree(tr._packed())
    before = {n: t.clone() for n, t in named}
    g = torch.Generator(device=DEV).manual_seed(6)
    with torch.no_grad():
        for p in tr.P.values():
            p.add_(torch.randn(p.shape, generator=g, device=DEV) * 0.02)
    tr._packed()                               # kernel refresh, in place
    torch.cuda.synchronize()
    kept = [(n, t.clone()) for n, t in named]
    fresh = dict(tree(tr._pack()))             # packing.* from the updated masters (rebinds tr.dec.*)
    assert len(kept) > 100
    for n, t in kept:
        assert torch.equal(t, fresh[n]), n
        assert not torch.equal(t, before[n]), n


@pytest.mark.parametrize("z,wseed", [(16, 4), (128, 6)])
def test_vae_train_step_against_oracle_autograd(z, wseed):
    """VAE training: SpeechVAEEncoderDecoder.forward -> (mse, lm_pred, kl) and the gradients of the criterion's
    loss 0.1 LS-NLL/ntokens + 10 mse + 1e-4 kl w.r.t. all 274 VAE tensors, vs the oracle's autograd (pinned to the reference)."""
    import torch.nn.functional as F
    from diffnorm_b200.plugin.latent_module import SpeechVAEEncoderDecoder
    from diffnorm_b200.train_vae import VaeTrainer
    B, T, lengths, dseed, drop_p = 2, 24, [24, 15], 31, 0.1
    arch = O.Arch(latent_dim=z)
    sd = O.init_state_dict(arch, seed=wseed, gains=O.PARITY_GAINS)
    vae = SpeechVAEEncoderDecoder(768, z)
    vae.load_state_dict({k[len("speech_decoder."):]: v for k, v in sd.items() if k.startswith("speech_decoder.")}, strict=True)
    vae = vae.to(DEV)
    audio, units, mask, eps_vae, _, _, keeps = O.train_case_inputs(z, B, T, lengths, dseed, drop_p, depth=arch.vae_depth)
    keys = [k for k in sd if k.startswith("speech_decoder.") and sd[k].is_floating_point()]
    for k in keys:
        sd[k].requires_grad_(True)
    ref = O.vae_train_loss(sd, arch, audio, units, mask, eps_vae, keeps, drop_p)
    ref["loss"].backward()
    tr = VaeTrainer(vae, drop_p=drop_p)
    mse, logits, kl = tr.forward(audio.to(DEV), torch.tensor(lengths, dtype=i32, device=DEV), eps_vae,
                                 [pack_keep_bits(k).to(DEV) for k in keeps])
    # the criterion's arithmetic on the returned logits (speech_vae_decoder_loss.py:60-82), plain torch
    lg = logits.detach().clone().requires_grad_(True)
    lprobs = F.log_softmax(lg, dim=-1).view(-1, lg.shape[-1])
    u = units.to(DEV).view(-1)
    ls, nll = O.label_smoothed_nll_loss(lprobs, u, 0.1, 0)
    ntok = int(u.ne(0).sum())
    (0.1 * ls / ntok).backward()
    loss = 0.1 * ls.detach() / ntok + 10 * mse + 1e-4 * kl
    for name, a, b in (("mse", mse, ref["mse_loss"]), ("kl", kl, ref["kl_loss"]), ("nll", nll / ntok, ref["nll_loss"]), ("loss", loss, ref["loss"])):
        print(f"[parity] vae train {name}: cuda {float(a):.6f} oracle {float(b):.6f}")
        assert abs(float(a) - float(b)) <= 2e-2 * abs(float(b)) + 1e-4
    grads = tr.backward(10.0, lg.grad, 1e-4)
    torch.cuda.synchronize()
    assert sorted("speech_decoder." + k for k in grads) == sorted(keys)
    worst, num, den = ("", 0.0), 0.0, 0.0
    for k in keys:
        gg, gr = grads[k[len("speech_decoder."):]].cpu().reshape(sd[k].shape), sd[k].grad
        e = rel(gg, gr)
        num += float((gg.double() - gr.double()).pow(2).sum())
        den += float(gr.double().pow(2).sum())
        if e > worst[1]:
            worst = (k, e)
        assert e < 0.25, (k, e, float(gr.norm()))
    print(f"[parity] vae train grads: global rel err {np.sqrt(num / den):.3e}, worst {worst}")
    assert np.sqrt(num / den) < 4e-2


def test_vae_plugin_criterion_backward():
    """speech_decoder task surface: criterion(model, sample) -> loss; loss.backward() fills .grad of all 274 VAE tensors."""
    import argparse
    from diffnorm_b200.plugin import compat
    z, B, T, lengths = 16, 2, 24, [24, 15]
    args = argparse.Namespace(task="speech_decoder", arch="speech_vae_decoder", target_is_code=True, target_code_size=1000,
                              latent_dim=z, criterion="speech_vae_decoder_loss")
    task = compat.setup_task(args)
    model = task.build_model(args).to(DEV).train()
    crit = task.build_criterion(args)
    audio, units, mask, eps_vae, _, _, keeps = O.train_case_inputs(z, B, T, lengths, 31, 0.1, depth=6)
    lens = torch.tensor(lengths, device=DEV)
    sample = {"net_input": {"src_tokens": audio.to(DEV), "src_lengths": lens}, "reduce_target": audio.to(DEV),
              "reduce_target_unit": units.to(DEV), "reduce_target_lengths": lens, "ntokens": int(units.ne(0).sum()),
              "nsentences": B}
    loss, sample_size, log = crit(model, sample)
    loss.backward()
    n = sum(1 for p_ in model.parameters() if p_.grad is not None and torch.isfinite(p_.grad).all() and float(p_.grad.abs().sum()) > 0)
    print(f"[parity] vae plugin: loss {float(loss):.4f} log {log}; tensors with gradient {n}")
    assert n == 274 and sample_size == B and set(log) >= {"loss", "nll_loss", "mse_loss", "kl_loss", "acc"}
    model.eval()
    with torch.no_grad():
        loss_eval, _, _ = crit(model, sample)
    assert torch.isfinite(loss_eval)


@pytest.mark.parametrize("z", [32, 128])
def test_other_latent_dims_inference_and_training(z):
    """latent_dim 32 / 128 (LM:1044-1051 chan_mults; 128 is the only value the reference's scripts deploy): one denoiser
    call + decode vs the oracle, and the training step's loss + gradients vs the oracle's autograd."""
    B, T, lengths, times = 2, 24, [24, 13], [61, 120]
    arch, sd, ldm = _build(z, 5)
    audio, units, mask, eps_vae, eps0, eps, keeps = O.train_case_inputs(z, B, T, lengths, 41, 0.1)
    eng = ldm._engine()
    lens = torch.tensor(lengths, dtype=i32, device=DEV)
    x = torch.randn(B, T, z, generator=torch.Generator().manual_seed(3))
    tt = torch.tensor(times)
    want = O.denoiser(sd, arch, x, torch.full((B,), 77), mask)
    got = eng.denoise(eng.stage_latent(x.to(DEV)), lens, B, T, torch.tensor([77], dtype=i32, device=DEV)).view(B, T, -1)[..., :z].cpu()
    assert float((got - want)[mask].abs().max()) <= 5e-2 * float(want[mask].std()) + 2e-2
    rec_w, log_w = O.vae_decode(sd, arch, x, mask)
    rec_g, log_g = eng.decode(eng.stage_latent(x.to(DEV)), lens, B, T)
    assert float((log_g.cpu()[..., :arch.vocab] - log_w)[mask].abs().max()) <= 5e-2 * float(log_w[mask].std())
    keys = [k for k in sd if k.startswith("model.") and sd[k].is_floating_point() and "pos_embed" not in k]
    for k in keys:
        sd[k].requires_grad_(True)
    ref = O.train_loss(sd, arch, audio, units, mask, tt, eps_vae, eps0, eps, keeps, 0.1, False)
    ref["total_loss"].backward()
    out, grads = DenoiserTrainer(ldm, drop_p=0.1).step(audio.to(DEV), units.to(DEV), lens, times=tt, noise={"vae": eps_vae, "eps0": eps0, "eps": eps},
                                                        keep_bits=[pack_keep_bits(k).to(DEV) for k in keeps])
    assert abs(float(out["noise_loss"]) - float(ref["noise_loss"])) <= 2e-2 * float(ref["noise_loss"]) + 1e-4
    num = sum(float((grads[k[6:]].cpu().reshape(sd[k].shape).double() - sd[k].grad.double()).pow(2).sum()) for k in keys)
    den = sum(float(sd[k].grad.double().pow(2).sum()) for k in keys)
    print(f"[parity] z={z}: training grads global rel err {np.sqrt(num / den):.3e}")
    assert np.sqrt(num / den) < 4e-2


@pytest.mark.parametrize("task_name,arch,criterion,n_grad", [("speech_decoder", "speech_vae_decoder", "speech_vae_decoder_loss", 274),
                                                              ("speech_diffusion_discrete", "diff_discrete", "ddpm_discrete_loss", 377)])
def test_load_dataset_then_criterion_runs(tmp_path, task_name, arch, criterion, n_grad):
    """The fairseq-train data path end to end: task.load_dataset("train") on an on-disk corpus (manifests + .npy features + unit
    TSV, the formats of repr_to_repr_unit_dataset.py:309-369), the dataset's own batch sampler + collater, then
    criterion(model, sample) and loss.backward() on the CUDA training step — for the VAE task and for the diffusion task."""
    import argparse
    from diffnorm_b200.plugin import compat
    from oracle.dataset_fixture import write_corpus
    src_dir, tgt_dir, tsv_dir = write_corpus(str(tmp_path), dim=768)
    args = argparse.Namespace(task=task_name, arch=arch, data=tsv_dir, src_feat_dir=src_dir, tgt_feat_dir=tgt_dir,
                              target_is_code=True, target_code_size=1000, latent_dim=16, criterion=criterion, dummy_config=None,
                              seed=1)
    task = compat.setup_task(args)
    task.load_dataset("train")
    ds = task.dataset("train")
    model = task.build_model(args).to(DEV).train()
    crit = task.build_criterion(args)
    batches = ds.batch_sampler(max_tokens=200)
    assert len(batches) >= 2 and sum(len(b) for b in batches) == len(ds)
    sample = ds.collater([ds[int(i)] for i in batches[0]])

    def to_dev(x):
        if torch.is_tensor(x):
            return x.to(DEV)
        if isinstance(x, dict):
            return {k: to_dev(v) for k, v in x.items()}
        return x

    sample = to_dev(sample)
    # what the collater promises the criterion: reduced targets, 0-padded, ntokens = sum of the reduced lengths
    assert sample["ntokens"] == int(sample["reduce_target_lengths"].sum()) == int(sample["reduce_target_unit"].ne(0).sum())
    loss, sample_size, log = crit(model, sample)
    assert torch.isfinite(loss) and sample_size == len(batches[0]) and log["ntokens"] == sample["ntokens"]
    loss.backward()
    train_params = [p_ for n_, p_ in model.named_parameters() if p_.requires_grad and p_.grad is not None]
    n = sum(1 for p_ in train_params if torch.isfinite(p_.grad).all() and float(p_.grad.abs().sum()) > 0)
    print(f"[parity] {task_name}: loss {float(loss):.4f}, {n} tensors with gradient, log {log}")
    assert n == n_grad
