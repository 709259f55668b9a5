"""Drop-in counterparts of the reference's ``latent_module`` classes for the normalization path.

Same class names, constructor arguments, public methods and ``state_dict`` keys as
fairseq/models/text_to_speech/latent_module.py (SpeechVAEEncoderDecoder :1035, Model :709,
LatentDiscreteModel :1300, DDPMScheduler :1241), so reference checkpoints load with ``strict=True`` — but the
modules are *parameter containers*: all arithmetic of ``encode_feature`` / ``decode_feature`` / ``ddim_sample``
runs in the sm_100a kernels through DiffNormEngine.  There is no torch / CPU implementation behind these
methods: on a machine without CUDA they raise.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch
from torch import nn

from ..config import DiffNormConfig
from ..schedule import DDPMScheduler  # noqa: F401  (re-exported under the reference's name)
from .compat import FairseqEncoder

# (key, shape, kind, fan_in); kind in {"w", "b", "ones", "randn", "buf"}
Spec = Tuple[str, Tuple[int, ...], str, int]


def _lin(key: str, out_f: int, in_f: int, k: int = 0, bias: bool = True) -> List[Spec]:
    shape = (out_f, in_f, k) if k > 0 else (out_f, in_f)
    fan_in = in_f * max(k, 1)
    s = [(key + ".weight", shape, "w", fan_in)]
    if bias:
        s.append((key + ".bias", (out_f,), "b", fan_in))
    return s


def _wavenet_spec(pre: str, cin: int, c: int, stacks: int, layers: int, dim_time: Optional[int]) -> List[Spec]:
    s = _lin(pre + "init_conv", c, cin, 3)
    for st in range(stacks):
        for b in range(layers):
            p = f"{pre}stacks.{st}.blocks.{b}."
            if dim_time is not None:
                s += _lin(p + "to_time_cond", 2 * c, dim_time)
            s += _lin(p + "conv", c, c, 3) + _lin(p + "res_conv", c, c, 1)
            if st == stacks - 1:
                s += _lin(p + "skip_conv", c, c, 1)
    return s + _lin(pre + "final_conv", c, c, 1)


def _transformer_spec(pre: str, dim: int, depth: int, heads: int, dim_head: int, dim_time: Optional[int]) -> List[Spec]:
    inner = DiffNormConfig.ff_inner(dim)
    hd = heads * dim_head
    s: List[Spec] = []
    for l in range(depth):
        p = f"{pre}layers.{l}."

        def norm(n):
            return _lin(p + f"{n}.to_gamma_beta", 2 * dim, dim_time) if dim_time is not None else [(p + f"{n}.gamma", (dim,), "ones", 0)]
        # the reference's ModuleList order (LM:666-674): norm, attention, norm, feed-forward — parameters() (and with it the
        # optimizer state of a resumed checkpoint) follows registration order
        s += norm(0)
        s += _lin(p + "1.to_q", hd, dim, bias=False) + _lin(p + "1.to_kv", 2 * hd, dim, bias=False)
        s += _lin(p + "1.to_out", dim, hd, bias=False)
        s += norm(4)
        s += _lin(p + "5.0", 2 * inner, dim) + _lin(p + "5.2.1", inner, inner, 3) + _lin(p + "5.3", dim, inner)
    s.append((pre + "to_pred.0.gamma", (dim,), "ones", 0))
    return s + _lin(pre + "to_pred.1", dim, dim, bias=False)


def vae_spec(cfg: DiffNormConfig, pre: str = "") -> List[Spec]:
    s: List[Spec] = []
    for i, (cin, cout) in enumerate(cfg.enc_widths()):
        s += _wavenet_spec(f"{pre}encoder_wave.{i}.", cin, cout, cfg.vae_stacks, cfg.vae_layers, None)
    for i, (cin, cout) in enumerate(cfg.dec_widths()):
        s += _wavenet_spec(f"{pre}decoder_wave.{i}.", cin, cout, cfg.vae_stacks, cfg.vae_layers, None)
    s += _transformer_spec(pre + "decoder_tf.", cfg.feat_dim, cfg.vae_depth, cfg.vae_heads, cfg.vae_dim_head, None)
    return s + _lin(pre + "decoder_lm", cfg.vocab, cfg.feat_dim)


def denoiser_spec(cfg: DiffNormConfig, pre: str = "") -> List[Spec]:
    s = _lin(pre + "init_conv", cfg.hid, cfg.latent_dim, 1)
    s.append((pre + "to_time_cond.0.weights", (cfg.hid // 2,), "randn", 0))
    s += _lin(pre + "to_time_cond.1", cfg.dim_time, cfg.hid + 1)
    s.append((pre + "pos_embed._float_tensor", (1,), "buf", 0))
    s += _wavenet_spec(pre + "wavenet.", cfg.hid, cfg.hid, cfg.wn_stacks, cfg.wn_layers, cfg.dim_time)
    s += _transformer_spec(pre + "transformer.", cfg.hid, cfg.depth, cfg.heads, cfg.dim_head, cfg.dim_time)
    return s + _lin(pre + "final_proj", cfg.latent_dim, cfg.hid)


def _materialise(root: nn.Module, spec: Iterable[Spec]):
    """Create the nested sub-module tree named by the dotted keys and register parameters with torch's default
    init laws (nn.Linear / nn.Conv1d: U(+-1/sqrt(fan_in)) for weight and bias; gamma = 1; LM:108 randn)."""
    for key, shape, kind, fan_in in spec:
        *path, leaf = key.split(".")
        mod = root
        for name in path:
            nxt = mod._modules.get(name)
            if nxt is None:
                nxt = nn.Module()
                mod.add_module(name, nxt)
            mod = nxt
        if kind == "buf":
            mod.register_buffer(leaf, torch.zeros(shape))
            continue
        t = torch.empty(shape)
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t.uniform_(-bound, bound)
        elif kind == "ones":
            t.fill_(1.0)
        else:
            t.normal_()
        mod.register_parameter(leaf, nn.Parameter(t))


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: diffnorm_b200 runs on CUDA (sm_100a) only; move the model and inputs to a GPU")


class SpeechVAEEncoderDecoder(FairseqEncoder):
    """LM:1035-1142.  ``encode_feature`` / ``decode_feature`` keep the reference's signatures and layouts."""

    def __init__(self, dim: int = 768, latent_dim: int = 16):
        super().__init__(None)
        self.dim, self.latent_dim = dim, latent_dim
        self.cfg = DiffNormConfig(latent_dim=latent_dim, feat_dim=dim)
        _materialise(self, vae_spec(self.cfg))
        self._owner = None  # set by LatentDiscreteModel so both share one engine
        self._eng = None    # stand-alone VAE (speech_vae_decoder checkpoints): its own VAE-only engine
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        self._eng = None
        self._vae_trainer_obj = None

    def train(self, mode: bool = True):
        # optimizers update weights through p.data (fairseq/optim/adam.py:183-237), which no version counter sees: any
        # switch into training mode drops the packed inference weights, they are re-packed on the next inference call
        if mode:
            self._eng = None
        return super().train(mode)

    def _engine(self):
        if self._owner is not None and self._owner() is not None:
            return self._owner()._engine()
        p = next(self.parameters())
        _require_cuda(p, "SpeechVAEEncoderDecoder")
        if self._eng is None:
            from ..engine import DiffNormEngine
            sd = {"speech_decoder." + k: v for k, v in self.state_dict().items()}
            self._eng = DiffNormEngine(sd, device=str(p.device), cfg=self.cfg, vae_only=True)
        return self._eng

    @torch.no_grad()
    def encode_feature(self, feature: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """feature [B,T,768] -> posterior sample, CHANNEL-FIRST [B,z,T] like the reference (LM:1099-1107).
        ``noise`` ([B,z,T], optional) replays the reference's CPU draw (distributions.py:38)."""
        _require_cuda(feature, "encode_feature")
        B, T, _ = feature.shape
        if noise is None:
            noise = torch.randn(B, self.latent_dim, T).to(feature.device)  # CPU draw then copy, as the reference
        z = self._engine().encode(feature.float().contiguous(), noise.to(feature.device).float())
        return z.transpose(1, 2)

    @torch.no_grad()
    def decode_feature(self, latent: torch.Tensor, mask: torch.Tensor):
        """latent [B,T,z], mask [B,T] bool -> (decoded_feature [B,T,768], lm_result [B,T,1004]) (LM:1109-1116)."""
        _require_cuda(latent, "decode_feature")
        eng = self._engine()
        B, T, _ = latent.shape
        lens = _mask_to_lengths(mask)
        xb = eng.stage_latent(latent.float())
        recon, logits = eng.decode(xb, lens, B, T)
        return recon.clone(), logits[..., : eng.cfg.vocab].clone()

    # ---- training entry (LM:1118-1142) ------------------------------------------------------------------------
    def _vae_trainer(self):
        if getattr(self, "_vae_trainer_obj", None) is None:
            from ..train_vae import VaeTrainer
            self._vae_trainer_obj = VaeTrainer(self, drop_p=0.1)
        return self._vae_trainer_obj

    def forward(self, input_feature, input_token, mask, _replay: Optional[Dict[str, object]] = None):
        """(mse_loss, lm_result [B,T,1004], kl_loss) like the reference; forward and backward run in the sm_100a kernels
        (diffnorm_b200/train_vae.py).  The three outputs carry an autograd node: whatever the criterion builds from them
        (speech_vae_decoder_loss.py:60-82) back-propagates into the CUDA backward, which fills the parameter gradients.
        ``_replay`` (extension) = {"eps_vae", "keep_bits"} replays the random draws for parity runs."""
        _require_cuda(input_feature, "SpeechVAEEncoderDecoder.forward")
        tr = self._vae_trainer()
        lens = _mask_to_lengths(mask)
        names = list(tr.P.keys())
        params = [tr.P[n] for n in names]
        need_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in params)
        return _VaeStepFn.apply(tr, input_feature, lens, _replay or {}, self.training, need_grad, names, *params)


class _VaeStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer, feat, lens, replay, training, need_grad, names, *params):
        mse, logits, kl = trainer.forward(feat, lens, eps_vae=replay.get("eps_vae"), keep_bits=replay.get("keep_bits"),
                                          train=training)
        ctx.trainer, ctx.names, ctx.shapes, ctx.need_grad = trainer, names, [p.shape for p in params], need_grad
        return mse.clone(), logits.clone(), kl.clone()

    @staticmethod
    def backward(ctx, g_mse, g_logits, g_kl):
        if not ctx.need_grad:
            raise RuntimeError("backward through a VAE step that ran without gradients (eval mode / no_grad)")
        grads = ctx.trainer.backward(0.0 if g_mse is None else float(g_mse), g_logits, 0.0 if g_kl is None else float(g_kl))
        return (None,) * 7 + tuple(grads[n].reshape(sh) for n, sh in zip(ctx.names, ctx.shapes))


def _mask_to_lengths(mask: torch.Tensor) -> torch.Tensor:
    """Right-padded boolean mask -> int32 lengths.  Non-prefix masks are rejected (the kernels take lengths)."""
    lens = mask.sum(dim=1)
    expect = torch.arange(mask.shape[1], device=mask.device)[None, :] < lens[:, None]
    if not torch.equal(expect, mask.bool()):
        raise ValueError("input_mask must be a right-padded prefix mask (lengths_to_mask form)")
    return lens.to(torch.int32).contiguous()


class Model(nn.Module):
    """Denoiser eps-predictor (LM:709-876): parameter container; executed by DiffNormEngine.denoise."""

    def __init__(self, dim: int, latent_dim: int, cfg: Optional[DiffNormConfig] = None):
        super().__init__()
        self.dim, self.latent_dim = dim, latent_dim
        self.cfg = cfg or DiffNormConfig(latent_dim=latent_dim, hid=dim)
        _materialise(self, denoiser_spec(self.cfg))


class LatentDiscreteModel(FairseqEncoder):
    """LM:1300-1613.  Constructor mirrors the reference: ``speech_decoder`` is the VAE *model wrapper* whose
    ``.encoder`` is a SpeechVAEEncoderDecoder (LM:1321)."""

    def __init__(self, speech_decoder, dim: int, latent_dim: int, timesteps: int = 1000, multitask: bool = True,
                 use_cond: bool = False, **unused):
        super().__init__(None)
        if use_cond:
            raise NotImplementedError("condition_on_prompt=True (LM:1326-1333) is not on the shipped path")
        self.speech_decoder = speech_decoder.encoder
        self.use_cond, self.multitask = use_cond, multitask
        self.cfg = DiffNormConfig(latent_dim=latent_dim, hid=dim, timesteps=timesteps)
        self.model = Model(dim, latent_dim, self.cfg)
        self.scheduler = DDPMScheduler(timesteps)
        self.dim, self.timesteps = dim, timesteps
        self._eng = None
        self._eng_version = None
        self._trainer_obj = None
        import weakref
        self.speech_decoder._owner = weakref.ref(self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # ---- engine management -------------------------------------------------------------------------------
    def _invalidate(self):
        self._eng = None
        self._trainer_obj = None   # holds the packed frozen VAE

    def train(self, mode: bool = True):
        # fairseq's optimizers write through p.data (optim/adam.py:183-237): no version counter changes, so staleness of
        # the packed inference weights is keyed on the mode switch instead: entering training mode (and every training
        # forward, below) drops the engine; the next inference call re-packs from the current parameters.
        if mode:
            self._eng = None
        return super().train(mode)

    def _engine(self):
        p = next(self.model.parameters())
        _require_cuda(p, "LatentDiscreteModel")
        dev = (p.device, p.data_ptr())
        if self._eng is None or self._eng_version != dev:   # also re-pack after .to(device) / a re-allocated parameter
            from ..engine import DiffNormEngine
            sd = {k: v for k, v in self.state_dict().items()}
            self._eng = DiffNormEngine(sd, device=str(p.device), cfg=self.cfg)
            self._eng_version = dev
        return self._eng

    @property
    def device(self):
        return next(self.model.parameters()).device

    # ---- inference entry used by the driver (diff_norm_synthesis.py:204) -------------------------------------
    @torch.no_grad()
    def ddim_sample(self, tgt_feature, prompt=None, prompt_mask=None, input_mask=None, cond_scale=1., ref_units=None,
                    start_step=50, noise: Optional[Dict[str, torch.Tensor]] = None):
        """LM:1386-1471.  Returns (out_tokens: list of B int64 tensors, match, total, recon_feature [B,T,768]).
        ``noise`` (extension) = {"vae": [B,z,T], "q": [B,T,z]} replays the reference's two draws for parity runs."""
        _require_cuda(tgt_feature, "ddim_sample")
        eng = self._engine()
        B, T, _ = tgt_feature.shape
        lens = _mask_to_lengths(input_mask)
        noise = noise or {}
        ev = noise.get("vae")
        if ev is None:
            ev = torch.randn(B, self.cfg.latent_dim, T)  # the reference draws this one on the CPU
        out = eng.normalize(tgt_feature.float().contiguous(), lens, int(start_step), ev.to(tgt_feature.device).float(),
                            None if noise.get("q") is None else noise["q"].to(tgt_feature.device).float(),
                            ref_units=None if ref_units is None else ref_units.to(torch.int64), reduce=False)
        units = out["units"]
        match = total = 0
        if ref_units is not None:
            match, total = (int(v) for v in out["acc"].tolist())  # one sync instead of the reference's two .item()
        lens_h = lens.tolist()
        out_tokens = [units[i, : lens_h[i]].clone() for i in range(B)]
        return out_tokens, match, total, out["recon"].clone()

    @torch.no_grad()
    def normalize_units(self, tgt_feature, lengths, start_step=50, noise=None):
        """Extension over the reference API: the fused tail of the driver loop (diff_norm_synthesis.py:204-216):
        returns the device tensors (units, dedup, duration, index_to_keep, counts) without per-utterance host work."""
        eng = self._engine()
        noise = noise or {}
        out = eng.normalize(tgt_feature.float().contiguous(), lengths.to(torch.int32), int(start_step),
                            noise.get("vae"), noise.get("q"))
        return out

    # ---- training entry (LM:1514-1613) ------------------------------------------------------------------------
    def _trainer(self):
        if self._trainer_obj is None:
            from ..train import DenoiserTrainer
            self._trainer_obj = DenoiserTrainer(self, drop_p=0.1)
        return self._trainer_obj

    def forward(self, audio, audio_units, src_feature=None, src_mask=None, tgt_mask=None, prompt=None, pitch=None,
                *args, _replay: Optional[Dict[str, object]] = None, **kwargs):
        """LatentDiscreteModel.forward: the denoiser training loss dict {total_loss, nll_loss, recon_mse_loss,
        noise_loss, acc}.  The whole step (forward AND backward) runs in the sm_100a kernels (diffnorm_b200/train.py);
        ``total_loss`` is tied to the denoiser parameters through an autograd Function whose backward hands the
        already-computed gradients to torch, so ``optimizer.backward(loss)`` (fairseq) / ``loss.backward()`` work
        unchanged.  In eval mode (valid_step) there is no dropout and no backward.
        ``_replay`` (extension) = {"times", "noise", "keep_bits"} replays the random draws for parity runs."""
        _require_cuda(audio, "LatentDiscreteModel.forward")
        if tgt_mask is None:
            raise ValueError("tgt_mask is required (diff_discrete.py:46-54 always passes it)")
        tr = self._trainer()
        tr.drop_p = 0.1 if self.training else 0.0
        if self.training:
            self._eng = None      # the weights are about to move: the packed inference copy is stale from here on
        lens = _mask_to_lengths(tgt_mask)
        names = list(tr.P.keys())
        params = [tr.P[n] for n in names]
        need_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in params)
        rp = _replay or {}
        outs = _TrainStepFn.apply(tr, audio, audio_units, lens, rp, need_grad, names, *params)
        total, nll, mse, acc = outs
        noise = total.detach() - (50.0 * mse + nll) / self.timesteps if self.multitask else total.detach()
        return {"total_loss": total, "nll_loss": nll, "recon_mse_loss": mse, "noise_loss": noise, "acc": acc}


class _TrainStepFn(torch.autograd.Function):
    """Forward = the whole CUDA training step (losses + parameter gradients); backward = scale and return them."""

    @staticmethod
    def forward(ctx, trainer, audio, units, lens, replay, need_grad, names, *params):
        out, grads = trainer.step(audio, units, lens, times=replay.get("times"), noise=replay.get("noise"),
                                  keep_bits=replay.get("keep_bits"), backward=need_grad)
        ctx.grads = [grads[n].reshape(p.shape) for n, p in zip(names, params)] if need_grad else None
        zero = torch.zeros((), device=audio.device)
        res = (out["total_loss"].clone(), out.get("nll_loss", zero).clone(), out.get("recon_mse_loss", zero).clone(),
               out.get("acc", zero).clone())
        ctx.mark_non_differentiable(*res[1:])
        return res

    @staticmethod
    def backward(ctx, g_total, *unused):
        if ctx.grads is None:
            raise RuntimeError("backward through a step that ran without gradients (eval mode / no_grad)")
        return (None,) * 7 + tuple(g * g_total for g in ctx.grads)
