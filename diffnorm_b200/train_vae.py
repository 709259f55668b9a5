"""Training step of the VAE: ``SpeechVAEEncoderDecoder.forward`` (LM:1118-1142) — WaveNet encoder ->
diagonal-Gaussian posterior (sample + masked KL) -> WaveNet decoder -> 6-layer transformer (attention dropout 0.1 in train
mode) -> to_pred / decoder_lm — and its backward over all 274 tensors, on the same sm_100a kernels as the denoiser step.

``forward`` returns what the reference module returns, ``(mse_loss, lm_pred, kl_loss)``; ``backward`` takes the upstream
gradients of those three (the criterion combines them as 0.1 LS-NLL/ntokens + 10 mse + 1e-4 kl,
speech_vae_decoder_loss.py:60-82) and returns fp32 gradients keyed like ``SpeechVAEEncoderDecoder.named_parameters()``.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

from . import ops
from .config import DiffNormConfig
from .packing import rup
from .repack import PackTable
from .train import FrozenDecoderTrain, VaeBlocksTrain, ZeroArena, _GradDict, _pack_vae_wavenet

bf16, f32, i32, i64 = torch.bfloat16, torch.float32, torch.int32, torch.int64


class VaeTrainer:
    def __init__(self, vae, drop_p: float = 0.1, seed: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("VaeTrainer needs a CUDA device: the product has no CPU path")
        self.vae_mod = vae
        self.cfg: DiffNormConfig = vae.cfg
        self.P: Dict[str, torch.nn.Parameter] = dict(vae.named_parameters())
        self.dev = next(iter(self.P.values())).device
        self.drop_p, self.seed, self.step_no = drop_p, seed, 0
        self.ws: Dict[tuple, torch.Tensor] = {}
        c = self.cfg
        self.G, self.S = c.vae_layers, c.vae_stacks
        self.arena = ZeroArena(self.dev)
        self.enc = VaeBlocksTrain(self.buf, self.G, self.S, "e", self.arena.zeros)
        self.dec = FrozenDecoderTrain(None, c, self.dev, self.buf, zeros=self.arena.zeros)
        self._pack_graph, self._pack_ptrs, self.enc_blocks, self._pack_table = None, None, None, None
        self.ctx = None

    def buf(self, name: str, rows: int, width: int, dtype=bf16, zero: bool = False) -> torch.Tensor:
        key = (name, width, dtype)
        t = self.ws.get(key)
        if t is None or t.shape[0] < rows:
            t = torch.zeros(rows, width, dtype=dtype, device=self.dev)
            self.ws[key] = t
        v = t[:rows]
        if zero:
            v.zero_()
        return v

    # ------------------------------------------------------------------------------------------------ packing
    def _pack(self, rec=None):
        c = self.cfg
        w = lambda k: self.P[k].detach().float()
        blocks, cin_pad = [], c.feat_dim
        enc_w = c.enc_widths()
        for i, (cin, cout) in enumerate(enc_w):
            b = _pack_vae_wavenet(w, f"encoder_wave.{i}.", cin, cout, cin_pad, self.G, self.S, i == len(enc_w) - 1, self.dev, rec)
            blocks.append(b)
            cin_pad = b.cp
        self.dec.pack(w, rec)
        return blocks

    def _packed(self):
        """Packed once with diffnorm_b200.packing; every later step refreshes the same packed tensors in place from the
        fp32 masters with ONE dn_pack_weights launch (see DenoiserTrainer._packed; DN_REPACK=torch = the CUDA-graph replay of
        the torch indexing kernels)."""
        ptrs = tuple(p.data_ptr() for p in self.P.values())
        if os.environ.get("DN_REPACK", "kernel") == "torch":
            if self._pack_graph is None or ptrs != self._pack_ptrs:
                self._pack()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.enc_blocks = self._pack()
                self._pack_graph, self._pack_ptrs = g, ptrs
            self._pack_graph.replay()
            return self.enc_blocks
        if self._pack_table is None or ptrs != self._pack_ptrs:
            rec = PackTable()
            self.enc_blocks = self._pack(rec)
            self._pack_table, self._pack_ptrs = rec.finalize(self.dev), ptrs
            return self.enc_blocks
        self._pack_table.run()
        return self.enc_blocks

    # ------------------------------------------------------------------------------------------------ forward / backward
    @torch.no_grad()
    def forward(self, feat: torch.Tensor, lengths: torch.Tensor, eps_vae: Optional[torch.Tensor] = None,
                keep_bits: Optional[Sequence[torch.Tensor]] = None, train: bool = True):
        """feat fp32 [B,T,768] cuda, lengths int32 [B] -> (mse_loss, lm_pred fp32 [B,T,vocab], kl_loss).
        eps_vae [B,z,T] replays the posterior draw (distributions.py:38 draws it on the CPU); keep_bits = per decoder layer
        int32 [B,H,T,ceil(T/32)] replays the attention dropout (drawn with Philox when None and train)."""
        c, dev = self.cfg, self.dev
        B, T, Cf = feat.shape
        M, z = B * T, c.latent_dim
        lens = lengths.to(device=dev, dtype=i32).contiguous()
        feat = feat.float().contiguous()
        eps = (torch.randn(B, z, T).to(dev) if eps_vae is None else eps_vae.to(dev)).float().contiguous()
        drop = train and (self.drop_p > 0 or keep_bits is not None)
        keep_scale = 1.0 / (1.0 - self.drop_p) if drop else 1.0
        if drop and keep_bits is None:
            Tw = (T + 31) // 32
            keep_bits = []
            for l in range(c.vae_depth):
                kb = torch.empty(B, c.vae_heads, T, Tw, dtype=i32, device=dev)
                ops.dropout_bits(kb, self.drop_p, self.seed, self.step_no * c.vae_depth + l)
                keep_bits.append(kb)
        self.step_no += 1
        enc_blocks = self._packed()
        a = ops.cast_pad_bf16(feat.view(M, Cf), Cf, out=self.buf("e.in", M, Cf))
        params = self.enc.forward(enc_blocks, a, B, T)                       # fp32 [M, rup(2z, 16)]: mean | logvar
        p3 = params.view(B, T, -1)
        zlat = ops.vae_reparam(p3, eps, z, True)
        kl = ops.vae_kl(p3, lens, z, torch.zeros(1, dtype=f32, device=dev))
        xb = ops.cast_pad_bf16(zlat.view(M, z), self.dec.zp, out=self.buf("v.zb", M, self.dec.zp))
        recon, logits = self.dec.forward(xb, lens, B, T, keep_bits if drop else None, keep_scale)
        units0 = self.buf("units0", M, 1, i64).view(-1)
        st = ops.decode_losses(recon, feat.view(M, Cf), logits, c.vocab, units0, lens, B, T)   # [0] sum sq err, [5] valid frames
        mse = (st[0] / (st[5].clamp(min=1) * Cf)).float()
        self.ctx = dict(B=B, T=T, lens=lens, feat=feat.view(M, Cf), eps=eps, params=p3, recon=recon, stats=st, enc_blocks=enc_blocks)
        return mse, logits.view(B, T, -1)[..., : c.vocab], kl[0]

    @torch.no_grad()
    def backward(self, g_mse: float, g_logits: Optional[torch.Tensor], g_kl: float, grad_hook=None) -> Dict[str, torch.Tensor]:
        """Upstream gradients of (mse_loss, lm_pred, kl_loss) -> {parameter name: fp32 gradient}."""
        cx, c, dev = self.ctx, self.cfg, self.dev
        B, T, lens = cx["B"], cx["T"], cx["lens"]
        M, z = B * T, c.latent_dim
        grads = _GradDict(grad_hook)
        self.arena.reset()                # gradient tensors are views into the arena: valid until the next backward
        dlogits = self.buf("v.dlogits", M, self.dec.vl, zero=g_logits is None)
        if g_logits is not None:
            ops.cast_pad_bf16(g_logits.float().contiguous().view(M, -1), self.dec.vl, out=dlogits)
        dz = self.dec.backward(dlogits, cx["recon"], cx["feat"], lens, cx["stats"], float(g_mse), B, T, grads=grads)
        dparams = ops.vae_reparam_bwd(cx["params"], cx["eps"], True, dz, lens, z, float(g_kl), self.buf("e.dparams", M, rup(2 * z, 64)))
        self.enc.backward(cx["enc_blocks"], dparams, B, T, grads=grads, first_needs_dx=False)
        return grads
