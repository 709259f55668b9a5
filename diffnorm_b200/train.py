"""Training step of the denoiser: ``LatentDiscreteModel.forward`` (LM:1514-1613, multitask=False:
the latent-noise MSE) and its backward, hand-scheduled over the sm_100a kernels.  torch is memory / streams only.

Per step (one batch of B utterances x T frames, per-utterance timestep t_b):
  forward   frozen VAE encode (inference engine) -> x_t -> time MLP -> gamma/beta of the 56 conditioned modules ->
            WaveNet (GEMM -> un-fused FiLM/gate so u and res are kept) -> 12 transformer layers (norm, QKV GEMM,
            attention with dropout + saved row statistics, out GEMM; norm, GEGLU GEMM -> un-fused GEGLU, conv GEMM,
            out GEMM) -> to_pred -> eps_hat -> weighted masked MSE (+ the decode branch's logging losses).
  backward  the same GEMM kernel over transposed weight packings (data gradients; conv taps read t + shift),
            dn_wgrad (weight gradients straight from the row-major activations), column sums (biases), the
            elementwise backward kernels, the two attention-backward kernels, and the fp32 time-MLP backward.
Gradients are returned as fp32 tensors keyed like ``Model.named_parameters()``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .config import DiffNormConfig
from .ops import GemmPlan
from .packing import (BK, WT, geglu_row_map, pack_conv3, pack_geglu, pack_linear, pack_skip_sum, pack_wavenet_level,
                      pack_wavenet_level_dgrad, rup)
from .repack import PackTable
from .schedule import DDPMScheduler

bf16, f32, i32, i64 = torch.bfloat16, torch.float32, torch.int32, torch.int64


def pack_keep_bits(keep: torch.Tensor) -> torch.Tensor:
    """bool [B,H,T,T] (True = kept) -> int32 words [B,H,T,ceil(T/32)]: bit k%32 of word k/32 = key k."""
    B, H, T, _ = keep.shape
    Tw = (T + 31) // 32
    k = torch.zeros(B, H, T, Tw * 32, dtype=torch.int64)
    k[..., :T] = keep.to(torch.int64).cpu()
    w = (k.view(B, H, T, Tw, 32) << torch.arange(32, dtype=torch.int64)).sum(-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32).contiguous()


class _Plans:
    pass


class ZeroArena:
    """Gradient storage of one training step: views into a few large fp32 buffers that are cleared with ONE memset per
    buffer at the start of backward (a step hands out ~600 accumulators; clearing them one by one was 600 fill launches).
    Views stay valid until the next step's `reset()`."""

    def __init__(self, device, chunk_elems: int = 1 << 26):
        self.dev, self.chunk_elems = device, chunk_elems
        self.chunks, self.used = [], []
        self.cur = 0

    def reset(self):
        for c, u in zip(self.chunks, self.used):
            if u:
                c[:u].zero_()
        self.used = [0] * len(self.chunks)
        self.cur = 0

    def zeros(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        need = (n + 3) & ~3                       # keep every view 16-byte aligned (TMA / float4)
        while True:
            if self.cur == len(self.chunks):
                self.chunks.append(torch.zeros(max(self.chunk_elems, need), dtype=f32, device=self.dev))
                self.used.append(0)
            c, u = self.chunks[self.cur], self.used[self.cur]
            if u + need <= c.numel():
                self.used[self.cur] = u + need
                return c[u:u + n].view(*shape)
            self.cur += 1


class _GradDict(dict):
    """name -> gradient; tells the data-parallel reducer about every gradient the moment it is final."""

    def __init__(self, hook=None):
        super().__init__()
        self._hook = hook

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        if self._hook is not None:
            self._hook(k, v)


def _pack_vae_wavenet(w, p: str, cin: int, cout: int, cin_pad: int, G: int, S: int, last_f32: bool, dev, rec=None) -> _Plans:
    """Forward (un-fused) and data-gradient packings of one VAE WavenetEncoder block (LM:1003-1032): init conv k3
    cin -> cout, S levels x G chains (dilation 2^g, unconditioned), skip sum, final 1x1.  Channel extents are padded to 128.
    `rec` (repack.PackTable) records each packing for the one-launch per-step refresh."""
    cp = rup(cout, 128)
    b = _Plans()
    b.p, b.cin, b.cout, b.cp, b.cin_pad, b.last = p, cin, cout, cp, cin_pad, last_f32
    b.init = pack_conv3(w(p + "init_conv.weight"), w(p + "init_conv.bias"), cin_pad=cin_pad, n_pad=cp, name=p + "init")
    b.init_T = pack_conv3(w(p + "init_conv.weight").permute(1, 0, 2), None, cin_pad=cp, n_pad=cin_pad, shift_sign=-1, name=p + "init^T")
    if rec is not None:
        rec.conv3(w(p + "init_conv.weight"), b.init.W, cin_pad)
        rec.vector(w(p + "init_conv.bias"), b.init.bias)
        rec.conv3(w(p + "init_conv.weight"), b.init_T.W, cp, transposed=True)
    b.lvl, b.lvl_T = [], []
    tiles = cp // 128
    for s_ in range(S):
        blk = [f"{p}stacks.{s_}.blocks.{g}." for g in range(G)]
        convs, ress = [w(k + "conv.weight") for k in blk], [w(k + "res_conv.weight") for k in blk]
        conv_b, res_b = [w(k + "conv.bias") for k in blk], [w(k + "res_conv.bias") for k in blk]
        lv = pack_wavenet_level(convs, conv_b, ress, res_b, cp)
        bi = torch.stack([lv.bias.view(G, tiles, 128), lv.bias2.view(G, tiles, 128)], dim=2).reshape(-1).contiguous()
        b.lvl.append(GemmPlan(lv.W, lv.segs, 2 * cp, tiles, _lib.EPI_BF16, bias=bi, groups=G, g_w_row=lv.g_w_row,
                              g_bias=2 * cp, dilation=1, dilation_shl_group=1, name=f"{p}lvl{s_}.ur"))
        lvT = pack_wavenet_level_dgrad(convs, ress, cp, name=f"{p}lvl{s_}^T")
        b.lvl_T.append(lvT)
        if rec is not None:
            rec.wavenet_level(convs, conv_b, ress, res_b, lv.W, cp, bi=bi)
            rec.wavenet_level_dgrad(convs, ress, lvT.W, cp)
    lastb = [f"{p}stacks.{S - 1}.blocks.{g}." for g in range(G)]
    skips, skip_b = [w(k + "skip_conv.weight") for k in lastb], [w(k + "skip_conv.bias") for k in lastb]
    b.skip = pack_skip_sum(skips, skip_b, cp, name=p + "skip")
    WsT = torch.zeros(G * cp, cp, device=dev)
    for g in range(G):
        WsT[g * cp:g * cp + cout, :cout] = skips[g].reshape(cout, cout).t()
    b.skip_T = pack_linear(WsT, None, k_pad=cp, n_pad=G * cp, name=p + "skip^T")
    b.n_final = rup(cout, 16) if last_f32 else cp
    b.final = pack_linear(w(p + "final_conv.weight"), w(p + "final_conv.bias"), epi=_lib.EPI_F32 if last_f32 else _lib.EPI_BF16,
                          k_pad=cp, n_pad=b.n_final, name=p + "final")
    b.final_T = pack_linear(w(p + "final_conv.weight").reshape(cout, cout).t(), None, k_pad=rup(cout, 64) if last_f32 else cp,
                            n_pad=cp, name=p + "final^T")
    if rec is not None:
        rec.skip_sum(skips, skip_b, b.skip.W, b.skip.bias, cp)
        for g, sk in enumerate(skips):
            rec.linear(sk, b.skip_T.W, transposed=True, row0=g * cp)
        rec.linear(w(p + "final_conv.weight"), b.final.W)
        rec.vector(w(p + "final_conv.bias"), b.final.bias)
        rec.linear(w(p + "final_conv.weight"), b.final_T.W, transposed=True)
    return b


class VaeBlocksTrain:
    """Training-mode execution of VAE WaveNet blocks (encoder or decoder side): un-fused forward keeping the pre-gate
    pairs, data-gradient backward, and (when `grads` is given) the weight / bias gradients of every conv in the block."""

    def __init__(self, buf, G: int, S: int, tag: str, zeros=None):
        self.buf, self.G, self.S, self.tag, self.zeros = buf, G, S, tag, zeros
        self.sv: Dict[object, object] = {}

    def forward(self, blocks, a, B, T, x_out=None):
        buf, G, S, tag = self.buf, self.G, self.S, self.tag
        M = B * T
        sv = self.sv = {}
        for i, b in enumerate(blocks):
            cp = b.cp
            sv[("in", i)] = a
            h = b.init.run(a, buf(f"{tag}.h{i}", M, cp), B, T)
            sv[("h", i)] = h
            src, g_a_col = h, 0
            for s_ in range(S):
                ur = b.lvl[s_].run(src, buf(f"{tag}.ur{i}.{s_}", M, G * 2 * cp), B, T, g_a_col=g_a_col, g_out_col=2 * cp)
                y = ops.wn_gate_fwd(ur, buf(f"{tag}.y{i}.{s_}", M, G * cp), B, T, cp, G)
                sv[("ur", i, s_)], sv[("y", i, s_)] = ur, y
                src, g_a_col = y, cp
            sk = b.skip.run(src, buf(f"{tag}.sk{i}", M, cp), B, T)
            sv[("sk", i)] = sk
            if b.last:
                out = x_out if x_out is not None else buf(f"{tag}.out", M, b.n_final, f32)
            else:
                out = buf(f"{tag}.o{i}", M, cp)
            a = b.final.run(sk, out, B, T)
        return a

    def backward(self, blocks, dcur, B, T, grads=None, first_needs_dx: bool = True):
        """dcur bf16 [M, >= cout of the last block] = gradient of the last block's output; returns the gradient w.r.t. the
        first block's input (bf16 [M, cin_pad]) or None when first_needs_dx is False."""
        buf, G, S, tag, sv = self.buf, self.G, self.S, self.tag, self.sv
        M = B * T
        dev = dcur.device
        zeros = self.zeros or (lambda *shape: torch.zeros(*shape, dtype=f32, device=dev))
        for i in reversed(range(len(blocks))):
            b = blocks[i]
            cp, c, cin, p = b.cp, b.cout, b.cin, b.p
            if grads is not None:
                dWf = zeros(c, rup(cp, 4))
                ops.wgrad(dcur, sv[("sk", i)], dWf, 1, M, c, cp)
                grads[p + "final_conv.weight"] = dWf[:, :c].reshape(c, c, 1)
                grads[p + "final_conv.bias"] = ops.colsum(dcur, 0, rup(c, 8), zeros(rup(c, 8)))[:c]
            dsk = b.final_T.run(dcur, buf(f"{tag}.dsk{i}", M, cp), B, T)
            if grads is not None:
                dWs = zeros(cp, G * cp)
                ops.wgrad(dsk, sv[("y", i, S - 1)], dWs, 1, M, cp, G * cp)
                dbs = ops.colsum(dsk, 0, cp, zeros(cp))[:c]
                for g in range(G):
                    k_ = f"{p}stacks.{S - 1}.blocks.{g}."
                    grads[k_ + "skip_conv.weight"] = dWs[:c, g * cp:g * cp + c].reshape(c, c, 1)
                    grads[k_ + "skip_conv.bias"] = dbs
            dy = b.skip_T.run(dsk, buf(f"{tag}.dyA{i}", M, G * cp), B, T)
            for s_ in reversed(range(S)):
                dur = ops.wn_gate_bwd(sv[("ur", i, s_)], dy, buf(f"{tag}.dur{i}", M, G * 2 * cp), B, T, cp, G)
                if grads is not None:
                    inp = sv[("y", i, s_ - 1)] if s_ > 0 else sv[("h", i)]
                    gx = cp if s_ > 0 else 0
                    tapsG = []
                    for k in range(3):
                        dWg = zeros(G, cp, cp)
                        ops.wgrad(dur, inp, dWg, B, T, cp, cp, 0, 0, 2 - k, groups=G, g_dy_col=2 * cp, g_x_col=gx, shift_shl_group=True)
                        tapsG.append(dWg[:, :c, :c])
                    dWc = torch.stack(tapsG, dim=-1)
                    dWr = zeros(G, cp, cp)
                    ops.wgrad(dur, inp, dWr, B, T, cp, cp, cp, 0, 0, groups=G, g_dy_col=2 * cp, g_x_col=gx)
                    dbg = ops.colsum(dur, 0, G * 2 * cp, zeros(G * 2 * cp)).view(G, 2, cp)
                    for g in range(G):
                        k_ = f"{p}stacks.{s_}.blocks.{g}."
                        grads[k_ + "conv.weight"], grads[k_ + "conv.bias"] = dWc[g], dbg[g, 0, :c]
                        grads[k_ + "res_conv.weight"], grads[k_ + "res_conv.bias"] = dWr[g, :c, :c].reshape(c, c, 1), dbg[g, 1, :c]
                if s_ > 0:
                    dy = b.lvl_T[s_].run(dur, buf(f"{tag}.dyB{i}", M, G * cp), B, T, g_a_col=2 * cp, g_out_col=cp)
                else:
                    d32 = buf(f"{tag}.dh32.{i}", M, cp, f32, zero=True)
                    b.lvl_T[0].run(dur, d32, B, T, g_a_col=2 * cp, g_out_col=0, epi=_lib.EPI_RESID)
                    dh0 = ops.cast_pad_bf16(d32, cp, out=buf(f"{tag}.dh0.{i}", M, cp))
            if grads is not None:
                a_in = sv[("in", i)]
                taps = []
                for k in range(3):
                    dWi = zeros(cp, rup(b.cin_pad, 4))
                    ops.wgrad(dh0, a_in, dWi, B, T, cp, b.cin_pad, 0, 0, 2 - k)
                    taps.append(dWi[:c, :cin])
                grads[p + "init_conv.weight"] = torch.stack(taps, dim=-1)
                grads[p + "init_conv.bias"] = ops.colsum(dh0, 0, cp, zeros(cp))[:c]
            if i > 0 or first_needs_dx:
                dcur = b.init_T.run(dh0, buf(f"{tag}.din{i}", M, b.cin_pad), B, T)
            else:
                dcur = None
        return dcur


class FrozenDecoderTrain:
    """decode_feature (LM:1109-1116) in training form: an un-fused forward that keeps the pre-activations, the
    data-gradient backward down to the latent, and optionally (VAE training) every weight gradient.  Inside a multitask
    diffusion step the VAE is frozen (diff_discrete.py:79-81): packed once, no dropout, no weight gradients."""

    def __init__(self, sd: Optional[Dict[str, torch.Tensor]], cfg: DiffNormConfig, dev, buf, pre: str = "speech_decoder.",
                 zeros=None):
        self.cfg, self.dev, self.buf, self.zeros = cfg, dev, buf, zeros
        c = cfg
        self.zp = rup(c.latent_dim, 64)
        self.D, self.H, self.dh = c.feat_dim, c.vae_heads, c.vae_dim_head
        self.inner = DiffNormConfig.ff_inner(self.D)
        self.ip = rup(self.inner, 128)
        self.vl = rup(c.vocab, 64)                 # dlogits row width (K of the lm-head data gradient)
        self.G, self.S = c.vae_layers, c.vae_stacks
        self.wn = VaeBlocksTrain(buf, self.G, self.S, "v", zeros)
        rm = geglu_row_map(self.inner)
        self.geglu_src, self.geglu_dst = torch.nonzero(rm >= 0).squeeze(1).to(dev), rm[rm >= 0].to(dev)
        self.sv: Dict[object, object] = {}
        if sd is not None:
            self.pack(lambda k: sd[pre + k].detach().to(dev).float())

    def pack(self, w, rec=None):
        c, dev, D, ip = self.cfg, self.dev, self.D, self.ip
        self.blocks = []
        cin_pad = self.zp
        dec_w = c.dec_widths()
        for i, (cin, cout) in enumerate(dec_w):
            b = _pack_vae_wavenet(w, f"decoder_wave.{i}.", cin, cout, cin_pad, self.G, self.S, i == len(dec_w) - 1, dev, rec)
            self.blocks.append(b)
            cin_pad = b.cp

        def lin(name, wk, bk=None, T=False, **kw):
            W = w(wk)
            plan = pack_linear(W.t() if T else W, None if bk is None else w(bk), name=name, **kw)
            if rec is not None:
                rec.linear(W, plan.W, transposed=T)
                if bk is not None:
                    rec.vector(w(bk), plan.bias)
            return plan

        def conv(name, wk, bk=None, T=False, **kw):
            W = w(wk)
            plan = pack_conv3(W.permute(1, 0, 2) if T else W, None if bk is None else w(bk), name=name,
                              shift_sign=-1 if T else 1, **kw)
            if rec is not None:
                rec.conv3(W, plan.W, kw["cin_pad"], transposed=T)
                if bk is not None:
                    rec.vector(w(bk), plan.bias)
            return plan

        def gamma(k):
            g = w(k).float().contiguous().clone()
            if rec is not None:
                rec.vector(w(k), g)
            return g

        self.layers = []
        for l in range(c.vae_depth):
            p = f"decoder_tf.layers.{l}."
            L = _Plans()
            L.p = p
            wq, wkv = w(p + "1.to_q.weight"), w(p + "1.to_kv.weight")
            wqkv = torch.cat([wq, wkv], 0)
            L.qkv, L.qkv_T = pack_linear(wqkv, None, name=p + "qkv"), pack_linear(wqkv.t(), None, name=p + "qkv^T")
            if rec is not None:
                rec.linear(wq, L.qkv.W)
                rec.linear(wkv, L.qkv.W, row0=wq.shape[0])
                rec.linear(wq, L.qkv_T.W, transposed=True)
                rec.linear(wkv, L.qkv_T.W, transposed=True, col0=wq.shape[0])
            L.out = lin(p + "to_out", p + "1.to_out.weight", epi=_lib.EPI_RESID)
            L.out_T = lin(p + "to_out^T", p + "1.to_out.weight", T=True)
            g = pack_geglu(w(p + "5.0.weight"), w(p + "5.0.bias"))
            L.ff1 = GemmPlan(g.W, g.segs, 2 * ip, g.n_tiles, _lib.EPI_BF16, bias=g.bias, name=p + "ff.h")
            L.ff1_T = GemmPlan(g.W.t().contiguous(), [(0, 0, 2 * ip // BK, 0, 0)], D, (D + WT - 1) // WT, _lib.EPI_BF16, name=p + "ff.h^T")
            if rec is not None:
                rec.geglu(w(p + "5.0.weight"), w(p + "5.0.bias"), g.W, g.bias, L.ff1_T.W)
            L.ffc = conv(p + "ff.conv", p + "5.2.1.weight", p + "5.2.1.bias", cin_pad=ip, n_pad=ip)
            L.ffc_T = conv(p + "ff.conv^T", p + "5.2.1.weight", T=True, cin_pad=ip, n_pad=ip)
            L.ff3 = lin(p + "ff.out", p + "5.3.weight", p + "5.3.bias", epi=_lib.EPI_RESID, k_pad=ip)
            L.ff3_T = lin(p + "ff.out^T", p + "5.3.weight", T=True, k_pad=D, n_pad=ip)
            L.g1, L.g2 = gamma(p + "0.gamma"), gamma(p + "4.gamma")
            self.layers.append(L)
        self.pred_gamma = gamma("decoder_tf.to_pred.0.gamma")
        self.pred = lin("vae.to_pred", "decoder_tf.to_pred.1.weight", epi=_lib.EPI_F32)
        self.pred_T = lin("vae.to_pred^T", "decoder_tf.to_pred.1.weight", T=True)
        self.lm = lin("vae.lm", "decoder_lm.weight", "decoder_lm.bias", epi=_lib.EPI_F32, n_pad=rup(c.vocab, 16))
        self.lm_T = lin("vae.lm^T", "decoder_lm.weight", T=True, epi=_lib.EPI_F32, k_pad=self.vl, n_pad=D)

    def forward(self, xb, lens, B, T, keep_bits=None, keep_scale: float = 1.0):
        """xb bf16 [B*T, zp] -> (recon fp32 [B*T, 768], logits fp32 [B*T, vp]); keeps activations for backward()."""
        c, buf = self.cfg, self.buf
        M, D, H, dh, ip = B * T, self.D, self.H, self.dh, self.ip
        sv = self.sv = {}
        x = self.wn.forward(self.blocks, xb, B, T, x_out=buf("v.x", M, D, f32))
        for l, L in enumerate(self.layers):
            xs1 = buf(f"v.xs1.{l}", M, D, f32)
            xs1.copy_(x)
            hb1 = ops.adarmsnorm(x, buf(f"v.hb1.{l}", M, D), B, T, L.g1)
            qkv = L.qkv.run(hb1, buf(f"v.qkv.{l}", M, 3 * H * dh), B, T)
            lse = buf(f"v.lse.{l}", B * H, T, f32)
            kb = None if keep_bits is None else keep_bits[l]
            ao = ops.attention_train(qkv, buf(f"v.ao.{l}", M, H * dh), lse, lens, kb, keep_scale, B, T, H, dh)
            L.out.run(ao, x, B, T)
            xs2 = buf(f"v.xs2.{l}", M, D, f32)
            xs2.copy_(x)
            hb2 = ops.adarmsnorm(x, buf(f"v.hb2.{l}", M, D), B, T, L.g2)
            hh = L.ff1.run(hb2, buf(f"v.h.{l}", M, 2 * ip), B, T)
            m1 = ops.geglu_fwd(hh, buf(f"v.m1.{l}", M, ip))
            m2 = L.ffc.run(m1, buf(f"v.m2.{l}", M, ip), B, T)
            L.ff3.run(m2, x, B, T)
            sv[l] = (xs1, hb1, qkv, lse, ao, xs2, hb2, hh, m1, m2, kb)
        hbf = ops.adarmsnorm(x, buf("v.hbf", M, D), B, T, self.pred_gamma)
        recon = self.pred.run(hbf, buf("v.recon", M, D, f32), B, T)
        rb = ops.cast_pad_bf16(recon, D, out=buf("v.rb", M, D))
        logits = self.lm.run(rb, buf("v.logits", M, rup(c.vocab, 16), f32), B, T)
        sv["x"], sv["hbf"], sv["rb"], sv["keep_scale"] = x, hbf, rb, keep_scale
        return recon, logits

    def backward(self, dlogits, recon, audio, lens, stats, mse_scale: float, B, T, grads=None):
        """dlogits bf16 [B*T, vl] (+ the masked-MSE term built here) -> d latent bf16 [B*T, zp].
        grads (dict) given => also every weight / bias / gamma gradient of the decoder (VAE training)."""
        buf, sv, c = self.buf, self.sv, self.cfg
        M, D, H, dh, ip, inner = B * T, self.D, self.H, self.dh, self.ip, self.inner
        dev = dlogits.device
        zeros = self.zeros or (lambda *shape: torch.zeros(*shape, dtype=f32, device=dev))
        wg = None
        if grads is not None:
            def wg(dY, X, n_rows, k_cols, shift=0):
                dW = zeros(n_rows, rup(k_cols, 4))
                if shift == 0:
                    ops.wgrad(dY, X, dW, 1, M, n_rows, k_cols)
                else:
                    ops.wgrad(dY, X, dW, B, T, n_rows, k_cols, 0, 0, shift)
                return dW
            cs = lambda src, cols: ops.colsum(src, 0, cols, zeros(cols))
            grads["decoder_lm.weight"] = wg(dlogits, sv["rb"], c.vocab, D)[:, :D]
            grads["decoder_lm.bias"] = cs(dlogits, rup(c.vocab, 8))[:c.vocab]
        d_lm = self.lm_T.run(dlogits, buf("v.dlm", M, D, f32), B, T)
        drec = ops.recon_grad(recon, audio, d_lm, lens, B, T, stats, mse_scale, buf("v.drec", M, D))
        if grads is not None:
            grads["decoder_tf.to_pred.1.weight"] = wg(drec, sv["hbf"], D, D)
        dhb = self.pred_T.run(drec, buf("v.dhb", M, D), B, T)
        dx = buf("v.dx", M, D, f32, zero=True)
        dxb = buf("v.dxb", M, D)
        dg = zeros(D) if grads is not None else None
        ops.adarmsnorm_bwd(sv["x"], dhb, dx, dxb, B, T, gamma_p=self.pred_gamma, dgamma_p=dg)
        if grads is not None:
            grads["decoder_tf.to_pred.0.gamma"] = dg
        for l in reversed(range(len(self.layers))):
            L = self.layers[l]
            p = L.p
            xs1, hb1, qkv, lse, ao, xs2, hb2, hh, m1, m2, kb = sv[l]
            if grads is not None:
                grads[p + "5.3.weight"] = wg(dxb, m2, D, ip)[:, :inner]
                grads[p + "5.3.bias"] = cs(dxb, D)
            dm2 = L.ff3_T.run(dxb, buf("v.dm2", M, ip), B, T)
            if grads is not None:
                grads[p + "5.2.1.weight"] = torch.stack([wg(dm2, m1, ip, ip, shift=2 - k)[:inner, :inner] for k in range(3)], dim=-1)
                grads[p + "5.2.1.bias"] = cs(dm2, ip)[:inner]
            dm1 = L.ffc_T.run(dm2, buf("v.dm1", M, ip), B, T)
            dh_ = ops.geglu_bwd(hh, dm1, buf("v.dh", M, 2 * ip))
            if grads is not None:
                dW1p, db1p = wg(dh_, hb2, 2 * ip, D), cs(dh_, 2 * ip)
                grads[p + "5.0.weight"] = zeros(2 * inner, D).index_copy_(0, self.geglu_dst, dW1p.index_select(0, self.geglu_src))
                grads[p + "5.0.bias"] = zeros(2 * inner).index_copy_(0, self.geglu_dst, db1p.index_select(0, self.geglu_src))
            dhb = L.ff1_T.run(dh_, buf("v.dhb", M, D), B, T)
            dg = zeros(D) if grads is not None else None
            ops.adarmsnorm_bwd(xs2, dhb, dx, dxb, B, T, gamma_p=L.g2, dgamma_p=dg)
            if grads is not None:
                grads[p + "4.gamma"] = dg
                grads[p + "1.to_out.weight"] = wg(dxb, ao, D, H * dh)
            dao = L.out_T.run(dxb, buf("v.dao", M, H * dh), B, T)
            dqkv = ops.attention_bwd(qkv, ao, dao, lse, lens, kb, sv["keep_scale"], buf("v.dqkv", M, 3 * H * dh),
                                     buf("v.delta", B * H, T, f32), B, T, H, dh)
            if grads is not None:
                dWqkv = wg(dqkv, hb1, 3 * H * dh, D)
                grads[p + "1.to_q.weight"], grads[p + "1.to_kv.weight"] = dWqkv[:H * dh], dWqkv[H * dh:]
            dhb = L.qkv_T.run(dqkv, buf("v.dhb", M, D), B, T)
            dg = zeros(D) if grads is not None else None
            ops.adarmsnorm_bwd(xs1, dhb, dx, dxb, B, T, gamma_p=L.g1, dgamma_p=dg)
            if grads is not None:
                grads[p + "0.gamma"] = dg
        return self.wn.backward(self.blocks, dxb, B, T, grads=grads)


class DenoiserTrainer:
    def __init__(self, ldm, drop_p: float = 0.1, seed: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("DenoiserTrainer needs a CUDA device: the product has no CPU path")
        self.multitask = bool(getattr(ldm, "multitask", False))   # LM:1600-1604: + (50 mse + nll) / T through decode_feature
        self.ldm = ldm
        self.cfg: DiffNormConfig = ldm.cfg
        self.P: Dict[str, torch.nn.Parameter] = dict(ldm.model.named_parameters())
        self.dev = next(iter(self.P.values())).device
        self.drop_p, self.seed, self.step_no = drop_p, seed, 0
        from .engine import DiffNormEngine
        sd = {k: v for k, v in ldm.state_dict().items()}
        # frozen VAE encoder of the training step: bf16 operands (the latent is about to be noised at sigma >= 0.016; the
        # split-precision encode of the normalization pass costs 3x here for nothing — 35.0 vs 29.5 ms per step)
        self.vae = DiffNormEngine(sd, device=str(self.dev), cfg=self.cfg, vae_only=True, vae_fmt="bf16")
        c = self.cfg
        s = DDPMScheduler(c.timesteps)
        self.sched = s
        coef = np.zeros((c.timesteps, 4), dtype=np.float32)
        sa, s1 = s.sqrt_alphas_cumprod.astype(np.float32), s.sqrt_one_minus_alphas_cumprod.astype(np.float32)
        snr = (sa ** 2) / (s1 ** 2)                       # LM:1283-1286, fp32 like the reference
        coef[:, 0], coef[:, 1], coef[:, 2] = sa, s1, np.minimum(snr, np.float32(5.0)) / snr
        self.coef = torch.from_numpy(coef).to(self.dev)
        self.beta0 = float(np.float32(s.betas[0]))
        self.ws: Dict[tuple, torch.Tensor] = {}
        self.cond_names = self.cond_layer_names(c)
        self.n_cond, self.gbw = len(self.cond_names), 2 * c.hid
        self.inner = DiffNormConfig.ff_inner(c.hid)
        self.ip = rup(self.inner, 128)
        rm = geglu_row_map(self.inner)
        self.geglu_src = torch.nonzero(rm >= 0).squeeze(1).to(self.dev)   # packed GEGLU rows that hold a real weight row
        self.geglu_dst = rm[rm >= 0].to(self.dev)                         # ... and the reference row each one maps to
        self.zp, self.zn = rup(c.latent_dim, 64), rup(c.latent_dim, 16)
        self._pe: Dict[int, torch.Tensor] = {}
        self._pack_graph = None
        self._pack_table = None
        self._pack_plans = None
        self._pack_ptrs = None
        self.arena = ZeroArena(self.dev)
        self.dec = FrozenDecoderTrain(sd, self.cfg, self.dev, self.buf) if self.multitask else None

    # ------------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def cond_layer_names(c: DiffNormConfig) -> List[str]:
        """The time-conditioned layers in table order: WaveNet blocks (LM:500-507), then per transformer layer the two
        adaptive norms (LM:666-674)."""
        names = [f"wavenet.stacks.{st}.blocks.{i}.to_time_cond" for st in range(c.wn_stacks) for i in range(c.wn_layers)]
        for l in range(c.depth):
            names += [f"transformer.layers.{l}.0.to_gamma_beta", f"transformer.layers.{l}.4.to_gamma_beta"]
        return names

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

    def _run(self, plan: GemmPlan, A, out, B, T, **kw):
        return plan.run(A, out, B, T, **kw)

    # ------------------------------------------------------------------------------------------------ packing (per step)
    def _pack(self, rec=None) -> _Plans:
        """Packs every weight with diffnorm_b200.packing.  `rec` (a repack.PackTable) additionally records, next to each
        pack call, which parameter feeds which packed tensor, so later steps refresh the same tensors with one launch."""
        c, P, ip, dev = self.cfg, self.P, self.ip, self.dev
        C = c.hid
        pl = _Plans()
        w = lambda k: P[k].detach()

        def lin(name, wk, bk=None, T=False, **kw):
            """pack_linear of parameter wk (transposed if T) + its record."""
            W = w(wk)
            plan = pack_linear(W.reshape(W.shape[0], -1).t() if T else W, None if bk is None else w(bk), name=name, **kw)
            if rec is not None:
                rec.linear(W, plan.W, transposed=T)
                if bk is not None:
                    rec.vector(w(bk), plan.bias)
            return plan

        def conv(name, wk, bk=None, T=False, **kw):
            W = w(wk)
            plan = pack_conv3(W.permute(1, 0, 2) if T else W, None if bk is None else w(bk), name=name,
                              shift_sign=-1 if T else 1, **kw)
            if rec is not None:
                rec.conv3(W, plan.W, kw["cin_pad"], transposed=T)
                if bk is not None:
                    rec.vector(w(bk), plan.bias)
            return plan

        pl.init = lin("init_conv", "init_conv.weight", "init_conv.bias", k_pad=self.zp)
        pl.wn_init = conv("wn.init", "wavenet.init_conv.weight", "wavenet.init_conv.bias", cin_pad=C, n_pad=C)
        pl.wn_init_T = conv("wn.init^T", "wavenet.init_conv.weight", T=True, cin_pad=C, n_pad=C)
        pl.lvl, pl.lvl_T = [], []
        G = c.wn_layers
        for s in range(c.wn_stacks):
            blk = [f"wavenet.stacks.{s}.blocks.{i}." for i in range(G)]
            convs, ress = [w(b + "conv.weight") for b in blk], [w(b + "res_conv.weight") for b in blk]
            conv_b, res_b = [w(b + "conv.bias") for b in blk], [w(b + "res_conv.bias") for b in blk]
            lv = pack_wavenet_level(convs, conv_b, ress, res_b, C, name=f"wn.lvl{s}")
            # un-fused form: plain +bias epilogue over the same packed tiles -> [conv 128 | res 128] column blocks
            tiles = C // 128
            bi = torch.stack([lv.bias.view(G, tiles, 128), lv.bias2.view(G, tiles, 128)], dim=2).reshape(-1).contiguous()
            fwd = GemmPlan(lv.W, lv.segs, 2 * C, tiles, _lib.EPI_BF16, bias=bi, groups=G, g_w_row=lv.g_w_row, g_bias=2 * C,
                           dilation=1, dilation_shl_group=1, name=f"wn.lvl{s}.ur")
            pl.lvl.append(fwd)
            lvT = pack_wavenet_level_dgrad(convs, ress, C, name=f"wn.lvl{s}^T")
            pl.lvl_T.append(lvT)
            if rec is not None:
                rec.wavenet_level(convs, conv_b, ress, res_b, lv.W, C, bi=bi)
                rec.wavenet_level_dgrad(convs, ress, lvT.W, C)
        last = [f"wavenet.stacks.{c.wn_stacks - 1}.blocks.{i}." for i in range(G)]
        skips, skip_b = [w(b + "skip_conv.weight") for b in last], [w(b + "skip_conv.bias") for b in last]
        pl.skip = pack_skip_sum(skips, skip_b, C, name="wn.skip")
        pl.skip_T = pack_linear(torch.cat([sk.reshape(C, C).t() for sk in skips], 0), None, k_pad=C, n_pad=G * C, name="wn.skip^T")
        if rec is not None:
            rec.skip_sum(skips, skip_b, pl.skip.W, pl.skip.bias, C)
            for g, sk in enumerate(skips):
                rec.linear(sk, pl.skip_T.W, transposed=True, row0=g * C)
        pl.wn_final = lin("wn.final", "wavenet.final_conv.weight", "wavenet.final_conv.bias", epi=_lib.EPI_F32, k_pad=C, n_pad=C)
        pl.wn_final_T = lin("wn.final^T", "wavenet.final_conv.weight", T=True, k_pad=C, n_pad=C)
        pl.layers = []
        for l in range(c.depth):
            p = f"transformer.layers.{l}."
            L = _Plans()
            wq, wkv = w(p + "1.to_q.weight"), w(p + "1.to_kv.weight")
            wqkv = torch.cat([wq, wkv], 0)
            L.qkv = pack_linear(wqkv, None, name=p + "qkv")
            L.qkv_T = pack_linear(wqkv.t(), None, name=p + "qkv^T")
            if rec is not None:
                rec.linear(wq, L.qkv.W)
                rec.linear(wkv, L.qkv.W, row0=wq.shape[0])
                rec.linear(wq, L.qkv_T.W, transposed=True)
                rec.linear(wkv, L.qkv_T.W, transposed=True, col0=wq.shape[0])
            L.out = lin(p + "to_out", p + "1.to_out.weight", epi=_lib.EPI_RESID)
            L.out_T = lin(p + "to_out^T", p + "1.to_out.weight", T=True)
            g = pack_geglu(w(p + "5.0.weight"), w(p + "5.0.bias"), name=p + "ff.geglu")
            L.ff1 = GemmPlan(g.W, g.segs, 2 * ip, g.n_tiles, _lib.EPI_BF16, bias=g.bias, name=p + "ff.h")
            L.ff1_T = GemmPlan(g.W.t().contiguous(), [(0, 0, 2 * ip // BK, 0, 0)], C, (C + WT - 1) // WT, _lib.EPI_BF16,
                               name=p + "ff.h^T")
            if rec is not None:
                rec.geglu(w(p + "5.0.weight"), w(p + "5.0.bias"), g.W, g.bias, L.ff1_T.W)
            L.ffc = conv(p + "ff.conv", p + "5.2.1.weight", p + "5.2.1.bias", cin_pad=ip, n_pad=ip)
            L.ffc_T = conv(p + "ff.conv^T", p + "5.2.1.weight", T=True, cin_pad=ip, n_pad=ip)
            L.ff3 = lin(p + "ff.out", p + "5.3.weight", p + "5.3.bias", epi=_lib.EPI_RESID, k_pad=ip)
            L.ff3_T = lin(p + "ff.out^T", p + "5.3.weight", T=True, k_pad=C, n_pad=ip)
            pl.layers.append(L)
        pl.pred = lin("to_pred", "transformer.to_pred.1.weight")
        pl.pred_T = lin("to_pred^T", "transformer.to_pred.1.weight", T=True)
        pl.proj = lin("final_proj", "final_proj.weight", "final_proj.bias", epi=_lib.EPI_F32, n_pad=self.zn)
        pl.proj_T = lin("final_proj^T", "final_proj.weight", T=True, k_pad=self.zp, n_pad=C)
        pl.Wcat = torch.cat([w(n + ".weight") for n in self.cond_names], 0).contiguous()       # [56 * 1024, 2048]
        pl.bcat = torch.cat([w(n + ".bias") for n in self.cond_names], 0).contiguous()
        if rec is not None:
            r0 = 0
            for n in self.cond_names:
                Wn, bn = w(n + ".weight"), w(n + ".bias")
                rec.linear(Wn, pl.Wcat, row0=r0)
                rec.vector(bn, pl.bcat, row0=r0)
                r0 += Wn.shape[0]
        return pl

    def _packed(self) -> _Plans:
        """Weights are packed with diffnorm_b200.packing once (which also zero-fills the padding); every later step
        refreshes the same packed tensors in place from the fp32 masters (updated in place by the optimizer) with ONE
        dn_pack_weights launch over the recorded descriptor table (diffnorm_b200.repack).  A change of any parameter's
        storage (e.g. `.to()`, a new tensor assigned) packs and records again.  DN_REPACK=torch keeps the previous form:
        the ~500 torch indexing kernels captured in a CUDA graph and replayed."""
        ptrs = tuple(p.data_ptr() for p in self.P.values())
        if os.environ.get("DN_REPACK", "kernel") == "torch":
            if self._pack_graph is None or ptrs != self._pack_ptrs:
                self._pack()                      # warm-up outside capture (allocator, lazy init)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._pack_plans = self._pack()
                self._pack_graph, self._pack_ptrs = g, ptrs
            self._pack_graph.replay()
            return self._pack_plans
        if self._pack_table is None or ptrs != self._pack_ptrs:
            rec = PackTable()
            self._pack_plans = self._pack(rec)
            self._pack_table, self._pack_ptrs = rec.finalize(self.dev), ptrs
            return self._pack_plans               # fresh from packing.*: nothing to refresh this step
        self._pack_table.run()
        return self._pack_plans

    def pe_table(self, T: int) -> torch.Tensor:
        t = self._pe.get(T)
        if t is None:
            half = self.cfg.hid // 2
            e = np.log(10000.0) / (half - 1)
            freq = torch.exp(torch.arange(half, dtype=torch.float) * -e)
            ang = torch.arange(T + 1, dtype=torch.float)[:, None] * freq[None, :]
            tab = torch.cat([ang.sin(), ang.cos()], dim=1)
            tab[0] = 0
            t = tab.to(self.dev).contiguous()
            self._pe[T] = t
        return t

    # ------------------------------------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, audio: torch.Tensor, units: Optional[torch.Tensor], lengths: torch.Tensor,
             times: Optional[torch.Tensor] = None, noise: Optional[Dict[str, torch.Tensor]] = None,
             keep_bits: Optional[Sequence[torch.Tensor]] = None, backward: bool = True, decode_losses: bool = True,
             grad_scale: float = 1.0, grad_hook=None):
        """audio fp32 [B,T,768] cuda; units int64 [B,T] (0 = pad, unit k -> k+4) or None; lengths int32 [B].
        times int [B] in [1, timesteps) (drawn if None, LM:1528); noise = {"vae": [B,z,T], "eps0": [B,T,z], "eps": [B,T,z]}
        replays the draws; keep_bits = per layer int32 [B,H,T,ceil(T/32)] (drawn with Philox if None and drop_p > 0).
        grad_hook(name, tensor) is called as each gradient becomes final (diffnorm_b200.dist.GradAllReducer.hook).
        Returns (loss dict of 0-d tensors, grads dict name -> fp32 tensor) ; grads = {} when backward=False."""
        c, dev = self.cfg, self.dev
        B, T, _ = audio.shape
        M, z, C, G, S = B * T, c.latent_dim, c.hid, c.wn_layers, c.wn_stacks
        H, dh, ip, zp, zn = c.heads, c.dim_head, self.ip, self.zp, self.zn
        noise = noise or {}
        lens = lengths.to(device=dev, dtype=i32).contiguous()
        t_idx = (torch.randint(1, c.timesteps, (B,), device=dev) if times is None else times.to(dev)).to(i32).contiguous()
        eps_vae = noise.get("vae")
        eps_vae = torch.randn(B, z, T, device=dev) if eps_vae is None else eps_vae.to(dev).float().contiguous()
        eps0 = noise.get("eps0")
        eps0 = torch.randn(B, T, z, device=dev) if eps0 is None else eps0.to(dev).float().contiguous()
        eps = noise.get("eps")
        eps = torch.randn(B, T, z, device=dev) if eps is None else eps.to(dev).float().contiguous()
        drop = self.drop_p > 0 or keep_bits is not None
        keep_scale = 1.0 / (1.0 - self.drop_p) if drop else 1.0
        Tw = (T + 31) // 32
        if drop and keep_bits is None:
            keep_bits = []
            for l in range(c.depth):
                kb = torch.empty(B, H, T, Tw, dtype=i32, device=dev)
                ops.dropout_bits(kb, self.drop_p, self.seed, self.step_no * c.depth + l)
                keep_bits.append(kb)
        self.step_no += 1
        pl = self._packed()
        rows = torch.arange(B, dtype=i32, device=dev)
        gstride = self.n_cond * self.gbw

        # ---- frozen VAE encode + noising (LM:1521-1535)
        zlat = self.vae.encode(audio.float().contiguous(), eps_vae)
        x_t = self.buf("x_t", M, z, f32)
        xb = self.buf("xb", M, zp)
        ops.train_noise(zlat.contiguous(), eps0, eps, self.beta0, self.coef, t_idx, x_t, xb)

        # ---- time conditioning (LM:104-116, :741-745, :507, :624)
        P = self.P
        tw = P["to_time_cond.0.weights"].detach().float().contiguous()
        W1, b1 = P["to_time_cond.1.weight"].detach().contiguous(), P["to_time_cond.1.bias"].detach().contiguous()
        feats = ops.time_features(t_idx, tw)
        pre = ops.linear_f32(feats, W1, b1, act=0)
        temb = ops.silu(pre, torch.empty_like(pre))
        gball = ops.linear_f32(temb, pl.Wcat, pl.bcat)                 # [B, 56 * 1024]
        gflat = gball.view(-1)
        gbk = dict(gb_t_stride=gstride, t_idx=rows, t_idx_stride=1)

        # ---- denoiser forward (LM:828-876), keeping what backward needs
        sv = {}
        h0 = self._run(pl.init, xb, self.buf("h0", M, C), B, T)
        hw = self._run(pl.wn_init, h0, self.buf("hw", M, C), B, T)
        src, g_a_col = hw, 0
        for s in range(S):
            ur = self._run(pl.lvl[s], src, self.buf(f"ur{s}", M, G * 2 * C), B, T, g_a_col=g_a_col, g_out_col=2 * C)
            y = ops.wn_gate_fwd(ur, self.buf(f"y{s}", M, G * C), B, T, C, G, gb=gflat[s * G * self.gbw:], g_gb=self.gbw, **gbk)
            sv[f"ur{s}"], sv[f"y{s}"] = ur, y
            src, g_a_col = y, C
        sk = self._run(pl.skip, src, self.buf("sk", M, C), B, T)
        x = self._run(pl.wn_final, sk, self.buf("x", M, C, f32), B, T, pe=self.pe_table(T), lengths=lens)
        n0 = S * G
        for l, L in enumerate(pl.layers):
            xs1 = self.buf(f"xs1.{l}", M, C, f32)
            xs1.copy_(x)
            hb1 = ops.adarmsnorm(x, self.buf(f"hb1.{l}", M, C), B, T, None, gflat[(n0 + 2 * l) * self.gbw:], gstride, rows, 1)
            qkv = self._run(L.qkv, hb1, self.buf(f"qkv.{l}", M, 3 * H * dh), B, T)
            lse = self.buf(f"lse.{l}", B * H, T, f32)
            ao = ops.attention_train(qkv, self.buf(f"ao.{l}", M, H * dh), lse, lens, keep_bits[l] if drop else None, keep_scale,
                                     B, T, H, dh)
            self._run(L.out, ao, x, B, T)
            xs2 = self.buf(f"xs2.{l}", M, C, f32)
            xs2.copy_(x)
            hb2 = ops.adarmsnorm(x, self.buf(f"hb2.{l}", M, C), B, T, None, gflat[(n0 + 2 * l + 1) * self.gbw:], gstride, rows, 1)
            hh = self._run(L.ff1, hb2, self.buf(f"h.{l}", M, 2 * ip), B, T)
            m1 = ops.geglu_fwd(hh, self.buf(f"m1.{l}", M, ip))
            m2 = self._run(L.ffc, m1, self.buf(f"m2.{l}", M, ip), B, T)
            self._run(L.ff3, m2, x, B, T)
            sv[l] = (xs1, hb1, qkv, lse, ao, xs2, hb2, hh, m1, m2)
        gpred = P["transformer.to_pred.0.gamma"].detach().float().contiguous()
        hbf = ops.adarmsnorm(x, self.buf("hbf", M, C), B, T, gpred)
        pb = self._run(pl.pred, hbf, self.buf("pb", M, C), B, T)
        eh = self._run(pl.proj, pb, self.buf("eh", M, zn, f32), B, T)

        # ---- losses (LM:1563-1611)
        loss = torch.zeros(1, dtype=f32, device=dev)
        dpred = self.buf("dpred", M, zp) if backward else None
        out = {}
        dx1 = None
        need_decode = (decode_losses or self.multitask) and units is not None
        if self.multitask and units is None:
            raise ValueError("multitask training needs the target units (LM:1583-1597)")
        if need_decode:
            xb1 = ops.pred_x1(x_t, eh, self.coef, t_idx, B, T, z, self.buf("xb1", M, zp))
            audio_f = audio.float().contiguous().view(M, -1)
            units_f = units.to(dev).to(i64).contiguous().view(-1)
            if self.multitask:
                recon, logits = self.dec.forward(xb1, lens, B, T)
            else:
                recon, logits = self.vae.decode(xb1, lens, B, T)
                recon, logits = recon.contiguous().view(M, -1), logits.contiguous().view(M, -1)
            st = ops.decode_losses(recon, audio_f, logits, c.vocab, units_f, lens, B, T)
            e_i = 0.1 / (c.vocab - 1)
            ntok = st[4].clamp(min=1)
            out["recon_mse_loss"] = (st[0] / (st[5].clamp(min=1) * c.feat_dim)).float()
            out["nll_loss"] = (((1.0 - 0.1 - e_i) * st[1] + e_i * st[2]) / ntok).float()
            out["acc"] = (st[3] / ntok).float()
            if self.multitask and backward:
                dlogits = ops.lsnll_bwd(logits, c.vocab, units_f, st, 0.1, grad_scale / c.timesteps, self.buf("v.dlogits", M, self.dec.vl))
                dx1 = self.dec.backward(dlogits, recon, audio_f, lens, st, 50.0 * grad_scale / c.timesteps, B, T)
        ops.noise_loss(eh, eps, lens, self.coef, t_idx, B, T, z, loss, dpred, grad_scale, dx1)
        out["noise_loss"] = loss[0]
        out["total_loss"] = loss[0]
        if self.multitask and need_decode:
            out["total_loss"] = loss[0] + (50.0 * out["recon_mse_loss"] + out["nll_loss"]) / c.timesteps
        out["pred_noise"] = eh.view(B, T, zn)[..., :z]
        if not backward:
            return out, {}

        # ================================================================================================ backward
        grads: Dict[str, torch.Tensor] = _GradDict(grad_hook)
        self.arena.reset()
        zeros = self.arena.zeros          # gradient tensors are views into the arena: valid until the next step

        def wg(dY, X, n_rows, k_cols, dy_col0=0, x_col0=0, shift=0, flat=True):
            dW = zeros(n_rows, rup(k_cols, 4))
            if flat and shift == 0:
                ops.wgrad(dY, X, dW, 1, M, n_rows, k_cols, dy_col0, x_col0, 0)
            else:
                ops.wgrad(dY, X, dW, B, T, n_rows, k_cols, dy_col0, x_col0, shift)
            return dW

        def cs(src, col0, cols):
            return ops.colsum(src, col0, cols, zeros(cols))

        dgball = zeros(B, gstride)
        # final_proj (Linear 512 -> z) and to_pred
        grads["final_proj.weight"] = wg(dpred, pb, z, C)[:, :C]
        grads["final_proj.bias"] = cs(dpred, 0, zn)[:z]
        dpb = self._run(pl.proj_T, dpred, self.buf("dpb", M, C), B, T)
        grads["transformer.to_pred.1.weight"] = wg(dpb, hbf, C, C)
        dhbf = self._run(pl.pred_T, dpb, self.buf("dhb", M, C), B, T)
        dx = self.buf("dx", M, C, f32, zero=True)
        dxb = self.buf("dxb", M, C)
        dgam = zeros(C)
        ops.adarmsnorm_bwd(x, dhbf, dx, dxb, B, T, gamma_p=gpred, dgamma_p=dgam)
        grads["transformer.to_pred.0.gamma"] = dgam
        for l in reversed(range(c.depth)):
            L = pl.layers[l]
            p = f"transformer.layers.{l}."
            xs1, hb1, qkv, lse, ao, xs2, hb2, hh, m1, m2 = sv[l]
            # feed-forward branch
            grads[p + "5.3.weight"] = wg(dxb, m2, C, ip)[:, :self.inner]
            grads[p + "5.3.bias"] = cs(dxb, 0, C)
            dm2 = self._run(L.ff3_T, dxb, self.buf("dm2", M, ip), B, T)
            taps = [wg(dm2, m1, ip, ip, shift=2 - k, flat=False)[:self.inner, :self.inner] for k in range(3)]
            grads[p + "5.2.1.weight"] = torch.stack(taps, dim=-1)
            grads[p + "5.2.1.bias"] = cs(dm2, 0, ip)[:self.inner]
            dm1 = self._run(L.ffc_T, dm2, self.buf("dm1", M, ip), B, T)
            dh_ = ops.geglu_bwd(hh, dm1, self.buf("dh", M, 2 * ip))
            dW1p = wg(dh_, hb2, 2 * ip, C)
            db1p = cs(dh_, 0, 2 * ip)
            gw = zeros(2 * self.inner, C).index_copy_(0, self.geglu_dst, dW1p.index_select(0, self.geglu_src))
            gb_ = zeros(2 * self.inner).index_copy_(0, self.geglu_dst, db1p.index_select(0, self.geglu_src))
            grads[p + "5.0.weight"], grads[p + "5.0.bias"] = gw, gb_
            dhb = self._run(L.ff1_T, dh_, self.buf("dhb", M, C), B, T)
            ops.adarmsnorm_bwd(xs2, dhb, dx, dxb, B, T, gb=gflat[(n0 + 2 * l + 1) * self.gbw:], dgb=dgball.view(-1)[(n0 + 2 * l + 1) * self.gbw:],
                               dgb_b_stride=gstride, **gbk)
            # attention branch
            grads[p + "1.to_out.weight"] = wg(dxb, ao, C, H * dh)
            dao = self._run(L.out_T, dxb, self.buf("dao", M, H * dh), B, T)
            dqkv = ops.attention_bwd(qkv, ao, dao, lse, lens, keep_bits[l] if drop else None, keep_scale,
                                     self.buf("dqkv", M, 3 * H * dh), self.buf("delta", B * H, T, f32), B, T, H, dh)
            dWqkv = wg(dqkv, hb1, 3 * H * dh, C)
            grads[p + "1.to_q.weight"], grads[p + "1.to_kv.weight"] = dWqkv[:H * dh], dWqkv[H * dh:]
            dhb = self._run(L.qkv_T, dqkv, self.buf("dhb", M, C), B, T)
            ops.adarmsnorm_bwd(xs1, dhb, dx, dxb, B, T, gb=gflat[(n0 + 2 * l) * self.gbw:], dgb=dgball.view(-1)[(n0 + 2 * l) * self.gbw:],
                               dgb_b_stride=gstride, **gbk)
        # WaveNet: final 1x1 conv (+PE), skip sum, 4 levels x 8 chains, init conv
        grads["wavenet.final_conv.weight"] = wg(dxb, sk, C, C).view(C, C, 1)
        grads["wavenet.final_conv.bias"] = cs(dxb, 0, C)
        dsk = self._run(pl.wn_final_T, dxb, self.buf("dsk", M, C), B, T)
        dWs = wg(dsk, sv[f"y{S - 1}"], C, G * C)
        dbs = cs(dsk, 0, C)
        for g in range(G):
            b_ = f"wavenet.stacks.{S - 1}.blocks.{g}."
            grads[b_ + "skip_conv.weight"] = dWs[:, g * C:(g + 1) * C].reshape(C, C, 1)
            grads[b_ + "skip_conv.bias"] = dbs
        dy = self._run(pl.skip_T, dsk, self.buf("dyA", M, G * C), B, T)
        dy_next = "dyB"
        for s in reversed(range(S)):
            dur = ops.wn_gate_bwd(sv[f"ur{s}"], dy, self.buf("dur", M, G * 2 * C), B, T, C, G, gb=gflat[s * G * self.gbw:],
                                  g_gb=self.gbw, dgb=dgball.view(-1)[s * G * self.gbw:], dgb_b_stride=gstride, g_dgb=self.gbw, **gbk)
            inp = sv[f"y{s - 1}"] if s > 0 else hw
            gx = C if s > 0 else 0
            # all 8 chains per launch: 3 conv taps (shift (2-k) 2^g) + the 1x1 res conv; biases from one column sum
            tapsG = []
            for k in range(3):
                dWg = zeros(G, C, C)
                ops.wgrad(dur, inp, dWg, B, T, C, C, 0, 0, 2 - k, groups=G, g_dy_col=2 * C, g_x_col=gx, shift_shl_group=True)
                tapsG.append(dWg)
            dWc = torch.stack(tapsG, dim=-1)                                     # [G, C, C, 3]
            dWr = zeros(G, C, C)
            ops.wgrad(dur, inp, dWr, B, T, C, C, C, 0, 0, groups=G, g_dy_col=2 * C, g_x_col=gx)
            dbg = cs(dur, 0, G * 2 * C).view(G, 2, C)
            for g in range(G):
                b_ = f"wavenet.stacks.{s}.blocks.{g}."
                grads[b_ + "conv.weight"], grads[b_ + "conv.bias"] = dWc[g], dbg[g, 0]
                grads[b_ + "res_conv.weight"], grads[b_ + "res_conv.bias"] = dWr[g].view(C, C, 1), dbg[g, 1]
            if s > 0:
                dy = self._run(pl.lvl_T[s], dur, self.buf(dy_next, M, G * C), B, T, g_a_col=2 * C, g_out_col=C)
                dy_next = "dyA" if dy_next == "dyB" else "dyB"
            else:
                dhw32 = self.buf("dhw32", M, C, f32, zero=True)
                self._run(pl.lvl_T[0], dur, dhw32, B, T, g_a_col=2 * C, g_out_col=0, epi=_lib.EPI_RESID)
                dhw = ops.cast_pad_bf16(dhw32, C, out=self.buf("dhw", M, C))
        taps = [wg(dhw, h0, C, C, shift=2 - k, flat=False) for k in range(3)]
        grads["wavenet.init_conv.weight"] = torch.stack(taps, dim=-1)
        grads["wavenet.init_conv.bias"] = cs(dhw, 0, C)
        dh0 = self._run(pl.wn_init_T, dhw, self.buf("dh0", M, C), B, T)
        grads["init_conv.weight"] = wg(dh0, xb, C, z)[:, :z].reshape(C, z, 1)
        grads["init_conv.bias"] = cs(dh0, 0, C)
        # time conditioning: 56 Linear(2048 -> 1024), SiLU, Linear(513 -> 2048), learned sinusoid weights
        dWcat, dbcat = zeros(*pl.Wcat.shape), zeros(*pl.bcat.shape)
        dtemb = zeros(B, temb.shape[1])
        ops.linear_f32_bwd(dgball, temb, pl.Wcat, dW=dWcat, db=dbcat, dX=dtemb)
        for i, n in enumerate(self.cond_names):
            grads[n + ".weight"] = dWcat[i * self.gbw:(i + 1) * self.gbw]
            grads[n + ".bias"] = dbcat[i * self.gbw:(i + 1) * self.gbw]
        dpre = ops.silu_bwd(pre, dtemb, torch.empty_like(pre))
        dW1, db1_, dfeat = zeros(*W1.shape), zeros(*b1.shape), zeros(B, feats.shape[1])
        ops.linear_f32_bwd(dpre, feats, W1, dW=dW1, db=db1_, dX=dfeat)
        grads["to_time_cond.1.weight"], grads["to_time_cond.1.bias"] = dW1, db1_
        grads["to_time_cond.0.weights"] = ops.time_features_bwd(t_idx, tw, dfeat, zeros(tw.shape[0]))
        return out, grads
