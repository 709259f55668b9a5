"""Host-side engine of the normalization pass: packs a reference ``state_dict`` once, owns the HBM workspace, and
drives the sm_100a kernels (through the C ABI) for VAE encode -> q_sample -> denoiser loop -> VAE decode ->
argmax -> run-length reduce.  torch is used only for memory, streams and CUDA-graph capture.

Data layout in HBM: every activation is row-major [B, T, C] (frames x channels, C padded to the GEMM granules),
bf16 for GEMM operands, fp32 for the transformer residual stream, the latent state and logits.  There are no
transposes anywhere (the reference flips between [B,C,T] and [B,T,C] around every conv, LM:863-866, :892-896).
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .config import UNIT_OFFSET, DiffNormConfig
from .packing import (pack_conv3, pack_geglu, pack_linear, pack_skip_sum, pack_wavenet_level, rup)
from .schedule import DDPMScheduler

bf16, f32, i32, i64 = torch.bfloat16, torch.float32, torch.int32, torch.int64


class _Wavenet:
    """Packed WavenetEncoder / Wavenet (LM:585-617, :1003-1032)."""

    def __init__(self, sd, pre: str, stacks: int, layers: int, cin_pad: int, final_epi: int, final_n_pad: int,
                 cond: bool, fmt: str = "bf16"):
        w = sd[pre + "init_conv.weight"]
        self.c = w.shape[0]
        self.c_pad = rup(self.c, 128)
        self.G, self.stacks, self.cond, self.fmt = layers, stacks, cond, fmt
        self.init = pack_conv3(w, sd[pre + "init_conv.bias"], cin_pad=cin_pad, n_pad=self.c_pad, name=pre + "init_conv",
                               fmt=fmt)
        self.levels = []
        for s in range(stacks):
            blk = [f"{pre}stacks.{s}.blocks.{i}." for i in range(layers)]
            self.levels.append(pack_wavenet_level(
                [sd[b + "conv.weight"] for b in blk], [sd[b + "conv.bias"] for b in blk],
                [sd[b + "res_conv.weight"] for b in blk], [sd[b + "res_conv.bias"] for b in blk], self.c_pad,
                name=f"{pre}stacks.{s}", fmt=fmt))
        last = [f"{pre}stacks.{stacks - 1}.blocks.{i}." for i in range(layers)]
        self.skip = pack_skip_sum([sd[b + "skip_conv.weight"] for b in last], [sd[b + "skip_conv.bias"] for b in last],
                                  self.c_pad, name=pre + "skip_sum", fmt=fmt)
        self.final = pack_linear(sd[pre + "final_conv.weight"], sd[pre + "final_conv.bias"], epi=final_epi,
                                 k_pad=self.c_pad, n_pad=final_n_pad, name=pre + "final_conv", fmt=fmt)

    def plans(self):
        return [self.init, *self.levels, self.skip, self.final]


class _TfLayer:
    pass


class DiffNormEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device: str = "cuda", cfg: Optional[DiffNormConfig] = None,
                 vae_only: bool = False, wfmt: Optional[str] = None, vae_fmt: Optional[str] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("DiffNormEngine needs a CUDA device: the product has no CPU path")
        sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        self.cfg = cfg or DiffNormConfig.from_state_dict(sd)
        self.dev = torch.device(device)
        self.ws: Dict[tuple, torch.Tensor] = {}
        # residual GEMM + following adaptive RMSNorm in one kernel (dn_gemm_resid_norm); DN_FUSE_NORM=0 keeps the pair
        self.fuse_norm = os.environ.get("DN_FUSE_NORM", "0") == "1"
        # Operand formats (DESIGN.md "operand formats"; evidence profiles/r02_a3_precision_probe.jsonl).  A bf16 weight rounded
        # ONCE carries the same error into all 99 calls of the sampler loop and that error adds up coherently in the latent (x0
        # error 0.18 %, 97.3 % unit agreement); an activation's rounding is fresh every call and averages out.  Two cures:
        #   "bf16sr" (default): bf16 operands, the weights re-rounded STOCHASTICALLY from their fp32 masters at every step
        #            (dn_sround_bf16, one launch inside the step graph): x0 error 0.036 %, 99.96 % agreement, bf16's power draw;
        #   "f16":   fp16 weights and activations (a tcgen05 kind::f16 MMA takes both in one format), stores saturate at
        #            +-65504: x0 error 0.022 %, but 4 % of clock under the 1000 W cap;
        #   "bf16":  round 1's format (A/B reference).
        # The once-per-pass VAE encoder / decoder have no averaging at all, and a 1 % logit error flips 2.5 % of near-tie
        # units: they run split precision (hi | lo bf16 pairs, 3 MMAs per K block).
        self.wfmt = wfmt or os.environ.get("DN_WFMT", "bf16" if self.fuse_norm else "bf16sr")
        self.vae_fmt = vae_fmt or os.environ.get("DN_VAE_FMT", "split")
        if self.wfmt not in ("bf16", "bf16sr", "f16") or self.vae_fmt not in ("bf16", "split"):
            raise ValueError("wfmt must be bf16 | bf16sr | f16 and vae_fmt bf16 | split")
        # "bf16sr": bf16 operands whose WEIGHTS are re-rounded stochastically from their fp32 masters at every sampler step
        # (dn_sround_bf16 inside the step graph): the weight error stops being the same in all 99 calls and averages out like
        # the activations' does, at bf16's power draw (fp16 costs 4 % of clock under the 1000 W cap)
        self.sr_seed = int(os.environ.get("DN_SR_SEED", "20240518"))
        self._sr32 = self._sr16 = None
        self.vsplit = self.vae_fmt == "split"
        self.adt = torch.float16 if self.wfmt == "f16" else bf16     # 16-bit activation format of the sampler loop
        # the DDIM update runs in the epilogue of the denoiser's last GEMM (DN_EPI_DDIM); DN_FUSE_DDIM=0 keeps dn_ddim_step
        self.fuse_ddim = os.environ.get("DN_FUSE_DDIM", "1") != "0" and self.cfg.latent_dim % 16 == 0
        self.gemm_impl = None   # None = automatic (CTA-pair kernel when the launch has >= 74 pair tiles); tests force others
        self._graphs: Dict[tuple, object] = {}
        self._shape_seen: Dict[tuple, int] = {}
        self.graph_after = 1     # eager passes of a (B, T) shape before its sampler step is captured (0 = capture at once)
        self._graph_kernels: Dict[tuple, int] = {}
        self.replayed_kernels = 0   # kernels executed through CUDA-graph replays (not visible to dn_launch_count)
        self._prof = None
        self._reserve_rows = 0
        c = self.cfg
        self.zp = rup(c.latent_dim, 64)        # latent staging width (K of the first GEMMs)
        self.xw = 2 * self.zp                  # the staging row is a split-precision pair [hi (zp) | lo (zp)]
        self.zn = rup(c.latent_dim, 16)        # eps_hat row width
        self.vp = rup(c.vocab, 16)             # logits row width
        self.vae_only = vae_only   # training keeps a frozen-VAE engine; the denoiser weights change every step
        if not vae_only:
            self._pack_denoiser(sd)
        self._pack_vae(sd)
        if not vae_only:
            self._build_time_table(sd)
        self.sched = DDPMScheduler(c.timesteps)
        self.ddim_rows = torch.from_numpy(self.sched.ddim_rows()).to(self.dev)
        self._pe_cache: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------ packing
    def _dev(self, plan):
        return plan.to(self.dev)

    def _pack_tf(self, sd, pre: str, dim: int, depth: int, cond: bool, fmt: str):
        layers = []
        ip = rup(DiffNormConfig.ff_inner(dim), 128)
        for l in range(depth):
            p = f"{pre}layers.{l}."
            L = _TfLayer()
            L.qkv = self._dev(pack_linear(torch.cat([sd[p + "1.to_q.weight"], sd[p + "1.to_kv.weight"]], 0), None,
                                          name=p + "qkv", fmt=fmt))
            L.out = self._dev(pack_linear(sd[p + "1.to_out.weight"], None, epi=_lib.EPI_RESID, name=p + "to_out", fmt=fmt))
            L.ff1 = self._dev(pack_geglu(sd[p + "5.0.weight"], sd[p + "5.0.bias"], name=p + "ff.geglu", fmt=fmt))
            L.ffc = self._dev(pack_conv3(sd[p + "5.2.1.weight"], sd[p + "5.2.1.bias"], cin_pad=ip, n_pad=ip,
                                         name=p + "ff.conv", fmt=fmt))
            L.ff3 = self._dev(pack_linear(sd[p + "5.3.weight"], sd[p + "5.3.bias"], epi=_lib.EPI_RESID, k_pad=ip,
                                          name=p + "ff.out", fmt=fmt))
            if not cond:
                L.g1 = sd[p + "0.gamma"].float().to(self.dev)
                L.g2 = sd[p + "4.gamma"].float().to(self.dev)
            layers.append(L)
        return layers, ip

    def _pack_denoiser(self, sd):
        c = self.cfg
        w = self.wfmt
        # the first GEMM reads the fp32 latent as a split-precision pair (K = 64: three K blocks instead of one)
        self.d_init = self._dev(pack_linear(sd["model.init_conv.weight"], sd["model.init_conv.bias"], k_pad=self.zp,
                                            name="model.init_conv", fmt="split"))
        self.d_wn = _Wavenet(sd, "model.wavenet.", c.wn_stacks, c.wn_layers, cin_pad=c.hid, final_epi=_lib.EPI_F32,
                             final_n_pad=c.hid, cond=True, fmt=w)
        for p in self.d_wn.plans():
            self._dev(p)
        self.d_layers, self.d_ip = self._pack_tf(sd, "model.transformer.", c.hid, c.depth, cond=True, fmt=w)
        self.d_pred_gamma = sd["model.transformer.to_pred.0.gamma"].float().to(self.dev)
        self.d_pred = self._dev(pack_linear(sd["model.transformer.to_pred.1.weight"], None, name="model.to_pred", fmt=w))
        self.d_proj = self._dev(pack_linear(sd["model.final_proj.weight"], sd["model.final_proj.bias"],
                                            epi=_lib.EPI_F32, n_pad=self.zn, name="model.final_proj", fmt=w))
        if w == "bf16sr":
            self._gather_sr_weights()

    def _sr_plans(self):
        plans = [*self.d_wn.plans(), self.d_pred, self.d_proj]
        for L in self.d_layers:
            plans += [L.qkv, L.out, L.ff1, L.ffc, L.ff3]
        return [p for p in plans if p.W32 is not None]

    def _gather_sr_weights(self):
        """Move every re-rounded plan's fp32 master and bf16 weight into ONE flat buffer each (offsets 256-byte aligned), so
        that a sampler step re-rounds all of them with a single launch; the plans keep views."""
        plans = self._sr_plans()
        offs, tot = [], 0
        for p in plans:
            offs.append(tot)
            tot += (p.W32.numel() + 127) // 128 * 128
        self._sr32 = torch.zeros(tot, dtype=f32, device=self.dev)
        self._sr16 = torch.zeros(tot, dtype=bf16, device=self.dev)
        for p, o in zip(plans, offs):
            n = p.W32.numel()
            self._sr32[o:o + n].copy_(p.W32.reshape(-1))
            self._sr16[o:o + n].copy_(p.W.reshape(-1))
            p.W = self._sr16[o:o + n].view(p.W.shape)
            p.W32 = self._sr32[o:o + n].view(p.W.shape)

    def reround_weights(self, t_idx):
        if self._sr32 is not None:
            ops.sround_bf16(self._sr32, self._sr16, self.sr_seed, t_idx)

    def _pack_vae(self, sd):
        c = self.cfg
        pre = "speech_decoder."
        self.enc: List[_Wavenet] = []
        cin_pad = c.feat_dim
        enc_w = c.enc_widths()
        for i, (cin, cout) in enumerate(enc_w):
            last = i == len(enc_w) - 1
            cp = rup(cout, 128)
            wn = _Wavenet(sd, f"{pre}encoder_wave.{i}.", c.vae_stacks, c.vae_layers, cin_pad=cin_pad,
                          final_epi=_lib.EPI_F32 if last else _lib.EPI_BF16, final_n_pad=rup(cout, 16) if last else cp,
                          cond=False, fmt=self.vae_fmt)
            for p in wn.plans():
                self._dev(p)
            self.enc.append(wn)
            cin_pad = cp
        self.dec: List[_Wavenet] = []
        cin_pad = self.zp
        dec_w = c.dec_widths()
        for i, (cin, cout) in enumerate(dec_w):
            last = i == len(dec_w) - 1
            cp = rup(cout, 128)
            wn = _Wavenet(sd, f"{pre}decoder_wave.{i}.", c.vae_stacks, c.vae_layers, cin_pad=cin_pad,
                          final_epi=_lib.EPI_F32 if last else _lib.EPI_BF16, final_n_pad=cout if last else cp, cond=False,
                          fmt=self.vae_fmt)
            for p in wn.plans():
                self._dev(p)
            self.dec.append(wn)
            cin_pad = cp
        self.v_layers, self.v_ip = self._pack_tf(sd, pre + "decoder_tf.", c.feat_dim, c.vae_depth, cond=False,
                                                 fmt=self.vae_fmt)
        self.v_pred_gamma = sd[pre + "decoder_tf.to_pred.0.gamma"].float().to(self.dev)
        self.v_pred = self._dev(pack_linear(sd[pre + "decoder_tf.to_pred.1.weight"], None, epi=_lib.EPI_F32,
                                            name="vae.to_pred", fmt=self.vae_fmt))
        self.v_lm = self._dev(pack_linear(sd[pre + "decoder_lm.weight"], sd[pre + "decoder_lm.bias"], epi=_lib.EPI_F32,
                                          n_pad=self.vp, name="vae.decoder_lm", fmt=self.vae_fmt))

    def _build_time_table(self, sd):
        """gamma/beta of all 32 WaveNet FiLMs + 24 adaptive norms depend only on t (LM:507,:624,:741-745):
        precompute table[t, layer, 2C] in fp32 with the library's own kernels."""
        c = self.cfg
        dev = self.dev
        steps = torch.arange(c.timesteps, dtype=i32, device=dev)
        feats = ops.time_features(steps, sd["model.to_time_cond.0.weights"].float().to(dev).contiguous())
        t_emb = ops.linear_f32(feats, sd["model.to_time_cond.1.weight"].float().to(dev).contiguous(),
                               sd["model.to_time_cond.1.bias"].float().to(dev).contiguous(), act=1)
        names = [f"model.wavenet.stacks.{s}.blocks.{i}.to_time_cond" for s in range(c.wn_stacks)
                 for i in range(c.wn_layers)]
        for l in range(c.depth):
            names += [f"model.transformer.layers.{l}.0.to_gamma_beta", f"model.transformer.layers.{l}.4.to_gamma_beta"]
        W = torch.cat([sd[n + ".weight"].float() for n in names], 0).to(dev).contiguous()
        b = torch.cat([sd[n + ".bias"].float() for n in names], 0).to(dev).contiguous()
        self.n_cond = len(names)
        self.gb_w = 2 * c.hid
        self.time_table = ops.linear_f32(t_emb, W, b)          # [T, n_cond * 2C]
        self.table_flat = self.time_table.view(-1)
        self.gb_t_stride = self.n_cond * self.gb_w
        self.t_emb = t_emb
        del W, b
        torch.cuda.synchronize()

    # ------------------------------------------------------------------------------------------------ workspace
    def buf(self, name: str, rows: int, width: int, dtype=bf16, frames: bool = True) -> torch.Tensor:
        """Named workspace buffer [rows, width].  One allocation per (name, width, dtype) is shared by every batch
        shape: a request returns the leading `rows` rows of the largest buffer allocated so far (row-major, so the
        prefix is contiguous).  Buffers are zero-initialised and the GEMM pad columns are only ever written with
        zeros, so the padding invariants survive sharing.  Growing a buffer drops the captured CUDA graphs (they hold
        the old pointers); call `reserve(max_rows)` first to avoid re-capturing."""
        key = (name, width, dtype)
        t = self.ws.get(key)
        want = max(rows, self._reserve_rows) if frames else rows  # frames=True: rows scale with B*T
        if t is None or t.shape[0] < rows:
            if t is not None:  # replacing a buffer: captured graphs hold the old pointer
                self._graphs.clear()
                self._graph_kernels.clear()
            t = torch.zeros(want, width, dtype=dtype, device=self.dev)
            self.ws[key] = t
        return t[:rows]

    def reserve(self, max_rows: int):
        """Size the shared workspace for batches of up to `max_rows` = B*T frames."""
        self._reserve_rows = max(self._reserve_rows, int(max_rows))

    def pe_table(self, T: int) -> torch.Tensor:
        """Sinusoidal table rows 0..T (row 0 = padding = zeros), sinusoidal_positional_embedding.py:36-58."""
        t = self._pe_cache.get(T)
        if t is None:
            dim = self.cfg.hid
            half = dim // 2
            e = np.log(10000.0) / (half - 1)
            freq = torch.exp(torch.arange(half, dtype=torch.float) * -e)
            ang = torch.arange(T + 1, dtype=torch.float)[:, None] * freq[None, :]
            tab = torch.cat([ang.sin(), ang.cos()], dim=1)
            tab[0] = 0
            t = tab.to(self.dev).contiguous()
            self._pe_cache[T] = t
        return t

    # ------------------------------------------------------------------------------------------------ sub-networks
    def _run(self, plan, A, out, B, T, **kw):
        prof = self._prof
        if prof is not None and prof["match"] in plan.name:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(A, out, B, T, impl=self.gemm_impl, **kw)
            e1.record()
            prof["events"].append((plan.name, e0, e1))
            return out
        return plan.run(A, out, B, T, impl=self.gemm_impl, **kw)

    def profile_launches(self, match: str, fn):
        """Run fn() eagerly with CUDA events around every dn_gemm launch whose plan name contains `match`;
        returns [(name, ms)] (bench.py's live per-launch timing of the dominant kernel)."""
        self._prof = {"match": match, "events": []}
        try:
            fn()
            torch.cuda.synchronize()
            return [(n, a.elapsed_time(b)) for n, a, b in self._prof["events"]]
        finally:
            self._prof = None

    def _wavenet(self, wn: _Wavenet, A, out, B, T, tag: str, gb_layer0: Optional[int] = None, t_idx=None,
                 t_idx_stride=0, pe=None, lengths=None, out_split: bool = False):
        """split-precision nets (wn.fmt == "split"): every bf16 activation is a [hi | lo] pair of twice the width."""
        M = B * T
        cp, G = wn.c_pad, wn.G
        sp = wn.fmt == "split"
        w2 = 2 if sp else 1
        dt = torch.float16 if wn.fmt == "f16" else bf16
        h = self.buf(tag + ".h", M, w2 * cp, dt)
        self._run(wn.init, A, h, B, T, out_split=sp)
        ys = [self.buf(tag + ".y0", M, w2 * G * cp, dt), self.buf(tag + ".y1", M, w2 * G * cp, dt)]
        src, g_a_col = h, 0
        for s, lvl in enumerate(wn.levels):
            dst = ys[s & 1]
            kw = {}
            if gb_layer0 is not None:
                kw = dict(gb=self.table_flat[(gb_layer0 + s * G) * self.gb_w:], gb_t_stride=self.gb_t_stride,
                          g_gb=self.gb_w, gb_half=self.gb_w // 2, t_idx=t_idx, t_idx_stride=t_idx_stride)
            self._run(lvl, src, dst, B, T, g_a_col=g_a_col, g_out_col=cp, out_split=sp, **kw)
            src, g_a_col = dst, cp
        sk = self.buf(tag + ".s", M, w2 * cp, dt)
        self._run(wn.skip, src, sk, B, T, out_split=sp)
        self._run(wn.final, sk, out, B, T, pe=pe, lengths=lengths, out_split=out_split)
        return out

    def _transformer(self, layers, ip, x, B, T, dim, heads, dh, lengths, tag, cond_layer0=None, t_idx=None,
                     t_idx_stride=0, final_gamma=None, split: bool = False):
        """split (the precise VAE decoder): the bf16 operands hb, ao, m1, m2 are [hi | lo] pairs; q, k, v are fp16."""
        M = B * T
        w2 = 2 if split else 1
        dt = bf16 if split else (torch.float16 if layers[0].qkv.fmt == "f16" else bf16)
        hb = self.buf(tag + ".h", M, w2 * dim, dt)
        qkv = self.buf(tag + ".qkv", M, 3 * heads * dh, torch.float16 if split else dt)
        ao = self.buf(tag + ".ao", M, w2 * heads * dh, dt)
        m1 = self.buf(tag + ".m1", M, w2 * ip, dt)
        m2 = self.buf(tag + ".m2", M, w2 * ip, dt)
        if final_gamma is not None:
            # row-complete fused form (shared timestep, width 512): every residual GEMM also writes the next norm's output
            def cond(i):
                return self.table_flat[(cond_layer0 + i) * self.gb_w:]
            ops.adarmsnorm(x, hb, B, T, None, cond(0), self.gb_t_stride, t_idx, 0)
            for l, L in enumerate(layers):
                self._run(L.qkv, hb, qkv, B, T)
                ops.attention(qkv, ao, lengths, B, T, heads, dh)
                ops.gemm_resid_norm(L.out, ao, x, hb, None, cond(2 * l + 1), self.gb_t_stride, t_idx)
                self._run(L.ff1, hb, m1, B, T)
                self._run(L.ffc, m1, m2, B, T)
                if l + 1 < len(layers):
                    ops.gemm_resid_norm(L.ff3, m2, x, hb, None, cond(2 * l + 2), self.gb_t_stride, t_idx)
                else:
                    ops.gemm_resid_norm(L.ff3, m2, x, hb, final_gamma)
            return x
        for l, L in enumerate(layers):
            for which in (0, 1):
                if cond_layer0 is not None:
                    gb = self.table_flat[(cond_layer0 + 2 * l + which) * self.gb_w:]
                    ops.adarmsnorm(x, hb, B, T, None, gb, self.gb_t_stride, t_idx, t_idx_stride)
                else:
                    ops.adarmsnorm(x, hb, B, T, L.g1 if which == 0 else L.g2, split=split)
                if which == 0:
                    self._run(L.qkv, hb, qkv, B, T, out_f16=split)
                    ops.attention(qkv, ao, lengths, B, T, heads, dh, out_split=split)
                    self._run(L.out, ao, x, B, T)
                else:
                    self._run(L.ff1, hb, m1, B, T, out_split=split)
                    self._run(L.ffc, m1, m2, B, T, out_split=split)
                    self._run(L.ff3, m2, x, B, T)
        return x

    def denoise(self, xb: torch.Tensor, lengths: torch.Tensor, B: int, T: int, t_idx: torch.Tensor,
                t_idx_stride: int = 0, ddim_into: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Model.forward (LM:828-876).  xb bf16 [B*T, 2*zp] latent staging (split pair); returns eps_hat fp32 [B*T, zn].
        ddim_into = the fp32 latent x [B*T, z]: the last GEMM's epilogue applies the DDIM update to x and xb instead of
        writing eps_hat (shared timestep only); returns x."""
        c = self.cfg
        M = B * T
        h0 = self.buf("d.h0", M, c.hid, self.adt)
        self._run(self.d_init, xb, h0, B, T)
        x = self.buf("d.x", M, c.hid, f32)
        self._wavenet(self.d_wn, h0, x, B, T, "d.wn", gb_layer0=0, t_idx=t_idx, t_idx_stride=t_idx_stride,
                      pe=self.pe_table(T), lengths=lengths)
        fuse = self.fuse_norm and t_idx_stride == 0 and c.hid == 512
        self._transformer(self.d_layers, self.d_ip, x, B, T, c.hid, c.heads, c.dim_head, lengths, "d.tf",
                          cond_layer0=c.wn_stacks * c.wn_layers, t_idx=t_idx, t_idx_stride=t_idx_stride,
                          final_gamma=self.d_pred_gamma if fuse else None)
        hb = self.buf("d.tf.h", M, c.hid, self.adt)
        if not fuse:
            ops.adarmsnorm(x, hb, B, T, self.d_pred_gamma)
        pb = self.buf("d.pred", M, c.hid, self.adt)
        self._run(self.d_pred, hb, pb, B, T)
        if ddim_into is not None:
            assert t_idx_stride == 0
            self._run(self.d_proj, pb, ddim_into, B, T, ddim=(self.ddim_rows, t_idx, xb, self.zp))
            return ddim_into
        eh = self.buf("d.eps", M, self.zn, f32)
        self._run(self.d_proj, pb, eh, B, T)
        return eh

    def encode_params(self, feat: torch.Tensor) -> torch.Tensor:
        """WaveNet encoder stack (LM:1100-1103): feat fp32 [B,T,768] -> posterior params fp32 [B,T,2z]."""
        B, T, Cf = feat.shape
        M = B * T
        sp = self.vsplit
        a = self.buf("e.in", M, (2 if sp else 1) * Cf)
        ops.cast_split(feat.contiguous().view(M, Cf), a, Cf if sp else 0)
        for i, wn in enumerate(self.enc):
            last = i == len(self.enc) - 1
            out = self.buf("e.params", M, wn.final.n_out, f32) if last else self.buf(f"e.o{i}", M, (2 if sp else 1) * wn.c_pad)
            self._wavenet(wn, a, out, B, T, f"e.wn{i}", out_split=sp and not last)
            a = out
        return a.view(B, T, -1)

    def encode(self, feat: torch.Tensor, eps: torch.Tensor, eps_channel_first: bool = True) -> torch.Tensor:
        """encode_feature + transpose (LM:1099-1107, :1397): -> z fp32 [B,T,z]."""
        params = self.encode_params(feat)
        return ops.vae_reparam(params, eps.contiguous(), self.cfg.latent_dim, eps_channel_first)

    def decode(self, xb: torch.Tensor, lengths: torch.Tensor, B: int, T: int, units_only: bool = False):
        """decode_feature (LM:1109-1116): latent staging bf16 [B*T, 2*zp] (split pair) -> (recon fp32 [B,T,768], logits
        fp32 [B,T,vp]).  units_only: the unit head's epilogue takes the argmax itself (LM:1450-1451) and the second value
        returned is units int64 [B,T] — the 257 MB of logits are never written."""
        c = self.cfg
        M = B * T
        sp = self.vsplit
        w2 = 2 if sp else 1
        a = xb if sp else self._hi_only(xb, M)
        x = self.buf("v.x", M, c.feat_dim, f32)
        for i, wn in enumerate(self.dec):
            last = i == len(self.dec) - 1
            out = x if last else self.buf(f"v.o{i}", M, w2 * wn.c_pad)
            self._wavenet(wn, a, out, B, T, f"v.wn{i}", out_split=sp and not last)
            a = out
        self._transformer(self.v_layers, self.v_ip, x, B, T, c.feat_dim, c.vae_heads, c.vae_dim_head, lengths, "v.tf",
                          split=sp)
        hb = self.buf("v.tf.h", M, w2 * c.feat_dim)
        ops.adarmsnorm(x, hb, B, T, self.v_pred_gamma, split=sp)
        recon = self.buf("v.recon", M, c.feat_dim, f32)
        self._run(self.v_pred, hb, recon, B, T)
        rb = self.buf("v.recon_bf16", M, w2 * c.feat_dim)
        ops.cast_split(recon, rb, c.feat_dim if sp else 0)
        if units_only:
            parts = self.buf("v.argmax_parts", M, 4 * self.v_lm.n_tiles, f32)
            self._run(self.v_lm, rb, parts, B, T, argmax_classes=c.vocab)
            return recon.view(B, T, -1), ops.argmax_combine(parts, UNIT_OFFSET).view(B, T)
        logits = self.buf("v.logits", M, self.vp, f32)
        self._run(self.v_lm, rb, logits, B, T)
        return recon.view(B, T, -1), logits.view(B, T, -1)

    def _hi_only(self, xb, M):
        """bf16-format VAE (DN_VAE_FMT=bf16, A/B runs only): a plain-width copy of the staging row's hi half."""
        a = self.buf("s.xb_hi", M, self.zp)
        a.copy_(xb[:, : self.zp])
        return a

    # ------------------------------------------------------------------------------------------------ the pass
    def _ddim_step(self, B, T):
        M, z = B * T, self.cfg.latent_dim
        x, xb = self.buf("s.x", M, z, f32), self.buf("s.xb", M, self.xw)
        t_idx, lens = self.buf("s.t", 1, 1, i32, frames=False).view(-1), self.buf("s.len", B, 1, i32, frames=False).view(-1)
        self.reround_weights(t_idx)
        if self.fuse_ddim:
            self.denoise(xb, lens, B, T, t_idx, ddim_into=x)
        else:
            eh = self.denoise(xb, lens, B, T, t_idx)
            ops.ddim_step(x, eh, self.ddim_rows, t_idx, 0, xb, self.zp)
        ops.advance_step(t_idx, -1)

    def _ddim_graph(self, B, T):
        """One sampler step (denoiser call + DDIM update + device-side t -= 1) captured as a CUDA graph; the step
        index lives in device memory so the same graph serves every step."""
        key = ("ddim", B, T)
        g = self._graphs.get(key)
        if g is None:
            t_idx, lens = self.buf("s.t", 1, 1, i32, frames=False).view(-1), self.buf("s.len", B, 1, i32, frames=False).view(-1)
            t_idx.fill_(1)
            lens.fill_(T)
            self._ddim_step(B, T)  # warm-up outside capture: function attributes, workspace allocation
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self._ddim_step(B, T)
            self._graph_kernels[key] = _lib.launch_count() - n0
            self._graphs[key] = g
        return g

    @torch.no_grad()
    def normalize(self, feat: torch.Tensor, lengths: torch.Tensor, start_step: int, eps_vae: Optional[torch.Tensor] = None,
                  eps_q: Optional[torch.Tensor] = None, ref_units: Optional[torch.Tensor] = None, sampler: str = "ddim",
                  step_noise: Optional[Sequence[torch.Tensor]] = None, timesteps: Optional[Sequence[int]] = None,
                  large_var: bool = False, use_graph: bool = True, collect: bool = False, reduce: bool = True,
                  logits: bool = False):
        """The whole pass on resident inputs (ddim_sample, LM:1386-1471, + the driver's reduce,
        diff_norm_synthesis.py:211-216).  feat fp32 [B,T,768] cuda, lengths int32 [B] cuda.
        logits (or collect): also return the [B,T,1004] logits; by default the unit head reduces them to units in its own
        epilogue and they are never materialised (the units are bit-identical either way)."""
        c = self.cfg
        B, T, _ = feat.shape
        M = B * T
        z = c.latent_dim
        if not (0 < start_step < c.timesteps):
            raise ValueError(f"start_step must be in (0, {c.timesteps}) (LM:1405 indexes the schedule with it)")
        if eps_vae is None:    # the library's own Philox draws (seeded from torch's generator, so torch.manual_seed replays them)
            eps_vae = ops.randn((B, z, T), self.dev, self._noise_seed(), 0)
        if eps_q is None:
            eps_q = ops.randn((B, T, z), self.dev, self._noise_seed(), 0)
        # a CUDA graph pays its warm-up step + capture (GPU idle meanwhile) only for shapes that come back: length-bucketed
        # batches of a dataset are mostly one-off (B, T) shapes and run eager (the host stays ~20x ahead of a 20 ms step)
        seen = self._shape_seen.get((B, T), 0)
        self._shape_seen[(B, T)] = seen + 1
        want_graph = use_graph and sampler == "ddim" and start_step > 2   # (a graph replays t -= 1 steps; 1-2 calls run eager)
        graph = self._ddim_graph(B, T) if want_graph and (seen >= self.graph_after or ("ddim", B, T) in self._graphs) else None
        graph_kernels = self._graph_kernels.get(("ddim", B, T), 0)
        lens = self.buf("s.len", B, 1, i32, frames=False).view(-1)
        lens.copy_(lengths)
        out = {}
        zlat = self.encode(feat, eps_vae)
        x = self.buf("s.x", M, z, f32)
        xb = self.buf("s.xb", M, self.xw)
        s = self.sched
        ops.q_sample(zlat, eps_q.contiguous(), float(np.float32(s.sqrt_alphas_cumprod[start_step])),
                     float(np.float32(s.sqrt_one_minus_alphas_cumprod[start_step])), x, xb, self.zp)
        if collect:
            out["z"] = zlat.clone()
            out["x_start"] = x.view(B, T, z).clone()
        t_idx = self.buf("s.t", 1, 1, i32, frames=False).view(-1)
        calls = 0
        if sampler == "ddim":
            t_idx.fill_(start_step - 1)
            # t = start-1 .. 1; the loop breaks after t = 1 (LM:1402,1444), so t = 0 runs only when start_step == 1
            n = max(start_step - 1, 1)
            if collect and n > 0:
                out["eps_first"] = self.denoise(xb, lens, B, T, t_idx).view(B, T, -1)[..., :z].clone()
            for _ in range(n):
                if graph is not None:
                    graph.replay()
                    self.replayed_kernels += graph_kernels
                else:
                    self._ddim_step(B, T)
            calls = n
        elif sampler == "ddpm":
            rows = torch.from_numpy(s.ddpm_rows(large_var)).to(self.dev)
            t_idx.fill_(start_step - 1)
            n = start_step - 1
            for k in range(n):
                self.reround_weights(t_idx)
                eh = self.denoise(xb, lens, B, T, t_idx)
                noise = step_noise[k] if step_noise is not None else torch.randn(B, T, z, device=self.dev, dtype=f32)
                ops.ddpm_step(x, eh, noise.contiguous(), rows, t_idx, xb, self.zp)
                ops.advance_step(t_idx, -1)
            calls = n
        elif sampler == "ddim_strided":
            sp, tmap = s.spaced(timesteps)
            rows = torch.from_numpy(sp.ddim_rows()).to(self.dev)
            r_idx = self.buf("s.r", 1, 1, i32, frames=False).view(-1)
            for i in range(len(tmap) - 1, 0, -1):
                t_idx.fill_(tmap[i])          # the model sees the original step (respace.py:117-129)
                r_idx.fill_(i)                # the update uses the respaced table row
                self.reround_weights(t_idx)
                eh = self.denoise(xb, lens, B, T, t_idx)
                ops.ddim_step(x, eh, rows, r_idx, 1, xb, self.zp)
                calls += 1
        else:
            raise ValueError(f"unknown sampler {sampler!r}")
        out["calls"] = calls
        # xb mirrors x as a split-precision pair (written by q_sample / the step kernels)
        if logits or collect:
            recon, lg = self.decode(xb, lens, B, T)
            units = ops.argmax_units(lg, c.vocab, UNIT_OFFSET)
            out["logits"] = lg
        else:
            recon, units = self.decode(xb, lens, B, T, units_only=True)
        out.update(x0=x.view(B, T, z), recon=recon, units=units)
        if ref_units is not None:
            out["acc"] = ops.unit_accuracy(units, ref_units.contiguous(), lens)
        if reduce:
            out["dedup"], out["duration"], out["index_to_keep"], out["counts"] = ops.reduce_tgt(units, lens)
        return out

    @staticmethod
    def _noise_seed() -> int:
        """A fresh 63-bit seed from torch's default CPU generator: `torch.manual_seed(s)` makes a pass reproducible."""
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def stage_latent(self, latent: torch.Tensor) -> torch.Tensor:
        """fp32 [B,T,z] latent -> split-precision bf16 staging buffer [B*T, 2*zp] (for decode / denoise on caller-supplied
        latents)."""
        B, T, z = latent.shape
        xb = self.buf("s.xb", B * T, self.xw)
        ops.cast_split(latent.contiguous().view(B * T, z), xb, self.zp)
        return xb

    def precision_mode(self) -> str:
        what = {"bf16sr": "bf16 operands, weights re-rounded stochastically every step", "f16": "fp16 operands",
                "bf16": "bf16 operands"}[self.wfmt]
        return f"loop: {what} (fp32 accumulate, fp32 latent in); VAE encode/decode: {self.vae_fmt}"
