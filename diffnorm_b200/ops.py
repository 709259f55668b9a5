"""Thin torch-tensor wrappers over the C ABI (include/diffnorm_b200.h).  torch is used for device memory and
streams only; every op below launches hand-written sm_100a kernels from libdiffnorm_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import GemmDesc, GemmSeg, check, lib

bf16, f16, f32, i32, i64 = torch.bfloat16, torch.float16, torch.float32, torch.int32, torch.int64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return t


def set_sm_limit(n: int):
    """Cap the persistent kernels' grids at n SMs (0 = all): leaves SMs to a concurrent NCCL collective."""
    check(lib.dn_set_sm_limit(int(n)), "dn_set_sm_limit")


# ------------------------------------------------------------------------------------------------ integer ops
def reduce_tgt(units: torch.Tensor, lengths: torch.Tensor):
    """Batched run-length reduction.  units [B,T] int64, lengths [B] int32 ->
    (dedup [B,T], duration [B,T], index_to_keep [B,T], counts [B] int32); row b valid for j < counts[b]."""
    _chk(units, i64, "units"), _chk(lengths, i32, "lengths")
    B, T = units.shape
    dedup = torch.empty_like(units)
    dur = torch.empty_like(units)
    keep = torch.empty_like(units)
    counts = torch.empty(B, dtype=i32, device=units.device)
    check(lib.dn_reduce_tgt(_p(units), _p(lengths), B, T, _p(dedup), _p(dur), _p(keep), _p(counts), _stream()),
          "dn_reduce_tgt")
    return dedup, dur, keep, counts


def argmax_units(logits: torch.Tensor, n_classes: int, offset: int = 4, out: Optional[torch.Tensor] = None):
    """logits [..., ld] fp32|bf16 (ld >= n_classes) -> units [...] int64 = argmax - offset."""
    if logits.dtype not in (f32, bf16):
        raise ValueError("logits must be fp32 or bf16")
    _chk(logits, logits.dtype, "logits")
    ld = logits.shape[-1]
    rows = logits.numel() // ld
    if out is None:
        out = torch.empty(logits.shape[:-1], dtype=i64, device=logits.device)
    check(lib.dn_argmax_units(_p(logits), int(logits.dtype == bf16), rows, n_classes, ld, offset, _p(out), _stream()),
          "dn_argmax_units")
    return out


def argmax_combine(partials: torch.Tensor, offset: int = 4, out: Optional[torch.Tensor] = None):
    """partials fp32 [rows, 2 * parts] written by a GEMM with the argmax epilogue -> units [rows] int64."""
    _chk(partials, f32, "partials")
    rows, w = partials.shape
    if out is None:
        out = torch.empty(rows, dtype=i64, device=partials.device)
    check(lib.dn_argmax_combine(_p(partials), rows, w // 2, offset, _p(out), _stream()), "dn_argmax_combine")
    return out


def unit_accuracy(units, ref_units, lengths):
    _chk(units, i64, "units"), _chk(ref_units, i64, "ref_units"), _chk(lengths, i32, "lengths")
    B, T = units.shape
    out = torch.empty(2, dtype=i64, device=units.device)
    check(lib.dn_unit_accuracy(_p(units), _p(ref_units), _p(lengths), B, T, _p(out), _stream()), "dn_unit_accuracy")
    return out


def gather_pack(src, src_row0, index_to_keep, counts, T: int, ldd: Optional[int] = None, dst_dtype=f32):
    _chk(src, f32, "src"), _chk(src_row0, i64, "src_row0"), _chk(index_to_keep, i64, "index_to_keep")
    _chk(counts, i32, "counts")
    B = counts.shape[0]
    Csrc = src.shape[-1]
    ldd = Csrc if ldd is None else ldd
    assert index_to_keep.shape == (B, T)
    dst = torch.empty(B, T, ldd, dtype=dst_dtype, device=src.device)
    check(lib.dn_gather_pack(_p(src), _p(src_row0), _p(index_to_keep), _p(counts), B, T, Csrc, _p(dst), ldd,
                             int(dst_dtype == bf16), _stream()), "dn_gather_pack")
    return dst


# ------------------------------------------------------------------------------------------------ elementwise
def cast_pad_bf16(src: torch.Tensor, ldo: int, out: Optional[torch.Tensor] = None):
    _chk(src, f32, "src")
    Cc = src.shape[-1]
    rows = src.numel() // Cc
    if out is None:
        out = torch.empty(*src.shape[:-1], ldo, dtype=bf16, device=src.device)
    check(lib.dn_cast_pad_bf16(_p(src), rows, Cc, Cc, _p(out), ldo, _stream()), "dn_cast_pad_bf16")
    return out


def cast_split(src: torch.Tensor, out: torch.Tensor, lo_col: int):
    """fp32 [rows, C] -> split-precision bf16 operand [rows, ldo]: hi in [0, C), lo = bf16(x - hi) in [lo_col, lo_col + C)."""
    _chk(src, f32, "src"), _chk(out, bf16, "out")
    Cc = src.shape[-1]
    check(lib.dn_cast_split(_p(src), src.numel() // Cc, Cc, Cc, _p(out), out.shape[-1], lo_col, _stream()), "dn_cast_split")
    return out


def sround_bf16(src: torch.Tensor, dst: torch.Tensor, seed: int, step: Optional[torch.Tensor] = None):
    """dst (bf16) = stochastic rounding of src (fp32), fresh for every value of the device step counter."""
    _chk(src, f32, "src"), _chk(dst, bf16, "dst")
    check(lib.dn_sround_bf16(_p(src), _p(dst), src.numel(), int(seed) & 0xFFFFFFFF, _p(step), _stream()), "dn_sround_bf16")
    return dst


def vae_reparam(params, eps, z: int, eps_channel_first: bool, out=None):
    _chk(params, f32, "params"), _chk(eps, f32, "eps")
    B, T, ldp = params.shape
    if out is None:
        out = torch.empty(B, T, z, dtype=f32, device=params.device)
    check(lib.dn_vae_reparam(_p(params), ldp, _p(eps), int(eps_channel_first), B, T, z, _p(out), _stream()),
          "dn_vae_reparam")
    return out


def q_sample(z_lat, eps, sqrt_ab: float, sqrt_1m_ab: float, x_out, x_bf16=None, x_lo_col: int = 0):
    _chk(z_lat, f32, "z"), _chk(eps, f32, "eps"), _chk(x_out, f32, "x")
    z = z_lat.shape[-1]
    rows = z_lat.numel() // z
    ldx = 0 if x_bf16 is None else x_bf16.shape[-1]
    check(lib.dn_q_sample(_p(z_lat), _p(eps), sqrt_ab, sqrt_1m_ab, rows, z, _p(x_out), _p(x_bf16), ldx, x_lo_col, _stream()),
          "dn_q_sample")
    return x_out


def ddim_step(x, eps_hat, coef_table, t_idx, mode: int = 0, x_bf16=None, x_lo_col: int = 0):
    _chk(x, f32, "x"), _chk(eps_hat, f32, "eps_hat"), _chk(coef_table, f32, "coef"), _chk(t_idx, i32, "t_idx")
    z = x.shape[-1]
    rows = x.numel() // z
    ldx = 0 if x_bf16 is None else x_bf16.shape[-1]
    check(lib.dn_ddim_step(_p(x), _p(eps_hat), eps_hat.shape[-1], _p(coef_table), _p(t_idx), rows, z, mode, _p(x_bf16),
                           ldx, x_lo_col, _stream()), "dn_ddim_step")
    return x


def ddpm_step(x, eps_hat, noise, coef_table, t_idx, x_bf16=None, x_lo_col: int = 0):
    _chk(x, f32, "x"), _chk(eps_hat, f32, "eps_hat"), _chk(noise, f32, "noise"), _chk(coef_table, f32, "coef")
    z = x.shape[-1]
    rows = x.numel() // z
    ldx = 0 if x_bf16 is None else x_bf16.shape[-1]
    check(lib.dn_ddpm_step(_p(x), _p(eps_hat), eps_hat.shape[-1], _p(noise), _p(coef_table), _p(t_idx), rows, z,
                           _p(x_bf16), ldx, x_lo_col, _stream()), "dn_ddpm_step")
    return x


def advance_step(t_idx, delta: int):
    check(lib.dn_advance_step(_p(t_idx), delta, _stream()), "dn_advance_step")


def adarmsnorm(x, out, B: int, T: int, gamma_p=None, gb=None, gb_t_stride: int = 0, t_idx=None, t_idx_stride: int = 0,
               split: bool = False):
    """out bf16 or fp16 [rows, C]; split: out is a split-precision bf16 operand [rows, 2C] = [hi | lo]."""
    _chk(x, f32, "x"), _chk(out, bf16 if split or out.dtype != f16 else f16, "out")
    Cc = x.shape[-1]
    if out.shape[-1] != (2 * Cc if split else Cc):
        raise ValueError("adarmsnorm: out width must be C (or 2C for a split-precision operand)")
    check(lib.dn_adarmsnorm(_p(x), _p(out), B, T, Cc, _p(gamma_p), _p(gb), gb_t_stride, _p(t_idx), t_idx_stride,
                            Cc if split else 0, _lib.FMT_F16 if out.dtype == f16 else _lib.FMT_BF16, _stream()), "dn_adarmsnorm")
    return out


def wavenet_gate(u, res, y, B: int, T: int, gb=None, gb_t_stride: int = 0, t_idx=None, t_idx_stride: int = 0):
    _chk(u, bf16, "u"), _chk(res, bf16, "res"), _chk(y, bf16, "y")
    check(lib.dn_wavenet_gate(_p(u), _p(res), _p(y), B, T, u.shape[-1], _p(gb), gb_t_stride, _p(t_idx), t_idx_stride,
                              _stream()), "dn_wavenet_gate")
    return y


def linear_f32(inp, W, bias=None, act: int = 0, out=None):
    _chk(inp, f32, "in"), _chk(W, f32, "W")
    M, K = inp.shape
    N = W.shape[0]
    assert W.shape[1] == K
    if out is None:
        out = torch.empty(M, N, dtype=f32, device=inp.device)
    check(lib.dn_linear_f32(_p(inp), _p(W), _p(bias), _p(out), M, N, K, act, _stream()), "dn_linear_f32")
    return out


def time_features(steps, w):
    _chk(steps, i32, "steps"), _chk(w, f32, "w")
    M, half = steps.shape[0], w.shape[0]
    out = torch.empty(M, 2 * half + 1, dtype=f32, device=w.device)
    check(lib.dn_time_features(_p(steps), _p(w), M, half, _p(out), _stream()), "dn_time_features")
    return out


def attention(qkv, out, lengths, B: int, T: int, H: int, dh: int, out_split: bool = False):
    """qkv bf16 or fp16 [B*T, 3*H*dh]; out [B*T, H*dh] in qkv's format, or (dh 96) the split-precision bf16 pair
    [B*T, 2*H*dh]."""
    if qkv.dtype not in (bf16, f16):
        raise ValueError("qkv must be bf16 or fp16")
    _chk(qkv, qkv.dtype, "qkv"), _chk(out, bf16 if out_split else qkv.dtype, "out")
    if out.shape[-1] != (2 if out_split else 1) * H * dh:
        raise ValueError("attention: out width")
    check(lib.dn_attention(_p(qkv), _p(out), _p(lengths), B, T, H, dh, _lib.FMT_F16 if qkv.dtype == f16 else _lib.FMT_BF16,
                           H * dh if out_split else 0, _stream()), "dn_attention")
    return out


# ------------------------------------------------------------------------------------------------ GEMM
import os as _os

_USE_2CTA = _os.environ.get("DN_GEMM_2CTA", "1") != "0"   # A/B switch for measurements; results are bit-identical
_WASTE_AWARE = _os.environ.get("DN_GEMM_WASTE_AWARE", "1") != "0"
# packed rows: M tiles of the per-utterance launches (shifted taps, per-utterance epilogue inputs) are filled with chunks of
# this many frames across utterance boundaries (0 = every tile stays inside one utterance); results are bit-identical
_ROW_CHUNK = int(_os.environ.get("DN_ROW_CHUNK", "32"))


class GemmPlan:
    """A packed weight (bf16 [w_rows, ldw], 256 rows per N tile) + its K-segment program + epilogue vectors.
    Built once per layer by diffnorm_b200.packing; `run` fills a dn_gemm_desc and calls dn_gemm."""

    def __init__(self, W, segs: Sequence[Sequence[int]], n_out: int, n_tiles: int, epi: int, bias=None, bias2=None,
                 groups: int = 1, g_w_row: int = 0, g_bias: int = 0, dilation: int = 1, dilation_shl_group: int = 0,
                 name: str = "", fmt: str = "bf16"):
        """fmt: operand format of this plan.  "bf16" / "f16": W is that 16-bit type, A is bf16.  "split": W = [W_hi | W_lo]
        (bf16 pairs side by side along K) and A = [A_hi | A_lo] (two halves of the row): every K segment runs three times
        — (A_hi, W_hi), (A_hi, W_lo), (A_lo, W_hi) — which contracts to ~2^-17 relative instead of 2^-9."""
        if fmt not in ("bf16", "bf16sr", "f16", "split"):
            raise ValueError(fmt)
        self.W32 = None     # "bf16sr": the fp32 packed master the engine re-rounds stochastically every sampler step
        if fmt == "bf16sr":
            fmt = "bf16"
        self.fmt = fmt
        self.W, self.segs, self.n_out, self.n_tiles, self.epi = W, [tuple(s) for s in segs], n_out, n_tiles, epi
        if fmt == "split" and 3 * len(self.segs) > _lib.MAX_SEGS:
            raise ValueError("too many K segments for a split-precision plan")
        self.bias, self.bias2, self.groups, self.g_w_row, self.g_bias = bias, bias2, groups, g_w_row, g_bias
        self.dilation, self.dilation_shl_group, self.name = dilation, dilation_shl_group, name
        self._flat_ok = groups == 1 and all(sg[1] == 0 for sg in self.segs)

    def to(self, device):
        self.W = self.W.to(device)
        self.W32 = None if self.W32 is None else self.W32.to(device)
        self.bias = None if self.bias is None else self.bias.to(device)
        self.bias2 = None if self.bias2 is None else self.bias2.to(device)
        return self

    def run(self, A, out, B: int, T: int, *, g_a_col: int = 0, g_out_col: int = 0, gb=None, gb_t_stride: int = 0,
            g_gb: int = 0, gb_half: int = 0, t_idx=None, t_idx_stride: int = 0, pe=None, lengths=None,
            epi: Optional[int] = None, impl: Optional[int] = None, a_cols: Optional[int] = None, out_split: bool = False,
            out_f16: bool = False, ddim=None, argmax_classes: Optional[int] = None, row_chunk: Optional[int] = None):
        """out_split: 16-bit output written as a split-precision pair (out is [.., 2 * width]: hi | lo).  out_f16: fp16 output."""
        epi = self.epi if epi is None else epi
        if ddim is not None:
            # (coef_rows fp32 [steps, 8], t_idx int32 device, staging bf16 [M, ld], lo_col): `out` is the fp32 latent, updated
            # in place by the DDIM step that consumes this GEMM's eps_hat in registers (DN_EPI_DDIM)
            epi = _lib.EPI_DDIM
            coef, t_idx, aux, aux_lo = ddim
            _chk(coef, f32, "coef"), _chk(aux, bf16, "staging")
        if argmax_classes is not None:
            epi = _lib.EPI_ARGMAX      # `out` = fp32 partials [M, 4 * n_tiles]; the logits themselves are never stored
            if out.shape[-1] != 4 * self.n_tiles:
                raise ValueError("argmax epilogue: out must be [rows, 4 * n_tiles]")
        if self._flat_ok and gb is None and pe is None:
            # no frame shifts and no per-utterance epilogue inputs: treat the batch as one long utterance so M tiles
            # run across utterance boundaries (T = 1000 would otherwise waste 24 of every 1024 tile rows)
            B, T = 1, B * T
        if row_chunk is None:
            row_chunk = _ROW_CHUNK if (B > 1 and T % 256) else 0
        if impl is None:
            # CTA-pair form (tcgen05 cta_group::2, M = 256 tiles) whenever it still fills the machine: 74 pairs of SMs ...
            pair_tiles = self.groups * B * ((T + 255) // 256) * self.n_tiles
            impl = _lib.GEMM_TCGEN05_2CTA if (_USE_2CTA and pair_tiles >= 74) else _lib.GEMM_TCGEN05
            # ... unless its 256-row tiles, which cannot cross utterances when taps are shifted, pad a ragged T by more than
            # the pair form gains (~10 % per row): T = 600 is 3 x 256 = 768 rows as pairs but 5 x 128 = 640 rows single
            if impl == _lib.GEMM_TCGEN05_2CTA and _WASTE_AWARE and not row_chunk and (T + 127) // 128 * 128 * 1.10 < (T + 255) // 256 * 256:
                impl = _lib.GEMM_TCGEN05
        wdt = f16 if self.fmt == "f16" else bf16
        _chk(A, wdt, "A")     # one 16-bit format per MMA: fp16 plans take fp16 activations
        f32out = epi in (_lib.EPI_F32, _lib.EPI_RESID, _lib.EPI_DDIM, _lib.EPI_ARGMAX)
        out_f16 = out_f16 or (not f32out and not out_split and out.dtype == f16)
        _chk(out, f32 if f32out else (f16 if out_f16 else bf16), "out")
        _chk(self.W, wdt, "W")
        d = GemmDesc()
        d.B, d.T, d.groups = B, T, self.groups
        lda = A.shape[-1]
        d.A, d.lda, d.a_cols, d.a_batch_stride, d.g_a_col = _p(A), lda, (lda if a_cols is None else a_cols), T * lda, g_a_col
        d.W, d.ldw, d.w_rows, d.g_w_row = _p(self.W), self.W.shape[1], self.W.shape[0], self.g_w_row
        segs = self.segs
        if self.fmt == "split":
            a_lo, w_lo = lda // 2, self.W.shape[1] // 2
            segs = [sg for (a0, sh, kb, w0, nm) in self.segs
                    for sg in ((a0, sh, kb, w0, nm), (a0, sh, kb, w0 + w_lo, nm), (a0 + a_lo, sh, kb, w0, nm))]
        d.num_segs = len(segs)
        for i, s in enumerate(segs):
            d.seg[i] = GemmSeg(*s)
        d.a_fmt = d.w_fmt = _lib.FMT_F16 if self.fmt == "f16" else _lib.FMT_BF16
        d.out_fmt = _lib.FMT_F16 if out_f16 else _lib.FMT_BF16
        d.out_lo_col = out.shape[-1] // 2 if out_split else 0
        d.dilation, d.dilation_shl_group = self.dilation, self.dilation_shl_group
        d.n_tiles, d.n_out, d.epi = self.n_tiles, self.n_out, epi
        d.bias, d.bias2, d.g_bias = _p(self.bias), _p(self.bias2), self.g_bias
        d.gb, d.gb_t_stride, d.g_gb, d.gb_half = _p(gb), gb_t_stride, g_gb, gb_half
        d.t_idx, d.t_idx_stride = _p(t_idx), t_idx_stride
        ldo = out.shape[-1]
        d.out, d.ldo, d.out_batch_stride, d.g_out_col = _p(out), ldo, T * ldo, g_out_col
        d.pe, d.lengths = _p(pe), _p(lengths)
        d.n_classes = 0 if argmax_classes is None else int(argmax_classes)
        d.row_chunk = int(row_chunk)
        if ddim is not None:
            d.coef, d.aux, d.aux_ld, d.aux_lo_col = _p(coef), _p(aux), aux.shape[-1], aux_lo
            if out.shape[-1] != self.n_out:
                raise ValueError("DDIM epilogue: the latent width must equal the plan's output width (z % 16 == 0)")
        check(lib.dn_gemm(C.byref(d), impl, _stream()), f"dn_gemm[{self.name}]")
        return out


def gemm_resid_norm(plan: GemmPlan, A, x, hb, gamma_p=None, gb=None, gb_t_stride: int = 0, t_idx=None):
    """x += A W^T + bias and hb = adaptive-RMSNorm(x) in one kernel (dn_gemm_resid_norm): `plan` is a plain
    EPI_RESID linear of width 512; gb / t_idx select ONE table row for the whole batch (shared timestep)."""
    _chk(A, bf16, "A"), _chk(x, f32, "x"), _chk(hb, bf16, "hb")
    assert plan.n_out == 512 and x.shape[-1] == 512 and hb.shape[-1] == 512 and plan._flat_ok and len(plan.segs) == 1
    assert plan.fmt == "bf16", "dn_gemm_resid_norm multiplies bf16 weights"
    d = _lib.ResidNormDesc()
    d.M, d.A, d.lda, d.k_blocks = x.shape[0], _p(A), A.shape[-1], plan.segs[0][2]
    d.W, d.ldw, d.bias, d.x, d.hb = _p(plan.W), plan.W.shape[1], _p(plan.bias), _p(x), _p(hb)
    d.gamma_p, d.gb, d.gb_t_stride, d.t_idx = _p(gamma_p), _p(gb), gb_t_stride, _p(t_idx)
    check(lib.dn_gemm_resid_norm(C.byref(d), _stream()), f"dn_gemm_resid_norm[{plan.name}]")
    return x


# ------------------------------------------------------------------------------------------------ training step
def wgrad(dY, X, dW, B: int, T: int, n_rows: int, k_cols: int, dy_col0: int = 0, x_col0: int = 0, shift: int = 0,
          splits: int = 0, groups: int = 1, g_dy_col: int = 0, g_x_col: int = 0, shift_shl_group: bool = False):
    """dW[n, c] += sum_{b,t} dY[b,t,dy_col0+n] * X[b,t-shift,x_col0+c].  dY, X bf16 [B*T, ld]; dW fp32 [>=n_rows, ldw]
    (groups > 1: dW [groups, n_rows, ldw], group g at column offsets g*g_dy_col / g*g_x_col, shift << g if asked)."""
    _chk(dY, bf16, "dY"), _chk(X, bf16, "X"), _chk(dW, f32, "dW")
    d = _lib.WgradDesc()
    d.B, d.T = B, T
    d.dY, d.ldy, d.dy_batch_stride, d.dy_col0 = _p(dY), dY.shape[-1], T * dY.shape[-1], dy_col0
    d.X, d.ldx, d.x_batch_stride, d.x_col0, d.x_shift = _p(X), X.shape[-1], T * X.shape[-1], x_col0, shift
    d.n_rows, d.k_cols, d.dW, d.ldw, d.splits = n_rows, k_cols, _p(dW), dW.shape[-1], splits
    d.groups, d.g_dy_col, d.g_x_col, d.shift_shl_group = groups, g_dy_col, g_x_col, int(shift_shl_group)
    d.g_dw_stride = dW.stride(0) if groups > 1 else 0
    check(lib.dn_wgrad(C.byref(d), _stream()), "dn_wgrad")
    return dW


def colsum(src, col0: int, cols: int, out):
    _chk(src, bf16, "src"), _chk(out, f32, "out")
    ld = src.shape[-1]
    check(lib.dn_colsum_bf16(_p(src), src.numel() // ld, ld, col0, cols, _p(out), _stream()), "dn_colsum_bf16")
    return out


def geglu_fwd(h, m):
    _chk(h, bf16, "h"), _chk(m, bf16, "m")
    check(lib.dn_geglu_fwd(_p(h), m.shape[0], m.shape[1], _p(m), _stream()), "dn_geglu_fwd")
    return m


def geglu_bwd(h, dm, dh):
    _chk(h, bf16, "h"), _chk(dm, bf16, "dm"), _chk(dh, bf16, "dh")
    check(lib.dn_geglu_bwd(_p(h), _p(dm), dm.shape[0], dm.shape[1], _p(dh), _stream()), "dn_geglu_bwd")
    return dh


def wn_gate_fwd(ur, y, B, T, Cc, G, gb=None, gb_t_stride=0, g_gb=0, t_idx=None, t_idx_stride=0):
    _chk(ur, bf16, "ur"), _chk(y, bf16, "y")
    check(lib.dn_wn_gate_fwd(_p(ur), _p(y), B, T, Cc, G, _p(gb), gb_t_stride, g_gb, _p(t_idx), t_idx_stride, _stream()),
          "dn_wn_gate_fwd")
    return y


def wn_gate_bwd(ur, dy, dur, B, T, Cc, G, gb=None, gb_t_stride=0, g_gb=0, t_idx=None, t_idx_stride=0, dgb=None,
                dgb_b_stride=0, g_dgb=0):
    _chk(ur, bf16, "ur"), _chk(dy, bf16, "dy"), _chk(dur, bf16, "dur")
    check(lib.dn_wn_gate_bwd(_p(ur), _p(dy), _p(dur), B, T, Cc, G, _p(gb), gb_t_stride, g_gb, _p(t_idx), t_idx_stride,
                             _p(dgb), dgb_b_stride, g_dgb, _stream()), "dn_wn_gate_bwd")
    return dur


def adarmsnorm_bwd(x, dy, dx, dx_bf16, B, T, gamma_p=None, dgamma_p=None, gb=None, gb_t_stride=0, t_idx=None,
                   t_idx_stride=0, dgb=None, dgb_b_stride=0):
    _chk(x, f32, "x"), _chk(dy, bf16, "dy"), _chk(dx, f32, "dx")
    check(lib.dn_adarmsnorm_bwd(_p(x), _p(dy), _p(dx), _p(dx_bf16), B, T, x.shape[-1], _p(gamma_p), _p(dgamma_p), _p(gb),
                                gb_t_stride, _p(t_idx), t_idx_stride, _p(dgb), dgb_b_stride, _stream()), "dn_adarmsnorm_bwd")
    return dx


def train_noise(z_lat, eps0, eps, beta0: float, coef, t_idx, x_t, xb):
    B, T, z = z_lat.shape
    check(lib.dn_train_noise(_p(_chk(z_lat, f32, "z")), _p(_chk(eps0, f32, "eps0")), _p(_chk(eps, f32, "eps")), beta0,
                             _p(coef), _p(t_idx), B, T, z, _p(x_t), _p(xb), xb.shape[-1], _stream()), "dn_train_noise")
    return x_t


def noise_loss(pred, eps, lengths, coef, t_idx, B, T, z, loss, dpred=None, grad_scale: float = 1.0, dx1=None):
    _chk(pred, f32, "pred"), _chk(eps, f32, "eps"), _chk(loss, f32, "loss")
    check(lib.dn_noise_loss(_p(pred), pred.shape[-1], _p(eps), _p(lengths), _p(coef), _p(t_idx), B, T, z, _p(loss),
                            _p(dpred), z if dpred is None else dpred.shape[-1], grad_scale, _p(dx1),
                            0 if dx1 is None else dx1.shape[-1], _stream()), "dn_noise_loss")
    return loss


def lsnll_bwd(logits, vocab: int, units, stats, eps_ls: float, nll_scale: float, dlogits):
    _chk(logits, f32, "logits"), _chk(units, i64, "units"), _chk(dlogits, bf16, "dlogits")
    check(lib.dn_lsnll_bwd(_p(logits), logits.shape[-1], vocab, _p(units), units.numel(), _p(stats), eps_ls, nll_scale,
                           _p(dlogits), dlogits.shape[-1], _stream()), "dn_lsnll_bwd")
    return dlogits


def recon_grad(recon, audio, d_lm, lengths, B, T, stats, mse_scale: float, out):
    _chk(recon, f32, "recon"), _chk(audio, f32, "audio"), _chk(d_lm, f32, "d_lm"), _chk(out, bf16, "out")
    check(lib.dn_recon_grad(_p(recon), _p(audio), _p(d_lm), _p(lengths), B, T, recon.shape[-1], _p(stats), mse_scale,
                            _p(out), _stream()), "dn_recon_grad")
    return out


def pred_x1(x_t, pred, coef, t_idx, B, T, z, xb):
    check(lib.dn_pred_x1(_p(x_t), _p(pred), pred.shape[-1], _p(coef), _p(t_idx), B, T, z, _p(xb), xb.shape[-1], _stream()),
          "dn_pred_x1")
    return xb


def decode_losses(recon, audio, logits, vocab: int, units, lengths, B, T):
    _chk(recon, f32, "recon"), _chk(audio, f32, "audio"), _chk(logits, f32, "logits"), _chk(units, i64, "units")
    out = torch.empty(6, dtype=torch.float64, device=recon.device)
    check(lib.dn_decode_losses(_p(recon), _p(audio), recon.shape[-1], _p(logits), logits.shape[-1], vocab, _p(units),
                               _p(lengths), B, T, _p(out), _stream()), "dn_decode_losses")
    return out


def randn(shape, device, seed: int, offset: int) -> torch.Tensor:
    """fp32 N(0, 1) tensor drawn by the library's own Philox kernel."""
    out = torch.empty(*shape, dtype=f32, device=device)
    check(lib.dn_randn(_p(out), out.numel(), int(seed), int(offset), _stream()), "dn_randn")
    return out


def dropout_bits(bits, p: float, seed: int, offset: int):
    check(lib.dn_dropout_bits(_p(bits), bits.numel(), p, seed, offset, _stream()), "dn_dropout_bits")
    return bits


def attention_train(qkv, out, lse2, lengths, keep_bits, keep_scale: float, B, T, H, dh):
    _chk(qkv, bf16, "qkv"), _chk(out, bf16, "out"), _chk(lse2, f32, "lse2")
    check(lib.dn_attention_train(_p(qkv), _p(out), _p(lse2), _p(lengths), _p(keep_bits), keep_scale, B, T, H, dh, _stream()),
          "dn_attention_train")
    return out


def attention_bwd(qkv, out, dout, lse2, lengths, keep_bits, keep_scale: float, dqkv, delta_ws, B, T, H, dh):
    _chk(qkv, bf16, "qkv"), _chk(out, bf16, "out"), _chk(dout, bf16, "dout"), _chk(dqkv, bf16, "dqkv")
    check(lib.dn_attention_bwd(_p(qkv), _p(out), _p(dout), _p(lse2), _p(lengths), _p(keep_bits), keep_scale, _p(dqkv),
                               _p(delta_ws), B, T, H, dh, _stream()), "dn_attention_bwd")
    return dqkv


def silu(pre, out):
    check(lib.dn_silu(_p(pre), _p(out), pre.numel(), _stream()), "dn_silu")
    return out


def silu_bwd(pre, dout, dpre):
    check(lib.dn_silu_bwd(_p(pre), _p(dout), _p(dpre), pre.numel(), _stream()), "dn_silu_bwd")
    return dpre


def linear_f32_bwd(dY, X, W, dW=None, db=None, dX=None):
    """dY fp32 [M, N] (row stride dY.stride(0)); dW[N,K] += dY^T X, db[N] += sum_m dY; dX[M,K] += dY W."""
    M, N = dY.shape
    K = X.shape[1] if X is not None else W.shape[1]
    check(lib.dn_linear_f32_bwd(_p(dY), dY.stride(0), _p(X), _p(W), M, N, K, _p(dW), _p(db), _p(dX), _stream()),
          "dn_linear_f32_bwd")


def time_features_bwd(steps, w, dfeat, dw):
    check(lib.dn_time_features_bwd(_p(steps), _p(w), _p(dfeat), steps.shape[0], w.shape[0], _p(dw), _stream()),
          "dn_time_features_bwd")
    return dw


def add_bf16_to_f32(src, col0: int, Cc: int, dst, accumulate: bool):
    _chk(src, bf16, "src"), _chk(dst, f32, "dst")
    ld = src.shape[-1]
    check(lib.dn_add_bf16_to_f32(_p(src), src.numel() // ld, ld, col0, Cc, _p(dst), dst.shape[-1], int(accumulate), _stream()),
          "dn_add_bf16_to_f32")
    return dst


def vae_kl(params, lengths, z: int, kl):
    _chk(params, f32, "params"), _chk(kl, f32, "kl")
    B, T, ldp = params.shape
    check(lib.dn_vae_kl(_p(params), ldp, _p(lengths), B, T, z, _p(kl), _stream()), "dn_vae_kl")
    return kl


def vae_reparam_bwd(params, eps, eps_channel_first: bool, dz, lengths, z: int, kl_scale: float, dparams):
    _chk(params, f32, "params"), _chk(eps, f32, "eps"), _chk(dz, bf16, "dz"), _chk(dparams, bf16, "dparams")
    B, T, ldp = params.shape
    check(lib.dn_vae_reparam_bwd(_p(params), ldp, _p(eps), int(eps_channel_first), _p(dz), dz.shape[-1], _p(lengths), B, T, z,
                                 kl_scale, _p(dparams), dparams.shape[-1], _stream()), "dn_vae_reparam_bwd")
    return dparams


def split_bf16x3(src, out=None):
    """fp32 [rows, C] -> bf16 [rows, 3C] = [hi | hi | lo]."""
    _chk(src, f32, "src")
    rows, Cc = src.shape
    if out is None:
        out = torch.empty(rows, 3 * Cc, dtype=bf16, device=src.device)
    check(lib.dn_split_bf16x3(_p(src), rows, Cc, _p(out), _stream()), "dn_split_bf16x3")
    return out
