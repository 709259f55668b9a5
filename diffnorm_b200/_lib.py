"""ctypes binding of libdiffnorm_b200.so (the C ABI declared in include/diffnorm_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly with build instructions, and every
wrapper raises on a non-zero status.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C diffnorm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DN_LIB") or os.path.join(_HERE, "csrc", "libdiffnorm_b200.so")   # DN_LIB: A/B experiment builds


class DiffNormLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise DiffNormLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU / torch fallback exists). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C diffnorm_b200/csrc`.")
    return C.CDLL(LIB_PATH)


lib = _load()

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmSeg(C.Structure):
    _fields_ = [("a_col0", i32), ("shift_mul", i32), ("k_blocks", i32), ("w_k0", i32), ("n_mma", i32)]


MAX_SEGS = 12   # DN_MAX_SEGS


class GemmDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("T", i32), ("groups", i32),
        ("A", vp), ("lda", i32), ("a_cols", i32), ("a_batch_stride", i64), ("g_a_col", i32),
        ("W", vp), ("ldw", i32), ("w_rows", i32), ("g_w_row", i32),
        ("num_segs", i32), ("seg", GemmSeg * MAX_SEGS),
        ("dilation", i32), ("dilation_shl_group", i32), ("n_tiles", i32), ("n_out", i32), ("epi", i32),
        ("bias", vp), ("bias2", vp), ("g_bias", i32),
        ("gb", vp), ("gb_t_stride", i64), ("g_gb", i32), ("gb_half", i32), ("t_idx", vp), ("t_idx_stride", i32),
        ("out", vp), ("ldo", i32), ("out_batch_stride", i64), ("g_out_col", i32),
        ("pe", vp), ("lengths", vp),
        ("a_fmt", i32), ("w_fmt", i32), ("out_fmt", i32), ("out_lo_col", i32),
        ("coef", vp), ("aux", vp), ("aux_ld", i32), ("aux_lo_col", i32), ("n_classes", i32), ("row_chunk", i32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("T", i32),
        ("dY", vp), ("ldy", i32), ("dy_batch_stride", i64), ("dy_col0", i32),
        ("X", vp), ("ldx", i32), ("x_batch_stride", i64), ("x_col0", i32), ("x_shift", i32),
        ("n_rows", i32), ("k_cols", i32),
        ("dW", vp), ("ldw", i32), ("splits", i32),
        ("groups", i32), ("g_dy_col", i32), ("g_x_col", i32), ("shift_shl_group", i32), ("g_dw_stride", i64),
    ]


class PackOp(C.Structure):
    _fields_ = [
        ("src", vp), ("dst", vp), ("s_row", i64), ("s_col", i64), ("s_tap", i64),
        ("rows", i32), ("cols", i32), ("taps", i32), ("ldd", i32),
        ("row0", i32), ("rblk", i32), ("rblk_stride", i32),
        ("col0", i32), ("cblk", i32), ("cblk_stride", i32), ("tap_cols", i32),
        ("tap_pos", i32 * 3), ("src_r_fastest", i32), ("out_f32", i32), ("tile0", i32), ("tiles_c", i32),
    ]


class ResidNormDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("A", vp), ("lda", i32), ("k_blocks", i32), ("W", vp), ("ldw", i32), ("bias", vp), ("x", vp), ("hb", vp),
        ("gamma_p", vp), ("gb", vp), ("gb_t_stride", i64), ("t_idx", vp),
    ]


EPI_BF16, EPI_F32, EPI_RESID, EPI_GEGLU, EPI_WN_GATE, EPI_DDIM, EPI_ARGMAX = range(7)
GEMM_TCGEN05, GEMM_SIMT_CHECK, GEMM_TCGEN05_2CTA = 0, 1, 2
FMT_BF16, FMT_F16 = 0, 1
ABI_VERSION = 2

# name -> argtypes ; every function returns int status except the two info calls
_SIGS = {
    "dn_set_sm_limit": [i32],
    "dn_reduce_tgt": [vp, vp, i32, i32, vp, vp, vp, vp, vp],
    "dn_argmax_units": [vp, i32, i64, i32, i32, i32, vp, vp],
    "dn_argmax_combine": [vp, i64, i32, i32, vp, vp],
    "dn_unit_accuracy": [vp, vp, vp, i32, i32, vp, vp],
    "dn_gather_pack": [vp, vp, vp, vp, i32, i32, i32, vp, i32, i32, vp],
    "dn_cast_pad_bf16": [vp, i64, i32, i32, vp, i32, vp],
    "dn_cast_split": [vp, i64, i32, i32, vp, i32, i32, vp],
    "dn_vae_reparam": [vp, i32, vp, i32, i32, i32, i32, vp, vp],
    "dn_split_bf16x3": [vp, i64, i32, vp, vp],
    "dn_sround_bf16": [vp, vp, i64, C.c_uint32, vp, vp],
    "dn_q_sample": [vp, vp, f32, f32, i64, i32, vp, vp, i32, i32, vp],
    "dn_ddim_step": [vp, vp, i32, vp, vp, i64, i32, i32, vp, i32, i32, vp],
    "dn_ddpm_step": [vp, vp, i32, vp, vp, vp, i64, i32, vp, i32, i32, vp],
    "dn_advance_step": [vp, i32, vp],
    "dn_adarmsnorm": [vp, vp, i32, i32, i32, vp, vp, i64, vp, i32, i32, i32, vp],
    "dn_wavenet_gate": [vp, vp, vp, i32, i32, i32, vp, i64, vp, i32, vp],
    "dn_linear_f32": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "dn_time_features": [vp, vp, i32, i32, vp, vp],
    "dn_gemm": [C.POINTER(GemmDesc), i32, vp],
    "dn_gemm_tile_rows": [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],   # host-only: the kernel's M tiling
    "dn_attention": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    # unit vocoder (the step after the pass)
    "dn_voc_conv1d": [vp, i32, i32, vp, vp, i32, i32, i32, i32, f32, i32, vp, f32, i32, vp, vp],
    "dn_voc_conv_transpose1d": [vp, i32, i32, vp, vp, i32, i32, i32, i32, f32, vp, vp],
    "dn_voc_layernorm": [vp, i32, i32, vp, vp, vp, vp],
    "dn_voc_durations": [vp, i32, vp, vp, vp],
    "dn_voc_embed_repeat": [vp, i32, vp, i32, vp, i32, vp, vp],
    # training step
    "dn_wgrad": [C.POINTER(WgradDesc), vp],
    "dn_gemm_resid_norm": [C.POINTER(ResidNormDesc), vp],
    "dn_pack_weights": [vp, i32, i32, vp],
    "dn_colsum_bf16": [vp, i64, i32, i32, i32, vp, vp],
    "dn_geglu_fwd": [vp, i64, i32, vp, vp],
    "dn_geglu_bwd": [vp, vp, i64, i32, vp, vp],
    "dn_wn_gate_fwd": [vp, vp, i32, i32, i32, i32, vp, i64, i32, vp, i32, vp],
    "dn_wn_gate_bwd": [vp, vp, vp, i32, i32, i32, i32, vp, i64, i32, vp, i32, vp, i64, i32, vp],
    "dn_adarmsnorm_bwd": [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i64, vp, i32, vp, i64, vp],
    "dn_train_noise": [vp, vp, vp, f32, vp, vp, i32, i32, i32, vp, vp, i32, vp],
    "dn_noise_loss": [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, f32, vp, i32, vp],
    "dn_vae_kl": [vp, i32, vp, i32, i32, i32, vp, vp],
    "dn_vae_reparam_bwd": [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, f32, vp, i32, vp],
    "dn_lsnll_bwd": [vp, i32, i32, vp, i64, vp, f32, f32, vp, i32, vp],
    "dn_recon_grad": [vp, vp, vp, vp, i32, i32, i32, vp, f32, vp, vp],
    "dn_pred_x1": [vp, vp, i32, vp, vp, i32, i32, i32, vp, i32, vp],
    "dn_decode_losses": [vp, vp, i32, vp, i32, i32, vp, vp, i32, i32, vp, vp],
    "dn_randn": [vp, i64, C.c_uint64, C.c_uint64, vp],
    "dn_dropout_bits": [vp, i64, f32, C.c_uint64, C.c_uint64, vp],
    "dn_attention_train": [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp],
    "dn_attention_bwd": [vp, vp, vp, vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, vp],
    "dn_silu": [vp, vp, i64, vp],
    "dn_silu_bwd": [vp, vp, vp, i64, vp],
    "dn_linear_f32_bwd": [vp, i64, vp, vp, i32, i64, i32, vp, vp, vp, vp],
    "dn_time_features_bwd": [vp, vp, vp, i32, i32, vp, vp],
    "dn_add_bf16_to_f32": [vp, i64, i32, i32, i32, vp, i32, i32, vp],
}
EXPORTS = sorted(list(_SIGS) + ["dn_abi_version", "dn_launch_count", "dn_batch_by_size"])

for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.dn_abi_version.restype = C.c_int
lib.dn_abi_version.argtypes = []
lib.dn_launch_count.restype = C.c_ulonglong
lib.dn_launch_count.argtypes = []
lib.dn_batch_by_size.restype = C.c_int64
lib.dn_batch_by_size.argtypes = [vp, i64, i64, i64, i32, vp]


def check(status: int, what: str):
    if status != 0:
        kind = "argument error" if status < 0 else "cudaError"
        raise DiffNormLibraryError(f"{what} failed: {kind} {status}")


def launch_count() -> int:
    return int(lib.dn_launch_count())
