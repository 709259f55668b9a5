/* diffnorm_b200 — C ABI of the sm_100a kernels behind DiffNorm's latent-diffusion normalization pass.
 *
 * One shared library (diffnorm_b200/csrc/libdiffnorm_b200.so), loaded with ctypes by the Python host layer
 * (diffnorm_b200/_lib.py).  Conventions for EVERY entry point:
 *   - raw device pointers + plain integer sizes; no torch / C++ types cross the boundary;
 *   - asynchronous w.r.t. the host: work is enqueued on `stream` (a cudaStream_t passed as void*), the caller
 *     synchronises; no allocation, no ownership transfer, no global state (except a launch counter);
 *   - returns 0 on success, a positive cudaError_t, or a negative DN_E* argument error;
 *   - activations are row-major [B, T, C] ("frames x channels"), C contiguous; bf16 where noted.
 * "LM" below = fairseq/models/text_to_speech/latent_module.py of the reference; every function cites the
 * reference code it replaces.
 */
#ifndef DIFFNORM_B200_H
#define DIFFNORM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_OK 0
#define DN_EINVAL (-1)   /* bad argument (null pointer, misaligned, unsupported size) */
#define DN_EDRIVER (-2)  /* driver entry point (cuTensorMapEncodeTiled) unavailable */

/* ---- library info ------------------------------------------------------------------------------------ */
int dn_abi_version(void);                 /* = 2 */
unsigned long long dn_launch_count(void); /* kernels launched by this library so far (bench.py gpu_launches) */
/* The persistent kernels (dn_gemm, dn_wgrad) size their grids to the SM count and stride tiles statically over their CTAs:
 * a concurrent kernel that takes SMs away (NCCL's all-reduce of the gradient buckets during the training step's backward)
 * delays whole CTAs and with them the launch.  n > 0 (even) caps the grids at n SMs so that the rest stays free for the
 * collective; 0 restores all SMs.  Process-wide; takes effect for launches enqueued afterwards. */
int dn_set_sm_limit(int32_t n);

/* ---- integer kernels --------------------------------------------------------------------------------- */

/* Run-length reduction of unit streams, batched over utterances.
 * Replaces research/TranSpeech/diff_norm_synthesis.py:25-46 (reduce_token) ==
 * fairseq/data/audio/repr_to_repr_unit_dataset.py:92-113 (_reduce_tgt), called twice per utterance
 * (diff_norm_synthesis.py:150 and :216).
 * units [B, T] int64 (row stride T), lengths [B] int32 valid tokens per row.
 * Outputs (row stride T): dedup[b, j], duration[b, j], index_to_keep[b, j] for j < counts[b].
 * lengths[b] == 0 gives counts[b] = 0 and duration[b,0] = 1 (the reference's quirk, :45), needs T >= 1. */
int dn_reduce_tgt(const int64_t* units, const int32_t* lengths, int32_t B, int32_t T, int64_t* dedup,
                  int64_t* duration, int64_t* index_to_keep, int32_t* counts, void* stream);

/* units[r] = argmax_c logits[r, c] - offset, first maximum wins (torch.argmax tie rule), NaN counts as max.
 * Replaces LM:1450-1451.  logits [rows, ld] fp32 (logits_bf16 = 0) or bf16 (= 1), C classes (C <= ld). */
int dn_argmax_units(const void* logits, int32_t logits_bf16, int64_t rows, int32_t C, int32_t ld, int32_t offset,
                    int64_t* units, void* stream);

/* Second stage of DN_EPI_ARGMAX: units[r] = index of the best of the row's `parts` (value, index) partials - offset
 * (same ordering as dn_argmax_units).  partials fp32 [rows, 2 * parts]. */
int dn_argmax_combine(const float* partials, int64_t rows, int32_t parts, int32_t offset, int64_t* units, void* stream);

/* match = #{valid (b,t): units == ref_units}, total = #valid.  Replaces LM:1453-1454 (two .item() syncs).
 * out2[0] = match, out2[1] = total (int64, device). */
int dn_unit_accuracy(const int64_t* units, const int64_t* ref_units, const int32_t* lengths, int32_t B, int32_t T,
                     int64_t* out2, void* stream);

/* Row gather + zero pad: dst[b, j, :] = src[src_row0[b] + index_to_keep[b, j], :] for j < counts[b], else 0.
 * Replaces diff_norm_synthesis.py:151,164-169.  src fp32 [*, C]; dst fp32 (dst_bf16 = 0) or bf16 [B, T, ldd]
 * (columns C..ldd are zero-filled). */
int dn_gather_pack(const float* src, const int64_t* src_row0, const int64_t* index_to_keep, const int32_t* counts,
                   int32_t B, int32_t T, int32_t C, void* dst, int32_t ldd, int32_t dst_bf16, void* stream);

/* ---- fused elementwise kernels (HBM-bound) ------------------------------------------------------------- */

/* fp32 [rows, C] -> bf16 [rows, ldo] with zero fill of the pad columns (operand staging for the GEMMs). */
int dn_cast_pad_bf16(const float* src, int64_t rows, int32_t C, int32_t lds, void* dst, int32_t ldo, void* stream);

/* Split-precision operand staging: fp32 [rows, C] (row stride lds) -> bf16 [rows, ldo] with hi = bf16(x) in columns
 * [0, C) and lo = bf16(x - hi) in columns [lo_col, lo_col + C); every other column is written with zeros.  A GEMM whose
 * K-segment program multiplies (hi, W_hi), (hi, W_lo) and (lo, W_hi) then contracts to ~2^-17 relative: the once-per-pass
 * VAE encoder / decoder run this way (their rounding error decides near-tie units; the 99-call loop averages its own).
 * lo_col = 0: plain cast (= dn_cast_pad_bf16). */
int dn_cast_split(const float* src, int64_t rows, int32_t C, int32_t lds, void* dst, int32_t ldo, int32_t lo_col,
                  void* stream);

/* Stochastic re-rounding of packed GEMM weights: dst[i] = bf16(src[i]) rounded up or down with probability proportional to
 * the distance (16 random bits from a counter hash of (i, seed, step[0]); deterministic given the seed; E[dst] = src).
 * A bf16 weight rounded ONCE carries the same error into all 99 denoiser calls of a pass, and that error accumulates
 * coherently in the latent; re-rounding the weights every sampler step makes it independent from step to step,
 * so it averages out like the activations' rounding does — at bf16's power draw instead of fp16's.  step is a DEVICE int32 (the
 * sampler's step counter) so the launch sits inside the captured step graph; n % 4 == 0, 16-byte aligned pointers. */
int dn_sround_bf16(const float* src, void* dst, int64_t n, uint32_t seed, const int32_t* step, void* stream);

/* fp32 [rows, C] -> bf16 [rows, 3C] = [hi | hi | lo] (hi = bf16(x), lo = bf16(x - hi)).  Against weights packed as
 * [hi | lo | hi] a single dn_gemm over K = 3C gives the contraction to ~2^-16 relative: used by the k-means unit
 * quantiser (examples/textless_nlp/gslm/speech2unit/clustering/quantize_with_kmeans.py:109-121). */
int dn_split_bf16x3(const float* src, int64_t rows, int32_t C, void* dst, void* stream);

/* VAE posterior reparameterisation.  Replaces distributions.py:24-41 + the caller's transpose LM:1397.
 * params [B, T, ldp] fp32 row-major with mean in columns [0,z) and logvar in [z,2z) (channel-last);
 * eps: eps_channel_first = 1 -> [B, z, T] (the reference's draw order), 0 -> [B, T, z].
 * z_out [B, T, z] fp32 = mean + exp(0.5 * clamp(logvar, -30, 20)) * eps. */
int dn_vae_reparam(const float* params, int32_t ldp, const float* eps, int32_t eps_channel_first, int32_t B,
                   int32_t T, int32_t z, float* z_out, void* stream);

/* Per-step coefficient table row layout (float32 x 8), built by the host from the float64 schedule:
 *   [0] sqrt_ab  [1] sqrt_1m_ab  [2] sqrt(ab_prev)  [3] sqrt(1-ab_prev)      (DDIM, LM:1419-1438)
 *   [4] sqrt_recip_ab  [5] sqrt_recipm1_ab  -> x0 = c4*x - c5*eps              (generic lib)
 *   [6] unused  [7] unused
 * DDPM rows: [0] sqrt_recip_ab [1] sqrt_recipm1_ab [2] coef1 [3] coef2 [4] exp(0.5*logvar) (0 when t == 0).
 * `t_idx` is a DEVICE int32 holding the current table row, so one captured CUDA graph serves every step.
 * The three update kernels move 4 latent channels per thread (16-byte accesses): z, lde and ldx must be multiples
 * of 4 and the fp32 pointers 16-byte aligned (z is 16, 32 or 128 on this path, LM:1044-1051), else DN_EINVAL.
 * x_lo_col != 0: the staging copy is split-precision: hi = bf16(x) in columns [0, z), lo = bf16(x - hi) in columns
 * [x_lo_col, x_lo_col + z) (the denoiser's first GEMM then sees the fp32 latent to ~2^-17). */

/* x = c0 * z + c1 * eps  (q_sample, LM:1405-1409).  Also writes the bf16 staging copy x_bf16 [n/z, ldx] if
 * non-null (zero padded to ldx columns). */
int dn_q_sample(const float* z_lat, const float* eps, float sqrt_ab, float sqrt_1m_ab, int64_t rows, int32_t z,
                float* x, void* x_bf16, int32_t ldx, int32_t x_lo_col, void* stream);

/* Inline DDIM eta=0 update (LM:1419-1442), in place on x [rows, z]; eps_hat [rows, lde] fp32.
 * mode 0: reference inline form with the 1e-10 clamps; mode 1: generic-lib form (gaussian_diffusion.py:513-560). */
int dn_ddim_step(float* x, const float* eps_hat, int32_t lde, const float* coef_table, const int32_t* t_idx,
                 int64_t rows, int32_t z, int32_t mode, void* x_bf16, int32_t ldx, int32_t x_lo_col, void* stream);

/* Ancestral DDPM update (gaussian_diffusion.py:232-252,295-344,402-417), in place; noise [rows, z] fp32. */
int dn_ddpm_step(float* x, const float* eps_hat, int32_t lde, const float* noise, const float* coef_table,
                 const int32_t* t_idx, int64_t rows, int32_t z, void* x_bf16, int32_t ldx, int32_t x_lo_col, void* stream);

/* t_idx[0] += delta (device-side step counter so the sampler loop needs no host round trip). */
int dn_advance_step(int32_t* t_idx, int32_t delta, void* stream);

/* (Adaptive) RMSNorm, LM:629-639:  out = x / max(||x||, 1e-12) * sqrt(C) * gamma_p  [* gamma_t + beta_t].
 * x fp32 [B*T, C]; out bf16 [B*T, C]; gamma_p [C] or null; gb (null = unconditioned) points at a table whose
 * row for utterance b is gb + t_idx[b * t_idx_stride] * gb_t_stride, holding gamma_t[0..C) then beta_t[C..2C).
 * out_lo_col = 0: out bf16 [B*T, C].  out_lo_col != 0 (>= C): out bf16 [B*T, 2 * out_lo_col], hi in columns [0, C), lo in
 * [out_lo_col, out_lo_col + C) (split-precision operand of the GEMM that follows).  out_fmt: DN_FMT_BF16 | DN_FMT_F16
 * (un-split outputs only). */
int dn_adarmsnorm(const float* x, void* out, int32_t B, int32_t T, int32_t C, const float* gamma_p, const float* gb,
                  int64_t gb_t_stride, const int32_t* t_idx, int32_t t_idx_stride, int32_t out_lo_col, int32_t out_fmt,
                  void* stream);

/* Residual GEMM fused with the adaptive RMSNorm that follows it in the transformer block (LM:692 -> :629-639 of the
 * next sub-layer; LM:704 -> the next layer's first norm / the final to_pred norm), hidden width C = 512 only:
 *     x[m, :] += A[m, :] W^T + bias ;   hb[m, :] = bf16(adaptive-RMSNorm(x[m, :]))        in one pass over x.
 * A bf16 [M, lda] (k_blocks*64 columns used), W bf16 [512, ldw] (K-major, as dn_gemm packs it), bias fp32 [512] or null,
 * x fp32 [M, 512] in/out, hb bf16 [M, 512] out.  gamma_p / gb as in dn_adarmsnorm, with ONE table row for the whole batch
 * (gb + t_idx[0] * gb_t_stride: the sampler's shared timestep); per-utterance timesteps use the un-fused pair. */
typedef struct {
    int32_t M;
    const void* A; int32_t lda; int32_t k_blocks;
    const void* W; int32_t ldw;
    const float* bias;
    float* x;
    void* hb;
    const float* gamma_p;
    const float* gb; int64_t gb_t_stride; const int32_t* t_idx;
} dn_resid_norm_desc;
int dn_gemm_resid_norm(const dn_resid_norm_desc* d, void* stream);

/* Stand-alone WaveNet gate (LM:524-533): y = tanh(u') * sigmoid(u') + res, u' = u * gamma + beta (gb optional).
 * u, res, y bf16 [B*T, C].  (The denoiser path uses the same math fused into dn_gemm's epilogue.) */
int dn_wavenet_gate(const void* u, const void* res, void* y, int32_t B, int32_t T, int32_t C, const float* gb,
                    int64_t gb_t_stride, const int32_t* t_idx, int32_t t_idx_stride, void* stream);

/* Small-M fp32 linear: out[m, n] = act(sum_k in[m, k] * W[n, k] + bias[n]); act 0 = none, 1 = SiLU.
 * Used for the time-conditioning MLP and the [T, 56, 1024] gamma/beta table (LM:741-745, :507, :624). */
int dn_linear_f32(const float* in, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                  int32_t act, void* stream);

/* Learned sinusoidal time features (LM:104-116): out[m, :] = [t, sin(t w 2pi), cos(t w 2pi)], t = steps[m]. */
int dn_time_features(const int32_t* steps, const float* w, int32_t M, int32_t half, float* out, void* stream);

/* ---- tensor-core GEMM / implicit causal convolution ----------------------------------------------------- */

/* One K-segment of the contraction: A columns [a_col0, a_col0 + 64*k_blocks) read at rows (t - shift_mul * d),
 * against W columns [w_k0, ...).  Rows with t - shift < 0 read as zero (causal left padding, LM:476-488). */
typedef struct {
    int32_t a_col0;
    int32_t shift_mul; /* tap k of a K=3 conv: 2, 1, 0 (multiplied by the dilation) */
    int32_t k_blocks;  /* segment length / 64 */
    int32_t w_k0;
    int32_t n_mma;     /* 0 = all rows of the W tile; 128 = only the first 128 rows (wavenet: taps 0,1) */
} dn_gemm_seg;

enum {
    DN_EPI_BF16 = 0,     /* out_bf16 = acc + bias */
    DN_EPI_F32 = 1,      /* out_f32 = acc + bias (+ pe[pos(b,t)] when pe != null; LM:867-868) */
    DN_EPI_RESID = 2,    /* out_f32 += acc + bias   (residual stream, LM:692,704) */
    DN_EPI_GEGLU = 3,    /* W tile = 128 "x" rows then 128 "gate" rows: out = gelu_erf(gate) * x (LM:881-885) */
    DN_EPI_WN_GATE = 4,  /* W tile = 128 conv rows then 128 res rows: y = tanh(u')sigmoid(u') + res (LM:513-536) */
    /* eps_hat = acc + bias is consumed in registers by the sampler's DDIM eta = 0 update (LM:1419-1442, dn_ddim_step mode 0):
     * out = the fp32 latent x [B*T, ldo] updated IN PLACE, aux = its split-precision bf16 staging copy.  The denoiser's
     * last GEMM (final_proj, n_out = z <= 128: one W tile) runs this, so eps_hat never goes to HBM and the separate update
     * kernel (a 16 MB launch at z = 16, pure latency) leaves the loop.  n_out % 16 == 0, groups == 1. */
    DN_EPI_DDIM = 5,
    /* The unit head (decoder_lm, LM:1095-1096,1450): the 1004 logits of a frame are never written; every epilogue
     * warpgroup reduces its 128 accumulator columns (+ bias) to one (max value, first index) pair with torch.argmax's
     * ordering (NaN greatest, first occurrence wins) and writes it to out = fp32 [B*T, 2 * (2 * n_tiles)]: partial p of a
     * row at columns 2p (value) and 2p + 1 (the index, as int32 bits).  dn_argmax_combine folds the partials of a row.
     * Columns >= n_classes are excluded.  The max is taken over the SAME fp32 values DN_EPI_F32 would have stored, so the
     * units are bit-identical to dn_argmax_units over materialised logits. */
    DN_EPI_ARGMAX = 6
};
enum { DN_GEMM_TCGEN05 = 0, DN_GEMM_SIMT_CHECK = 1, DN_GEMM_TCGEN05_2CTA = 2 };
enum { DN_FMT_BF16 = 0, DN_FMT_F16 = 1 };   /* 16-bit operand / output formats (same tensor-core rate) */
#define DN_MAX_SEGS 12

typedef struct {
    int32_t B, T;               /* utterances, frames per utterance */
    int32_t groups;             /* independent problems in one launch (the 8 wavenet chains); >= 1 */
    /* A: bf16 activations [B, T, lda]; tensor-map extent in columns = a_cols */
    const void* A;
    int32_t lda, a_cols;
    int64_t a_batch_stride;     /* elements between utterances */
    int32_t g_a_col;            /* + g * g_a_col columns for group g */
    /* W: bf16 [w_rows, ldw], K contiguous (packed by the host: 256 rows per N tile) */
    const void* W;
    int32_t ldw, w_rows;
    int32_t g_w_row;            /* + g * g_w_row rows for group g */
    int32_t num_segs;
    dn_gemm_seg seg[DN_MAX_SEGS];
    int32_t dilation;           /* d for group 0 */
    int32_t dilation_shl_group; /* 1: d << g (wavenet chain i has dilation 2^i, LM:553) */
    int32_t n_tiles;            /* W tiles (of 256 rows) per group */
    int32_t n_out;              /* logical output columns per group */
    int32_t epi;
    const float* bias;          /* per output column (GEGLU: per packed W row), may be null */
    const float* bias2;         /* WN: res_conv bias */
    int32_t g_bias;             /* + g * g_bias for group g */
    const float* gb;            /* WN: gamma/beta table base (null = unconditioned) */
    int64_t gb_t_stride;
    int32_t g_gb, gb_half;      /* per-group offset; beta = gamma + gb_half */
    const int32_t* t_idx;       /* device: table row per utterance */
    int32_t t_idx_stride;       /* 0: all utterances share t_idx[0] */
    void* out;
    int32_t ldo;
    int64_t out_batch_stride;
    int32_t g_out_col;
    const float* pe;            /* [T + 1, n_out] sinusoidal table (row 0 = zeros) or null */
    const int32_t* lengths;     /* [B] valid frames (pe positions), may be null = all valid */
    /* operand formats; a_fmt must equal w_fmt (a kind::f16 MMA takes both operands in one 16-bit format).  The sampler loop
     * runs fp16: a weight's rounding error repeats identically in all 99 calls and adds up coherently (x0 error 0.18 % with
     * bf16 weights against 0.02-0.03 % with fp16, profiles/r02_a3_precision_probe.jsonl); fp16 carries 3 more mantissa bits
     * at the same MMA rate, and every fp16 store saturates at +-65504 instead of overflowing. */
    int32_t a_fmt, w_fmt;       /* DN_FMT_* of A and of W */
    int32_t out_fmt;            /* DN_FMT_* of 16-bit outputs (BF16 / GEGLU / WN_GATE epilogues) */
    /* != 0: split-precision 16-bit output: column c holds hi = rn(v), column c + out_lo_col holds rn(v - hi); the GEGLU /
     * WaveNet-gate epilogues then use full-precision erf / tanh / exp.  n_out must be a multiple of 64. */
    int32_t out_lo_col;
    /* DN_EPI_DDIM only */
    const float* coef;          /* per-step coefficient rows (float32 x 8, see dn_ddim_step); row t_idx[0] is used */
    void* aux;                  /* bf16 staging copy of x: hi in columns [0, n_out), lo in [aux_lo_col, aux_lo_col + n_out) */
    int32_t aux_ld, aux_lo_col;
    int32_t n_classes;          /* DN_EPI_ARGMAX: classes that take part in the argmax (<= n_out) */
    /* Packed rows (ragged batches): 0 = every M tile lies inside one utterance (a T that is not a multiple of the tile
     * height pads each utterance up to it); 8 | 16 | 32 | 64 = M tiles are filled with row chunks of that many frames taken
     * across utterance boundaries — a tile that straddles a boundary loads / stores one TMA box per chunk, each with its own
     * (utterance, frame) origin so shifted taps still zero-fill at t < 0 — and only the last chunk of an utterance pads.
     * Results are bit-identical to row_chunk = 0. */
    int32_t row_chunk;
} dn_gemm_desc;

/* out = epilogue(A (*) W^T): bf16 operands, fp32 accumulation in TMEM (tcgen05.mma fed by TMA).
 * Replaces the cuBLAS / cuDNN calls behind nn.Linear / CausalConv1d on the path.
 * impl = DN_GEMM_SIMT_CHECK runs a slow one-thread-per-output CUDA kernel with identical semantics
 * (test checker for the tensor-core kernel; never used by the engine).
 * impl = DN_GEMM_TCGEN05_2CTA runs the same kernel as clusters of two CTAs on M = 256 tiles (tcgen05.mma.cta_group::2:
 * each CTA stages half of the W tile, the leader issues the MMAs for both SMs); results are bit-identical. */
int dn_gemm(const dn_gemm_desc* d, int32_t impl, void* stream);

/* Host-only (no CUDA call): the M tiling dn_gemm's kernel uses for a [B, T] launch — m_tiles = M tiles per group and N tile, and
 * for CTA `cta_rank` (< ctas, 1 or 2) of M tile `m_tile` the (utterance, frame) of each of its 128 accumulator rows
 * (b_out / t_out, 128 entries each, may be null) and the number of TMA boxes its A tile is loaded with (1, or 128 / row_chunk when
 * the tile straddles utterances).  Rows with b == B or t >= T are padding.  Lets the packed-rows map be checked without a GPU. */
int dn_gemm_tile_rows(int32_t B, int32_t T, int32_t row_chunk, int32_t ctas, int32_t m_tile, int32_t cta_rank,
                      int32_t* m_tiles, int32_t* b_out, int32_t* t_out, int32_t* boxes);

/* ---- attention ------------------------------------------------------------------------------------------ */

/* Non-causal multi-head attention with key-padding mask (LM:299-343, :945-949), flash-style (no N x N matrix).
 * qkv bf16 [B, T, 3*H*dh] = [q | k | v], each head-major (h d); out bf16 [B, T, H*dh].
 * Keys j >= lengths[b] get exactly zero weight.  dh in {64, 96}.
 * fmt = DN_FMT_F16: q, k, v, the probabilities and (dh 64) the output are fp16;
 * out_lo_col != 0 (dh 96 only): out is split-precision bf16 [B, T, 2 * out_lo_col] (hi | lo, as dn_adarmsnorm). */
int dn_attention(const void* qkv, void* out, const int32_t* lengths, int32_t B, int32_t T, int32_t H, int32_t dh,
                 int32_t fmt, int32_t out_lo_col, void* stream);

/* ---- denoiser training step (LatentDiscreteModel.forward, LM:1514-1613, and its backward) ----------------- */
/* Forward runs the same GEMM / conv / norm kernels as the normalization pass with per-utterance timesteps
 * (t_idx_stride = 1) and the un-fused epilogues below, which keep the pre-activations backward needs.
 * Backward = dn_gemm over transposed weight packings (dgrad) + dn_wgrad + the elementwise kernels below.
 * Accumulating outputs ("+=") must be zeroed by the caller at the start of a step. */

/* Weight-gradient GEMM:  dW[n, c] += sum_{b,t} dY[b, t, dy_col0 + n] * X[b, t - x_shift, x_col0 + c]
 * (bf16 operands in their row-major [B, T, C] layout, fp32 accumulation, split-K with TMA reduce-add).
 * Rows with t - x_shift outside [0, T) read as zero: a conv tap k of dilation d uses x_shift = (2-k) d. */
typedef struct {
    int32_t B, T;
    const void* dY; int32_t ldy; int64_t dy_batch_stride; int32_t dy_col0;
    const void* X;  int32_t ldx; int64_t x_batch_stride;  int32_t x_col0; int32_t x_shift;
    int32_t n_rows, k_cols;     /* extent of dW */
    float* dW; int32_t ldw;
    int32_t splits;             /* <= 0: chosen by the library */
    /* groups > 1: independent problems in one launch (the chains of a WaveNet level): group g reads columns
     * + g * g_dy_col / + g * g_x_col, writes dW + g * g_dw_stride (elements), and uses x_shift << g when
     * shift_shl_group is set (chain g has dilation 2^g, LM:553). */
    int32_t groups, g_dy_col, g_x_col, shift_shl_group;
    int64_t g_dw_stride;
} dn_wgrad_desc;
int dn_wgrad(const dn_wgrad_desc* d, void* stream);

/* out[c] += sum_r src[r, col0 + c]  (bias gradients); src bf16 [rows, ld]. */
int dn_colsum_bf16(const void* src, int64_t rows, int32_t ld, int32_t col0, int32_t cols, float* out, void* stream);

/* GEGLU (LM:881-885) on the packed pre-activation h [rows, 2*ip] (tile j: 128 "x" columns, then 128 gate columns):
 * m = gelu_erf(gate) * x, bf16 [rows, ip];  backward writes dh in h's layout. */
int dn_geglu_fwd(const void* h, int64_t rows, int32_t ip, void* m, void* stream);
int dn_geglu_bwd(const void* h, const void* dm, int64_t rows, int32_t ip, void* dh, void* stream);

/* WaveNet FiLM + gate (LM:513-536) for G chains at once.  ur bf16 [B*T, G*2*C]: per chain, tile j = 128 conv columns
 * then 128 res columns (what dn_gemm writes with DN_EPI_BF16 over a wavenet-level packing); y bf16 [B*T, G*C].
 * gamma/beta of (b, g): gb + t_idx[b*t_idx_stride]*gb_t_stride + g*g_gb (gamma[C], beta[C]); gb may be null.
 * Backward writes dur bf16 [B*T, G*2*C] per chain as [du (C) | dres (C)] and dgb[b*dgb_b_stride + g*g_dgb + ...] =
 * (dgamma[C], dbeta[C]) summed over the utterance's frames (overwritten, not accumulated; dgb may be null). */
int dn_wn_gate_fwd(const void* ur, void* y, int32_t B, int32_t T, int32_t C, int32_t G, const float* gb,
                   int64_t gb_t_stride, int32_t g_gb, const int32_t* t_idx, int32_t t_idx_stride, void* stream);
int dn_wn_gate_bwd(const void* ur, const void* dy, void* dur, int32_t B, int32_t T, int32_t C, int32_t G, const float* gb,
                   int64_t gb_t_stride, int32_t g_gb, const int32_t* t_idx, int32_t t_idx_stride, float* dgb,
                   int64_t dgb_b_stride, int32_t g_dgb, void* stream);

/* Backward of dn_adarmsnorm: dx (fp32 residual-stream gradient) += d(norm)/dx . dy; optional bf16 copy of the updated
 * dx; dgamma_p[C] += (unconditioned norms); dgb[b*dgb_b_stride + (gamma | beta)] += (conditioned norms).  C = 512 | 768. */
int dn_adarmsnorm_bwd(const float* x, const void* dy, float* dx, void* dx_bf16, int32_t B, int32_t T, int32_t C,
                      const float* gamma_p, float* dgamma_p, const float* gb, int64_t gb_t_stride, const int32_t* t_idx,
                      int32_t t_idx_stride, float* dgb, int64_t dgb_b_stride, void* stream);

/* Per-timestep training coefficients, float32 x 4 per step: [sqrt_ab, sqrt(1-ab), min(snr,5)/snr, 0] (LM:1530-1566).
 * x_t = sa[t_b] (z + beta0 eps0) + s1[t_b] eps  (+ bf16 staging copy, zero padded to ldx). */
int dn_train_noise(const float* z_lat, const float* eps0, const float* eps, float beta0, const float* coef,
                   const int32_t* t_idx, int32_t B, int32_t T, int32_t z, float* x_t, void* x_bf16, int32_t ldx,
                   void* stream);
/* loss[0] += mean_b( w_b mean_{T,z}( mask (pred - eps)^2 ) ) (LM:1563-1569); dpred bf16 [B*T, ldd] = grad_scale *
 * d loss / d pred (zero on padded frames and pad columns; may be null).  dx1 (bf16 [B*T, ld1], optional) = gradient
 * w.r.t. x1_hat from the decode branch (multitask): adds -s1 / max(sa, 1e-10) * dx1 on valid frames (LM:1572). */
int dn_noise_loss(const float* pred, int32_t lde, const float* eps, const int32_t* lengths, const float* coef,
                  const int32_t* t_idx, int32_t B, int32_t T, int32_t z, float* loss, void* dpred, int32_t ldd,
                  float grad_scale, const void* dx1, int32_t ld1, void* stream);
/* x1_hat = (x_t - s1 pred) / max(sa, 1e-10) as the bf16 staging input of decode_feature (LM:1572). */
int dn_pred_x1(const float* x_t, const float* pred, int32_t lde, const float* coef, const int32_t* t_idx, int32_t B,
               int32_t T, int32_t z, void* x_bf16, int32_t ldx, void* stream);
/* Logging losses of the decode branch (LM:1573-1597), forward only.  out6 (double, overwritten) = [sum over valid
 * frames of (recon-audio)^2, nll sum, smoothing sum (-sum_c lprob), correct argmax, target tokens, valid frames];
 * units int64 [B*T] with 0 = padding (ignored), logits fp32 [B*T, ld] over V classes. */
int dn_decode_losses(const float* recon, const float* audio, int32_t C, const float* logits, int32_t ld, int32_t V,
                     const int64_t* units, const int32_t* lengths, int32_t B, int32_t T, double* out6, void* stream);

/* Backward of the decode branch's losses (multitask, LM:1576-1604); stats = out6 of dn_decode_losses (device).
 * dlogits bf16 [rows, ldd] = nll_scale / n_tokens * d(label-smoothed NLL, eps_ls) / d logits (0 on rows with unit 0 and on
 * pad columns); dn_recon_grad: out bf16 [B*T, C] = d_lm + mse_scale * d(masked MSE) / d recon. */
int dn_lsnll_bwd(const float* logits, int32_t ld, int32_t V, const int64_t* units, int64_t rows, const double* stats,
                 float eps_ls, float nll_scale, void* dlogits, int32_t ldd, void* stream);
int dn_recon_grad(const float* recon, const float* audio, const float* d_lm, const int32_t* lengths, int32_t B, int32_t T,
                  int32_t C, const double* stats, float mse_scale, void* out, void* stream);

/* VAE training (LM:1118-1142): kl[0] += mean_b(0.5 mean_{z,T}(mask (mu^2 + exp(lv) - 1 - lv))) (distributions.py:62-74),
 * and the backward of the reparameterisation + KL: dparams bf16 [B*T, ldd] from dz bf16 [B*T, ldz] (channel-last),
 * kl_scale = d loss / d kl.  params as in dn_vae_reparam. */
int dn_vae_kl(const float* params, int32_t ldp, const int32_t* lengths, int32_t B, int32_t T, int32_t z, float* kl,
              void* stream);
int dn_vae_reparam_bwd(const float* params, int32_t ldp, const float* eps, int32_t eps_channel_first, const void* dz,
                       int32_t ldz, const int32_t* lengths, int32_t B, int32_t T, int32_t z, float kl_scale, void* dparams,
                       int32_t ldd, void* stream);

/* out[0..n) ~ N(0, 1) (Philox4x32-10 + Box-Muller; `offset` advances the counter so that successive calls with one seed do not
 * repeat).  Replaces the reference's torch.randn draws of the posterior sample (distributions.py:37-41, drawn on the CPU and
 * copied) and of q_sample (LM:1409) when the caller supplies no noise. */
int dn_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* Attention-dropout keep bits (LM:338): n_words uint32, each bit kept with probability 1-p (Philox4x32-10). */
int dn_dropout_bits(uint32_t* bits, int64_t n_words, float p, uint64_t seed, uint64_t offset, void* stream);

/* dn_attention with dropout and the saved row statistic L2 = m + log2(l) (fp32 [B, H, T]); dh = 64 (denoiser) or
 * dh = 96 (VAE decoder).
 * keep_bits [B, H, T, ceil(T/32)] (bit k%32 of word k/32 = key k kept) or null; keep_scale = 1/(1-p). */
int dn_attention_train(const void* qkv, void* out, float* lse2, const int32_t* lengths, const uint32_t* keep_bits,
                       float keep_scale, int32_t B, int32_t T, int32_t H, int32_t dh, void* stream);
/* Backward of dn_attention_train (dh = 64 | 96): dqkv bf16 [B, T, 3*H*dh] (fully written); delta_ws fp32 [B, H, T]. */
int dn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, const int32_t* lengths,
                     const uint32_t* keep_bits, float keep_scale, void* dqkv, float* delta_ws, int32_t B, int32_t T,
                     int32_t H, int32_t dh, void* stream);

/* Time-conditioning MLP pieces (LM:104-116, :741-745, :507, :624), fp32, M = utterances. */
int dn_silu(const float* pre, float* out, int64_t n, void* stream);
int dn_silu_bwd(const float* pre, const float* dout, float* dpre, int64_t n, void* stream);
/* dW[N,K] += dY^T X, db[N] += colsum(dY) (when dW != null); dX[M,K] += dY W (when dX != null).  dY [M, ldy]. */
int dn_linear_f32_bwd(const float* dY, int64_t ldy, const float* X, const float* W, int32_t M, int64_t N, int32_t K,
                      float* dW, float* db, float* dX, void* stream);
int dn_time_features_bwd(const int32_t* steps, const float* w, const float* dfeat, int32_t M, int32_t half, float* dw,
                         void* stream);
/* dst fp32 [rows, ldd] (= or +=) src bf16 [rows, ld] columns [col0, col0 + C). */
int dn_add_bf16_to_f32(const void* src, int64_t rows, int32_t ld, int32_t col0, int32_t C, float* dst, int32_t ldd,
                       int32_t accumulate, void* stream);

/* Per-step weight re-packing of the training step: the optimizer updates the fp32 master weights in place, the GEMMs read
 * bf16 K-major tiles (and their transposes for the data gradients).  ONE launch walks a table of copy descriptors that lives
 * in device memory: dst[rowmap(r), colmap(c, k)] = cast(src[r * s_row + c * s_col + k * s_tap]) for r < rows, c < cols,
 * k < taps, with  rowmap(r) = row0 + (r / rblk) * rblk_stride + r % rblk,
 *                 colmap(c, k) = col0 + tap_pos[k] * tap_cols + (c / cblk) * cblk_stride + c % cblk
 * (nn.Linear / conv weights, optionally transposed, the 3 taps of a k = 3 conv side by side, the 128-row interleave of the
 * GEGLU and WaveNet tiles; padding rows / columns are never written and stay zero).  Tiles of 64 x 64 (x taps): tile0 =
 * number of tiles of all earlier ops, tiles_c = ceil(cols / 64).  Replaces the ~500 torch indexing kernels of
 * diffnorm_b200.packing on the training path (same bits: fp32 -> bf16 round-to-nearest-even, or a plain fp32 copy). */
typedef struct {
    const float* src;
    void* dst;
    int64_t s_row, s_col, s_tap;          /* element strides of src */
    int32_t rows, cols, taps;             /* taps <= 3 */
    int32_t ldd;                          /* dst leading dimension (elements) */
    int32_t row0, rblk, rblk_stride;
    int32_t col0, cblk, cblk_stride, tap_cols;
    int32_t tap_pos[3];
    int32_t src_r_fastest;                /* 1: src is contiguous along r (and k): transposed through shared memory */
    int32_t out_f32;                      /* dst dtype: 0 = bf16, 1 = fp32 */
    int32_t tile0, tiles_c;
} dn_pack_op;
int dn_pack_weights(const dn_pack_op* ops_device, int32_t n_ops, int32_t total_tiles, void* stream);

/* ---- the step after the pass: duration-aware unit vocoder ------------------------------------ */
/* CodeHiFiGAN generator + duration predictor on the reduced units the pass writes (fairseq/models/text_to_speech/
 * codehifigan.py:49-76, hifigan.py:20-179, fastspeech2.py:117-151; driver examples/speech_to_speech/
 * generate_waveform_from_code.py:78-96).  fp32, ONE utterance, channel-first activations [C, L] like the reference. */

/* y[co, l] = act(bias[co] + sum_ci sum_k w[co, ci, k] * lrelu(x[ci, l + k*dilation - pad], in_slope)) (+ res[co, l]), times
 * out_scale, stored or (accumulate != 0) added to y.  x [Cin, L], w [Cout, Cin, K] (nn.Conv1d layout, weight norm already
 * folded), y [Cout, L] ("same" padding is the caller's pad = dilation (K-1)/2).  in_slope = 1 leaves x as is (hifigan.py:93-97
 * applies leaky_relu 0.1 before every conv); out_act: 0 none, 1 ReLU (duration predictor), 2 tanh (conv_post). K <= 16. */
int dn_voc_conv1d(const float* x, int32_t L, int32_t Cin, const float* w, const float* bias, int32_t Cout, int32_t K,
                  int32_t dilation, int32_t pad, float in_slope, int32_t out_act, const float* res, float out_scale,
                  int32_t accumulate, float* y, void* stream);
/* nn.ConvTranspose1d(Cin, Cout, K, stride, padding = (K - stride) / 2) after leaky_relu(in_slope): x [Cin, L] -> y [Cout, L*stride];
 * w [Cin, Cout, K] (hifigan.py:128-140, :157-158). */
int dn_voc_conv_transpose1d(const float* x, int32_t L, int32_t Cin, const float* w, const float* bias, int32_t Cout, int32_t K,
                            int32_t stride, int32_t pad, float in_slope, float* y, void* stream);
/* LayerNorm over the C channels at every position, eps 1e-5 (fastspeech2.py:147,149); x, y [C, L]. */
int dn_voc_layernorm(const float* x, int32_t C, int32_t L, const float* gamma, const float* beta, float* y, void* stream);
/* dur[t] = max(round(exp(log_dur[t]) - 1), 1) (round half to even, codehifigan.py:66-68; log_dur null = all ones) and
 * start[0..T] = its exclusive prefix sum (start[T] = number of output frames). */
int dn_voc_durations(const float* log_dur, int32_t T, int64_t* dur, int64_t* start, void* stream);
/* x[c, j] = table[code[u], c] for start[u] <= j < start[u+1]: embedding lookup + repeat_interleave (codehifigan.py:57,69);
 * table [num_embeddings, dim], x [dim, Lo]. */
int dn_voc_embed_repeat(const int64_t* code, int32_t T, const float* table, int32_t dim, const int64_t* start, int32_t Lo,
                        float* x, void* stream);

/* ---- host-side helper (no CUDA) --------------------------------------------------------------------------- */

/* Length-bucketed batching under a padded-token budget; same contract and results as the reference's Cython
 * batch_by_size_vec (fairseq/data/data_utils_fast.pyx:20-101).  num_tokens[i] = length of the i-th utterance in
 * iteration (normally length-sorted) order; batch k = [batch_ends[k-1], batch_ends[k]) with batch_ends[-1] = 0.
 * batch_ends must hold n + 1 entries.  Returns the number of batches, or DN_EINVAL (an utterance longer than
 * max_tokens, bad arguments).  max_tokens / max_sentences <= 0 disable that limit. */
int64_t dn_batch_by_size(const int64_t* num_tokens, int64_t n, int64_t max_tokens, int64_t max_sentences,
                         int32_t bsz_mult, int64_t* batch_ends);

#ifdef __cplusplus
}
#endif
#endif /* DIFFNORM_B200_H */
