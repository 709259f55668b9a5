// dn_attention (dh = 64): flash attention on tcgen05 tensor cores with S and the per-block P.V product in TMEM.
//
// One CTA = 128 queries of one (utterance, head); 2 CTAs per SM hide each other's softmax / MMA latency.
//   warp 0      TMA producer : Q once, then K_j / V_j tiles (box {64 dh, 128 frames, 1 utt} of the [3*H*dh, T, B] map)
//   warp 1      MMA issuer   : S_j = Q K_j^T  (M128 N128 K64, both K-major)            -> TMEM cols [0,128)
//                              O  += P_j V_j  (M128 N64 K128, A = P from TMEM, B = V MN-major) -> TMEM cols [128,192)
//                              S_{j+1} is issued as soon as the softmax threads hold S_j in registers.
//   warps 2..9  softmax      : two threads per query row (warps w and w+4 share rows and split the 128 key columns /
//                              64 output columns).  S_j is read from TMEM once (TMEM reads are 64 B/clk/SM); row max
//                              with one smem exchange; exp2 / row sum with packed f32x2 FFMA2/FADD2 + FMNMX3 +
//                              MUFU.EX2; P_j goes back to TMEM as packed bf16 (cols [192,256)); O accumulates in TMEM
//                              and is only rescaled when the running max jumps by more than 2^8 (lazy rescale).
//   K and V tiles are double buffered in smem (5 x 16 KB with Q).
// Keys j >= lengths[b] get exactly zero weight (LM:333-335); key blocks past the length are skipped.
#include "common.cuh"

namespace dn {

struct FalseTag { static constexpr bool value = false; };
struct TrueTag { static constexpr bool value = true; };

constexpr int TA_BM = 128, TA_BN = 128, TA_DH = 64;
constexpr int TA_THREADS = 320;                           // TMA warp + MMA warp + 8 softmax warps
constexpr int TA_TILE = TA_BM * TA_DH * 2;                  // 16 KB: Q and each K / V stage
constexpr int TA_SMEM = 5 * TA_TILE + 1024 + 128 + 6 * 128 * 4;  // Q, K(2), V(2) + align slack + barriers + exchange

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// instruction descriptor: 16-bit x 16-bit -> fp32 (both operands bf16, or both fp16), M = 128, runtime N, A K-major,
// B K-major (0) or MN-major (1)
__device__ __forceinline__ uint32_t ta_idesc(uint32_t n, uint32_t b_mn_major, bool f16) {
    const uint32_t fm = f16 ? 0u : 1u;
    return (1u << 4) | (fm << 7) | (fm << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// TRAIN: additionally applies the attention-dropout keep bits (LM:338; keep [B,H,T,ceil(T/32)] words, bit k%32 of word
// k/32 = key k kept; P is scaled by keep_scale = 1/(1-p) after the row sum) and saves L2 = m + log2(l) per query row.
// F16: q, k, v, the probabilities P and the output are fp16 instead of bf16 (the sampler loop's format).
// OPT: key blocks after the first exponentiate OPTIMISTICALLY against the running max they already have, while the second
// half of S is still on its way from TMEM and before the two threads of a row have exchanged their maxima; the block is
// redone from the registers in the rare case that its true max exceeds the running one by more than 2^8 (the same lazy
// threshold the un-optimistic path applies before it raises the max).  The dependent chain of a block shrinks from
// LDTM -> max -> barrier -> exp2 -> STTM to LDTM/exp2 overlapped -> barrier -> STTM.
template <bool TRAIN, bool F16, bool OPT>
__global__ void __launch_bounds__(TA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out,
                    const int* __restrict__ lengths, int T, int H, float scale_log2, float* __restrict__ lse2,
                    const uint32_t* __restrict__ keep, float keep_scale) {
    extern __shared__ uint8_t ta_smem_raw[];
    uint8_t* smem = ta_smem_raw + ((1024u - (smem_u32(ta_smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TA_TILE;      // two stages
    uint8_t* sV = smem + 3 * TA_TILE;  // two stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TA_TILE);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // S_j complete in TMEM
    uint64_t* s_free = bars + 10;   // S_j copied to registers by all softmax threads -> S_{j+1} may overwrite it
    uint64_t* p_full = bars + 11;   // P_j in TMEM (and O rescaled if needed) -> PV_j may be issued
    uint64_t* p_empty = bars + 12;  // PV_j complete: P columns reusable, O stable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* xmax = reinterpret_cast<float*>(bars + 16);   // [2 parities][2 halves][128 rows] partial row maxima
    float* lsum = xmax + 4 * TA_BM;                      // [2 halves][128 rows] partial row sums

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TA_BM, h = blockIdx.y, b = blockIdx.z;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    const int nkb = (len + TA_BN - 1) / TA_BN;
    const int qcol = h * TA_DH, kcol = (H + h) * TA_DH, vcol = (2 * H + h) * TA_DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(k_full + i, 1);
            mbar_init(k_empty + i, 1);
            mbar_init(v_full + i, 1);
            mbar_init(v_empty + i, 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 256);
        mbar_init(p_full, 256);
        mbar_init(p_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();       // qkv comes from the kernel before (lengths, read above, is an input of the whole pass)
    // TMEM columns: S fp32 [0,128) | O fp32 [128,192) | P bf16x2 [192,256) (two keys per 32-bit column)
    const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            mbar_expect_tx(q_full, TA_TILE);
            tma_load_3d(&tmQKV, q_full, sQ, qcol, q0, b);
            for (int j = 0; j < nkb; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(k_empty + st, ph ^ 1);
                mbar_expect_tx(k_full + st, TA_TILE);
                tma_load_3d(&tmQKV, k_full + st, sK + st * TA_TILE, kcol, j * TA_BN, b);
                mbar_wait(v_empty + st, ph ^ 1);
                mbar_expect_tx(v_full + st, TA_TILE);
                tma_load_3d(&tmQKV, v_full + st, sV + st * TA_TILE, vcol, j * TA_BN, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t id_s = ta_idesc(TA_BN, 0, F16), id_o = ta_idesc(TA_DH, 1, F16);
            const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
            auto issue_s = [&](int j) {
                const int st = j & 1;
                mbar_wait(k_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128(smem_u32(sK + st * TA_TILE));
#pragma unroll
                for (int k = 0; k < TA_DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, id_s, k > 0);
                umma_commit(k_empty + st);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkb; ++j) {
                if (j + 1 < nkb) {
                    // S_{j+1} as soon as the softmax threads hold S_j in registers: it runs under their exp phase
                    mbar_wait(s_free, j & 1);
                    issue_s(j + 1);
                }
                const int st = j & 1;
                mbar_wait(p_full, j & 1);            // P_j in TMEM (and O rescaled when the row max jumped)
                mbar_wait(v_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * TA_TILE));  // MN-major: 8-key groups 1 KB apart
#pragma unroll
                for (int k = 0; k < TA_BN / 16; ++k) {
                    // A = P from TMEM: 16 keys = 8 columns per k-step; B = V (MN-major): 16 rows of 128 B per k-step
                    umma_bf16_ts(tO, tP + 8 * k, dv + (uint64_t)((k * 16 * 128) >> 4), id_o, (j > 0) || (k > 0));
                }
                umma_commit(v_empty + st);
                umma_commit(p_empty);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: two threads per query row
        // warps 2..5 own key columns [0,64) of each block and output columns [0,32); warps 6..9 the other halves.
        // (A warp may only touch TMEM lanes 32*(warp%4)..+31, so warps w and w+4 share rows and split columns.)
        // S_j is read ONCE into 64 registers per thread; P_j goes back to TMEM as packed bf16 (the PV MMA reads its
        // A operand from TMEM, so there is no smem round trip and no proxy fence); O stays in TMEM, accumulated by the
        // MMA itself.  The running max m is only raised when a block's max exceeds it by more than 2^RESCALE_LOG2
        // (P then stays <= 2^8: exact in fp32 sums, same relative precision in bf16); only then is O rescaled in
        // TMEM (tcgen05.ld -> mul -> tcgen05.st), which is rare after the first block.
        constexpr float RESCALE_LOG2 = 8.f;
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        float m = -INFINITY, l = 0.f;
        const uint64_t scale2 = pack2(scale_log2, scale_log2);
        const int Tw = (T + 31) >> 5;
        const uint32_t* krow = nullptr;
        if (TRAIN && keep) {
            const int tq = q0 + row < T ? q0 + row : T - 1;
            krow = keep + (((long long)b * H + h) * T + tq) * Tw;
        }
        // Only the last key block of an utterance can contain masked keys: the block body is instantiated twice so
        // the hot (unmasked) copy carries no predicated compare/select instructions at all.
        auto block = [&](const int j, auto masked_tag) {
            constexpr bool masked = decltype(masked_tag)::value;
            float s0[32], s1[32];
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            tmem_ld32(tS + lane_off + half * 64, s0);
            tmem_ld32(tS + lane_off + half * 64 + 32, s1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            const int kbase = j * TA_BN + half * 64;
            if constexpr (masked) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    s0[i] = (kbase + i < len) ? s0[i] : -INFINITY;
                    s1[i] = (kbase + 32 + i < len) ? s1[i] : -INFINITY;
                }
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                mx[(i >> 1) & 1] = max3(mx[(i >> 1) & 1], s0[i], s0[i + 1]);
                mx[2 + ((i >> 1) & 1)] = max3(mx[2 + ((i >> 1) & 1)], s1[i], s1[i + 1]);
            }
            float mxx = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            xmax[((j & 1) * 2 + half) * TA_BM + row] = mxx;
            named_bar_sync(1, 256);
            mxx = fmaxf(mxx, xmax[((j & 1) * 2 + (half ^ 1)) * TA_BM + row]);
            const float mxs = mxx * scale_log2;            // finite: every processed block has a valid key
            const bool raise = mxs > m + RESCALE_LOG2;     // j == 0: m = -inf -> true
            const float m_old = m;
            if (raise) {
                l *= ex2_approx(m - mxs);                  // j == 0: l = 0 * 0
                m = mxs;
            }
            const uint64_t nm2 = pack2(-m, -m);
            // p = exp2(s * scale - m), partial row sum, packed bf16 pairs
            uint64_t rs2[2] = {0ull, 0ull};
            uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu;
            if (TRAIN && krow) {
                const int wi = kbase >> 5;
                kw0 = wi < Tw ? krow[wi] : 0u;
                kw1 = wi + 1 < Tw ? krow[wi + 1] : 0u;
            }
            auto half_row = [&](const float (&s)[32], uint32_t (&w)[16], const uint32_t kw) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0, e1;
                    unpack2(ffma2(pack2(s[i], s[i + 1]), scale2, nm2), e0, e1);
                    e0 = ex2_approx(e0);   // masked keys: ex2(-inf) = +0
                    e1 = ex2_approx(e1);
                    rs2[(i >> 1) & 1] = fadd2(rs2[(i >> 1) & 1], pack2(e0, e1));
                    if (TRAIN && krow) {   // dropout acts on the normalised probabilities: the row sum stays un-dropped
                        e0 = ((kw >> i) & 1u) ? e0 * keep_scale : 0.f;
                        e1 = ((kw >> (i + 1)) & 1u) ? e1 * keep_scale : 0.f;
                    }
                    w[i >> 1] = pack16<F16>(e0, e1);
                }
            };
            uint32_t w0[16], w1[16];
            half_row(s0, w0, kw0);
            if (j > 0) {
                mbar_wait(p_empty, (j - 1) & 1);           // PV_{j-1} done: P columns free, O stable
                tc_fence_after();
                if (__any_sync(0xffffffffu, raise)) {
                    const float a = raise ? ex2_approx(m_old - m) : 1.f;
#pragma unroll 1
                    for (int c8 = 0; c8 < 32; c8 += 8) {
                        float o[8];
                        tmem_ld8(tO + lane_off + half * 32 + c8, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] *= a;
                        tmem_st8(tO + lane_off + half * 32 + c8, o);
                    }
                }
            }
            tmem_st16(tP + lane_off + half * 32, w0);
            half_row(s1, w1, kw1);
            tmem_st16(tP + lane_off + half * 32 + 16, w1);
            float r0, r1, r2, r3;
            unpack2(rs2[0], r0, r1);
            unpack2(rs2[1], r2, r3);
            l += (r0 + r1) + (r2 + r3);
            tmem_st_wait();
            tc_fence_before();          // TMEM reads / writes ordered before the MMA that follows the barrier
            mbar_arrive(p_full);
        };
        // optimistic form of `block` for j >= 1 (m is finite): same outputs unless the rare redo path runs
        auto block_opt = [&](const int j, auto masked_tag) {
            constexpr bool masked = decltype(masked_tag)::value;
            float s0[32], s1[32];
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            tmem_ld32(tS + lane_off + half * 64, s0);
            tmem_ld_wait();
            tmem_ld32(tS + lane_off + half * 64 + 32, s1);   // in flight under the first half's exponentials
            const int kbase = j * TA_BN + half * 64;
            uint64_t nm2 = pack2(-m, -m);
            uint64_t rs2[2] = {0ull, 0ull};
            float mx[2] = {-INFINITY, -INFINITY};
            auto half_row = [&](float (&s)[32], uint32_t (&w)[16], const int k0, const bool with_max) {
                if constexpr (masked) {
                    if (with_max) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) s[i] = (k0 + i < len) ? s[i] : -INFINITY;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float e0, e1;
                    if (with_max) mx[(i >> 1) & 1] = max3(mx[(i >> 1) & 1], s[i], s[i + 1]);
                    unpack2(ffma2(pack2(s[i], s[i + 1]), scale2, nm2), e0, e1);
                    e0 = ex2_approx(e0);   // masked keys: ex2(-inf) = +0
                    e1 = ex2_approx(e1);
                    rs2[(i >> 1) & 1] = fadd2(rs2[(i >> 1) & 1], pack2(e0, e1));
                    w[i >> 1] = pack16<F16>(e0, e1);
                }
            };
            // P goes to TMEM half by half as soon as it exists (its columns are free once PV_{j-1} has completed, which it
            // has long before the first 32 exponentials are done), so only S stays in registers for the redo path
            uint32_t w[16];
            half_row(s0, w, kbase, true);
            mbar_wait(p_empty, (j - 1) & 1);              // PV_{j-1} done: P columns free, O stable
            tc_fence_after();
            tmem_st16(tP + lane_off + half * 32, w);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            half_row(s1, w, kbase + 32, true);
            tmem_st16(tP + lane_off + half * 32 + 16, w);
            float mxx = fmaxf(mx[0], mx[1]);
            xmax[((j & 1) * 2 + half) * TA_BM + row] = mxx;
            named_bar_sync(1, 256);
            mxx = fmaxf(mxx, xmax[((j & 1) * 2 + (half ^ 1)) * TA_BM + row]);
            const float mxs = mxx * scale_log2;
            const bool raise = mxs > m + RESCALE_LOG2;
            if (__any_sync(0xffffffffu, raise)) {         // rare: the optimistic exponentials used a stale max
                // the whole warp redoes the block (tcgen05.st / wait are warp-collective); lanes that did not raise
                // recompute the values they already had
                const float m_old = m;
                if (raise) {
                    l *= ex2_approx(m - mxs);
                    m = mxs;
                }
                nm2 = pack2(-m, -m);
                rs2[0] = rs2[1] = 0ull;
                tmem_st_wait();
                half_row(s0, w, kbase, false);
                tmem_st16(tP + lane_off + half * 32, w);
                half_row(s1, w, kbase + 32, false);
                tmem_st16(tP + lane_off + half * 32 + 16, w);
                const float a = raise ? ex2_approx(m_old - m) : 1.f;
#pragma unroll 1
                for (int c8 = 0; c8 < 32; c8 += 8) {
                    float o[8];
                    tmem_ld8(tO + lane_off + half * 32 + c8, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] *= a;
                    tmem_st8(tO + lane_off + half * 32 + c8, o);
                }
            }
            float r0, r1, r2, r3;
            unpack2(rs2[0], r0, r1);
            unpack2(rs2[1], r2, r3);
            l += (r0 + r1) + (r2 + r3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full);
        };
        const int n_full = len / TA_BN;  // key blocks without any masked key
        if constexpr (OPT && !TRAIN) {
            if (nkb > 0) {
                if (n_full > 0) block(0, FalseTag{}); else block(0, TrueTag{});   // j = 0 has no running max yet
                for (int j = 1; j < n_full; ++j) block_opt(j, FalseTag{});
                if (n_full < nkb && n_full > 0) block_opt(n_full, TrueTag{});
            }
        } else {
            for (int j = 0; j < n_full; ++j) block(j, FalseTag{});
            if (n_full < nkb) block(n_full, TrueTag{});
        }
        // combine the two partial row sums, normalise, store my 32 output columns
        lsum[half * TA_BM + row] = l;
        named_bar_sync(2, 256);
        l += lsum[(half ^ 1) * TA_BM + row];
        const int t = q0 + row;
        if (TRAIN && lse2 && half == 0 && t < T) lse2[((long long)b * H + h) * T + t] = m + log2f(l);
        if (nkb > 0) {
            mbar_wait(p_empty, (nkb - 1) & 1);   // last PV complete
            tc_fence_after();
            float o[32];
            tmem_ld32(tO + lane_off + half * 32, o);
            tmem_ld_wait();
            if (t < T) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                __nv_bfloat16* op = out + ((long long)b * T + t) * (H * TA_DH) + h * TA_DH + half * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    *reinterpret_cast<uint4*>(op + i) =
                        make_uint4(pack16<F16>(o[i] * inv, o[i + 1] * inv), pack16<F16>(o[i + 2] * inv, o[i + 3] * inv),
                                   pack16<F16>(o[i + 4] * inv, o[i + 5] * inv), pack16<F16>(o[i + 6] * inv, o[i + 7] * inv));
                }
            }
        } else if (t < T) {
            __nv_bfloat16* op = out + ((long long)b * T + t) * (H * TA_DH) + h * TA_DH + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(op + i) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);

int launch_attention_tc(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st,
                        float* lse2, const uint32_t* keep, float keep_scale, bool train, bool f16) {
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        attr_set = true;
    }
    static int opt = -1;       // DN_ATTN_OPT=0: the max-first softmax for every key block (A/B switch)
    if (opt < 0) {
        const char* e = getenv("DN_ATTN_OPT");
        opt = (e && e[0] == '0') ? 0 : 1;
    }
    if (train && f16) return DN_EINVAL;
    CUtensorMap m;
    const int ld = 3 * H * TA_DH;
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {TA_DH, TA_BM, 1};
    int r = encode_bf16_map(&m, qkv, 3, dims, str, box);
    if (r) return r;
    dim3 grid((T + TA_BM - 1) / TA_BM, H, B);
    const float scale_log2 = (1.0f / sqrtf((float)TA_DH)) * 1.4426950408889634f;
#define DN_ATT_LAUNCH(TR, FH, OP, L2, KP, KS)                                                                                 \
    DN_CUDA_OK(launch_ex(attention_tc_kernel<TR, FH, OP>, grid, TA_THREADS, TA_SMEM, st, 1, m, reinterpret_cast<__nv_bfloat16*>(out), \
                         lengths, T, H, scale_log2, L2, KP, KS))
    if (train) DN_ATT_LAUNCH(true, false, false, lse2, keep, keep_scale);
    else if (f16 && opt) DN_ATT_LAUNCH(false, true, true, (float*)nullptr, (const uint32_t*)nullptr, 1.f);
    else if (f16) DN_ATT_LAUNCH(false, true, false, (float*)nullptr, (const uint32_t*)nullptr, 1.f);
    else if (opt) DN_ATT_LAUNCH(false, false, true, (float*)nullptr, (const uint32_t*)nullptr, 1.f);
    else DN_ATT_LAUNCH(false, false, false, (float*)nullptr, (const uint32_t*)nullptr, 1.f);
#undef DN_ATT_LAUNCH
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

}  // namespace dn
