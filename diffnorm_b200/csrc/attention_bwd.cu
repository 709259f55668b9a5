// dn_attention_bwd (dh = 64 | 96): backward of the self-attention of the denoiser (dh 64) and of the frozen VAE decoder
// (dh 96, data gradients only, multitask training) (LM:299-343, :908-950) on tcgen05 tensor
// cores, flash style (nothing of size N x N touches HBM).  Forward (attention_tc.cu, TRAIN form) saved, per query
// row, L2 = m + log2(l) (log-sum-exp of the scaled scores in the log2 domain); with D[q] = sum_d dO[q,d] O[q,d]:
//     P  = exp2(S * scale_log2 - L2[q])            (keys >= length: 0)
//     dV = (P o keep * ks)^T dO                    (keep/ks: attention dropout, LM:338)
//     dS = scale * P o ((dO V^T) o keep * ks - D[q])
//     dQ = dS K        dK = dS^T Q
// Two kernels, both with the forward's role layout (warp 0 TMA producer, warp 1 MMA issuer, 8 softmax warps = two
// threads per TMEM lane) and operands in TMEM where the ISA allows it (A of the second GEMM of every pair):
//   attn_bwd_dq_kernel   CTA = 128 queries: per key block  S = Q K^T, dP = dO V^T  -> dS (TMEM) -> dQ += dS K
//   attn_bwd_dkv_kernel  CTA = 128 keys:    per query block S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T (TMEM)
//                                           -> dV += P^T dO, dK += dS^T Q
// K / V (resp. Q / dO) blocks are double buffered in smem; the same smem tile serves as a K-major operand of the first
// GEMM and as an MN-major operand of the second.  One CTA per SM (all 512 TMEM columns in the dKdV kernel).
#include "common.cuh"

namespace dn {

constexpr int AB_T = 128;             // tile edge (queries or keys)
constexpr int AB_THREADS = 320;
constexpr int AB_TILE = AB_T * 64 * 2;      // 16 KB: one {64 columns, 128 rows} TMA box
// an operand of head dim DH = ceil(DH/64) boxes (dh 96: the second box's upper 32 columns are never consumed)
template <int DH> struct AbCfg {
    static constexpr int NT = (DH + 63) / 64;
    static constexpr int OP = NT * AB_TILE;
    // resident pair + 2 stages x pair + align slack + barriers + per-query scalars + keep words
    static constexpr int SMEM = 6 * OP + 1024 + 256 + 2 * 2 * AB_T * 4 + 2 * AB_T * 4 * 4;
};
// MN-major SW128 operand whose 64-column atoms are AB_TILE bytes apart (LBO); 8-row groups 1 KB apart (SBO)
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((AB_TILE >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);

__device__ __forceinline__ float ab_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// M = 128, runtime N, A K-major (smem) or TMEM, B K-major (0) or MN-major (1)
__device__ __forceinline__ uint32_t ab_idesc(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[b, h, t] = sum_d dO[b, t, h, d] * O[b, t, h, d]   (one warp per (frame, head); 2 bf16 per lane)
template <int DH>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int B, int T,
                                  int H, float* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = (long long)B * T * H;
    for (long long w = warp0; w < total; w += nwarps) {
        const int h = (int)(w % H);
        const long long bt = w / H;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DH; c += 64) {
            if (c + lane * 2 < DH) {
                const long long off = (bt * H + h) * DH + c + lane * 2;
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + off));
                const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d_o + off));
                acc += a.x * g.x + a.y * g.y;
            }
        }
        const float s = warp_sum(acc);
        if (lane == 0) {
            const int b = (int)(bt / T), t = (int)(bt % T);
            delta[((long long)b * H + h) * T + t] = s;
        }
    }
}

struct AbBars {
    uint64_t* r_full;      // resident pair loaded
    uint64_t* st_full;     // [2] streamed pair loaded
    uint64_t* st_empty;    // [2] streamed pair consumed
    uint64_t* sdp_full;    // S and dP of block j complete in TMEM
    uint64_t* s_free;      // S and dP of block j copied to registers (256 arrivals)
    uint64_t* ds_full;     // P / dS of block j written to TMEM (256 arrivals)
    uint64_t* ds_empty;    // second GEMMs of block j complete
};

// loads one operand (NT boxes) of head dim DH: columns [col, col + 64 NT), rows [row0, row0 + 128) of utterance b
template <int DH>
__device__ __forceinline__ void ab_load(const CUtensorMap* m, uint64_t* bar, uint8_t* dst, int col, int row0, int b) {
#pragma unroll
    for (int i = 0; i < AbCfg<DH>::NT; ++i) tma_load_3d(m, bar, dst + i * AB_TILE, col + i * 64, row0, b);
}
// D[128 x 128] = A[128 x DH] B[128 x DH]^T, both K-major over DH (k-step k lives in box k/4)
template <int DH>
__device__ __forceinline__ void ab_mma_kk(uint32_t d_tmem, uint32_t sa, uint32_t sb, uint32_t idesc) {
#pragma unroll
    for (int k = 0; k < DH / 16; ++k) {
        const uint32_t off = (k >> 2) * AB_TILE + (k & 3) * 32;
        umma_bf16(d_tmem, umma_desc_sw128(sa + off), umma_desc_sw128(sb + off), idesc, k > 0);
    }
}
// D[128 x DH] (+)= A[128 x 128 from TMEM] B[128 x DH], B MN-major (rows = contraction index, 16 per k-step)
template <int DH>
__device__ __forceinline__ void ab_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t sb, uint32_t idesc, bool accumulate) {
#pragma unroll
    for (int k = 0; k < AB_T / 16; ++k)
        umma_bf16_ts(d_tmem, a_tmem + 8 * k, ab_desc_mn(sb + k * 16 * 128), idesc, accumulate || (k > 0));
}
// this thread's half (DH/2 columns) of an fp32 accumulator row -> bf16 global
template <int DH>
__device__ __forceinline__ void ab_store_half(uint32_t taddr, __nv_bfloat16* dst, bool valid_acc, bool write) {
    constexpr int HC = DH / 2;   // 32 or 48
    float o[48];
    if (valid_acc) {
        float a[32];
        tmem_ld32(taddr, a);
        if (HC == 48) {
            float c[16];
            tmem_ld16(taddr + 32, c);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[32 + i] = c[i];
        } else {
            tmem_ld_wait();
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = a[i];
    } else {
#pragma unroll
        for (int i = 0; i < 48; ++i) o[i] = 0.f;
    }
    if (write) {
#pragma unroll
        for (int i = 0; i < HC; i += 8)
            *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16(o[i], o[i + 1]), pack_bf16(o[i + 2], o[i + 3]),
                                                            pack_bf16(o[i + 4], o[i + 5]), pack_bf16(o[i + 6], o[i + 7]));
    }
}

// ---------------------------------------------------------------------------------------------------- dQ
template <int DH>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse2, const float* __restrict__ delta, const int* __restrict__ lengths,
                   const uint32_t* __restrict__ keep, float keep_scale, __nv_bfloat16* __restrict__ dqkv, int T, int H,
                   float scale, float scale_log2) {
    constexpr int OP = AbCfg<DH>::OP;
    extern __shared__ uint8_t ab_smem_raw[];
    uint8_t* smem = ab_smem_raw + ((1024u - (smem_u32(ab_smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS
    uint8_t* sQ = smem;                     // resident: Q tile, dO tile
    uint8_t* sDO = smem + OP;
    uint8_t* sStage = smem + 2 * OP;        // 2 stages x (K, V)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * OP);
    AbBars bb{bars + 0, bars + 1, bars + 3, bars + 5, bars + 6, bars + 7, bars + 8};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    const int nkb = (len + AB_T - 1) / AB_T;
    const int qcol = h * DH, kcol = (H + h) * DH, vcol = (2 * H + h) * DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(bb.r_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bb.st_full + i, 1);
            mbar_init(bb.st_empty + i, 1);
        }
        mbar_init(bb.sdp_full, 1);
        mbar_init(bb.s_free, 256);
        mbar_init(bb.ds_full, 256);
        mbar_init(bb.ds_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM: S fp32 [0,128) | dP fp32 [128,256) | dS bf16x2 [256,320) | dQ fp32 [320,320+DH)
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDS = tmem_base + 256, tDQ = tmem_base + 320;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            mbar_expect_tx(bb.r_full, 2 * OP);
            ab_load<DH>(&tmQKV, bb.r_full, sQ, qcol, q0, b);
            ab_load<DH>(&tmDO, bb.r_full, sDO, h * DH, q0, b);
            for (int j = 0; j < nkb; ++j) {
                const int st = j & 1;
                mbar_wait(bb.st_empty + st, ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(bb.st_full + st, 2 * OP);
                ab_load<DH>(&tmQKV, bb.st_full + st, sStage + st * 2 * OP, kcol, j * AB_T, b);
                ab_load<DH>(&tmQKV, bb.st_full + st, sStage + st * 2 * OP + OP, vcol, j * AB_T, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t id_s = ab_idesc(AB_T, 0), id_q = ab_idesc(DH, 1);
            auto issue_first = [&](int j) {
                const int st = j & 1;
                mbar_wait(bb.st_full + st, (j >> 1) & 1);
                tc_fence_after();
                ab_mma_kk<DH>(tS, smem_u32(sQ), smem_u32(sStage + st * 2 * OP), id_s);          // S  = Q  K^T
                ab_mma_kk<DH>(tDP, smem_u32(sDO), smem_u32(sStage + st * 2 * OP + OP), id_s);   // dP = dO V^T
                umma_commit(bb.sdp_full);
            };
            mbar_wait(bb.r_full, 0);
            issue_first(0);
            for (int j = 0; j < nkb; ++j) {
                if (j + 1 < nkb) {
                    mbar_wait(bb.s_free, j & 1);
                    issue_first(j + 1);
                }
                const int st = j & 1;
                mbar_wait(bb.ds_full, j & 1);
                tc_fence_after();
                ab_mma_ts<DH>(tDQ, tDS, smem_u32(sStage + st * 2 * OP), id_q, j > 0);            // dQ += dS K
                umma_commit(bb.st_empty + st);
                umma_commit(bb.ds_empty);
            }
        }
    } else {
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int t = q0 + row;
        const long long bh = (long long)b * H + h;
        const float L2 = t < T ? lse2[bh * T + t] : INFINITY;   // rows past T: P = 0
        const float Dq = t < T ? delta[bh * T + t] : 0.f;
        const int Tw = (T + 31) >> 5;
        const uint32_t* krow = keep ? keep + (bh * T + (t < T ? t : T - 1)) * Tw : nullptr;
        for (int j = 0; j < nkb; ++j) {
            float s0[32], s1[32], p0[32], p1[32];
            mbar_wait(bb.sdp_full, j & 1);
            tc_fence_after();
            tmem_ld32(tS + lane_off + half * 64, s0);
            tmem_ld32(tS + lane_off + half * 64 + 32, s1);
            tmem_ld32(tDP + lane_off + half * 64, p0);
            tmem_ld32(tDP + lane_off + half * 64 + 32, p1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bb.s_free);
            const int kbase = j * AB_T + half * 64;
            uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu;
            if (krow) {
                const int wi = kbase >> 5;
                kw0 = wi < Tw ? krow[wi] : 0u;
                kw1 = wi + 1 < Tw ? krow[wi + 1] : 0u;
            }
            uint32_t w0[16], w1[16];
            auto half_row = [&](const float (&s)[32], const float (&dp)[32], uint32_t kw, int kb0, uint32_t (&w)[16]) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float ds[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float pr = (kb0 + i + e < len) ? ab_ex2(fmaf(s[i + e], scale_log2, -L2)) : 0.f;
                        const float dpe = ((kw >> (i + e)) & 1u) ? dp[i + e] * keep_scale : 0.f;
                        ds[e] = scale * pr * (dpe - Dq);
                    }
                    w[i >> 1] = pack_bf16(ds[0], ds[1]);
                }
            };
            half_row(s0, p0, kw0, kbase, w0);
            half_row(s1, p1, kw1, kbase + 32, w1);
            if (j > 0) {
                mbar_wait(bb.ds_empty, (j - 1) & 1);
                tc_fence_after();
            }
            tmem_st16(tDS + lane_off + half * 32, w0);
            tmem_st16(tDS + lane_off + half * 32 + 16, w1);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bb.ds_full);
        }
        if (nkb > 0) {
            mbar_wait(bb.ds_empty, (nkb - 1) & 1);
            tc_fence_after();
        }
        ab_store_half<DH>(tDQ + lane_off + half * (DH / 2),
                          dqkv + ((long long)b * T + (t < T ? t : 0)) * (3 * H * DH) + qcol + half * (DH / 2), nkb > 0, t < T);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------- dK, dV
template <int DH>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const float* __restrict__ lse2, const float* __restrict__ delta, const int* __restrict__ lengths,
                    const uint32_t* __restrict__ keep, float keep_scale, __nv_bfloat16* __restrict__ dqkv, int T, int H,
                    float scale, float scale_log2) {
    constexpr int OP = AbCfg<DH>::OP;
    // dh 64: P^T / dS^T have their own TMEM columns, so S^T / dP^T of the next query block can be issued while this
    // block's softmax runs.  dh 96 needs 2 x 96 accumulator columns: P^T / dS^T then alias S^T / dP^T and blocks run
    // strictly one after the other.
    constexpr bool EARLY = DH == 64;
    extern __shared__ uint8_t ab_smem_raw[];
    uint8_t* smem = ab_smem_raw + ((1024u - (smem_u32(ab_smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS
    uint8_t* sK = smem;                     // resident: K tile, V tile
    uint8_t* sV = smem + OP;
    uint8_t* sStage = smem + 2 * OP;        // 2 stages x (Q, dO)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * OP);
    AbBars bb{bars + 0, bars + 1, bars + 3, bars + 5, bars + 6, bars + 7, bars + 8};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* colv = reinterpret_cast<float*>(bars + 32);   // [2 stages][L2 (128) | D (128)]
    uint32_t* kws = reinterpret_cast<uint32_t*>(colv + 2 * 2 * AB_T);   // [2 stages][128 queries][4 keep words of this key tile]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    // queries of every frame of the padded utterance attend (padded queries are computed by the reference too, LM:333);
    // keys past the length have P = 0, so key tiles past the length produce exact zeros.
    const int nqb = (k0 < len) ? (T + AB_T - 1) / AB_T : 0;
    const int qcol = h * DH, kcol = (H + h) * DH, vcol = (2 * H + h) * DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(bb.r_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bb.st_full + i, 1);
            mbar_init(bb.st_empty + i, 1);
        }
        mbar_init(bb.sdp_full, 1);
        mbar_init(bb.s_free, 256);
        mbar_init(bb.ds_full, 256);
        mbar_init(bb.ds_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM dh 64: S^T [0,128) | dP^T [128,256) | P^T bf16x2 [256,320) | dS^T bf16x2 [320,384) | dV [384,448) | dK [448,512)
    //      dh 96: S^T [0,128) | dP^T [128,256) | P^T over [0,64) | dS^T over [128,192) | dV [256,352) | dK [352,448)
    const uint32_t tS = tmem_base, tDP = tmem_base + 128;
    const uint32_t tP = EARLY ? tmem_base + 256 : tmem_base, tDS = EARLY ? tmem_base + 320 : tmem_base + 128;
    const uint32_t tDV = EARLY ? tmem_base + 384 : tmem_base + 256, tDK = tDV + DH;

    if (warp == 0) {
        if (lane == 0 && nqb > 0) {
            mbar_expect_tx(bb.r_full, 2 * OP);
            ab_load<DH>(&tmQKV, bb.r_full, sK, kcol, k0, b);
            ab_load<DH>(&tmQKV, bb.r_full, sV, vcol, k0, b);
            for (int i = 0; i < nqb; ++i) {
                const int st = i & 1;
                mbar_wait(bb.st_empty + st, ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(bb.st_full + st, 2 * OP);
                ab_load<DH>(&tmQKV, bb.st_full + st, sStage + st * 2 * OP, qcol, i * AB_T, b);
                ab_load<DH>(&tmDO, bb.st_full + st, sStage + st * 2 * OP + OP, h * DH, i * AB_T, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nqb > 0) {
            const uint32_t id_s = ab_idesc(AB_T, 0), id_o = ab_idesc(DH, 1);
            auto issue_first = [&](int i) {
                const int st = i & 1;
                mbar_wait(bb.st_full + st, (i >> 1) & 1);
                tc_fence_after();
                ab_mma_kk<DH>(tS, smem_u32(sK), smem_u32(sStage + st * 2 * OP), id_s);         // S^T  = K Q^T
                ab_mma_kk<DH>(tDP, smem_u32(sV), smem_u32(sStage + st * 2 * OP + OP), id_s);   // dP^T = V dO^T
                umma_commit(bb.sdp_full);
            };
            mbar_wait(bb.r_full, 0);
            issue_first(0);
            for (int i = 0; i < nqb; ++i) {
                if (EARLY && i + 1 < nqb) {
                    mbar_wait(bb.s_free, i & 1);
                    issue_first(i + 1);
                }
                const int st = i & 1;
                mbar_wait(bb.ds_full, i & 1);
                tc_fence_after();
                ab_mma_ts<DH>(tDV, tP, smem_u32(sStage + st * 2 * OP + OP), id_o, i > 0);      // dV += P^T dO
                ab_mma_ts<DH>(tDK, tDS, smem_u32(sStage + st * 2 * OP), id_o, i > 0);          // dK += dS^T Q
                umma_commit(bb.st_empty + st);
                umma_commit(bb.ds_empty);
                if (!EARLY && i + 1 < nqb) {
                    mbar_wait(bb.ds_empty, i & 1);   // P^T / dS^T alias the next block's S^T / dP^T
                    issue_first(i + 1);
                }
            }
        }
    } else {
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;                 // key row of this thread
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int key = k0 + row;
        const bool key_ok = key < len;
        const long long bh = (long long)b * H + h;
        const int Tw = (T + 31) >> 5;
        const int st_tid = threadIdx.x - 64;            // 0..255 among the softmax threads
        // per-query scalars (L2, D) and keep words of a query block: fetched one block ahead into registers so their
        // global-memory latency hides under the previous block's math, then published through smem
        float pf_v = 0.f;
        uint32_t pf_k[2] = {0u, 0u};
        auto prefetch = [&](int i) {
            const int qb0 = i * AB_T;
            const int qq = qb0 + (st_tid & 127);
            if (st_tid < 128) pf_v = qq < T ? lse2[bh * T + qq] : INFINITY;   // L2 = +inf for queries past T: P = 0
            else pf_v = qq < T ? delta[bh * T + qq] : 0.f;
            if (keep) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int idx = st_tid * 2 + e, qi = idx >> 2, w = idx & 3;
                    const int q2 = qb0 + qi, wi = (k0 >> 5) + w;
                    pf_k[e] = wi < Tw ? keep[(bh * T + (q2 < T ? q2 : T - 1)) * Tw + wi] : 0u;
                }
            }
        };
        if (nqb > 0) prefetch(0);
        for (int i = 0; i < nqb; ++i) {
            float* cv = colv + (i & 1) * 2 * AB_T;
            uint32_t* kw = kws + (i & 1) * AB_T * 4;
            cv[st_tid] = pf_v;
            kw[st_tid * 2] = pf_k[0];
            kw[st_tid * 2 + 1] = pf_k[1];
            named_bar_sync(1, 256);
            if (i + 1 < nqb) prefetch(i + 1);
            float s0[32], s1[32], p0[32], p1[32];
            mbar_wait(bb.sdp_full, i & 1);
            tc_fence_after();
            tmem_ld32(tS + lane_off + half * 64, s0);
            tmem_ld32(tS + lane_off + half * 64 + 32, s1);
            tmem_ld32(tDP + lane_off + half * 64, p0);
            tmem_ld32(tDP + lane_off + half * 64 + 32, p1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bb.s_free);
            uint32_t wp0[16], wp1[16], wd0[16], wd1[16];
            auto half_row = [&](const float (&s)[32], const float (&dp)[32], int qoff, uint32_t (&wp)[16], uint32_t (&wd)[16]) {
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    float pd[2], ds[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int qi = qoff + c + e;          // query index inside the block
                        const float L2 = cv[qi], Dq = cv[AB_T + qi];
                        const float pr = key_ok ? ab_ex2(fmaf(s[c + e], scale_log2, -L2)) : 0.f;
                        float kf = keep_scale;
                        if (keep) kf = ((kw[qi * 4 + qd] >> lane) & 1u) ? keep_scale : 0.f;
                        pd[e] = pr * kf;
                        ds[e] = scale * pr * (dp[c + e] * kf - Dq);
                    }
                    wp[c >> 1] = pack_bf16(pd[0], pd[1]);
                    wd[c >> 1] = pack_bf16(ds[0], ds[1]);
                }
            };
            half_row(s0, p0, half * 64, wp0, wd0);
            half_row(s1, p1, half * 64 + 32, wp1, wd1);
            if (EARLY) {
                if (i > 0) {
                    mbar_wait(bb.ds_empty, (i - 1) & 1);
                    tc_fence_after();
                }
            } else {
                named_bar_sync(2, 256);   // the partner thread of this row has read the S^T / dP^T columns P^T / dS^T overwrite
            }
            tmem_st16(tP + lane_off + half * 32, wp0);
            tmem_st16(tP + lane_off + half * 32 + 16, wp1);
            tmem_st16(tDS + lane_off + half * 32, wd0);
            tmem_st16(tDS + lane_off + half * 32 + 16, wd1);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bb.ds_full);
        }
        if (nqb > 0) {
            mbar_wait(bb.ds_empty, (nqb - 1) & 1);
            tc_fence_after();
        }
        __nv_bfloat16* base = dqkv + ((long long)b * T + (key < T ? key : 0)) * (3 * H * DH) + half * (DH / 2);
        ab_store_half<DH>(tDK + lane_off + half * (DH / 2), base + kcol, nqb > 0, key < T);
        ab_store_half<DH>(tDV + lane_off + half * (DH / 2), base + vcol, nqb > 0, key < T);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int DH>
static int launch_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, const int32_t* lengths,
                                const uint32_t* keep_bits, float keep_scale, void* dqkv, float* delta_ws, int B, int T, int H,
                                cudaStream_t st) {
    constexpr int SMEM = AbCfg<DH>::SMEM;
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    {
        const long long warps = (long long)B * T * H;
        long long blocks = (warps + 7) / 8;
        if (blocks > 148 * 16) blocks = 148 * 16;
        attn_delta_kernel<DH><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(out),
                                                           reinterpret_cast<const __nv_bfloat16*>(dout), B, T, H, delta_ws);
        DN_LAUNCH_CHECK();
        count_launch();
    }
    CUtensorMap mq, mdo;
    {
        const int ld = 3 * H * DH;
        cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
        cuuint32_t box[3] = {64, AB_T, 1};
        int r = encode_bf16_map(&mq, qkv, 3, dims, str, box);
        if (r) return r;
    }
    {
        const int ld = H * DH;
        cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
        cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
        cuuint32_t box[3] = {64, AB_T, 1};
        int r = encode_bf16_map(&mdo, dout, 3, dims, str, box);
        if (r) return r;
    }
    dim3 grid((T + AB_T - 1) / AB_T, H, B);
    const float scale = 1.0f / sqrtf((float)DH);
    const float scale_log2 = scale * 1.4426950408889634f;
    attn_bwd_dq_kernel<DH><<<grid, AB_THREADS, SMEM, st>>>(mq, mdo, lse2, delta_ws, lengths, keep_bits, keep_scale,
                                                           reinterpret_cast<__nv_bfloat16*>(dqkv), T, H, scale, scale_log2);
    DN_LAUNCH_CHECK();
    count_launch();
    attn_bwd_dkv_kernel<DH><<<grid, AB_THREADS, SMEM, st>>>(mq, mdo, lse2, delta_ws, lengths, keep_bits, keep_scale,
                                                            reinterpret_cast<__nv_bfloat16*>(dqkv), T, H, scale, scale_log2);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

}  // namespace dn

using namespace dn;

extern "C" int dn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, const int32_t* lengths,
                                const uint32_t* keep_bits, float keep_scale, void* dqkv, float* delta_ws, int32_t B, int32_t T,
                                int32_t H, int32_t dh, void* stream) {
    if (!qkv || !out || !dout || !lse2 || !dqkv || !delta_ws || B <= 0 || T <= 0 || H <= 0) return DN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15)
        return DN_EINVAL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dh == 64)
        return launch_attention_bwd<64>(qkv, out, dout, lse2, lengths, keep_bits, keep_scale, dqkv, delta_ws, B, T, H, st);
    if (dh == 96)
        return launch_attention_bwd<96>(qkv, out, dout, lse2, lengths, keep_bits, keep_scale, dqkv, delta_ws, B, T, H, st);
    return DN_EINVAL;
}
