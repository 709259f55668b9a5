// dn_attention (dh = 64): flash attention on tcgen05 tensor cores with S and the per-block P.V product in TMEM.
//
// One CTA = 128 queries of one (utterance, head); 2 CTAs per SM hide each other's softmax / MMA latency.
//   warp 0      TMA producer : Q once, then K_j / V_j tiles (box {64 dh, 128 frames, 1 utt} of the [3*H*dh, T, B] map)
//   warp 1      MMA issuer   : S_j = Q K_j^T  (M128 N128 K64, both K-major)  ->  TMEM cols [0,128)
//                              O_j = P_j V_j  (M128 N64 K128, A = P from smem, B = V MN-major) -> TMEM cols [128,192)
//   warps 2..9  softmax      : two threads per query row (tcgen05.ld 32x32b; warps w and w+4 share rows and split the
//                              128 key columns / 64 output columns): two passes over S_j in TMEM (row max with one
//                              smem exchange, then exp2 / row sum / bf16 P written to smem in the UMMA 128B-swizzle
//                              layout, packed f32x2 FFMA2/FADD2 + FMNMX3 + MUFU.EX2); running (m, l) and the fp32
//                              output half-row live in registers: O = O * alpha_j + O_j after each block.
// Keys j >= lengths[b] get exactly zero weight (LM:333-335); key blocks past the length are skipped.
#include "common.cuh"

namespace dn {

struct FalseTag { static constexpr bool value = false; };
struct TrueTag { static constexpr bool value = true; };

constexpr int TA_BM = 128, TA_BN = 128, TA_DH = 64;
constexpr int TA_THREADS = 320;                           // TMA warp + MMA warp + 8 softmax warps
constexpr int TA_TILE = TA_BM * TA_DH * 2;                  // 16 KB: Q, K, V tiles and each 64-key half of P
constexpr int TA_SMEM = 5 * TA_TILE + 1024 + 128 + 6 * 128 * 4;  // Q, K, V, P(2) + align slack + barriers + exchange

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
// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): halves the FMA-pipe instruction count of the softmax
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// instruction descriptor: bf16 x bf16 -> fp32, M = 128, runtime N, A K-major, B K-major (0) or MN-major (1)
__device__ __forceinline__ uint32_t ta_idesc(uint32_t n, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(TA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out,
                    const int* __restrict__ lengths, int T, int H, float scale_log2) {
    extern __shared__ uint8_t ta_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ta_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TA_TILE;
    uint8_t* sV = smem + 2 * TA_TILE;
    uint8_t* sP = smem + 3 * TA_TILE;  // two 64-key halves, 16 KB each
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TA_TILE);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 2;
    uint64_t* v_full = bars + 3;
    uint64_t* v_empty = bars + 4;
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint64_t* o_empty = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* xmax = reinterpret_cast<float*>(bars + 16);   // [2 parities][2 halves][128 rows] partial row maxima
    float* lsum = xmax + 4 * TA_BM;                      // [2 halves][128 rows] partial row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TA_BM, h = blockIdx.y, b = blockIdx.z;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    const int nkb = (len + TA_BN - 1) / TA_BN;
    const int qcol = h * TA_DH, kcol = (H + h) * TA_DH, vcol = (2 * H + h) * TA_DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 256);
        mbar_init(o_full, 1);
        mbar_init(o_empty, 256);
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
    const uint32_t tS = tmem_base, tO = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            mbar_expect_tx(q_full, TA_TILE);
            tma_load_3d(&tmQKV, q_full, sQ, qcol, q0, b);
            for (int j = 0; j < nkb; ++j) {
                const uint32_t ph = j & 1;
                mbar_wait(k_empty, ph ^ 1);
                mbar_expect_tx(k_full, TA_TILE);
                tma_load_3d(&tmQKV, k_full, sK, kcol, j * TA_BN, b);
                mbar_wait(v_empty, ph ^ 1);
                mbar_expect_tx(v_full, TA_TILE);
                tma_load_3d(&tmQKV, v_full, sV, vcol, j * TA_BN, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t id_s = ta_idesc(TA_BN, 0), id_o = ta_idesc(TA_DH, 1);
            const uint64_t dq = umma_desc_sw128(smem_u32(sQ));
            const uint64_t dk = umma_desc_sw128(smem_u32(sK));
            const uint64_t dv = umma_desc_sw128(smem_u32(sV));   // MN-major: 8-key groups 1024 B apart (SBO)
            const uint64_t dp = umma_desc_sw128(smem_u32(sP));
            auto issue_s = [&](int j) {
                mbar_wait(k_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < TA_DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, id_s, k > 0);
                umma_commit(k_empty);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkb; ++j) {
                mbar_wait(p_full, j & 1);            // P_j written, S_j fully read
                if (j > 0) mbar_wait(o_empty, (j - 1) & 1);
                mbar_wait(v_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < TA_BN / 16; ++k) {
                    // A = P: 64-key halves 16 KB apart, 32 B per k-step inside a half
                    const uint64_t a = dp + (uint64_t)(((k >> 2) * TA_TILE + (k & 3) * 32) >> 4);
                    // B = V (MN-major): 16 keys = 16 rows of 128 B per k-step
                    const uint64_t bb = dv + (uint64_t)((k * 16 * 128) >> 4);
                    umma_bf16(tO, a, bb, id_o, k > 0);
                }
                umma_commit(v_empty);
                umma_commit(o_full);
                if (j + 1 < nkb) issue_s(j + 1);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax + output: two threads per row
        // warps 2..5 own key columns [0,64) of each block and output columns [0,32); warps 6..9 the other halves.
        // (A warp may only touch TMEM lanes 32*(warp%4)..+31, so warps w and w+4 share rows and split columns.)
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        uint64_t o2[TA_DH / 4];  // this thread's 32 fp32 output columns as packed f32x2 pairs (FFMA2)
#pragma unroll
        for (int i = 0; i < TA_DH / 4; ++i) o2[i] = 0ull;
        float m = -INFINITY, l = 0.f;
        uint8_t* prow = sP + half * TA_TILE + row * 128;  // my 64-key half of the P tile
        const int sw = row & 7;
        const uint64_t scale2 = pack2(scale_log2, scale_log2);
        // Only the last key block of an utterance can contain masked keys: the block body is instantiated twice so
        // the hot (unmasked) copy carries no predicated compare/select instructions at all.
        auto block = [&](const int j, auto masked_tag) {
            constexpr bool masked = decltype(masked_tag)::value;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kbase = j * TA_BN + half * 64;
            // pass 1: max over my 64 columns, then exchange with the partner thread of this row
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float s[32];
                tmem_ld32(tS + lane_off + half * 64 + c * 32, s);
                tmem_ld_wait();
                if constexpr (masked) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) s[i] = (kbase + c * 32 + i < len) ? s[i] : -INFINITY;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) mx[(i >> 1) & 3] = max3(mx[(i >> 1) & 3], s[i], s[i + 1]);
            }
            float mxx = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            xmax[((j & 1) * 2 + half) * TA_BM + row] = mxx;
            named_bar_sync(1, 256);
            mxx = fmaxf(mxx, xmax[((j & 1) * 2 + (half ^ 1)) * TA_BM + row]);
            const float mn = fmaxf(m, mxx * scale_log2);   // finite: every processed block has a valid key
            const float alpha = ex2_approx(m - mn);
            m = mn;
            const uint64_t nmn2 = pack2(-mn, -mn);
            // pass 2: p = exp2(s * scale - m), partial row sum, bf16 P into the swizzled A-operand tile
            uint64_t rs2[2] = {0ull, 0ull};
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float s[32];
                tmem_ld32(tS + lane_off + half * 64 + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        float e0, e1;
                        unpack2(ffma2(pack2(s[g * 8 + i], s[g * 8 + i + 1]), scale2, nmn2), e0, e1);
                        e0 = ex2_approx(e0);
                        e1 = ex2_approx(e1);
                        if constexpr (masked) {
                            const int key = kbase + c * 32 + g * 8 + i;
                            e0 = (key < len) ? e0 : 0.f;
                            e1 = (key + 1 < len) ? e1 : 0.f;
                        }
                        rs2[(i >> 1) & 1] = fadd2(rs2[(i >> 1) & 1], pack2(e0, e1));
                        w[i >> 1] = pack_bf16(e0, e1);
                    }
                    const int chunk = c * 4 + g;  // 16-byte chunk (8 keys) within my 64-key half row
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            float r0, r1, r2, r3;
            unpack2(rs2[0], r0, r1);
            unpack2(rs2[1], r2, r3);
            l = fmaf(l, alpha, (r0 + r1) + (r2 + r3));
            tc_fence_before();          // S_j reads done before the next S MMA overwrites it
            fence_proxy_async_smem();   // generic-proxy P stores -> visible to the tensor core (async proxy)
            mbar_arrive(p_full);
            // accumulate my 32 columns of O_j
            const uint64_t alpha2 = pack2(alpha, alpha);
            mbar_wait(o_full, j & 1);
            tc_fence_after();
            {
                float pv[32];
                tmem_ld32(tO + lane_off + half * 32, pv);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) o2[i >> 1] = ffma2(o2[i >> 1], alpha2, pack2(pv[i], pv[i + 1]));
            }
            tc_fence_before();
            mbar_arrive(o_empty);
        };
        const int n_full = len / TA_BN;  // key blocks without any masked key
        for (int j = 0; j < n_full; ++j) block(j, FalseTag{});
        if (n_full < nkb) block(n_full, TrueTag{});
        // combine the two partial row sums, normalise, store my 32 output columns
        lsum[half * TA_BM + row] = l;
        named_bar_sync(2, 256);
        l += lsum[(half ^ 1) * TA_BM + row];
        const int t = q0 + row;
        if (t < T) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            __nv_bfloat16* op = out + ((long long)b * T + t) * (H * TA_DH) + h * TA_DH + half * 32;
#pragma unroll
            for (int i = 0; i < TA_DH / 4; i += 4) {
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a, bb;
                    unpack2(o2[i + k], a, bb);
                    w[k] = pack_bf16(a * inv, bb * inv);
                }
                *reinterpret_cast<uint4*>(op + i * 2) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);

int launch_attention_tc(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        attr_set = true;
    }
    CUtensorMap m;
    const int ld = 3 * H * TA_DH;
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {TA_DH, TA_BM, 1};
    int r = encode_bf16_map(&m, qkv, 3, dims, str, box);
    if (r) return r;
    dim3 grid((T + TA_BM - 1) / TA_BM, H, B);
    const float scale_log2 = (1.0f / sqrtf((float)TA_DH)) * 1.4426950408889634f;
    attention_tc_kernel<<<grid, TA_THREADS, TA_SMEM, st>>>(m, reinterpret_cast<__nv_bfloat16*>(out), lengths, T, H,
                                                          scale_log2);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

}  // namespace dn
