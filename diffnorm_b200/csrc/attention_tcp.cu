// dn_attention (dh = 64), PERSISTENT form of the tcgen05 flash-attention kernel of attention_tc.cu (inference only).
//
// attention_tc.cu launches one CTA per (128 queries, head, utterance): 4096 CTAs of ~8 key blocks each at config 2, and every
// one of them pays TMEM allocation, barrier initialisation, descriptor prefetch and the latency of its first Q / K loads
// (~8 % of its life) before its tensor core sees work.  Here 2 CTAs per SM stay resident and walk a static tile list
// (query tile fastest, so CTAs running at the same time share the K / V of one (utterance, head) in L2); TMEM, barriers and
// pipeline state live across tiles:
//   * Q is double buffered: the producer warp loads the NEXT tile's Q (and runs ahead into its K / V stages) while the
//     current tile is still in its last key blocks;
//   * the MMA warp issues S_0 of the next tile right behind the last P.V of the current one, so the softmax threads find it
//     complete when they return from writing the previous tile's output;
//   * every barrier phase is derived from GLOBAL counters (key blocks processed so far, tiles so far) that the three roles
//     advance identically from the same tile list.
// Softmax: two threads per query row, optimistic exponentiation for key blocks after the first (see attention_tc.cu OPT).
#include <stdlib.h>

#include "common.cuh"

namespace dn {

constexpr int TP_BM = 128, TP_BN = 128, TP_DH = 64;
constexpr int TP_THREADS = 320;                           // TMA warp + MMA warp + 8 softmax warps
constexpr int TP_TILE = TP_BM * TP_DH * 2;                // 16 KB
constexpr int TP_SMEM = 6 * TP_TILE + 1024 + 256 + 6 * 128 * 4;   // Q(2), K(2), V(2) + align slack + barriers + exchange

__device__ __forceinline__ float tp_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tp_max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ uint32_t tp_idesc(uint32_t n, uint32_t b_mn_major, bool f16) {
    const uint32_t fm = f16 ? 0u : 1u;
    return (1u << 4) | (fm << 7) | (fm << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct TpFalse { static constexpr bool value = false; };
struct TpTrue { static constexpr bool value = true; };

template <bool F16>
__global__ void __launch_bounds__(TP_THREADS, 2)
attention_tcp_kernel(const __grid_constant__ CUtensorMap tmQKV, uint16_t* __restrict__ out, const int* __restrict__ lengths,
                     int T, int H, int B, float scale_log2, int stagger_ns) {
    extern __shared__ uint8_t tp_smem_raw[];
    uint8_t* smem = tp_smem_raw + ((1024u - (smem_u32(tp_smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                  // two buffers
    uint8_t* sK = smem + 2 * TP_TILE;    // two stages
    uint8_t* sV = smem + 4 * TP_TILE;    // two stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TP_TILE);
    uint64_t* q_full = bars + 0;    // [2]
    uint64_t* q_empty = bars + 2;   // [2] all S MMAs of the tile that used this Q buffer have completed
    uint64_t* k_full = bars + 4;    // [2]
    uint64_t* k_empty = bars + 6;   // [2]
    uint64_t* v_full = bars + 8;    // [2]
    uint64_t* v_empty = bars + 10;  // [2]
    uint64_t* s_full = bars + 12;   // S of the current key block complete in TMEM
    uint64_t* s_free = bars + 13;   // S copied to registers by all softmax threads
    uint64_t* p_full = bars + 14;   // P in TMEM (and O rescaled if needed)
    uint64_t* p_empty = bars + 15;  // P.V complete: P columns reusable, O stable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    float* xmax = reinterpret_cast<float*>(bars + 32);   // [2 parities][2 halves][128 rows]
    float* lsum = xmax + 4 * TP_BM;                      // [2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqt = (T + TP_BM - 1) / TP_BM;
    const int n_tiles = nqt * H * B;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int i = 0; i < 2; ++i) {
            mbar_init(q_full + i, 1);
            mbar_init(q_empty + i, 1);
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
    const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;

    // tile w -> (query tile, head, utterance): query tile fastest
    auto tile_of = [&](int w, int& q0, int& h, int& b, int& len, int& nkb) {
        q0 = (w % nqt) * TP_BM;
        h = (w / nqt) % H;
        b = w / (nqt * H);
        len = lengths ? lengths[b] : T;
        len = len > T ? T : len;
        nkb = (len + TP_BN - 1) / TP_BN;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            uint32_t g = 0, tq = 0;   // key blocks / non-empty tiles so far
            for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
                int q0, h, b, len, nkb;
                tile_of(w, q0, h, b, len, nkb);
                if (nkb == 0) continue;
                const int qcol = h * TP_DH, kcol = (H + h) * TP_DH, vcol = (2 * H + h) * TP_DH;
                const uint32_t qb = tq & 1;
                mbar_wait(q_empty + qb, ((tq >> 1) & 1) ^ 1);
                mbar_expect_tx(q_full + qb, TP_TILE);
                tma_load_3d(&tmQKV, q_full + qb, sQ + qb * TP_TILE, qcol, q0, b);
                for (int j = 0; j < nkb; ++j, ++g) {
                    const uint32_t st = g & 1, ph = (g >> 1) & 1;
                    mbar_wait(k_empty + st, ph ^ 1);
                    mbar_expect_tx(k_full + st, TP_TILE);
                    tma_load_3d(&tmQKV, k_full + st, sK + st * TP_TILE, kcol, j * TP_BN, b);
                    mbar_wait(v_empty + st, ph ^ 1);
                    mbar_expect_tx(v_full + st, TP_TILE);
                    tma_load_3d(&tmQKV, v_full + st, sV + st * TP_TILE, vcol, j * TP_BN, b);
                }
                ++tq;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            const uint32_t id_s = tp_idesc(TP_BN, 0, F16), id_o = tp_idesc(TP_DH, 1, F16);
            uint32_t g = 0, tq = 0;
            for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
                int q0, h, b, len, nkb;
                tile_of(w, q0, h, b, len, nkb);
                if (nkb == 0) continue;
                const uint32_t qb = tq & 1;
                const uint64_t dq = umma_desc_sw128(smem_u32(sQ + qb * TP_TILE));
                auto issue_s = [&](uint32_t gg, bool last) {   // S of global block gg (its K stage = gg & 1)
                    const uint32_t st = gg & 1;
                    mbar_wait(k_full + st, (gg >> 1) & 1);
                    tc_fence_after();
                    const uint64_t dk = umma_desc_sw128(smem_u32(sK + st * TP_TILE));
#pragma unroll
                    for (int k = 0; k < TP_DH / 16; ++k) umma_bf16(tS, dq + 2 * k, dk + 2 * k, id_s, k > 0);
                    umma_commit(k_empty + st);
                    if (last) umma_commit(q_empty + qb);       // the tile's last read of this Q buffer
                    umma_commit(s_full);
                };
                mbar_wait(q_full + qb, (tq >> 1) & 1);
                if (g > 0) mbar_wait(s_free, (g - 1) & 1);       // S of the previous tile's last block is in registers
                issue_s(g, nkb == 1);
                for (int j = 0; j < nkb; ++j, ++g) {
                    if (j + 1 < nkb) {
                        mbar_wait(s_free, g & 1);
                        issue_s(g + 1, j + 2 == nkb);
                    }
                    const uint32_t st = g & 1;
                    mbar_wait(p_full, g & 1);
                    mbar_wait(v_full + st, (g >> 1) & 1);
                    tc_fence_after();
                    const uint64_t dv = umma_desc_sw128(smem_u32(sV + st * TP_TILE));
#pragma unroll
                    for (int k = 0; k < TP_BN / 16; ++k)
                        umma_bf16_ts(tO, tP + 8 * k, dv + (uint64_t)((k * 16 * 128) >> 4), id_o, (j > 0) || (k > 0));
                    umma_commit(v_empty + st);
                    umma_commit(p_empty);
                }
                ++tq;
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: two threads per query row
        constexpr float RESCALE_LOG2 = 8.f;
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint64_t scale2 = pack2(scale_log2, scale_log2);
        uint32_t g = 0;
        // The two resident CTAs of an SM start together and do identical work per key block: left alone they stay in
        // lockstep, both reading TMEM at the same time and both on the MUFU at the same time.  The second wave of CTAs
        // starts half a block period late so that one CTA's TMEM reads run under the other's exponentials.
        if (stagger_ns > 0 && blockIdx.x >= (gridDim.x + 1) / 2) __nanosleep(stagger_ns);
        for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
            int q0, h, b, len, nkb;
            tile_of(w, q0, h, b, len, nkb);
            const int t = q0 + row;
            uint16_t* op = out + ((long long)b * T + (t < T ? t : 0)) * (H * TP_DH) + h * TP_DH + half * 32;
            if (nkb == 0) {
                if (t < T) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(op + i) = make_uint4(0u, 0u, 0u, 0u);
                }
                continue;
            }
            float m = -INFINITY, l = 0.f;
            // one key block: j = index inside the tile (logic), gg = global block index (barrier phases)
            auto block = [&](const int j, const uint32_t gg, auto masked_tag) {
                constexpr bool masked = decltype(masked_tag)::value;
                float s0[32], s1[32];
                mbar_wait(s_full, gg & 1);
                tc_fence_after();
                tmem_ld32(tS + lane_off + half * 64, s0);
                tmem_ld_wait();
                tmem_ld32(tS + lane_off + half * 64 + 32, s1);   // in flight under the first half's work
                const int kbase = j * TP_BN + half * 64;
                uint64_t nm2 = pack2(-m, -m);
                uint64_t rs2[2] = {0ull, 0ull};
                float mx[2] = {-INFINITY, -INFINITY};
                auto half_max = [&](float (&s)[32], const int k0) {
                    if constexpr (masked) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) s[i] = (k0 + i < len) ? s[i] : -INFINITY;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mx[(i >> 1) & 1] = tp_max3(mx[(i >> 1) & 1], s[i], s[i + 1]);
                };
                auto half_exp = [&](const float (&s)[32], uint32_t (&wv)[16]) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float e0, e1;
                        unpack2(ffma2(pack2(s[i], s[i + 1]), scale2, nm2), e0, e1);
                        e0 = tp_ex2(e0);   // masked keys: ex2(-inf) = +0
                        e1 = tp_ex2(e1);
                        rs2[(i >> 1) & 1] = fadd2(rs2[(i >> 1) & 1], pack2(e0, e1));
                        wv[i >> 1] = pack16<F16>(e0, e1);
                    }
                };
                uint32_t wv[16];
                bool raise;
                float m_old = m;
                if (j == 0) {
                    // first block of a tile: no running max yet -> max first (both halves), then exponentiate
                    half_max(s0, kbase);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(s_free);
                    half_max(s1, kbase + 32);
                    float mxx = fmaxf(mx[0], mx[1]);
                    xmax[((gg & 1) * 2 + half) * TP_BM + row] = mxx;
                    named_bar_sync(1, 256);
                    mxx = fmaxf(mxx, xmax[((gg & 1) * 2 + (half ^ 1)) * TP_BM + row]);
                    m = mxx * scale_log2;                         // finite: every processed block has a valid key
                    nm2 = pack2(-m, -m);
                    half_exp(s0, wv);
                    tmem_st16(tP + lane_off + half * 32, wv);     // P columns are free: the previous tile's last P.V was awaited
                    half_exp(s1, wv);
                    tmem_st16(tP + lane_off + half * 32 + 16, wv);
                    raise = false;
                } else {
                    // optimistic: exponentiate against the running max while the second half of S is still in flight
                    half_max(s0, kbase);
                    half_exp(s0, wv);
                    mbar_wait(p_empty, (gg - 1) & 1);             // P.V of the previous block done: P columns free, O stable
                    tc_fence_after();
                    tmem_st16(tP + lane_off + half * 32, wv);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(s_free);
                    half_max(s1, kbase + 32);
                    half_exp(s1, wv);
                    tmem_st16(tP + lane_off + half * 32 + 16, wv);
                    float mxx = fmaxf(mx[0], mx[1]);
                    xmax[((gg & 1) * 2 + half) * TP_BM + row] = mxx;
                    named_bar_sync(1, 256);
                    mxx = fmaxf(mxx, xmax[((gg & 1) * 2 + (half ^ 1)) * TP_BM + row]);
                    const float mxs = mxx * scale_log2;
                    raise = mxs > m + RESCALE_LOG2;
                    if (__any_sync(0xffffffffu, raise)) {         // rare: redo the block against the new max (warp-uniform)
                        if (raise) {
                            l *= tp_ex2(m - mxs);
                            m = mxs;
                        }
                        nm2 = pack2(-m, -m);
                        rs2[0] = rs2[1] = 0ull;
                        tmem_st_wait();
                        half_exp(s0, wv);
                        tmem_st16(tP + lane_off + half * 32, wv);
                        half_exp(s1, wv);
                        tmem_st16(tP + lane_off + half * 32 + 16, wv);
                        const float a = raise ? tp_ex2(m_old - m) : 1.f;
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
                float r0, r1, r2, r3;
                unpack2(rs2[0], r0, r1);
                unpack2(rs2[1], r2, r3);
                l += (r0 + r1) + (r2 + r3);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_full);
            };
            const int n_full = len / TP_BN;
            for (int j = 0; j < nkb; ++j) {
                if (j < n_full) block(j, g + j, TpFalse{});
                else block(j, g + j, TpTrue{});
            }
            g += nkb;
            // combine the two partial row sums, normalise, store my 32 output columns
            lsum[half * TP_BM + row] = l;
            named_bar_sync(2, 256);
            l += lsum[(half ^ 1) * TP_BM + row];
            mbar_wait(p_empty, (g - 1) & 1);      // the tile's last P.V complete
            tc_fence_after();
            float o[32];
            tmem_ld32(tO + lane_off + half * 32, o);
            tmem_ld_wait();
            tc_fence_before();
            if (t < T) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    *reinterpret_cast<uint4*>(op + i) =
                        make_uint4(pack16<F16>(o[i] * inv, o[i + 1] * inv), pack16<F16>(o[i + 2] * inv, o[i + 3] * inv),
                                   pack16<F16>(o[i + 4] * inv, o[i + 5] * inv), pack16<F16>(o[i + 6] * inv, o[i + 7] * inv));
                }
            }
            named_bar_sync(2, 256);   // lsum is rewritten by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);
int num_sms();

int launch_attention_tcp(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st, bool f16) {
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tcp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tcp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
        attr_set = true;
    }
    CUtensorMap m;
    const int ld = 3 * H * TP_DH;
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {TP_DH, TP_BM, 1};
    int r = encode_bf16_map(&m, qkv, 3, dims, str, box);
    if (r) return r;
    const long long tiles = (long long)((T + TP_BM - 1) / TP_BM) * H * B;
    const int resident = 2 * num_sms();
    const int grid = (int)(tiles < resident ? tiles : resident);
    const float scale_log2 = (1.0f / sqrtf((float)TP_DH)) * 1.4426950408889634f;
    static int stagger = -1;
    if (stagger < 0) {
        const char* e = getenv("DN_ATTN_STAGGER_NS");
        stagger = e ? atoi(e) : 900;
    }
    if (f16)
        DN_CUDA_OK(launch_ex(attention_tcp_kernel<true>, grid, TP_THREADS, TP_SMEM, st, 1, m, reinterpret_cast<uint16_t*>(out), lengths,
                             T, H, B, scale_log2, stagger));
    else
        DN_CUDA_OK(launch_ex(attention_tcp_kernel<false>, grid, TP_THREADS, TP_SMEM, st, 1, m, reinterpret_cast<uint16_t*>(out), lengths,
                             T, H, B, scale_log2, stagger));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

}  // namespace dn
