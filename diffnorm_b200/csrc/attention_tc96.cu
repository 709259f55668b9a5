// dn_attention (dh = 96): the VAE decoder's attention (LM:1084-1094: 8 heads x 96) on tcgen05 tensor cores — S, P and
// O in TMEM like the dh = 64 kernel (attention_tc.cu), with the head dimension spread over TWO 64-column TMA boxes per
// operand (the second box's upper 32 columns are loaded but never consumed), so:
//   S_j = Q K_j^T     : 6 k-steps of 16, k-step k lives in box k / 4                      -> TMEM cols [0,128)
//   O  += P_j V_j     : M128 N96 K128, A = P from TMEM, B = V MN-major, its two 64-column atoms one box apart (LBO)
//                                                                                         -> TMEM cols [128,224)
//   P_j (packed 16 bit)                                                                   -> TMEM cols [256,320)
// One CTA per SM (512 TMEM columns, 5 x 32 KB of operand tiles).  Once per pass in the normalization path (6 layers),
// and the attention of VAE training (TRAIN: dropout bits + the saved row statistic).  Formats: bf16, or fp16 q/k/v/P
// with the output written as a split-precision bf16 pair [hi | lo] (the precise decoder, DESIGN.md "operand formats").
#include "common.cuh"

namespace dn {

constexpr int T9_BM = 128, T9_BN = 128, T9_DH = 96;
constexpr int T9_THREADS = 320;                       // TMA warp + MMA warp + 8 softmax warps
constexpr int T9_BOX = T9_BM * 64 * 2;                // 16 KB: one {64 columns, 128 rows} TMA box
constexpr int T9_OP = 2 * T9_BOX;                     // one operand (Q, or a K / V stage): two boxes
constexpr int T9_SMEM = 5 * T9_OP + 1024 + 128 + 6 * 128 * 4;

__device__ __forceinline__ float t9_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 16-bit x 16-bit -> fp32, M = 128, runtime N, A K-major (smem) or TMEM, B K-major (0) or MN-major (1)
__device__ __forceinline__ uint32_t t9_idesc(uint32_t n, uint32_t b_mn_major, bool f16) {
    const uint32_t fm = f16 ? 0u : 1u;
    return (1u << 4) | (fm << 7) | (fm << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// MN-major SW128 operand whose 64-column atoms are T9_BOX bytes apart (LBO); 8-row groups 1 KB apart (SBO)
__device__ __forceinline__ uint64_t t9_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((T9_BOX >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void t9_load(const CUtensorMap* m, uint64_t* bar, uint8_t* dst, int col, int row0, int b) {
    tma_load_3d(m, bar, dst, col, row0, b);
    tma_load_3d(m, bar, dst + T9_BOX, col + 64, row0, b);
}

struct T9False { static constexpr bool value = false; };
struct T9True { static constexpr bool value = true; };

// out_lo_col != 0: out is a split-precision bf16 pair with row stride 2 * out_lo_col (hi at column c, lo at out_lo_col + c)
template <bool TRAIN, bool F16>
__global__ void __launch_bounds__(T9_THREADS, 1)
attention_tc96_kernel(const __grid_constant__ CUtensorMap tmQKV, uint16_t* __restrict__ out, const int* __restrict__ lengths,
                      int T, int H, float scale_log2, float* __restrict__ lse2, const uint32_t* __restrict__ keep,
                      float keep_scale, int out_lo_col) {
    extern __shared__ uint8_t t9_smem_raw[];
    uint8_t* smem = t9_smem_raw + ((1024u - (smem_u32(t9_smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + T9_OP;        // two stages
    uint8_t* sV = smem + 3 * T9_OP;    // two stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * T9_OP);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // S_j complete in TMEM
    uint64_t* s_free = bars + 10;   // S_j copied to registers by all softmax threads
    uint64_t* p_full = bars + 11;   // P_j in TMEM (and O rescaled if needed)
    uint64_t* p_empty = bars + 12;  // PV_j complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* xmax = reinterpret_cast<float*>(bars + 16);   // [2 parities][2 halves][128 rows]
    float* lsum = xmax + 4 * T9_BM;                      // [2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * T9_BM, h = blockIdx.y, b = blockIdx.z;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    const int nkb = (len + T9_BN - 1) / T9_BN;
    const int qcol = h * T9_DH, kcol = (H + h) * T9_DH, vcol = (2 * H + h) * T9_DH;

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
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 256;

    if (warp == 0) {
        if (lane == 0 && nkb > 0) {
            mbar_expect_tx(q_full, T9_OP);
            t9_load(&tmQKV, q_full, sQ, qcol, q0, b);
            for (int j = 0; j < nkb; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(k_empty + st, ph ^ 1);
                mbar_expect_tx(k_full + st, T9_OP);
                t9_load(&tmQKV, k_full + st, sK + st * T9_OP, kcol, j * T9_BN, b);
                mbar_wait(v_empty + st, ph ^ 1);
                mbar_expect_tx(v_full + st, T9_OP);
                t9_load(&tmQKV, v_full + st, sV + st * T9_OP, vcol, j * T9_BN, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const uint32_t id_s = t9_idesc(T9_BN, 0, F16), id_o = t9_idesc(T9_DH, 1, F16);
            auto issue_s = [&](int j) {
                const int st = j & 1;
                mbar_wait(k_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK + st * T9_OP);
#pragma unroll
                for (int k = 0; k < T9_DH / 16; ++k) {
                    const uint32_t off = (k >> 2) * T9_BOX + (k & 3) * 32;   // k-step k: box k / 4, 32 B per step inside the atom
                    umma_bf16(tS, umma_desc_sw128(aq + off), umma_desc_sw128(ak + off), id_s, k > 0);
                }
                umma_commit(k_empty + st);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nkb; ++j) {
                if (j + 1 < nkb) {
                    mbar_wait(s_free, j & 1);
                    issue_s(j + 1);
                }
                const int st = j & 1;
                mbar_wait(p_full, j & 1);
                mbar_wait(v_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint32_t av = smem_u32(sV + st * T9_OP);
#pragma unroll
                for (int k = 0; k < T9_BN / 16; ++k)   // A = P from TMEM (16 keys = 8 columns per k-step), B = V: 16 key rows per k-step
                    umma_bf16_ts(tO, tP + 8 * k, t9_desc_mn(av + k * 16 * 128), id_o, (j > 0) || (k > 0));
                umma_commit(v_empty + st);
                umma_commit(p_empty);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax: two threads per query row
        // warps 2..5 own key columns [0,64) of each block and output columns [0,48); warps 6..9 the other halves
        constexpr float RESCALE_LOG2 = 8.f;
        constexpr int OC = T9_DH / 2;   // 48 output columns per thread
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
            const int kbase = j * T9_BN + half * 64;
            if constexpr (masked) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    s0[i] = (kbase + i < len) ? s0[i] : -INFINITY;
                    s1[i] = (kbase + 32 + i < len) ? s1[i] : -INFINITY;
                }
            }
            float mxx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) mxx = fmaxf(mxx, fmaxf(s0[i], s1[i]));
            xmax[((j & 1) * 2 + half) * T9_BM + row] = mxx;
            named_bar_sync(1, 256);
            mxx = fmaxf(mxx, xmax[((j & 1) * 2 + (half ^ 1)) * T9_BM + row]);
            const float mxs = mxx * scale_log2;            // finite: every processed block has a valid key
            const bool raise = mxs > m + RESCALE_LOG2;     // j == 0: m = -inf -> true
            const float m_old = m;
            if (raise) {
                l *= t9_ex2(m - mxs);
                m = mxs;
            }
            const uint64_t nm2 = pack2(-m, -m);
            float rs = 0.f;
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
                    e0 = t9_ex2(e0);   // masked keys: ex2(-inf) = +0
                    e1 = t9_ex2(e1);
                    rs += e0 + e1;
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
                    const float a = raise ? t9_ex2(m_old - m) : 1.f;
#pragma unroll 1
                    for (int c8 = 0; c8 < OC; c8 += 8) {
                        float o[8];
                        tmem_ld8(tO + lane_off + half * OC + c8, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] *= a;
                        tmem_st8(tO + lane_off + half * OC + c8, o);
                    }
                }
            }
            tmem_st16(tP + lane_off + half * 32, w0);
            half_row(s1, w1, kw1);
            tmem_st16(tP + lane_off + half * 32 + 16, w1);
            l += rs;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full);
        };
        const int n_full = len / T9_BN;
        for (int j = 0; j < n_full; ++j) block(j, T9False{});
        if (n_full < nkb) block(n_full, T9True{});
        lsum[half * T9_BM + row] = l;
        named_bar_sync(2, 256);
        l += lsum[(half ^ 1) * T9_BM + row];
        const int t = q0 + row;
        if (TRAIN && lse2 && half == 0 && t < T) lse2[((long long)b * H + h) * T + t] = m + log2f(l);
        const int ldo = out_lo_col ? 2 * out_lo_col : H * T9_DH;
        uint16_t* op = out + ((long long)b * T + (t < T ? t : 0)) * ldo + h * T9_DH + half * OC;
        float o[OC];
        if (nkb > 0) {
            mbar_wait(p_empty, (nkb - 1) & 1);   // last PV complete
            tc_fence_after();
            float a[32], c[16];
            tmem_ld32(tO + lane_off + half * OC, a);
            tmem_ld16(tO + lane_off + half * OC + 32, c);
            tmem_ld_wait();
            const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = a[i] * inv;
#pragma unroll
            for (int i = 0; i < 16; ++i) o[32 + i] = c[i] * inv;
        } else {
#pragma unroll
            for (int i = 0; i < OC; ++i) o[i] = 0.f;
        }
        if (t < T) {
#pragma unroll
            for (int i = 0; i < OC; i += 8) {
                if (out_lo_col) {
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split_bf16(o[i + 2 * q], o[i + 2 * q + 1], hh[q], ll[q]);
                    *reinterpret_cast<uint4*>(op + i) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<uint4*>(op + out_lo_col + i) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                } else {
                    *reinterpret_cast<uint4*>(op + i) =
                        make_uint4(pack16<F16>(o[i], o[i + 1]), pack16<F16>(o[i + 2], o[i + 3]), pack16<F16>(o[i + 4], o[i + 5]),
                                   pack16<F16>(o[i + 6], o[i + 7]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);

template <bool TRAIN, bool F16>
static int t9_launch(const CUtensorMap& m, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st, float* lse2,
                     const uint32_t* keep, float keep_scale, int out_lo_col) {
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attention_tc96_kernel<TRAIN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, T9_SMEM));
        attr_set = true;
    }
    dim3 grid((T + T9_BM - 1) / T9_BM, H, B);
    const float scale_log2 = (1.0f / sqrtf((float)T9_DH)) * 1.4426950408889634f;
    DN_CUDA_OK(launch_ex(attention_tc96_kernel<TRAIN, F16>, grid, T9_THREADS, T9_SMEM, st, 1, m, reinterpret_cast<uint16_t*>(out),
                         lengths, T, H, scale_log2, lse2, keep, keep_scale, out_lo_col));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

int launch_attention_tc96(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st, float* lse2,
                          const uint32_t* keep, float keep_scale, bool train, bool f16, int out_lo_col) {
    CUtensorMap m;
    const int ld = 3 * H * T9_DH;
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
    cuuint32_t box[3] = {64, T9_BM, 1};
    int r = encode_bf16_map(&m, qkv, 3, dims, str, box);
    if (r) return r;
    if (train) {
        if (f16 || out_lo_col) return DN_EINVAL;
        return t9_launch<true, false>(m, out, lengths, B, T, H, st, lse2, keep, keep_scale, 0);
    }
    if (f16) return t9_launch<false, true>(m, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, out_lo_col);
    return t9_launch<false, false>(m, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, out_lo_col);
}

}  // namespace dn
