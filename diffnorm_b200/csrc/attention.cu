// dn_attention: flash-style non-causal multi-head attention with key-padding mask.
// Replaces Attend.forward (LM:299-343) + the head split/merge of Attention.forward (LM:945-949): no N x N
// matrix is ever written to HBM (the reference materialises [B, 8, N, N] fp32 per layer).
// This file: the mma.sync form (64-query x 64-key tiles, bf16 m16n8k16 with fp32 accumulation, online softmax in
// registers in the exp2 domain, cp.async double-buffered K/V) used for the VAE decoder's dh = 96 heads (once per pass)
// and as the A/B reference of the tcgen05/TMEM kernel in attention_tc.cu, which serves the denoiser's dh = 64 heads.
#include <stdlib.h>

#include "common.cuh"

namespace dn {

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
template <bool F16>
__device__ __forceinline__ void mma_16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (F16)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int ATT_BM = 64, ATT_BN = 64, ATT_THREADS = 128;

// F16: q, k, v and the probabilities are fp16 (3 more mantissa bits than bf16; the precise VAE decoder).  out_lo_col != 0:
// the output row is a split-precision bf16 pair [hi | lo at out_lo_col] with row stride 2 * out_lo_col.
template <int DH, bool F16>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                 const int* __restrict__ lengths, int T, int H, float scale_log2, float* __restrict__ lse2,
                 const uint32_t* __restrict__ keep, float keep_scale, int out_lo_col) {
    constexpr int LDS = DH + 8;       // padded smem row (elements): conflict-free ldmatrix
    constexpr int CH = DH / 8;        // 16-byte chunks per row
    constexpr int KS = DH / 16;       // k-steps over the head dim
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sK = sQ + ATT_BM * LDS;
    __nv_bfloat16* sV = sK + 2 * ATT_BN * LDS;

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_BM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int len = lengths ? lengths[b] : T;
    len = len > T ? T : len;
    const int ld = 3 * H * DH;
    const __nv_bfloat16* base = qkv + (long long)b * T * ld;
    const int qoff = h * DH, koff = H * DH + h * DH, voff = 2 * H * DH + h * DH;

    for (int c = tid; c < ATT_BM * CH; c += ATT_THREADS) {
        const int r = c / CH, cc = (c % CH) * 8;
        const int t = q0 + r;
        const bool ok = t < T;
        cp_async16(sQ + r * LDS + cc, base + (long long)(ok ? t : 0) * ld + qoff + cc, ok ? 16 : 0);
    }
    auto load_kv = [&](int stage, int kb) {
        __nv_bfloat16* k = sK + stage * ATT_BN * LDS;
        __nv_bfloat16* v = sV + stage * ATT_BN * LDS;
        for (int c = tid; c < ATT_BN * CH; c += ATT_THREADS) {
            const int r = c / CH, cc = (c % CH) * 8;
            const int key = kb * ATT_BN + r;
            const bool ok = key < len;
            const __nv_bfloat16* src = base + (long long)(ok ? key : 0) * ld;
            cp_async16(k + r * LDS + cc, src + koff + cc, ok ? 16 : 0);
            cp_async16(v + r * LDS + cc, src + voff + cc, ok ? 16 : 0);
        }
    };
    const int nkb = (len + ATT_BN - 1) / ATT_BN;
    if (nkb > 0) load_kv(0, 0);
    cp_async_commit();

    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows lane/4 and lane/4 + 8 of this warp's 16
    uint32_t qf[KS][4];

    for (int kb = 0; kb < nkb; ++kb) {
        if (kb + 1 < nkb) {
            load_kv((kb + 1) & 1, kb + 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) {
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 16 + (lane >> 4) * 8;
                ldsm_x4(qf[kk], sQ + r * LDS + c);
            }
        }
        const __nv_bfloat16* k = sK + (kb & 1) * ATT_BN * LDS;
        const __nv_bfloat16* v = sV + (kb & 1) * ATT_BN * LDS;

        // ---- S = Q K^T : 16 x 64 per warp
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                uint32_t bf[4];
                const int r = nb2 * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(bf, k + r * LDS + c);
                mma_16<F16>(s[nb2 * 2], qf[kk], bf[0], bf[1]);
                mma_16<F16>(s[nb2 * 2 + 1], qf[kk], bf[2], bf[3]);
            }
        }
        // ---- mask + online softmax (exp2 domain)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int key = kb * ATT_BN + nb * 8 + (lane & 3) * 2;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = key + (e & 1) < len;
                s[nb][e] = ok ? s[nb][e] * scale_log2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);  // finite: every processed block has a valid key
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];  // P as A fragments for the 4 k16 steps over the 64 keys
        // attention dropout (training form): keep words of this thread's two query rows for the 64 keys of the block
        uint32_t kw0[2] = {0xffffffffu, 0xffffffffu}, kw1[2] = {0xffffffffu, 0xffffffffu};
        if (keep) {
            const int Tw = (T + 31) >> 5;
            const int ra = q0 + warp * 16 + (lane >> 2), rb = ra + 8;
            const uint32_t* k0p = keep + (((long long)b * H + h) * T + (ra < T ? ra : T - 1)) * Tw;
            const uint32_t* k1p = keep + (((long long)b * H + h) * T + (rb < T ? rb : T - 1)) * Tw;
            const int wi = (kb * ATT_BN) >> 5;
            kw0[0] = wi < Tw ? k0p[wi] : 0u;
            kw0[1] = wi + 1 < Tw ? k0p[wi + 1] : 0u;
            kw1[0] = wi < Tw ? k1p[wi] : 0u;
            kw1[1] = wi + 1 < Tw ? k1p[wi + 1] : 0u;
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            float p0 = exp2f(s[nb][0] - mn0), p1 = exp2f(s[nb][1] - mn0);
            float p2 = exp2f(s[nb][2] - mn1), p3 = exp2f(s[nb][3] - mn1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            if (keep) {   // dropout acts on the normalised probabilities: the row sums above stay un-dropped
                const int kk = nb * 8 + (lane & 3) * 2;   // key offset inside the block
                const uint32_t wa = kw0[kk >> 5], wb = kw1[kk >> 5];
                p0 = ((wa >> (kk & 31)) & 1u) ? p0 * keep_scale : 0.f;
                p1 = ((wa >> ((kk + 1) & 31)) & 1u) ? p1 * keep_scale : 0.f;
                p2 = ((wb >> (kk & 31)) & 1u) ? p2 * keep_scale : 0.f;
                p3 = ((wb >> ((kk + 1) & 31)) & 1u) ? p3 * keep_scale : 0.f;
            }
            pf[nb >> 1][(nb & 1) * 2 + 0] = pack16<F16>(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = pack16<F16>(p2, p3);
        }
        l0 = l0 * a0 + rs0;
        l1 = l1 * a1 + rs1;
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) {
            o[i][0] *= a0; o[i][1] *= a0;
            o[i][2] *= a1; o[i][3] *= a1;
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int db2 = 0; db2 < DH / 16; ++db2) {
                uint32_t bf[4];
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = db2 * 16 + (lane >> 4) * 8;
                ldsm_x4_t(bf, v + r * LDS + c);
                mma_16<F16>(o[db2 * 2], pf[kk], bf[0], bf[1]);
                mma_16<F16>(o[db2 * 2 + 1], pf[kk], bf[2], bf[3]);
            }
        }
        __syncthreads();
    }
    if (nkb == 0) cp_async_wait<0>();

    // ---- normalise and store
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    if (lse2 && (lane & 3) == 0) {   // training form: row statistic L2 = m + log2(l) for the backward kernels
        float* lp = lse2 + ((long long)b * H + h) * T;
        if (r0 < T) lp[r0] = m0 + log2f(l0);
        if (r1 < T) lp[r1] = m1 + log2f(l1);
    }
    const int ldo = out_lo_col ? 2 * out_lo_col : H * DH;
    __nv_bfloat16* ob = out + (long long)b * T * ldo + h * DH + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
        uint32_t h0, l0, h1, l1;
        split_bf16(o[i][0] * i0, o[i][1] * i0, h0, l0);
        split_bf16(o[i][2] * i1, o[i][3] * i1, h1, l1);
        if (r0 < T) {
            *reinterpret_cast<uint32_t*>(ob + (long long)r0 * ldo + i * 8) = h0;
            if (out_lo_col) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * ldo + out_lo_col + i * 8) = l0;
        }
        if (r1 < T) {
            *reinterpret_cast<uint32_t*>(ob + (long long)r1 * ldo + i * 8) = h1;
            if (out_lo_col) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * ldo + out_lo_col + i * 8) = l1;
        }
    }
}

template <int DH, bool F16 = false>
static int launch_attention(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st,
                            float* lse2 = nullptr, const uint32_t* keep = nullptr, float keep_scale = 1.f, int out_lo_col = 0) {
    constexpr int SMEM = (ATT_BM + 4 * ATT_BN) * (DH + 8) * 2;
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(attention_kernel<DH, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    dim3 grid((T + ATT_BM - 1) / ATT_BM, H, B);
    const float scale_log2 = (1.0f / sqrtf((float)DH)) * 1.4426950408889634f;
    attention_kernel<DH, F16><<<grid, ATT_THREADS, SMEM, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                              reinterpret_cast<__nv_bfloat16*>(out), lengths, T, H, scale_log2,
                                                              lse2, keep, keep_scale, out_lo_col);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

int launch_attention_tc(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st,
                        float* lse2, const uint32_t* keep, float keep_scale, bool train, bool f16);
int launch_attention_tc96(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st, float* lse2,
                          const uint32_t* keep, float keep_scale, bool train, bool f16, int out_lo_col);

int launch_attention_tcp(const void* qkv, void* out, const int32_t* lengths, int B, int T, int H, cudaStream_t st, bool f16);

// DN_ATTN_PERSIST=1 selects the persistent kernel of attention_tcp.cu (measured slower: 244 vs 198 us per layer, its softmax
// role spills inside the key-block loop at the 96-register cap of 2 CTAs / SM; kept as a tested experiment)
static bool use_persistent_attention() {
    const char* e = getenv("DN_ATTN_PERSIST");
    return e && e[0] == '1';
}

// DN_ATTN_IMPL=mma selects the mma.sync kernels (bring-up / A-B reference of the tcgen05 kernels)
static bool use_mma_attention() {
    const char* e = getenv("DN_ATTN_IMPL");
    return e && e[0] == 'm';
}

}  // namespace dn

extern "C" int dn_attention(const void* qkv, void* out, const int32_t* lengths, int32_t B, int32_t T, int32_t H,
                            int32_t dh, int32_t fmt, int32_t out_lo_col, void* stream) {
    if (!qkv || !out || B <= 0 || T <= 0 || H <= 0 || B > 65535 || H > 65535) return DN_EINVAL;
    if ((unsigned)fmt > 1u || (out_lo_col && dh != 96) || (fmt == DN_FMT_F16 && dh != 96 && dh != 64)) return DN_EINVAL;
    if (out_lo_col && (out_lo_col < H * dh || out_lo_col % 2)) return DN_EINVAL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dh == 96 && !dn::use_mma_attention() && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0)
        return dn::launch_attention_tc96(qkv, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, false, fmt == DN_FMT_F16, out_lo_col);
    if (dh == 96 && fmt == DN_FMT_F16)
        return dn::launch_attention<96, true>(qkv, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, out_lo_col);
    if (dh == 64) {
        // tcgen05/TMEM kernel (attention_tc.cu)
        const bool use_mma = dn::use_mma_attention();
        if (!use_mma && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) {
            if (dn::use_persistent_attention()) return dn::launch_attention_tcp(qkv, out, lengths, B, T, H, st, fmt == DN_FMT_F16);
            return dn::launch_attention_tc(qkv, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, false, fmt == DN_FMT_F16);
        }
        if (fmt == DN_FMT_F16) return DN_EINVAL;   // the mma.sync dh-64 kernel is the bf16 A/B reference only
        return dn::launch_attention<64>(qkv, out, lengths, B, T, H, st);
    }
    if (dh == 96) return dn::launch_attention<96>(qkv, out, lengths, B, T, H, st, nullptr, nullptr, 1.f, out_lo_col);
    if (dh == 32) return dn::launch_attention<32>(qkv, out, lengths, B, T, H, st);
    return DN_EINVAL;
}

extern "C" int dn_attention_train(const void* qkv, void* out, float* lse2, const int32_t* lengths, const uint32_t* keep_bits,
                                  float keep_scale, int32_t B, int32_t T, int32_t H, int32_t dh, void* stream) {
    if (!qkv || !out || !lse2 || B <= 0 || T <= 0 || H <= 0 || B > 65535 || H > 65535) return DN_EINVAL;
    if (reinterpret_cast<uintptr_t>(qkv) & 15) return DN_EINVAL;
    if (dh == 96) {   // VAE decoder: frozen / eval inside a diffusion step (no dropout), train mode in VAE training
        if (dn::use_mma_attention())
            return dn::launch_attention<96>(qkv, out, lengths, B, T, H, reinterpret_cast<cudaStream_t>(stream), lse2, keep_bits,
                                            keep_scale);
        return dn::launch_attention_tc96(qkv, out, lengths, B, T, H, reinterpret_cast<cudaStream_t>(stream), lse2, keep_bits,
                                         keep_scale, true, false, 0);
    }
    if (dh != 64) return DN_EINVAL;
    return dn::launch_attention_tc(qkv, out, lengths, B, T, H, reinterpret_cast<cudaStream_t>(stream), lse2, keep_bits,
                                   keep_scale, true, false);
}
