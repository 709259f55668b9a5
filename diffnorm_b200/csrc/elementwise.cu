// HBM-bound fused elementwise kernels of the normalization pass:
// adaptive RMSNorm, WaveNet gate, q_sample / DDIM / DDPM updates, VAE reparameterisation, gather+pad,
// fp32 -> bf16 operand staging, small-M fp32 linear for the time-conditioning tables.
// All are one-pass, 16-byte vectorised, coalesced along the channel dimension.
#include "common.cuh"

namespace dn {

constexpr int EW_THREADS = 256;

static inline int ew_grid(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---------------------------------------------------------------------------------------------- cast + pad
__global__ void cast_pad_kernel(const float* __restrict__ src, long long rows, int C, int lds,
                                __nv_bfloat16* __restrict__ dst, int ldo) {
    const int groups = ldo / 8;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (c + k < C) ? src[r * lds + c + k] : 0.f;
        *reinterpret_cast<uint4*>(dst + r * ldo + c) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

// ---------------------------------------------------------------------------------------------- VAE reparam
__global__ void vae_reparam_kernel(const float* __restrict__ params, int ldp, const float* __restrict__ eps,
                                   int eps_cf, int B, int T, int z, float* __restrict__ out) {
    const long long total = (long long)B * T * z;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % z);
        const long long bt = i / z;
        const int t = (int)(bt % T);
        const int b = (int)(bt / T);
        const float mean = params[bt * ldp + c];
        float lv = params[bt * ldp + z + c];
        lv = fminf(fmaxf(lv, -30.f), 20.f);
        const float e = eps_cf ? eps[((long long)b * z + c) * T + t] : eps[i];
        out[i] = mean + expf(0.5f * lv) * e;
    }
}

// ---------------------------------------------------------------------------------------------- diffusion updates
__device__ __forceinline__ void store_bf16_pad(__nv_bfloat16* xb, int ldx, long long r, int c, int z, float v) {
    if (xb) xb[r * ldx + c] = __float2bfloat16(v);
}

__global__ void q_sample_kernel(const float* __restrict__ z_lat, const float* __restrict__ eps, float ca, float cb,
                                long long rows, int z, float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int ldx) {
    const int zc = xb ? ldx : z;  // iterate over padded width so the pad columns get zeroed
    const long long total = rows * zc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / zc;
        const int c = (int)(i % zc);
        if (c < z) {
            const float v = ca * z_lat[r * z + c] + cb * eps[r * z + c];
            x[r * z + c] = v;
            store_bf16_pad(xb, ldx, r, c, z, v);
        } else {
            xb[r * ldx + c] = __float2bfloat16(0.f);
        }
    }
}

__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ eh, int lde,
                                 const float* __restrict__ table, const int* __restrict__ t_idx, long long rows, int z,
                                 int mode, __nv_bfloat16* __restrict__ xb, int ldx) {
    const float* cf = table + (long long)t_idx[0] * 8;
    const float c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3], c4 = cf[4], c5 = cf[5];
    const long long total = rows * z;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / z;
        const int c = (int)(i % z);
        const float xv = x[i];
        const float e = eh[r * lde + c];
        float x0, pn;
        if (mode == 0) {  // LM:1422-1424 incl. safe_div clamps
            x0 = (xv - c1 * e) / fmaxf(c0, 1e-10f);
            pn = (xv - c0 * x0) / fmaxf(c1, 1e-10f);
        } else {          // gaussian_diffusion.py:535-541
            x0 = c4 * xv - c5 * e;
            pn = (c4 * xv - x0) / c5;
        }
        const float v = x0 * c2 + c3 * pn;
        x[i] = v;
        store_bf16_pad(xb, ldx, r, c, z, v);
    }
}

__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eh, int lde,
                                 const float* __restrict__ noise, const float* __restrict__ table,
                                 const int* __restrict__ t_idx, long long rows, int z, __nv_bfloat16* __restrict__ xb,
                                 int ldx) {
    const float* cf = table + (long long)t_idx[0] * 8;
    const float c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3], c4 = cf[4];
    const long long total = rows * z;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / z;
        const int c = (int)(i % z);
        const float xv = x[i];
        const float x0 = c0 * xv - c1 * eh[r * lde + c];
        const float v = c2 * x0 + c3 * xv + c4 * noise[i];
        x[i] = v;
        store_bf16_pad(xb, ldx, r, c, z, v);
    }
}

__global__ void advance_step_kernel(int* t_idx, int delta) { t_idx[0] += delta; }

// ---------------------------------------------------------------------------------------------- adaptive RMSNorm
// one warp per frame; C in {512, 768} (any multiple of 128 up to 1024): each lane owns C/32 values as float4s
template <int C>
__global__ void __launch_bounds__(EW_THREADS)
adarmsnorm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int T,
                  const float* __restrict__ gamma_p, const float* __restrict__ gb, long long gb_t_stride,
                  const int* __restrict__ t_idx, int t_idx_stride) {
    constexpr int V = C / 128;  // float4 per lane
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)B * T;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float scale = sqrtf((float)C);
    for (long long r = warp0; r < rows; r += nwarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + r * C);
        float4 v[V];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            v[k] = xr[k * 32 + lane];
            ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
        }
        ss = warp_sum(ss);
        const float inv = scale / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (LM:631)
        const float* g = nullptr;
        if (gb) {
            const int b = (int)(r / T);
            g = gb + (long long)t_idx[(long long)b * t_idx_stride] * gb_t_stride;
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = (k * 32 + lane) * 4;
            float o[4] = {v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv};
            if (gamma_p) {
                const float4 gp = __ldg(reinterpret_cast<const float4*>(gamma_p + c));
                o[0] *= gp.x; o[1] *= gp.y; o[2] *= gp.z; o[3] *= gp.w;
            }
            if (g) {
                const float4 ga = __ldg(reinterpret_cast<const float4*>(g + c));
                const float4 be = __ldg(reinterpret_cast<const float4*>(g + C + c));
                o[0] = o[0] * ga.x + be.x; o[1] = o[1] * ga.y + be.y;
                o[2] = o[2] * ga.z + be.z; o[3] = o[3] * ga.w + be.w;
            }
            *reinterpret_cast<uint2*>(out + r * C + c) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
        }
    }
}

// ---------------------------------------------------------------------------------------------- WaveNet gate
__global__ void wavenet_gate_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ res,
                                    __nv_bfloat16* __restrict__ y, int B, int T, int C, const float* __restrict__ gb,
                                    long long gb_t_stride, const int* __restrict__ t_idx, int t_idx_stride) {
    const int groups = C / 8;
    const long long total = (long long)B * T * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        const uint4 uu = *reinterpret_cast<const uint4*>(u + r * C + c);
        const uint4 rr = *reinterpret_cast<const uint4*>(res + r * C + c);
        const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(&uu);
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(&rr);
        const float* g = nullptr;
        if (gb) g = gb + (long long)t_idx[(r / T) * t_idx_stride] * gb_t_stride;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float uv = __bfloat162float(up[k]);
            if (g) uv = uv * __ldg(g + c + k) + __ldg(g + C + c + k);
            o[k] = wn_gate(uv) + __bfloat162float(rp[k]);
        }
        *reinterpret_cast<uint4*>(y + r * C + c) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
}

// ---------------------------------------------------------------------------------------------- gather + pad
__global__ void gather_pack_kernel(const float* __restrict__ src, const long long* __restrict__ src_row0,
                                   const long long* __restrict__ keep, const int* __restrict__ counts, int B, int T,
                                   int C, void* __restrict__ dst, int ldd, int dst_bf16) {
    const int groups = ldd / 4;
    const long long total = (long long)B * T * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long bt = i / groups;
        const int c = (int)(i % groups) * 4;
        const int j = (int)(bt % T);
        const int b = (int)(bt / T);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < counts[b] && c < C) {
            const long long sr = src_row0[b] + keep[(long long)b * T + j];
            v = *reinterpret_cast<const float4*>(src + sr * C + c);
        }
        if (dst_bf16)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + bt * ldd + c) =
                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        else
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + bt * ldd + c) = v;
    }
}

// ---------------------------------------------------------------------------------------------- small-M fp32 linear
// out[m, n] = act(in[m, :] . W[n, :] + bias[n]).  One warp per (n, 8 rows of m): W row is read once per 8 outputs.
constexpr int LIN_MB = 8;
__global__ void __launch_bounds__(EW_THREADS)
linear_f32_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ out, int M, int N, int K, int act) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int mblocks = (M + LIN_MB - 1) / LIN_MB;
    const long long total = (long long)N * mblocks;
    for (long long w = warp0; w < total; w += nwarps) {
        const int n = (int)(w % N);
        const int m0 = (int)(w / N) * LIN_MB;
        float acc[LIN_MB];
#pragma unroll
        for (int i = 0; i < LIN_MB; ++i) acc[i] = 0.f;
        const float* wr = W + (long long)n * K;
        for (int k = lane; k < K; k += 32) {
            const float wv = __ldg(wr + k);
#pragma unroll
            for (int i = 0; i < LIN_MB; ++i)
                if (m0 + i < M) acc[i] += wv * in[(long long)(m0 + i) * K + k];
        }
#pragma unroll
        for (int i = 0; i < LIN_MB; ++i) {
            const float s = warp_sum(acc[i]);
            if (lane == 0 && m0 + i < M) {
                float v = s + (bias ? bias[n] : 0.f);
                if (act == 1) v = v / (1.f + expf(-v));
                out[(long long)(m0 + i) * N + n] = v;
            }
        }
    }
}

__global__ void time_features_kernel(const int* __restrict__ steps, const float* __restrict__ w, int M, int half,
                                     float* __restrict__ out) {
    const int width = 2 * half + 1;
    const long long total = (long long)M * width;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / width);
        const int c = (int)(i % width);
        const float t = (float)steps[m];
        float v;
        if (c == 0) {
            v = t;
        } else {
            const int j = (c - 1) % half;
            // LM:113: freqs = t * w * 2 * pi, evaluated left to right in fp32
            const float f = t * w[j] * 2.f * 3.14159265358979323846f;
            v = (c - 1 < half) ? sinf(f) : cosf(f);
        }
        out[i] = v;
    }
}

}  // namespace dn

using namespace dn;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dn_cast_pad_bf16(const float* src, int64_t rows, int32_t C, int32_t lds, void* dst, int32_t ldo,
                                void* stream) {
    if (!src || !dst || rows <= 0 || ldo % 8 || C > ldo) return DN_EINVAL;
    cast_pad_kernel<<<ew_grid(rows * (ldo / 8), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        src, rows, C, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldo);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_vae_reparam(const float* params, int32_t ldp, const float* eps, int32_t eps_channel_first, int32_t B,
                              int32_t T, int32_t z, float* z_out, void* stream) {
    if (!params || !eps || !z_out || B <= 0 || T <= 0 || z <= 0 || ldp < 2 * z) return DN_EINVAL;
    vae_reparam_kernel<<<ew_grid((long long)B * T * z, EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        params, ldp, eps, eps_channel_first, B, T, z, z_out);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_q_sample(const float* z_lat, const float* eps, float sqrt_ab, float sqrt_1m_ab, int64_t rows,
                           int32_t z, float* x, void* x_bf16, int32_t ldx, void* stream) {
    if (!z_lat || !eps || !x || rows <= 0 || z <= 0 || (x_bf16 && ldx < z)) return DN_EINVAL;
    q_sample_kernel<<<ew_grid(rows * (x_bf16 ? ldx : z), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        z_lat, eps, sqrt_ab, sqrt_1m_ab, rows, z, x, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_ddim_step(float* x, const float* eps_hat, int32_t lde, const float* coef_table, const int32_t* t_idx,
                            int64_t rows, int32_t z, int32_t mode, void* x_bf16, int32_t ldx, void* stream) {
    if (!x || !eps_hat || !coef_table || !t_idx || rows <= 0 || z <= 0 || lde < z) return DN_EINVAL;
    ddim_step_kernel<<<ew_grid(rows * z, EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        x, eps_hat, lde, coef_table, t_idx, rows, z, mode, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_ddpm_step(float* x, const float* eps_hat, int32_t lde, const float* noise, const float* coef_table,
                            const int32_t* t_idx, int64_t rows, int32_t z, void* x_bf16, int32_t ldx, void* stream) {
    if (!x || !eps_hat || !noise || !coef_table || !t_idx || rows <= 0 || z <= 0 || lde < z) return DN_EINVAL;
    ddpm_step_kernel<<<ew_grid(rows * z, EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        x, eps_hat, lde, noise, coef_table, t_idx, rows, z, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_advance_step(int32_t* t_idx, int32_t delta, void* stream) {
    if (!t_idx) return DN_EINVAL;
    advance_step_kernel<<<1, 1, 0, ST(stream)>>>(t_idx, delta);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_adarmsnorm(const float* x, void* out, int32_t B, int32_t T, int32_t C, const float* gamma_p,
                             const float* gb, int64_t gb_t_stride, const int32_t* t_idx, int32_t t_idx_stride,
                             void* stream) {
    if (!x || !out || B <= 0 || T <= 0 || (gb && !t_idx)) return DN_EINVAL;
    const long long rows = (long long)B * T;
    const int grid = ew_grid(rows, EW_THREADS / 32);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    switch (C) {
        case 128: adarmsnorm_kernel<128><<<grid, EW_THREADS, 0, ST(stream)>>>(x, o, B, T, gamma_p, gb, gb_t_stride, t_idx, t_idx_stride); break;
        case 256: adarmsnorm_kernel<256><<<grid, EW_THREADS, 0, ST(stream)>>>(x, o, B, T, gamma_p, gb, gb_t_stride, t_idx, t_idx_stride); break;
        case 512: adarmsnorm_kernel<512><<<grid, EW_THREADS, 0, ST(stream)>>>(x, o, B, T, gamma_p, gb, gb_t_stride, t_idx, t_idx_stride); break;
        case 768: adarmsnorm_kernel<768><<<grid, EW_THREADS, 0, ST(stream)>>>(x, o, B, T, gamma_p, gb, gb_t_stride, t_idx, t_idx_stride); break;
        case 1024: adarmsnorm_kernel<1024><<<grid, EW_THREADS, 0, ST(stream)>>>(x, o, B, T, gamma_p, gb, gb_t_stride, t_idx, t_idx_stride); break;
        default: return DN_EINVAL;
    }
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_wavenet_gate(const void* u, const void* res, void* y, int32_t B, int32_t T, int32_t C, const float* gb,
                               int64_t gb_t_stride, const int32_t* t_idx, int32_t t_idx_stride, void* stream) {
    if (!u || !res || !y || B <= 0 || T <= 0 || C % 8 || (gb && !t_idx)) return DN_EINVAL;
    wavenet_gate_kernel<<<ew_grid((long long)B * T * (C / 8), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(u), reinterpret_cast<const __nv_bfloat16*>(res),
        reinterpret_cast<__nv_bfloat16*>(y), B, T, C, gb, gb_t_stride, t_idx, t_idx_stride);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_gather_pack(const float* src, const int64_t* src_row0, const int64_t* index_to_keep,
                              const int32_t* counts, int32_t B, int32_t T, int32_t C, void* dst, int32_t ldd,
                              int32_t dst_bf16, void* stream) {
    if (!src || !src_row0 || !index_to_keep || !counts || !dst || B <= 0 || T <= 0 || C % 4 || ldd % 4 || ldd < C)
        return DN_EINVAL;
    gather_pack_kernel<<<ew_grid((long long)B * T * (ldd / 4), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        src, reinterpret_cast<const long long*>(src_row0), reinterpret_cast<const long long*>(index_to_keep), counts, B,
        T, C, dst, ldd, dst_bf16);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_linear_f32(const float* in, const float* W, const float* bias, float* out, int32_t M, int32_t N,
                             int32_t K, int32_t act, void* stream) {
    if (!in || !W || !out || M <= 0 || N <= 0 || K <= 0) return DN_EINVAL;
    const long long warps = (long long)N * ((M + LIN_MB - 1) / LIN_MB);
    linear_f32_kernel<<<ew_grid(warps, EW_THREADS / 32), EW_THREADS, 0, ST(stream)>>>(in, W, bias, out, M, N, K, act);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_time_features(const int32_t* steps, const float* w, int32_t M, int32_t half, float* out,
                                void* stream) {
    if (!steps || !w || !out || M <= 0 || half <= 0) return DN_EINVAL;
    time_features_kernel<<<ew_grid((long long)M * (2 * half + 1), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(steps, w, M,
                                                                                                         half, out);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
