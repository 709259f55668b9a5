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

// split-precision staging: hi = bf16(x) in [0, C), lo = bf16(x - hi) in [lo_col, lo_col + C), zeros elsewhere
__global__ void cast_split_kernel(const float* __restrict__ src, long long rows, int C, int lds,
                                  __nv_bfloat16* __restrict__ dst, int ldo, int lo_col) {
    const int groups = ldo / 8;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        const bool is_lo = c >= lo_col;
        const int cs = is_lo ? c - lo_col : c;      // source column
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = (cs + k < C) ? src[r * lds + cs + k] : 0.f;
            if (is_lo) v[k] -= round_bf16(v[k]);
        }
        *reinterpret_cast<uint4*>(dst + r * ldo + c) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

// ---------------------------------------------------------------------------------------------- stochastic weight rounding
// dst[i] = bf16 stochastic rounding of src[i]: add 16 uniform random bits below the kept mantissa, truncate.  The bits come
// from a counter hash of (element index, seed, step), so a step's weights are a deterministic function of the seed, and the
// rounding error of a weight is independent from step to step (E[dst] = src).
__device__ __forceinline__ uint32_t sr_hash(uint32_t x) {   // lowbias32 (avalanching 32-bit finaliser)
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(EW_THREADS)
sround_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4, uint32_t seed,
                   const int* __restrict__ step) {
    const uint32_t key = sr_hash(seed ^ (0x9E3779B9U * (uint32_t)(step ? step[0] + 1 : 1)));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
        const uint32_t r0 = sr_hash((uint32_t)(2 * i) ^ key), r1 = sr_hash((uint32_t)(2 * i + 1) ^ key ^ (uint32_t)(i >> 31));
        const uint32_t b0 = __float_as_uint(v.x) + (r0 & 0xFFFFu), b1 = __float_as_uint(v.y) + (r0 >> 16);
        const uint32_t b2 = __float_as_uint(v.z) + (r1 & 0xFFFFu), b3 = __float_as_uint(v.w) + (r1 >> 16);
        // (finite weights: the carry may reach the exponent, which is the correct round-up to the next binade)
        *reinterpret_cast<uint2*>(dst + 4 * i) = make_uint2((b0 >> 16) | (b1 & 0xFFFF0000u), (b2 >> 16) | (b3 & 0xFFFF0000u));
    }
}

// ---------------------------------------------------------------------------------------------- bf16 split staging
// fp32 [rows, C] -> bf16 [rows, 3C] = [hi | hi | lo] with hi = bf16(x), lo = bf16(x - hi): against weights packed as
// [hi | lo | hi] one bf16 GEMM over K = 3C computes x.w to ~2^-16 relative (the lo.lo term is dropped) — the fp32-grade
// contraction the k-means assignment needs (near-ties between centroids).
__global__ void split_bf16x3_kernel(const float* __restrict__ src, long long rows, int C, __nv_bfloat16* __restrict__ dst) {
    const int groups = C / 4;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * C + c);
        const float x[4] = {v.x, v.y, v.z, v.w};
        float hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hi[k] = __bfloat162float(__float2bfloat16(x[k]));
            lo[k] = x[k] - hi[k];
        }
        const uint2 h = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]));
        __nv_bfloat16* d = dst + r * 3 * C + c;
        *reinterpret_cast<uint2*>(d) = h;
        *reinterpret_cast<uint2*>(d + C) = h;
        *reinterpret_cast<uint2*>(d + 2 * C) = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]));
    }
}

// ---------------------------------------------------------------------------------------------- VAE reparam
// Generic form (any z, either eps layout): one element per thread.
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

// Channel-first eps [B, z, T] (the reference's draw order, distributions.py:38) against channel-last params / out:
// a block transposes a {z x 32 frames} eps tile through smem so that both the eps reads (along t) and the
// params / out accesses (along c, float4) are coalesced.  z % 4 == 0, z <= 128.
constexpr int RP_TT = 32;
__global__ void __launch_bounds__(EW_THREADS)
vae_reparam_cf_kernel(const float* __restrict__ params, int ldp, const float* __restrict__ eps, int B, int T, int z,
                      float* __restrict__ out) {
    __shared__ float tile[128][RP_TT + 1];
    const int tiles_t = (T + RP_TT - 1) / RP_TT;
    for (int blk = blockIdx.x; blk < B * tiles_t; blk += gridDim.x) {
        const int b = blk / tiles_t, t0 = (blk % tiles_t) * RP_TT;
        __syncthreads();
        for (int i = threadIdx.x; i < z * RP_TT; i += EW_THREADS) {
            const int c = i / RP_TT, tt = i % RP_TT;
            tile[c][tt] = (t0 + tt < T) ? eps[((long long)b * z + c) * T + t0 + tt] : 0.f;
        }
        __syncthreads();
        const int zg = z / 4;
        for (int i = threadIdx.x; i < RP_TT * zg; i += EW_THREADS) {
            const int tt = i / zg, c = (i % zg) * 4;
            if (t0 + tt >= T) continue;
            const long long bt = (long long)b * T + t0 + tt;
            const float4 mean = *reinterpret_cast<const float4*>(params + bt * ldp + c);
            const float4 lv = *reinterpret_cast<const float4*>(params + bt * ldp + z + c);
            float4 o;
            o.x = mean.x + expf(0.5f * fminf(fmaxf(lv.x, -30.f), 20.f)) * tile[c + 0][tt];
            o.y = mean.y + expf(0.5f * fminf(fmaxf(lv.y, -30.f), 20.f)) * tile[c + 1][tt];
            o.z = mean.z + expf(0.5f * fminf(fmaxf(lv.z, -30.f), 20.f)) * tile[c + 2][tt];
            o.w = mean.w + expf(0.5f * fminf(fmaxf(lv.w, -30.f), 20.f)) * tile[c + 3][tt];
            *reinterpret_cast<float4*>(out + bt * z + c) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------- diffusion updates
// All three updates process 4 latent channels per thread (16-byte loads / stores of the fp32 state, 8-byte stores of
// the bf16 staging copy, whose pad columns [z, ldx) are rewritten with zeros).  z % 4 == 0, lde % 4 == 0, ldx % 4 == 0.
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
}

struct QSampleOp {
    const float* __restrict__ z_lat;
    const float* __restrict__ eps;
    float ca, cb;
    __device__ __forceinline__ float4 operator()(long long r, int c, int z, const float4& /*xv*/) const {
        const float4 a = *reinterpret_cast<const float4*>(z_lat + r * z + c);
        const float4 e = *reinterpret_cast<const float4*>(eps + r * z + c);
        return make_float4(ca * a.x + cb * e.x, ca * a.y + cb * e.y, ca * a.z + cb * e.z, ca * a.w + cb * e.w);
    }
    static constexpr bool reads_x = false;
};

struct DdimOp {
    const float* __restrict__ eh;
    int lde, mode;
    float c0, c1, c2, c3, c4, c5;
    __device__ __forceinline__ float one(float xv, float e) const {
        float x0, pn;
        if (mode == 0) {  // LM:1422-1424 incl. safe_div clamps
            x0 = (xv - c1 * e) / fmaxf(c0, 1e-10f);
            pn = (xv - c0 * x0) / fmaxf(c1, 1e-10f);
        } else {          // gaussian_diffusion.py:535-541
            x0 = c4 * xv - c5 * e;
            pn = (c4 * xv - x0) / c5;
        }
        return x0 * c2 + c3 * pn;
    }
    __device__ __forceinline__ float4 operator()(long long r, int c, int /*z*/, const float4& xv) const {
        const float4 e = *reinterpret_cast<const float4*>(eh + r * lde + c);
        return make_float4(one(xv.x, e.x), one(xv.y, e.y), one(xv.z, e.z), one(xv.w, e.w));
    }
    static constexpr bool reads_x = true;
};

struct DdpmOp {
    const float* __restrict__ eh;
    const float* __restrict__ noise;
    int lde;
    float c0, c1, c2, c3, c4;
    __device__ __forceinline__ float one(float xv, float e, float n) const {
        const float x0 = c0 * xv - c1 * e;
        return c2 * x0 + c3 * xv + c4 * n;
    }
    __device__ __forceinline__ float4 operator()(long long r, int c, int z, const float4& xv) const {
        const float4 e = *reinterpret_cast<const float4*>(eh + r * lde + c);
        const float4 n = *reinterpret_cast<const float4*>(noise + r * z + c);
        return make_float4(one(xv.x, e.x, n.x), one(xv.y, e.y, n.y), one(xv.z, e.z, n.z), one(xv.w, e.w, n.w));
    }
    static constexpr bool reads_x = true;
};

// xlo != 0: the staging row is a split-precision pair: hi in [0, z), lo = bf16(v - hi) in [xlo, xlo + z); the columns in
// between and after are pad (zero).  The thread that owns latent columns [c, c+4) writes both halves.
template <class Op>
__device__ __forceinline__ void latent_update_loop(const Op& op, float* __restrict__ x, long long rows, int z,
                                                   __nv_bfloat16* __restrict__ xb, int ldx, int xlo) {
    const int zc = xb ? ldx : z;  // iterate over the padded width so the pad columns get zeroed
    const int groups = zc / 4;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        if (c < z) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (Op::reads_x) xv = *reinterpret_cast<const float4*>(x + r * z + c);
            const float4 v = op(r, c, z, xv);
            *reinterpret_cast<float4*>(x + r * z + c) = v;
            if (xb) {
                store_bf16x4(xb + r * ldx + c, v.x, v.y, v.z, v.w);
                if (xlo)
                    store_bf16x4(xb + r * ldx + xlo + c, v.x - round_bf16(v.x), v.y - round_bf16(v.y), v.z - round_bf16(v.z),
                                 v.w - round_bf16(v.w));
            }
        } else if (!xlo || c < xlo || c >= xlo + z) {
            store_bf16x4(xb + r * ldx + c, 0.f, 0.f, 0.f, 0.f);
        }
    }
}

__global__ void __launch_bounds__(EW_THREADS)
q_sample_kernel(const float* __restrict__ z_lat, const float* __restrict__ eps, float ca, float cb, long long rows, int z,
                float* __restrict__ x, __nv_bfloat16* __restrict__ xb, int ldx, int xlo) {
    latent_update_loop(QSampleOp{z_lat, eps, ca, cb}, x, rows, z, xb, ldx, xlo);
}

__global__ void __launch_bounds__(EW_THREADS)
ddim_step_kernel(float* __restrict__ x, const float* __restrict__ eh, int lde, const float* __restrict__ table,
                 const int* __restrict__ t_idx, long long rows, int z, int mode, __nv_bfloat16* __restrict__ xb, int ldx,
                 int xlo) {
    const float* cf = table + (long long)t_idx[0] * 8;
    latent_update_loop(DdimOp{eh, lde, mode, cf[0], cf[1], cf[2], cf[3], cf[4], cf[5]}, x, rows, z, xb, ldx, xlo);
}

__global__ void __launch_bounds__(EW_THREADS)
ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eh, int lde, const float* __restrict__ noise,
                 const float* __restrict__ table, const int* __restrict__ t_idx, long long rows, int z,
                 __nv_bfloat16* __restrict__ xb, int ldx, int xlo) {
    const float* cf = table + (long long)t_idx[0] * 8;
    latent_update_loop(DdpmOp{eh, noise, lde, cf[0], cf[1], cf[2], cf[3], cf[4]}, x, rows, z, xb, ldx, xlo);
}

__global__ void advance_step_kernel(int* t_idx, int delta) { t_idx[0] += delta; }

// ---------------------------------------------------------------------------------------------- adaptive RMSNorm
// one warp per frame; C in {512, 768} (any multiple of 128 up to 1024): each lane owns C/32 values as float4s
// OUT 0: bf16 [rows, C]; OUT 1: split pair, row = [hi (C) | ... | lo at lo_col (C)] with row stride 2 * lo_col; OUT 2: fp16
template <int C, int OUT>
__global__ void __launch_bounds__(EW_THREADS)
adarmsnorm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int T,
                  const float* __restrict__ gamma_p, const float* __restrict__ gb, long long gb_t_stride,
                  const int* __restrict__ t_idx, int t_idx_stride, int lo_col) {
    constexpr int V = C / 128;  // float4 per lane
    pdl_trigger();
    pdl_wait();
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
            if constexpr (OUT == 1) {
                uint32_t h0, l0, h1, l1;
                split_bf16(o[0], o[1], h0, l0);
                split_bf16(o[2], o[3], h1, l1);
                *reinterpret_cast<uint2*>(out + r * 2 * lo_col + c) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(out + r * 2 * lo_col + lo_col + c) = make_uint2(l0, l1);
            } else {
                *reinterpret_cast<uint2*>(out + r * C + c) = make_uint2(pack16<OUT == 2>(o[0], o[1]), pack16<OUT == 2>(o[2], o[3]));
            }
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
        float ga[8], be[8];
        if (g) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c)), g1 = __ldg(reinterpret_cast<const float4*>(g + c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(g + C + c)), b1 = __ldg(reinterpret_cast<const float4*>(g + C + c + 4));
            ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
            be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) { ga[k] = 1.f; be[k] = 0.f; }
        }
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = wn_gate(fmaf(__bfloat162float(up[k]), ga[k], be[k])) + __bfloat162float(rp[k]);
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
        if ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(in)) & 15) == 0) {
            // 16-byte loads: the weight row streams from HBM once per 8 outputs, the 8 input rows come from L1
            for (int k = lane * 4; k < K; k += 128) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
                for (int i = 0; i < LIN_MB; ++i)
                    if (m0 + i < M) {
                        const float4 xv = *reinterpret_cast<const float4*>(in + (long long)(m0 + i) * K + k);
                        acc[i] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
                    }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                const float wv = __ldg(wr + k);
#pragma unroll
                for (int i = 0; i < LIN_MB; ++i)
                    if (m0 + i < M) acc[i] += wv * in[(long long)(m0 + i) * K + k];
            }
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

extern "C" int dn_cast_split(const float* src, int64_t rows, int32_t C, int32_t lds, void* dst, int32_t ldo, int32_t lo_col,
                             void* stream) {
    if (!src || !dst || rows <= 0 || ldo % 8 || lo_col % 8 || lo_col < 0) return DN_EINVAL;
    if (lo_col ? (C > lo_col || lo_col + C > ldo) : C > ldo) return DN_EINVAL;
    cast_split_kernel<<<ew_grid(rows * (ldo / 8), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        src, rows, C, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldo, lo_col ? lo_col : ldo);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_sround_bf16(const float* src, void* dst, int64_t n, uint32_t seed, const int32_t* step, void* stream) {
    if (!src || !dst || n <= 0 || n % 4 || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15)) return DN_EINVAL;
    sround_bf16_kernel<<<ew_grid(n / 4, EW_THREADS), EW_THREADS, 0, ST(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n / 4,
                                                                                 seed, step);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_split_bf16x3(const float* src, int64_t rows, int32_t C, void* dst, void* stream) {
    if (!src || !dst || rows <= 0 || C <= 0 || C % 4 || (reinterpret_cast<uintptr_t>(src) & 15)) return DN_EINVAL;
    split_bf16x3_kernel<<<ew_grid(rows * (C / 4), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(src, rows, C,
                                                                                          reinterpret_cast<__nv_bfloat16*>(dst));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_vae_reparam(const float* params, int32_t ldp, const float* eps, int32_t eps_channel_first, int32_t B,
                              int32_t T, int32_t z, float* z_out, void* stream) {
    if (!params || !eps || !z_out || B <= 0 || T <= 0 || z <= 0 || ldp < 2 * z) return DN_EINVAL;
    if (eps_channel_first && z % 4 == 0 && z <= 128 && ldp % 4 == 0 &&
        !((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(z_out)) & 15))
        vae_reparam_cf_kernel<<<ew_grid((long long)B * ((T + RP_TT - 1) / RP_TT), 1), EW_THREADS, 0, ST(stream)>>>(
            params, ldp, eps, B, T, z, z_out);
    else
        vae_reparam_kernel<<<ew_grid((long long)B * T * z, EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
            params, ldp, eps, eps_channel_first, B, T, z, z_out);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_q_sample(const float* z_lat, const float* eps, float sqrt_ab, float sqrt_1m_ab, int64_t rows,
                           int32_t z, float* x, void* x_bf16, int32_t ldx, int32_t x_lo_col, void* stream) {
    if (!z_lat || !eps || !x || rows <= 0 || z <= 0 || z % 4 || (x_bf16 && (ldx < z || ldx % 4))) return DN_EINVAL;
    if (x_lo_col && (!x_bf16 || x_lo_col % 4 || x_lo_col < z || x_lo_col + z > ldx)) return DN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(z_lat) | reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(x)) & 15)
        return DN_EINVAL;
    q_sample_kernel<<<ew_grid(rows * ((x_bf16 ? ldx : z) / 4), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        z_lat, eps, sqrt_ab, sqrt_1m_ab, rows, z, x, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx, x_lo_col);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_ddim_step(float* x, const float* eps_hat, int32_t lde, const float* coef_table, const int32_t* t_idx,
                            int64_t rows, int32_t z, int32_t mode, void* x_bf16, int32_t ldx, int32_t x_lo_col, void* stream) {
    if (!x || !eps_hat || !coef_table || !t_idx || rows <= 0 || z <= 0 || lde < z) return DN_EINVAL;
    if (x_lo_col && (!x_bf16 || x_lo_col % 4 || x_lo_col < z || x_lo_col + z > ldx)) return DN_EINVAL;
    if (z % 4 || lde % 4 || (x_bf16 && (ldx < z || ldx % 4)) ||
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps_hat)) & 15))
        return DN_EINVAL;
    ddim_step_kernel<<<ew_grid(rows * ((x_bf16 ? ldx : z) / 4), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        x, eps_hat, lde, coef_table, t_idx, rows, z, mode, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx, x_lo_col);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_ddpm_step(float* x, const float* eps_hat, int32_t lde, const float* noise, const float* coef_table,
                            const int32_t* t_idx, int64_t rows, int32_t z, void* x_bf16, int32_t ldx, int32_t x_lo_col,
                            void* stream) {
    if (!x || !eps_hat || !noise || !coef_table || !t_idx || rows <= 0 || z <= 0 || lde < z) return DN_EINVAL;
    if (x_lo_col && (!x_bf16 || x_lo_col % 4 || x_lo_col < z || x_lo_col + z > ldx)) return DN_EINVAL;
    if (z % 4 || lde % 4 || (x_bf16 && (ldx < z || ldx % 4)) ||
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps_hat) | reinterpret_cast<uintptr_t>(noise)) & 15))
        return DN_EINVAL;
    ddpm_step_kernel<<<ew_grid(rows * ((x_bf16 ? ldx : z) / 4), EW_THREADS), EW_THREADS, 0, ST(stream)>>>(
        x, eps_hat, lde, noise, coef_table, t_idx, rows, z, reinterpret_cast<__nv_bfloat16*>(x_bf16), ldx, x_lo_col);
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
                             int32_t out_lo_col, int32_t out_fmt, void* stream) {
    if (!x || !out || B <= 0 || T <= 0 || (gb && !t_idx)) return DN_EINVAL;
    if (out_lo_col && (out_lo_col < C || out_lo_col % 4 || out_fmt != DN_FMT_BF16)) return DN_EINVAL;
    if ((unsigned)out_fmt > 1u) return DN_EINVAL;
    const long long rows = (long long)B * T;
    const int grid = ew_grid(rows, EW_THREADS / 32);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#define DN_NORM_CASE(CC)                                                                                                        \
    case CC:                                                                                                                    \
        if (out_lo_col)                                                                                                         \
            DN_CUDA_OK(launch_ex(adarmsnorm_kernel<CC, 1>, grid, EW_THREADS, 0, ST(stream), 1, x, o, B, T, gamma_p, gb,         \
                                 (long long)gb_t_stride, t_idx, t_idx_stride, out_lo_col));                                     \
        else if (out_fmt == DN_FMT_F16)                                                                                         \
            DN_CUDA_OK(launch_ex(adarmsnorm_kernel<CC, 2>, grid, EW_THREADS, 0, ST(stream), 1, x, o, B, T, gamma_p, gb,         \
                                 (long long)gb_t_stride, t_idx, t_idx_stride, 0));                                              \
        else                                                                                                                    \
            DN_CUDA_OK(launch_ex(adarmsnorm_kernel<CC, 0>, grid, EW_THREADS, 0, ST(stream), 1, x, o, B, T, gamma_p, gb,         \
                                 (long long)gb_t_stride, t_idx, t_idx_stride, 0));                                              \
        break;
    switch (C) {
        DN_NORM_CASE(128)
        DN_NORM_CASE(256)
        DN_NORM_CASE(512)
        DN_NORM_CASE(768)
        DN_NORM_CASE(1024)
        default: return DN_EINVAL;
    }
#undef DN_NORM_CASE
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
