// Training-step kernels of the denoiser (LatentDiscreteModel.forward, LM:1514-1613, and its backward): the
// HBM-bound forward pieces that keep what backward needs, the elementwise / reduction backward kernels, the loss
// kernels, and the small fp32 kernels of the time-conditioning MLP backward.  GEMM-shaped work (dgrad, wgrad,
// attention backward) lives in gemm.cu / wgrad.cu / attention_bwd.cu.
#include <curand_kernel.h>

#include "common.cuh"

namespace dn {

constexpr int TR_THREADS = 256;

static inline int tr_grid(long long items, int per_block) {
    long long g = (items + per_block - 1) / per_block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&o)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        o[2 * i] = f.x;
        o[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
    return make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}
__device__ __forceinline__ float block_sum(float v, float* red) {  // blockDim = TR_THREADS
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < TR_THREADS / 32; ++i) t += red[i];
    return t;
}

// ---------------------------------------------------------------------------------------------- GEGLU (LM:881-885)
// h [rows, 2*ip] packed like the GEGLU weight tiles: tile j = 128 "x" columns then the 128 matching gate columns.
__global__ void __launch_bounds__(TR_THREADS)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long rows, int ip, __nv_bfloat16* __restrict__ m) {
    const int groups = ip / 8;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        const __nv_bfloat16* hp = h + r * 2 * ip + (c >> 7) * 256 + (c & 127);
        float x[8], g[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(hp), x);
        unpack8(*reinterpret_cast<const uint4*>(hp + 128), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = gelu_erf(g[k]) * x[k];
        *reinterpret_cast<uint4*>(m + r * ip + c) = pack8(o);
    }
}

__global__ void __launch_bounds__(TR_THREADS)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dm, long long rows, int ip,
                 __nv_bfloat16* __restrict__ dh) {
    const int groups = ip / 8;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        const long long off = r * 2 * ip + (c >> 7) * 256 + (c & 127);
        float x[8], g[8], d[8], dx[8], dg[8];
        unpack8(*reinterpret_cast<const uint4*>(h + off), x);
        unpack8(*reinterpret_cast<const uint4*>(h + off + 128), g);
        unpack8(*reinterpret_cast<const uint4*>(dm + r * ip + c), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float cdf = 0.5f * (1.f + erff(g[k] * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * __expf(-0.5f * g[k] * g[k]);
            dx[k] = d[k] * g[k] * cdf;                 // d/dx [gelu(g) x] = gelu(g)
            dg[k] = d[k] * x[k] * (cdf + g[k] * pdf);  // gelu'(g) = Phi(g) + g phi(g)
        }
        *reinterpret_cast<uint4*>(dh + off) = pack8(dx);
        *reinterpret_cast<uint4*>(dh + off + 128) = pack8(dg);
    }
}

// ---------------------------------------------------------------------------------------------- WaveNet gate (LM:513-536)
// ur [B*T, G*2*C] as the GEMM wrote it (per group, tile j = 128 conv columns then 128 res columns);
// y [B*T, G*C].  gamma/beta row of utterance b, group g: gb + t_idx[b*stride]*gb_t_stride + g*g_gb (gamma, then beta at +C).
__global__ void __launch_bounds__(TR_THREADS)
wn_gate_fwd_kernel(const __nv_bfloat16* __restrict__ ur, __nv_bfloat16* __restrict__ y, int B, int T, int C, int G,
                   const float* __restrict__ gb, long long gb_t_stride, int g_gb, const int* __restrict__ t_idx,
                   int t_idx_stride) {
    const int cg = C / 8;
    const long long total = (long long)B * T * G * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cg) * 8;
        long long q = i / cg;
        const int g = (int)(q % G);
        const long long r = q / G;
        const __nv_bfloat16* up = ur + r * (long long)G * 2 * C + (long long)g * 2 * C + (c >> 7) * 256 + (c & 127);
        float u[8], rs[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(up), u);
        unpack8(*reinterpret_cast<const uint4*>(up + 128), rs);
        if (gb) {
            const float* gr = gb + (long long)t_idx[(r / T) * t_idx_stride] * gb_t_stride + (long long)g * g_gb;
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = fmaf(u[k], __ldg(gr + c + k), __ldg(gr + C + c + k));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = tanhf(u[k]) / (1.f + __expf(-u[k])) + rs[k];
        *reinterpret_cast<uint4*>(y + r * (long long)G * C + (long long)g * C + c) = pack8(o);
    }
}

// backward: block = (utterance b, group g, 128-channel tile j); 16 channel octets x 16 frame lanes.
// dur [B*T, G*2*C] is written NON-interleaved per group: [du (C) | dres (C)], the K layout the dgrad GEMM wants.
// dgb [B, dgb_b_stride] receives (=, the block owns its entries) dgamma at g*g_dgb + ch and dbeta at + C.
__global__ void __launch_bounds__(TR_THREADS)
wn_gate_bwd_kernel(const __nv_bfloat16* __restrict__ ur, const __nv_bfloat16* __restrict__ dy,
                   __nv_bfloat16* __restrict__ dur, int B, int T, int C, int G, const float* __restrict__ gb,
                   long long gb_t_stride, int g_gb, const int* __restrict__ t_idx, int t_idx_stride,
                   float* __restrict__ dgb, long long dgb_b_stride, int g_dgb) {
    __shared__ float red[2][16][128 + 4];
    const int tiles = C / 128;
    int blk = blockIdx.x;
    const int j = blk % tiles;
    blk /= tiles;
    const int g = blk % G;
    const int b = blk / G;
    const int oct = threadIdx.x & 15, tl = threadIdx.x >> 4;
    const int c = j * 128 + oct * 8;
    float ga[8], be[8], sg[8], sb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ga[k] = 1.f; be[k] = 0.f; sg[k] = 0.f; sb[k] = 0.f; }
    if (gb) {
        const float* gr = gb + (long long)t_idx[b * t_idx_stride] * gb_t_stride + (long long)g * g_gb;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ga[k] = gr[c + k]; be[k] = gr[C + c + k]; }
    }
    for (int t = tl; t < T; t += 16) {
        const long long r = (long long)b * T + t;
        float u[8], d[8], du[8];
        unpack8(*reinterpret_cast<const uint4*>(ur + r * (long long)G * 2 * C + (long long)g * 2 * C + j * 256 + oct * 8), u);
        const uint4 dyv = *reinterpret_cast<const uint4*>(dy + r * (long long)G * C + (long long)g * C + c);
        unpack8(dyv, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float up = fmaf(u[k], ga[k], be[k]);
            const float th = tanhf(up), s = 1.f / (1.f + __expf(-up));
            const float dup = d[k] * ((1.f - th * th) * s + th * s * (1.f - s));
            sg[k] += dup * u[k];
            sb[k] += dup;
            du[k] = dup * ga[k];
        }
        __nv_bfloat16* op = dur + r * (long long)G * 2 * C + (long long)g * 2 * C;
        *reinterpret_cast<uint4*>(op + c) = pack8(du);
        *reinterpret_cast<uint4*>(op + C + c) = dyv;
    }
    if (dgb) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            red[0][tl][oct * 8 + k] = sg[k];
            red[1][tl][oct * 8 + k] = sb[k];
        }
        __syncthreads();
        const int w = threadIdx.x >> 7, ch = threadIdx.x & 127;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[w][i][ch];
        dgb[(long long)b * dgb_b_stride + (long long)g * g_dgb + w * C + j * 128 + ch] = s;
    }
}

// ---------------------------------------------------------------------------------------------- RMSNorm backward (LM:629-639)
// out = n * gamma_eff + beta, n = x * sqrt(C) / max(||x||, 1e-12).  One warp per frame (same mapping as the forward);
// a block covers frames of ONE utterance so the per-utterance gamma/beta gradients reduce in registers, then smem,
// then one atomicAdd per channel per block.
template <int C>
__global__ void __launch_bounds__(TR_THREADS)
adarmsnorm_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dx,
                      __nv_bfloat16* __restrict__ dx_bf16, int B, int T, int rows_per_block,
                      const float* __restrict__ gamma_p, float* __restrict__ dgamma_p, const float* __restrict__ gb,
                      long long gb_t_stride, const int* __restrict__ t_idx, int t_idx_stride, float* __restrict__ dgb,
                      long long dgb_b_stride) {
    constexpr int V = C / 128;
    __shared__ float red[TR_THREADS / 32][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = (T + rows_per_block - 1) / rows_per_block;
    const int b = blockIdx.x / chunks;
    const int t_begin = (blockIdx.x % chunks) * rows_per_block;
    const int t_end = min(T, t_begin + rows_per_block);
    const float scale = sqrtf((float)C);
    const float* g = gb ? gb + (long long)t_idx[(long long)b * t_idx_stride] * gb_t_stride : nullptr;
    float4 ge[V], ag[V], ab[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = (k * 32 + lane) * 4;
        ge[k] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (gamma_p) ge[k] = *reinterpret_cast<const float4*>(gamma_p + c);
        if (g) {
            const float4 t4 = *reinterpret_cast<const float4*>(g + c);
            ge[k].x *= t4.x; ge[k].y *= t4.y; ge[k].z *= t4.z; ge[k].w *= t4.w;
        }
        ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[k] = ag[k];
    }
    for (int t = t_begin + warp; t < t_end; t += TR_THREADS / 32) {
        const long long r = (long long)b * T + t;
        const float4* xr = reinterpret_cast<const float4*>(x + r * C);
        float4 xv[V], dn[V];
        float ss = 0.f, dot = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            xv[k] = xr[k * 32 + lane];
            const uint2 dv = *reinterpret_cast<const uint2*>(dy + r * C + (k * 32 + lane) * 4);
            const float2 d0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.x));
            const float2 d1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dv.y));
            ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
            ab[k].x += d0.x; ab[k].y += d0.y; ab[k].z += d1.x; ab[k].w += d1.y;
            dn[k] = make_float4(d0.x, d0.y, d1.x, d1.y);   // holds dy for now
        }
        ss = warp_sum(ss);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        const float inv = scale / nrm;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            // d gamma_eff += dy * n ; dn = dy * gamma_eff
            ag[k].x += dn[k].x * xv[k].x * inv; ag[k].y += dn[k].y * xv[k].y * inv;
            ag[k].z += dn[k].z * xv[k].z * inv; ag[k].w += dn[k].w * xv[k].w * inv;
            dn[k].x *= ge[k].x; dn[k].y *= ge[k].y; dn[k].z *= ge[k].z; dn[k].w *= ge[k].w;
            dot += dn[k].x * xv[k].x + dn[k].y * xv[k].y + dn[k].z * xv[k].z + dn[k].w * xv[k].w;
        }
        dot = warp_sum(dot) / (nrm * nrm);   // (x_hat . dn) / ||x||
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = (k * 32 + lane) * 4;
            float4 o = *reinterpret_cast<const float4*>(dx + r * C + c);
            o.x += inv * (dn[k].x - xv[k].x * dot); o.y += inv * (dn[k].y - xv[k].y * dot);
            o.z += inv * (dn[k].z - xv[k].z * dot); o.w += inv * (dn[k].w - xv[k].w * dot);
            *reinterpret_cast<float4*>(dx + r * C + c) = o;
            if (dx_bf16) *reinterpret_cast<uint2*>(dx_bf16 + r * C + c) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        }
    }
    // block reduction of the per-channel sums (gamma sums, then beta sums, through one smem buffer)
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = (k * 32 + lane) * 4;
            *reinterpret_cast<float4*>(&red[warp][c]) = pass == 0 ? ag[k] : ab[k];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += TR_THREADS) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < TR_THREADS / 32; ++w) sum += red[w][c];
            if (g) {
                // gamma_eff = gamma_t (conditioned norms have no gamma parameter, LM:662)
                atomicAdd(dgb + (long long)b * dgb_b_stride + pass * C + c, sum);
            } else if (dgamma_p && pass == 0) {
                atomicAdd(dgamma_p + c, sum);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- column sums (bias grads)
__global__ void __launch_bounds__(TR_THREADS)
colsum_kernel(const __nv_bfloat16* __restrict__ src, long long rows, int ld, int col0, int cols, int rows_per_block,
              float* __restrict__ out) {
    __shared__ float red[8][32 * 8 + 8];
    const int cgs = (cols + 7) / 8;                // column octets
    const int cg_chunks = (cgs + 31) / 32;
    const int cg = (blockIdx.x % cg_chunks) * 32 + (threadIdx.x & 31);
    const long long r0 = (long long)(blockIdx.x / cg_chunks) * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    const int rl = threadIdx.x >> 5;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (cg < cgs) {
        for (long long r = r0 + rl; r < r1; r += 8) {
            float v[8];
            unpack8(*reinterpret_cast<const uint4*>(src + r * ld + col0 + cg * 8), v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[rl][(threadIdx.x & 31) * 8 + k] = acc[k];
    __syncthreads();
    const int c = threadIdx.x;  // 256 columns of this chunk
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][c];
    const int col = (blockIdx.x % cg_chunks) * 256 + c;
    if (col < cols) atomicAdd(out + col, s);
}

// ---------------------------------------------------------------------------------------------- noising / losses
// coef table row (float32 x 4) per timestep: [sqrt_ab, sqrt(1-ab), w = min(snr,5)/snr, unused]  (LM:1530-1534,1563-1566)
__global__ void __launch_bounds__(TR_THREADS)
train_noise_kernel(const float* __restrict__ zl, const float* __restrict__ eps0, const float* __restrict__ eps,
                   float beta0, const float* __restrict__ coef, const int* __restrict__ t_idx, int B, int T, int z,
                   float* __restrict__ x_t, __nv_bfloat16* __restrict__ xb, int ldx) {
    const int groups = ldx / 4;
    const long long total = (long long)B * T * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        if (c < z) {
            const float* cf = coef + (long long)t_idx[r / T] * 4;
            const float sa = cf[0], s1 = cf[1];
            const float4 a = *reinterpret_cast<const float4*>(zl + r * z + c);
            const float4 e0 = *reinterpret_cast<const float4*>(eps0 + r * z + c);
            const float4 e = *reinterpret_cast<const float4*>(eps + r * z + c);
            float4 v;
            v.x = sa * (a.x + e0.x * beta0) + s1 * e.x; v.y = sa * (a.y + e0.y * beta0) + s1 * e.y;
            v.z = sa * (a.z + e0.z * beta0) + s1 * e.z; v.w = sa * (a.w + e0.w * beta0) + s1 * e.w;
            *reinterpret_cast<float4*>(x_t + r * z + c) = v;
            *reinterpret_cast<uint2*>(xb + r * ldx + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        } else {
            *reinterpret_cast<uint2*>(xb + r * ldx + c) = make_uint2(0u, 0u);
        }
    }
}

// loss += sum_b w_b / (B T z) sum_{t < len_b, c} (pred - eps)^2 ; dpred = 2 w_b / (B T z) (pred - eps) [valid frames]
__global__ void __launch_bounds__(TR_THREADS)
noise_loss_kernel(const float* __restrict__ pred, int lde, const float* __restrict__ eps, const int* __restrict__ lengths,
                  const float* __restrict__ coef, const int* __restrict__ t_idx, int B, int T, int z,
                  float* __restrict__ loss, __nv_bfloat16* __restrict__ dpred, int ldd, float grad_scale,
                  const __nv_bfloat16* __restrict__ dx1, int ld1) {
    __shared__ float red[TR_THREADS / 32];
    const int groups = ldd / 4;
    const long long total = (long long)B * T * groups;
    const float norm = 1.f / ((float)B * (float)T * (float)z);
    float acc = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        const int b = (int)(r / T), t = (int)(r % T);
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < z && t < lengths[b]) {
            const float w = coef[(long long)t_idx[b] * 4 + 2] * norm;
            const float4 p = *reinterpret_cast<const float4*>(pred + r * lde + c);
            const float4 e = *reinterpret_cast<const float4*>(eps + r * z + c);
            d = make_float4(p.x - e.x, p.y - e.y, p.z - e.z, p.w - e.w);
            acc += w * (d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w);
            const float gsc = 2.f * w * grad_scale;
            d.x *= gsc; d.y *= gsc; d.z *= gsc; d.w *= gsc;
            if (dx1) {   // multitask: x1_hat = (x_t - s1 pred) / max(sa, 1e-10)  =>  d pred += -s1 / max(sa, 1e-10) * d x1_hat
                const float* cf = coef + (long long)t_idx[b] * 4;
                const float k = -cf[1] / fmaxf(cf[0], 1e-10f);
                const uint2 v = *reinterpret_cast<const uint2*>(dx1 + r * ld1 + c);
                const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
                const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
                d.x += k * a0.x; d.y += k * a0.y; d.z += k * a1.x; d.w += k * a1.y;
            }
        }
        if (dpred) *reinterpret_cast<uint2*>(dpred + r * ldd + c) = make_uint2(pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, acc);
}

// x1_hat = (x_t - s1 pred) / max(sa, 1e-10) -> bf16 staging for decode_feature (LM:1572)
__global__ void __launch_bounds__(TR_THREADS)
pred_x1_kernel(const float* __restrict__ x_t, const float* __restrict__ pred, int lde, const float* __restrict__ coef,
               const int* __restrict__ t_idx, int B, int T, int z, __nv_bfloat16* __restrict__ xb, int ldx) {
    const int groups = ldx / 4;
    const long long total = (long long)B * T * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < z) {
            const float* cf = coef + (long long)t_idx[r / T] * 4;
            const float inv = 1.f / fmaxf(cf[0], 1e-10f), s1 = cf[1];
            const float4 x = *reinterpret_cast<const float4*>(x_t + r * z + c);
            const float4 p = *reinterpret_cast<const float4*>(pred + r * lde + c);
            v = make_float4((x.x - s1 * p.x) * inv, (x.y - s1 * p.y) * inv, (x.z - s1 * p.z) * inv, (x.w - s1 * p.w) * inv);
        }
        *reinterpret_cast<uint2*>(xb + r * ldx + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

// decode-side logging losses (forward only; LM:1573-1597): one warp per valid frame.
// out[0] += sum (recon - audio)^2 ; out[1] += nll ; out[2] += smooth ; out[3] += correct ; out[4] += tokens ; out[5] += frames
__global__ void __launch_bounds__(TR_THREADS)
decode_losses_kernel(const float* __restrict__ recon, const float* __restrict__ audio, int C,
                     const float* __restrict__ logits, int ld, int V, const long long* __restrict__ units,
                     const int* __restrict__ lengths, int B, int T, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double a_se = 0, a_nll = 0, a_sm = 0, a_ok = 0, a_tok = 0, a_fr = 0;
    for (long long r = warp0; r < (long long)B * T; r += nwarps) {
        const int b = (int)(r / T), t = (int)(r % T);
        if (t < lengths[b]) {
            float se = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float d = recon[r * C + c] - audio[r * C + c];
                se += d * d;
            }
            a_se += (double)warp_sum(se);
            a_fr += 1;
        }
        const long long u = units[r];
        if (u != 0) {  // ignore_index = padding = 0
            const float* lg = logits + r * ld;
            float mx = -INFINITY;
            int bi = 0x7fffffff;
            for (int c = lane; c < V; c += 32) {
                const float v = lg[c];
                if (v > mx) { mx = v; bi = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > mx || (ov == mx && oi < bi)) { mx = ov; bi = oi; }
            }
            float se = 0.f, sl = 0.f;
            for (int c = lane; c < V; c += 32) {
                const float v = lg[c];
                se += __expf(v - mx);
                sl += v;
            }
            se = warp_sum(se);
            sl = warp_sum(sl);
            const float lse = mx + logf(se);
            a_nll += (double)(lse - lg[u]);
            a_sm += (double)((float)V * lse - sl);
            a_ok += (bi == (int)u) ? 1.0 : 0.0;
            a_tok += 1;
        }
    }
    if (lane == 0) {
        atomicAdd(out + 0, a_se);
        atomicAdd(out + 1, a_nll);
        atomicAdd(out + 2, a_sm);
        atomicAdd(out + 3, a_ok);
        atomicAdd(out + 4, a_tok);
        atomicAdd(out + 5, a_fr);
    }
}

// Backward of the decode branch's losses (multitask, LM:1576-1604): stats = dn_decode_losses' out6 (device doubles).
// d logits = nll_scale / n_tokens * [ (1 - eps - e_i) (p - onehot(u)) + e_i (V p - 1) ],  e_i = eps / (V - 1), rows with u != 0
__global__ void __launch_bounds__(TR_THREADS)
lsnll_bwd_kernel(const float* __restrict__ logits, int ld, int V, const long long* __restrict__ units, long long rows,
                 const double* __restrict__ stats, float eps_ls, float nll_scale, __nv_bfloat16* __restrict__ dlogits, int ldd) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float ntok = fmaxf((float)stats[4], 1.f);
    const float e_i = eps_ls / (float)(V - 1);
    const float a = (1.f - eps_ls - e_i), sc = nll_scale / ntok;
    for (long long r = warp0; r < rows; r += nwarps) {
        const long long u = units[r];
        const float* lg = logits + r * ld;
        __nv_bfloat16* dl = dlogits + r * ldd;
        if (u == 0) {
            for (int c = lane * 2; c < ldd; c += 64) *reinterpret_cast<uint32_t*>(dl + c) = 0u;
            continue;
        }
        float mx = -INFINITY;
        for (int c = lane; c < V; c += 32) mx = fmaxf(mx, lg[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < V; c += 32) se += __expf(lg[c] - mx);
        se = warp_sum(se);
        const float inv = 1.f / se;
        for (int c = lane * 2; c < ldd; c += 64) {
            float g[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = c + e;
                if (cc < V) {
                    const float p = __expf(lg[cc] - mx) * inv;
                    g[e] = sc * (a * (p - (cc == (int)u ? 1.f : 0.f)) + e_i * ((float)V * p - 1.f));
                } else {
                    g[e] = 0.f;
                }
            }
            *reinterpret_cast<uint32_t*>(dl + c) = pack_bf16(g[0], g[1]);
        }
    }
}

// d recon = d_lm (lm-head data gradient, fp32) + mse_scale * 2 (recon - audio) / (n_frames C) on valid frames -> bf16
__global__ void __launch_bounds__(TR_THREADS)
recon_grad_kernel(const float* __restrict__ recon, const float* __restrict__ audio, const float* __restrict__ d_lm,
                  const int* __restrict__ lengths, int B, int T, int C, const double* __restrict__ stats, float mse_scale,
                  __nv_bfloat16* __restrict__ out) {
    const int groups = C / 4;
    const long long total = (long long)B * T * groups;
    const float k = 2.f * mse_scale / (fmaxf((float)stats[5], 1.f) * (float)C);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 4;
        float4 g = *reinterpret_cast<const float4*>(d_lm + r * C + c);
        if ((int)(r % T) < lengths[r / T]) {
            const float4 a = *reinterpret_cast<const float4*>(recon + r * C + c);
            const float4 t = *reinterpret_cast<const float4*>(audio + r * C + c);
            g.x += k * (a.x - t.x); g.y += k * (a.y - t.y); g.z += k * (a.z - t.z); g.w += k * (a.w - t.w);
        } else {
            g = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<uint2*>(out + r * C + c) = make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
    }
}

// ---------------------------------------------------------------------------------------------- VAE posterior (training)
// kl[0] += mean_b( 0.5 * mean_{z,T}( mask * (mu^2 + exp(lv) - 1 - lv) ) ),  lv = clamp(logvar, -30, 20)
// (distributions.py:62-74 kl_3d + LM:1127-1128).  params fp32 [B*T, ldp]: mu in [0,z), logvar in [z,2z).
__global__ void __launch_bounds__(TR_THREADS)
vae_kl_kernel(const float* __restrict__ params, int ldp, const int* __restrict__ lengths, int B, int T, int z,
              float* __restrict__ kl) {
    __shared__ float red[TR_THREADS / 32];
    const long long total = (long long)B * T * z;
    float acc = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % z);
        const long long r = i / z;
        if ((int)(r % T) < lengths[r / T]) {
            const float mu = params[r * ldp + c];
            const float lv = fminf(fmaxf(params[r * ldp + z + c], -30.f), 20.f);
            acc += mu * mu + expf(lv) - 1.f - lv;
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(kl, acc * 0.5f / ((float)B * (float)T * (float)z));
}

// backward of z = mu + exp(0.5 lv) eps and of the KL term: dparams bf16 [B*T, ldd] (pad columns zero).
// dz bf16 [B*T, ldz] (gradient w.r.t. the latent, channel-last); eps channel-first [B, z, T] or channel-last.
__global__ void __launch_bounds__(TR_THREADS)
vae_reparam_bwd_kernel(const float* __restrict__ params, int ldp, const float* __restrict__ eps, int eps_cf,
                       const __nv_bfloat16* __restrict__ dz, int ldz, const int* __restrict__ lengths, int B, int T, int z,
                       float kl_scale, __nv_bfloat16* __restrict__ dparams, int ldd) {
    const long long total = (long long)B * T * ldd;
    const float ks = kl_scale * 0.5f / ((float)B * (float)T * (float)z);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % ldd);
        const long long r = i / ldd;
        float g = 0.f;
        if (col < 2 * z) {
            const int c = col < z ? col : col - z;
            const int b = (int)(r / T), t = (int)(r % T);
            const float d = __bfloat162float(dz[r * ldz + c]);
            const float valid = t < lengths[b] ? 1.f : 0.f;
            const float mu = params[r * ldp + c];
            const float lvr = params[r * ldp + z + c];
            if (col < z) {
                g = d + ks * valid * 2.f * mu;
            } else if (lvr >= -30.f && lvr <= 20.f) {   // clamp passes the gradient only inside its range
                const float e = eps_cf ? eps[((long long)b * z + c) * T + t] : eps[r * z + c];
                g = d * e * 0.5f * expf(0.5f * lvr) + ks * valid * (expf(lvr) - 1.f);
            }
        }
        dparams[i] = __float2bfloat16(g);
    }
}

// ---------------------------------------------------------------------------------------------- dropout keep bits
__global__ void dropout_bits_kernel(uint32_t* __restrict__ bits, long long n_words, float p, unsigned long long seed,
                                    unsigned long long offset) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)tid, offset, &st);
    for (long long w = tid; w < n_words; w += stride) {
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 u = curand_uniform4(&st);
            word |= (u.x > p ? 1u : 0u) << (q * 4) | (u.y > p ? 1u : 0u) << (q * 4 + 1) |
                    (u.z > p ? 1u : 0u) << (q * 4 + 2) | (u.w > p ? 1u : 0u) << (q * 4 + 3);
        }
        bits[w] = word;
    }
}

// ---------------------------------------------------------------------------------------------- N(0, 1) noise
// out[i] ~ N(0, 1): Philox4x32-10 counter streams (subsequence = thread, offset = `offset`), Box-Muller, 4 values per draw.
// The noise of the posterior sample and of q_sample (distributions.py:37-41, LM:1409) when the caller supplies none.
__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)tid, offset, &st);
    for (long long i = tid * 4; i < n; i += stride * 4) {
        const float4 v = curand_normal4(&st);
        if (i + 3 < n) {
            *reinterpret_cast<float4*>(out + i) = v;
        } else {
            const float e[4] = {v.x, v.y, v.z, v.w};
            for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = e[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------- small fp32 (time MLP)
__global__ void silu_kernel(const float* __restrict__ pre, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = pre[i];
        out[i] = v / (1.f + expf(-v));
    }
}
__global__ void silu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dout, float* __restrict__ dpre,
                                long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = pre[i], s = 1.f / (1.f + expf(-v));
        dpre[i] = dout[i] * s * (1.f + v * (1.f - s));
    }
}

// dW[n, k] += sum_m dY[m, n] X[m, k] ; db[n] += sum_m dY[m, n].  M is small (utterances per batch).
__global__ void __launch_bounds__(TR_THREADS)
linear_f32_wgrad_kernel(const float* __restrict__ dY, long long ldy, const float* __restrict__ X, int M, long long N,
                        int K, float* __restrict__ dW, float* __restrict__ db) {
    const int k4s = (K + 3) / 4;
    const long long total = N * k4s;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / k4s;
        const int k = (int)(i % k4s) * 4;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        float sb = 0.f;
        for (int m = 0; m < M; ++m) {
            const float d = dY[(long long)m * ldy + n];
            sb += d;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (k + q < K) a[q] += d * X[(long long)m * K + k + q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k + q < K) dW[n * K + k + q] += a[q];
        if (db && k == 0) db[n] += sb;
    }
}

// dX[m, k] += sum_n dY[m, n] W[n, k]: a block owns a (k chunk of 256, n chunk) and 8 rows of m at a time
constexpr int LB_NCH = 256;
__global__ void __launch_bounds__(TR_THREADS)
linear_f32_dgrad_kernel(const float* __restrict__ dY, long long ldy, const float* __restrict__ W, int M, long long N,
                        int K, float* __restrict__ dX) {
    __shared__ float sd[8][LB_NCH];
    const int kch = (K + TR_THREADS - 1) / TR_THREADS;
    const long long nch = (N + LB_NCH - 1) / LB_NCH;
    const int mch = (M + 7) / 8;
    for (long long blk = blockIdx.x; blk < (long long)kch * nch * mch; blk += gridDim.x) {
        const int kc = (int)(blk % kch);
        const long long nc = (blk / kch) % nch;
        const int m0 = (int)(blk / ((long long)kch * nch)) * 8;
        const int k = kc * TR_THREADS + threadIdx.x;
        const long long n0 = nc * LB_NCH;
        const int nn = (int)min((long long)LB_NCH, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * LB_NCH; i += TR_THREADS) {
            const int mi = i / LB_NCH, ni = i % LB_NCH;
            sd[mi][ni] = (m0 + mi < M && ni < nn) ? dY[(long long)(m0 + mi) * ldy + n0 + ni] : 0.f;
        }
        __syncthreads();
        if (k < K) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            for (int ni = 0; ni < nn; ++ni) {
                const float w = W[(n0 + ni) * K + k];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += sd[i][ni] * w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (m0 + i < M) atomicAdd(dX + (long long)(m0 + i) * K + k, acc[i]);
        }
    }
}

// d weights[j] += sum_m ( dfeat[m, 1+j] cos(f) - dfeat[m, 1+half+j] sin(f) ) * t_m * 2 pi,  f = t_m w_j 2 pi  (LM:111-116)
__global__ void time_features_bwd_kernel(const int* __restrict__ steps, const float* __restrict__ w,
                                         const float* __restrict__ dfeat, int M, int half, float* __restrict__ dw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= half) return;
    const int width = 2 * half + 1;
    float acc = 0.f;
    for (int m = 0; m < M; ++m) {
        const float t = (float)steps[m];
        const float f = t * w[j] * 2.f * 3.14159265358979323846f;
        const float tp = t * 2.f * 3.14159265358979323846f;
        acc += (dfeat[(long long)m * width + 1 + j] * cosf(f) - dfeat[(long long)m * width + 1 + half + j] * sinf(f)) * tp;
    }
    dw[j] += acc;
}

// fp32 [rows, C] (ld lds) += bf16 [rows, ld] columns [col0, col0 + C)   (gathering a bf16 GEMM result into an fp32 grad)
__global__ void add_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, long long rows, int ld, int col0, int C,
                                       float* __restrict__ dst, int ldd, int accumulate) {
    const int groups = C / 8;
    const long long total = rows * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / groups;
        const int c = (int)(i % groups) * 8;
        float v[8];
        unpack8(*reinterpret_cast<const uint4*>(src + r * ld + col0 + c), v);
        float4* d = reinterpret_cast<float4*>(dst + r * ldd + c);
        float4 a = accumulate ? d[0] : make_float4(0.f, 0.f, 0.f, 0.f), bq = accumulate ? d[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3];
        bq.x += v[4]; bq.y += v[5]; bq.z += v[6]; bq.w += v[7];
        d[0] = a;
        d[1] = bq;
    }
}

}  // namespace dn

using namespace dn;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
typedef __nv_bfloat16 bf;

extern "C" int dn_geglu_fwd(const void* h, int64_t rows, int32_t ip, void* m, void* stream) {
    if (!h || !m || rows <= 0 || ip <= 0 || ip % 128) return DN_EINVAL;
    geglu_fwd_kernel<<<tr_grid(rows * (ip / 8), TR_THREADS), TR_THREADS, 0, ST(stream)>>>((const bf*)h, rows, ip, (bf*)m);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_geglu_bwd(const void* h, const void* dm, int64_t rows, int32_t ip, void* dh, void* stream) {
    if (!h || !dm || !dh || rows <= 0 || ip <= 0 || ip % 128) return DN_EINVAL;
    geglu_bwd_kernel<<<tr_grid(rows * (ip / 8), TR_THREADS), TR_THREADS, 0, ST(stream)>>>((const bf*)h, (const bf*)dm, rows,
                                                                                         ip, (bf*)dh);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_wn_gate_fwd(const void* ur, void* y, int32_t B, int32_t T, int32_t C, int32_t G, const float* gb,
                              int64_t gb_t_stride, int32_t g_gb, const int32_t* t_idx, int32_t t_idx_stride,
                              void* stream) {
    if (!ur || !y || B <= 0 || T <= 0 || C <= 0 || C % 128 || G <= 0 || (gb && !t_idx)) return DN_EINVAL;
    wn_gate_fwd_kernel<<<tr_grid((long long)B * T * G * (C / 8), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        (const bf*)ur, (bf*)y, B, T, C, G, gb, gb_t_stride, g_gb, t_idx, t_idx_stride);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_wn_gate_bwd(const void* ur, const void* dy, void* dur, int32_t B, int32_t T, int32_t C, int32_t G,
                              const float* gb, int64_t gb_t_stride, int32_t g_gb, const int32_t* t_idx,
                              int32_t t_idx_stride, float* dgb, int64_t dgb_b_stride, int32_t g_dgb, void* stream) {
    if (!ur || !dy || !dur || B <= 0 || T <= 0 || C <= 0 || C % 128 || G <= 0 || (gb && !t_idx)) return DN_EINVAL;
    wn_gate_bwd_kernel<<<B * G * (C / 128), TR_THREADS, 0, ST(stream)>>>((const bf*)ur, (const bf*)dy, (bf*)dur, B, T, C, G,
                                                                         gb, gb_t_stride, g_gb, t_idx, t_idx_stride, dgb,
                                                                         dgb_b_stride, g_dgb);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_adarmsnorm_bwd(const float* x, const void* dy, float* dx, void* dx_bf16, int32_t B, int32_t T, int32_t C,
                                 const float* gamma_p, float* dgamma_p, const float* gb, int64_t gb_t_stride,
                                 const int32_t* t_idx, int32_t t_idx_stride, float* dgb, int64_t dgb_b_stride,
                                 void* stream) {
    if (!x || !dy || !dx || B <= 0 || T <= 0 || (gb && (!t_idx || !dgb))) return DN_EINVAL;
    int rpb = 64;
    const int chunks = (T + rpb - 1) / rpb;
    const int grid = B * chunks;
#define LAUNCH_NB(CC)                                                                                                   \
    adarmsnorm_bwd_kernel<CC><<<grid, TR_THREADS, 0, ST(stream)>>>(x, (const bf*)dy, dx, (bf*)dx_bf16, B, T, rpb, gamma_p, \
                                                                   dgamma_p, gb, gb_t_stride, t_idx, t_idx_stride, dgb,  \
                                                                   dgb_b_stride)
    switch (C) {
        case 512: LAUNCH_NB(512); break;
        case 768: LAUNCH_NB(768); break;
        default: return DN_EINVAL;
    }
#undef LAUNCH_NB
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_colsum_bf16(const void* src, int64_t rows, int32_t ld, int32_t col0, int32_t cols, float* out,
                              void* stream) {
    if (!src || !out || rows <= 0 || cols <= 0 || ld % 8 || col0 % 8 || cols % 8) return DN_EINVAL;
    const int cg_chunks = ((cols + 7) / 8 + 31) / 32;
    const int rpb = 256;
    const long long rchunks = (rows + rpb - 1) / rpb;
    colsum_kernel<<<(int)(cg_chunks * rchunks), TR_THREADS, 0, ST(stream)>>>((const bf*)src, rows, ld, col0, cols, rpb, out);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_train_noise(const float* z_lat, const float* eps0, const float* eps, float beta0, const float* coef,
                              const int32_t* t_idx, int32_t B, int32_t T, int32_t z, float* x_t, void* x_bf16, int32_t ldx,
                              void* stream) {
    if (!z_lat || !eps0 || !eps || !coef || !t_idx || !x_t || !x_bf16 || B <= 0 || T <= 0 || z <= 0 || z % 4 || ldx % 4 ||
        ldx < z)
        return DN_EINVAL;
    train_noise_kernel<<<tr_grid((long long)B * T * (ldx / 4), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        z_lat, eps0, eps, beta0, coef, t_idx, B, T, z, x_t, (bf*)x_bf16, ldx);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_noise_loss(const float* pred, int32_t lde, const float* eps, const int32_t* lengths, const float* coef,
                             const int32_t* t_idx, int32_t B, int32_t T, int32_t z, float* loss, void* dpred, int32_t ldd,
                             float grad_scale, const void* dx1, int32_t ld1, void* stream) {
    if (!pred || !eps || !lengths || !coef || !t_idx || !loss || B <= 0 || T <= 0 || z <= 0 || z % 4 || lde % 4 || ldd % 4 ||
        ldd < z || (dx1 && (ld1 % 4 || ld1 < z)))
        return DN_EINVAL;
    noise_loss_kernel<<<tr_grid((long long)B * T * (ldd / 4), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        pred, lde, eps, lengths, coef, t_idx, B, T, z, loss, (bf*)dpred, ldd, grad_scale, (const bf*)dx1, ld1);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_pred_x1(const float* x_t, const float* pred, int32_t lde, const float* coef, const int32_t* t_idx,
                          int32_t B, int32_t T, int32_t z, void* x_bf16, int32_t ldx, void* stream) {
    if (!x_t || !pred || !coef || !t_idx || !x_bf16 || B <= 0 || T <= 0 || z <= 0 || z % 4 || lde % 4 || ldx % 4 || ldx < z)
        return DN_EINVAL;
    pred_x1_kernel<<<tr_grid((long long)B * T * (ldx / 4), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        x_t, pred, lde, coef, t_idx, B, T, z, (bf*)x_bf16, ldx);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_decode_losses(const float* recon, const float* audio, int32_t C, const float* logits, int32_t ld,
                                int32_t V, const int64_t* units, const int32_t* lengths, int32_t B, int32_t T, double* out6,
                                void* stream) {
    if (!recon || !audio || !logits || !units || !lengths || !out6 || B <= 0 || T <= 0 || V <= 0 || ld < V) return DN_EINVAL;
    DN_CUDA_OK(cudaMemsetAsync(out6, 0, 6 * sizeof(double), ST(stream)));
    decode_losses_kernel<<<tr_grid((long long)B * T, TR_THREADS / 32), TR_THREADS, 0, ST(stream)>>>(
        recon, audio, C, logits, ld, V, (const long long*)units, lengths, B, T, out6);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_vae_kl(const float* params, int32_t ldp, const int32_t* lengths, int32_t B, int32_t T, int32_t z, float* kl,
                         void* stream) {
    if (!params || !lengths || !kl || B <= 0 || T <= 0 || z <= 0 || ldp < 2 * z) return DN_EINVAL;
    vae_kl_kernel<<<tr_grid((long long)B * T * z, TR_THREADS), TR_THREADS, 0, ST(stream)>>>(params, ldp, lengths, B, T, z, kl);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_vae_reparam_bwd(const float* params, int32_t ldp, const float* eps, int32_t eps_channel_first, const void* dz,
                                  int32_t ldz, const int32_t* lengths, int32_t B, int32_t T, int32_t z, float kl_scale,
                                  void* dparams, int32_t ldd, void* stream) {
    if (!params || !eps || !dz || !lengths || !dparams || B <= 0 || T <= 0 || z <= 0 || ldp < 2 * z || ldz < z || ldd < 2 * z)
        return DN_EINVAL;
    vae_reparam_bwd_kernel<<<tr_grid((long long)B * T * ldd, TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        params, ldp, eps, eps_channel_first, (const bf*)dz, ldz, lengths, B, T, z, kl_scale, (bf*)dparams, ldd);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_lsnll_bwd(const float* logits, int32_t ld, int32_t V, const int64_t* units, int64_t rows, const double* stats,
                            float eps_ls, float nll_scale, void* dlogits, int32_t ldd, void* stream) {
    if (!logits || !units || !stats || !dlogits || rows <= 0 || V <= 1 || ld < V || ldd < V || ldd % 2) return DN_EINVAL;
    lsnll_bwd_kernel<<<tr_grid(rows, TR_THREADS / 32), TR_THREADS, 0, ST(stream)>>>(logits, ld, V, (const long long*)units, rows,
                                                                                   stats, eps_ls, nll_scale, (bf*)dlogits, ldd);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_recon_grad(const float* recon, const float* audio, const float* d_lm, const int32_t* lengths, int32_t B,
                             int32_t T, int32_t C, const double* stats, float mse_scale, void* out, void* stream) {
    if (!recon || !audio || !d_lm || !lengths || !stats || !out || B <= 0 || T <= 0 || C <= 0 || C % 4) return DN_EINVAL;
    recon_grad_kernel<<<tr_grid((long long)B * T * (C / 4), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(
        recon, audio, d_lm, lengths, B, T, C, stats, mse_scale, (bf*)out);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    if (!out || n <= 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return DN_EINVAL;
    const long long threads = (n + 3) / 4;
    long long blocks = (threads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    randn_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, seed, offset);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_dropout_bits(uint32_t* bits, int64_t n_words, float p, uint64_t seed, uint64_t offset, void* stream) {
    if (!bits || n_words <= 0 || p < 0.f || p >= 1.f) return DN_EINVAL;
    dropout_bits_kernel<<<tr_grid(n_words, TR_THREADS * 4), TR_THREADS, 0, ST(stream)>>>(bits, n_words, p, seed, offset);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_silu(const float* pre, float* out, int64_t n, void* stream) {
    if (!pre || !out || n <= 0) return DN_EINVAL;
    silu_kernel<<<tr_grid(n, TR_THREADS), TR_THREADS, 0, ST(stream)>>>(pre, out, n);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_silu_bwd(const float* pre, const float* dout, float* dpre, int64_t n, void* stream) {
    if (!pre || !dout || !dpre || n <= 0) return DN_EINVAL;
    silu_bwd_kernel<<<tr_grid(n, TR_THREADS), TR_THREADS, 0, ST(stream)>>>(pre, dout, dpre, n);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_linear_f32_bwd(const float* dY, int64_t ldy, const float* X, const float* W, int32_t M, int64_t N,
                                 int32_t K, float* dW, float* db, float* dX, void* stream) {
    if (!dY || M <= 0 || N <= 0 || K <= 0 || (dW && !X) || (dX && !W)) return DN_EINVAL;
    if (dW) {
        linear_f32_wgrad_kernel<<<tr_grid(N * ((K + 3) / 4), TR_THREADS), TR_THREADS, 0, ST(stream)>>>(dY, ldy, X, M, N, K, dW,
                                                                                                     db);
        DN_LAUNCH_CHECK();
        count_launch();
    }
    if (dX) {
        const long long blocks = (long long)((K + TR_THREADS - 1) / TR_THREADS) * ((N + LB_NCH - 1) / LB_NCH) * ((M + 7) / 8);
        linear_f32_dgrad_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), TR_THREADS, 0, ST(stream)>>>(dY, ldy, W, M, N,
                                                                                                           K, dX);
        DN_LAUNCH_CHECK();
        count_launch();
    }
    return 0;
}

extern "C" int dn_time_features_bwd(const int32_t* steps, const float* w, const float* dfeat, int32_t M, int32_t half,
                                    float* dw, void* stream) {
    if (!steps || !w || !dfeat || !dw || M <= 0 || half <= 0) return DN_EINVAL;
    time_features_bwd_kernel<<<(half + 127) / 128, 128, 0, ST(stream)>>>(steps, w, dfeat, M, half, dw);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_add_bf16_to_f32(const void* src, int64_t rows, int32_t ld, int32_t col0, int32_t C, float* dst,
                                  int32_t ldd, int32_t accumulate, void* stream) {
    if (!src || !dst || rows <= 0 || C <= 0 || C % 8 || ld % 8 || col0 % 8 || ldd % 4) return DN_EINVAL;
    add_bf16_to_f32_kernel<<<tr_grid(rows * (C / 8), TR_THREADS), TR_THREADS, 0, ST(stream)>>>((const bf*)src, rows, ld, col0,
                                                                                             C, dst, ldd, accumulate);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
