// Integer tail of the normalization pass: warp-shuffle argmax over the 1004-way unit logits (LM:1450-1451),
// warp-ballot segmented-scan run-length reduction (_reduce_tgt, repr_to_repr_unit_dataset.py:92-113 ==
// diff_norm_synthesis.py:25-46) and the match/total accuracy counters (LM:1453-1454).  Bit-exact integer work.
#include "common.cuh"

namespace dn {

// ------------------------------------------------------------------------------------------------ argmax
// order: NaN > everything (torch.argmax), otherwise larger value, ties -> smaller index
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) return vn && (!bn || i < bi);
    return v > bv || (v == bv && i < bi);
}

template <bool BF16, bool VEC4>
__global__ void __launch_bounds__(256)
argmax_units_kernel(const void* __restrict__ logits, long long rows, int C, int ld, int offset,
                    long long* __restrict__ units) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        if (BF16) {
            const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(logits) + r * ld;
            for (int c = lane * 2; c < C; c += 64) {  // 4-byte loads, coalesced
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p + c);
                const float v0 = __low2float(h), v1 = __high2float(h);
                if (better(v0, c, bv, bi)) { bv = v0; bi = c; }
                if (c + 1 < C && better(v1, c + 1, bv, bi)) { bv = v1; bi = c + 1; }
            }
        } else if (VEC4) {
            // 16-byte loads, all of a row's loads in flight before the first compare (8 x float4 per lane at C = 1004)
            const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(logits) + r * ld);
            const int n4 = C >> 2;
            float4 v[8];
            for (int base = 0; base < n4; base += 256) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = base + k * 32 + lane;
                    v[k] = q < n4 ? __ldcs(p + q) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = (base + k * 32 + lane) * 4;
                    if (c < C) {  // padding lanes hold -inf, which must not beat an all -inf row's index 0
                        if (better(v[k].x, c, bv, bi)) { bv = v[k].x; bi = c; }
                        if (better(v[k].y, c + 1, bv, bi)) { bv = v[k].y; bi = c + 1; }
                        if (better(v[k].z, c + 2, bv, bi)) { bv = v[k].z; bi = c + 2; }
                        if (better(v[k].w, c + 3, bv, bi)) { bv = v[k].w; bi = c + 3; }
                    }
                }
            }
            const float* ps = reinterpret_cast<const float*>(logits) + r * ld;
            for (int c = (n4 << 2) + lane; c < C; c += 32) {
                const float x = ps[c];
                if (better(x, c, bv, bi)) { bv = x; bi = c; }
            }
        } else {
            const float* p = reinterpret_cast<const float*>(logits) + r * ld;
            for (int c = lane; c < C; c += 32) {
                const float v = p[c];
                if (better(v, c, bv, bi)) { bv = v; bi = c; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) units[r] = (long long)bi - offset;
    }
}

// second stage of the GEMM's argmax epilogue (DN_EPI_ARGMAX): fold a row's (value, index) partials
__global__ void argmax_combine_kernel(const float* __restrict__ partials, long long rows, int parts, int offset,
                                      long long* __restrict__ units) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const float2* p = reinterpret_cast<const float2*>(partials) + r * parts;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int k = 0; k < parts; ++k) {
            const float2 q = p[k];
            const int qi = __float_as_int(q.y);
            if (qi != 0x7fffffff && better(q.x, qi, bv, bi)) { bv = q.x; bi = qi; }
        }
        units[r] = (long long)bi - offset;
    }
}

// ------------------------------------------------------------------------------------------------ run-length reduce
constexpr int RL_THREADS = 256;
__global__ void __launch_bounds__(RL_THREADS)
reduce_tgt_kernel(const long long* __restrict__ units, const int* __restrict__ lengths, int T,
                  long long* __restrict__ dedup, long long* __restrict__ duration, long long* __restrict__ keep,
                  int* __restrict__ counts) {
    __shared__ int warp_tot[RL_THREADS / 32];
    const int b = blockIdx.x;
    const int n = lengths[b];
    const long long base = (long long)b * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) {  // reference quirk: the trailing duration is appended unconditionally (:45)
        if (tid == 0) {
            counts[b] = 0;
            duration[base] = 1;
        }
        return;
    }
    int carry = 0;  // runs found in previous chunks (uniform across the block)
    for (int start = 0; start < n; start += RL_THREADS) {
        const int i = start + tid;
        long long u = 0;
        bool flag = false;
        if (i < n) {
            u = units[base + i];
            flag = (i == 0) || (u != units[base + i - 1]);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        const int prefix = __popc(ballot & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(ballot);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < RL_THREADS / 32; ++w) {
            const int c = warp_tot[w];
            if (w < warp) woff += c;
            tot += c;
        }
        if (flag) {
            const int j = carry + woff + prefix;
            dedup[base + j] = u;
            keep[base + j] = i;
        }
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) counts[b] = carry;
    __threadfence_block();
    __syncthreads();
    for (int j = tid; j < carry; j += RL_THREADS) {
        const long long nxt = (j + 1 < carry) ? keep[base + j + 1] : (long long)n;
        duration[base + j] = nxt - keep[base + j];
    }
}

__global__ void unit_accuracy_kernel(const long long* __restrict__ units, const long long* __restrict__ ref,
                                     const int* __restrict__ lengths, int B, int T,
                                     unsigned long long* __restrict__ out2) {
    unsigned long long match = 0, total = 0;
    const long long n = (long long)B * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / T), t = (int)(i % T);
        if (t < lengths[b]) {
            ++total;
            match += units[i] == ref[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        match += __shfl_xor_sync(0xffffffffu, match, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if ((threadIdx.x & 31) == 0 && total) {
        atomicAdd(out2, match);
        atomicAdd(out2 + 1, total);
    }
}

}  // namespace dn

using namespace dn;

extern "C" int dn_argmax_units(const void* logits, int32_t logits_bf16, int64_t rows, int32_t C, int32_t ld,
                               int32_t offset, int64_t* units, void* stream) {
    if (!logits || !units || rows <= 0 || C <= 0 || ld < C || (logits_bf16 && (ld & 1))) return DN_EINVAL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long blocks = (rows + 7) / 8;
    const int grid = (int)(blocks > 148 * 16 ? 148 * 16 : blocks);
    if (logits_bf16)
        argmax_units_kernel<true, false><<<grid, 256, 0, st>>>(logits, rows, C, ld, offset, reinterpret_cast<long long*>(units));
    else if (ld % 4 == 0 && !(reinterpret_cast<uintptr_t>(logits) & 15))
        argmax_units_kernel<false, true><<<grid, 256, 0, st>>>(logits, rows, C, ld, offset, reinterpret_cast<long long*>(units));
    else
        argmax_units_kernel<false, false><<<grid, 256, 0, st>>>(logits, rows, C, ld, offset, reinterpret_cast<long long*>(units));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_reduce_tgt(const int64_t* units, const int32_t* lengths, int32_t B, int32_t T, int64_t* dedup,
                             int64_t* duration, int64_t* index_to_keep, int32_t* counts, void* stream) {
    if (!units || !lengths || !dedup || !duration || !index_to_keep || !counts || B <= 0 || T <= 0) return DN_EINVAL;
    reduce_tgt_kernel<<<B, RL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(units), lengths, T, reinterpret_cast<long long*>(dedup),
        reinterpret_cast<long long*>(duration), reinterpret_cast<long long*>(index_to_keep), counts);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_argmax_combine(const float* partials, int64_t rows, int32_t parts, int32_t offset, int64_t* units,
                                 void* stream) {
    if (!partials || !units || rows <= 0 || parts <= 0 || (reinterpret_cast<uintptr_t>(partials) & 7)) return DN_EINVAL;
    const long long blocks = (rows + 255) / 256;
    argmax_combine_kernel<<<(int)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        partials, rows, parts, offset, reinterpret_cast<long long*>(units));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_unit_accuracy(const int64_t* units, const int64_t* ref_units, const int32_t* lengths, int32_t B,
                                int32_t T, int64_t* out2, void* stream) {
    if (!units || !ref_units || !lengths || !out2 || B <= 0 || T <= 0) return DN_EINVAL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DN_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(int64_t), st));
    const long long n = (long long)B * T;
    const int grid = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    unit_accuracy_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(units),
                                               reinterpret_cast<const long long*>(ref_units), lengths, B, T,
                                               reinterpret_cast<unsigned long long*>(out2));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
