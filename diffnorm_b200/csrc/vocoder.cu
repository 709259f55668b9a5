// The step AFTER the normalization pass: the duration-aware unit vocoder — CodeHiFiGAN generator +
// duration predictor (fairseq/models/text_to_speech/codehifigan.py:49-76, hifigan.py:20-179, fastspeech2.py:117-151) on
// the reduced units the pass writes.  fp32 throughout (the reference runs it in fp32; a 16-bit tensor-core form is a
// later step): channel-first [C, L] activations of ONE utterance like the reference driver
// (examples/speech_to_speech/generate_waveform_from_code.py:78-96), CUDA-core kernels tiled through shared memory.
//   voc_conv1d_kernel    y[co, l] = act(b[co] + sum_ci sum_k w[co, ci, k] lrelu(x[ci, l + k d - pad])) (+ res) (accumulate)
//   voc_convt1d_kernel   ConvTranspose1d(k, stride u, padding (k - u) / 2): L -> u L
//   voc_layernorm_kernel LayerNorm over channels at every position (duration predictor)
//   voc_duration_kernel  dur = max(round_half_even(exp(log_dur) - 1), 1) and its exclusive prefix sum (one block)
//   voc_embed_repeat     x[c, j] = table[code[u(j)], c], u(j) = the unit whose duration interval covers frame j
#include "common.cuh"

namespace dn {

constexpr int VC_THREADS = 256;
constexpr int VC_LT = 64;      // output positions per block
constexpr int VC_CI = 8;       // input channels per shared-memory step

__device__ __forceinline__ float voc_lrelu(float v, float slope) { return v >= 0.f ? v : v * slope; }

// CO_T output channels x 64 positions per block; thread (ty, tx) owns CO_T/16 channels x 4 positions.
template <int CO_T>
__global__ void __launch_bounds__(VC_THREADS)
voc_conv1d_kernel(const float* __restrict__ x, int L, int Cin, const float* __restrict__ w, const float* __restrict__ bias,
                  int Cout, int K, int dil, int pad, float in_slope, int out_act, const float* __restrict__ res,
                  float out_scale, int accumulate, float* __restrict__ y) {
    constexpr int RC = CO_T / 16;          // channels per thread
    extern __shared__ float vsm[];
    const int span = VC_LT + (K - 1) * dil;            // input positions a block needs
    float* sx = vsm;                                    // [VC_CI][span]
    float* sw = vsm + VC_CI * span;                     // [CO_T][VC_CI][K]
    const int l0 = blockIdx.x * VC_LT, co0 = blockIdx.y * CO_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[RC][4];
#pragma unroll
    for (int r = 0; r < RC; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
    for (int ci0 = 0; ci0 < Cin; ci0 += VC_CI) {
        __syncthreads();
        for (int i = threadIdx.x; i < VC_CI * span; i += VC_THREADS) {
            const int c = i / span, s = i % span;
            const int ci = ci0 + c, l = l0 + s - pad;
            float v = 0.f;
            if (ci < Cin && l >= 0 && l < L) v = voc_lrelu(x[(long long)ci * L + l], in_slope);
            sx[i] = v;
        }
        for (int i = threadIdx.x; i < CO_T * VC_CI * K; i += VC_THREADS) {
            const int k = i % K, c = (i / K) % VC_CI, o = i / (K * VC_CI);
            const int co = co0 + o, ci = ci0 + c;
            sw[i] = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < VC_CI; ++c) {
            for (int k = 0; k < K; ++k) {
                float xv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) xv[q] = sx[c * span + tx + 16 * q + k * dil];
#pragma unroll
                for (int r = 0; r < RC; ++r) {
                    const float wv = sw[((ty + 16 * r) * VC_CI + c) * K + k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(wv, xv[q], acc[r][q]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RC; ++r) {
        const int co = co0 + ty + 16 * r;
        if (co >= Cout) continue;
        const float b = bias ? bias[co] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = l0 + tx + 16 * q;
            if (l >= L) continue;
            float v = acc[r][q] + b;
            if (out_act == 1) v = fmaxf(v, 0.f);
            else if (out_act == 2) v = tanhf(v);
            const long long o = (long long)co * L + l;
            if (res) v += res[o];
            v *= out_scale;
            y[o] = accumulate ? y[o] + v : v;
        }
    }
}

// ConvTranspose1d: y[co, j] = b[co] + sum_ci sum_{k : (j + pad - k) % stride == 0} lrelu(x[ci, (j + pad - k) / stride]) w[ci, co, k]
// One thread per (co, j); x rows are read through L1/L2 (each input element is reused K / stride times per channel).
__global__ void __launch_bounds__(VC_THREADS)
voc_convt1d_kernel(const float* __restrict__ x, int L, int Cin, const float* __restrict__ w, const float* __restrict__ bias,
                   int Cout, int K, int stride, int pad, float in_slope, float* __restrict__ y) {
    extern __shared__ float vsm[];                       // weights of this block's output channel: [Cin][K]
    const int Lo = L * stride;
    const int co = blockIdx.y;
    for (int i = threadIdx.x; i < Cin * K; i += VC_THREADS) vsm[i] = w[((long long)(i / K) * Cout + co) * K + (i % K)];
    __syncthreads();
    const int j = blockIdx.x * VC_THREADS + threadIdx.x;
    if (j >= Lo) return;
    // taps k with (j + pad - k) divisible by stride: k = (j + pad) % stride + m stride
    const int k0 = (j + pad) % stride;
    float acc = bias ? bias[co] : 0.f;
    for (int k = k0; k < K; k += stride) {
        const int l = (j + pad - k) / stride;
        if (l < 0 || l >= L) continue;
        float a = 0.f;
        for (int ci = 0; ci < Cin; ++ci) a = fmaf(voc_lrelu(x[(long long)ci * L + l], in_slope), vsm[ci * K + k], a);
        acc += a;
    }
    y[(long long)co * Lo + j] = acc;
}

// LayerNorm over the C channels of every position (x, y channel-first [C, L]); eps 1e-5 (torch default)
__global__ void voc_layernorm_kernel(const float* __restrict__ x, int C, int L, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    float mean = 0.f;
    for (int c = 0; c < C; ++c) mean += x[(long long)c * L + l];
    mean /= (float)C;
    float var = 0.f;
    for (int c = 0; c < C; ++c) {
        const float d = x[(long long)c * L + l] - mean;
        var += d * d;
    }
    const float inv = rsqrtf(var / (float)C + 1e-5f);
    for (int c = 0; c < C; ++c) y[(long long)c * L + l] = (x[(long long)c * L + l] - mean) * inv * gamma[c] + beta[c];
}

// dur[t] = max(round(exp(log_dur[t]) - 1), 1) with torch.round's half-to-even; start[t] = exclusive prefix sum, start[T] = total
__global__ void voc_duration_kernel(const float* __restrict__ log_dur, int T, long long* __restrict__ dur,
                                    long long* __restrict__ start) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long s = 0;
    for (int t = 0; t < T; ++t) {
        long long d = 1;
        if (log_dur) {
            d = (long long)rintf(expf(log_dur[t]) - 1.f);
            if (d < 1) d = 1;
        }
        dur[t] = d;
        start[t] = s;
        s += d;
    }
    start[T] = s;
}

// x[c, j] = table[code[u], c] for start[u] <= j < start[u + 1]
__global__ void voc_embed_repeat_kernel(const long long* __restrict__ code, int T, const float* __restrict__ table, int dim,
                                        const long long* __restrict__ start, int Lo, float* __restrict__ x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Lo) return;
    int lo = 0, hi = T - 1;                 // last u with start[u] <= j
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (start[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const float* row = table + code[lo] * (long long)dim;
    for (int c = 0; c < dim; ++c) x[(long long)c * Lo + j] = row[c];
}

}  // namespace dn

using namespace dn;
#define VST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dn_voc_conv1d(const float* x, int32_t L, int32_t Cin, const float* w, const float* bias, int32_t Cout, int32_t K,
                             int32_t dilation, int32_t pad, float in_slope, int32_t out_act, const float* res, float out_scale,
                             int32_t accumulate, float* y, void* stream) {
    if (!x || !w || !y || L <= 0 || Cin <= 0 || Cout <= 0 || K <= 0 || K > 16 || dilation <= 0 || pad < 0) return DN_EINVAL;
    const int span = VC_LT + (K - 1) * dilation;
    const int co_t = Cout >= 64 ? 64 : (Cout >= 32 ? 32 : 16);
    const size_t smem = (size_t)(VC_CI * span + co_t * VC_CI * K) * sizeof(float);
    dim3 grid((L + VC_LT - 1) / VC_LT, (Cout + co_t - 1) / co_t);
    if (co_t == 64)
        voc_conv1d_kernel<64><<<grid, VC_THREADS, smem, VST(stream)>>>(x, L, Cin, w, bias, Cout, K, dilation, pad, in_slope, out_act,
                                                                      res, out_scale, accumulate, y);
    else if (co_t == 32)
        voc_conv1d_kernel<32><<<grid, VC_THREADS, smem, VST(stream)>>>(x, L, Cin, w, bias, Cout, K, dilation, pad, in_slope, out_act,
                                                                      res, out_scale, accumulate, y);
    else
        voc_conv1d_kernel<16><<<grid, VC_THREADS, smem, VST(stream)>>>(x, L, Cin, w, bias, Cout, K, dilation, pad, in_slope, out_act,
                                                                      res, out_scale, accumulate, y);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_voc_conv_transpose1d(const float* x, int32_t L, int32_t Cin, const float* w, const float* bias, int32_t Cout,
                                       int32_t K, int32_t stride, int32_t pad, float in_slope, float* y, void* stream) {
    if (!x || !w || !y || L <= 0 || Cin <= 0 || Cout <= 0 || K <= 0 || stride <= 0 || pad < 0 || K - 2 * pad != stride)
        return DN_EINVAL;   // (L - 1) stride - 2 pad + K == L stride
    if ((size_t)Cin * K * sizeof(float) > 48 * 1024) return DN_EINVAL;
    dim3 grid((L * stride + VC_THREADS - 1) / VC_THREADS, Cout);
    voc_convt1d_kernel<<<grid, VC_THREADS, (size_t)Cin * K * sizeof(float), VST(stream)>>>(x, L, Cin, w, bias, Cout, K, stride, pad,
                                                                                          in_slope, y);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_voc_layernorm(const float* x, int32_t C, int32_t L, const float* gamma, const float* beta, float* y, void* stream) {
    if (!x || !y || !gamma || !beta || C <= 0 || L <= 0) return DN_EINVAL;
    voc_layernorm_kernel<<<(L + 127) / 128, 128, 0, VST(stream)>>>(x, C, L, gamma, beta, y);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_voc_durations(const float* log_dur, int32_t T, int64_t* dur, int64_t* start, void* stream) {
    if (!dur || !start || T <= 0) return DN_EINVAL;
    voc_duration_kernel<<<1, 32, 0, VST(stream)>>>(log_dur, T, reinterpret_cast<long long*>(dur), reinterpret_cast<long long*>(start));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

extern "C" int dn_voc_embed_repeat(const int64_t* code, int32_t T, const float* table, int32_t dim, const int64_t* start, int32_t Lo,
                                   float* x, void* stream) {
    if (!code || !table || !start || !x || T <= 0 || dim <= 0 || Lo <= 0) return DN_EINVAL;
    voc_embed_repeat_kernel<<<(Lo + 127) / 128, 128, 0, VST(stream)>>>(reinterpret_cast<const long long*>(code), T, table, dim,
                                                                      reinterpret_cast<const long long*>(start), Lo, x);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
