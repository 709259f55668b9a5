// MUFU.EX2 throughput probe: f32 vs packed f16x2 / bf16x2 (elements per clock per SM).  nvcc -arch=sm_100a -O3 -o mufu_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    // 8 independent chains per thread to cover MUFU latency
    float f[8];
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + 0.001f * (threadIdx.x + i);
        h[i] = 0x38003800u + threadIdx.x + i;  // two halves near 0.5
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            } else if (MODE == 1) {
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
            } else if (MODE == 2) {
                asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
            } else if (MODE == 3) {   // cvt pair + packed ex2 (the softmax form): 2 f32 -> f16x2 -> ex2
                uint32_t p;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(f[i]), "f"(f[i]));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p));
                h[i] ^= p;
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += f[i] + __uint_as_float(h[i]);
    if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int elems_per_op) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 4096, threads = 1024, blocks = sms * 2;
    k<MODE><<<blocks, threads>>>(out, 16, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * iters * 8;
    printf("%-34s %8.3f ms  %7.2f Gop/s  %7.2f Gelem/s  (= %.1f elem/clk/SM at the %.0f MHz max clock)\n", name, ms, ops / ms / 1e6,
           ops * elems_per_op / ms / 1e6, ops * elems_per_op / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3);
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.f16x2", 2);
    run<2>("ex2.approx.ftz.bf16x2", 2);
    run<3>("cvt.rn.f16x2.f32 + ex2.f16x2", 2);
    return 0;
}
