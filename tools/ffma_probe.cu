// FFMA issue-rate probe: same register-tiled sliding-window pattern as k_blur_tma's row pass, with the taps
// coming from (a) registers loaded from a kernel-parameter struct, (b) compile-time immediates, (c) packed
// fma.rn.f32x2.  Prints FMA per clock per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
struct Taps { float f[64]; };
__device__ constexpr float kT[27] = {0.0012f,0.0023f,0.0041f,0.0070f,0.0113f,0.0173f,0.0252f,0.0349f,0.0459f,0.0574f,0.0683f,0.0773f,0.0831f,0.0852f,
                                     0.0831f,0.0773f,0.0683f,0.0574f,0.0459f,0.0349f,0.0252f,0.0173f,0.0113f,0.0070f,0.0041f,0.0023f,0.0012f};
template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(float *out, Taps taps, int iters, long long *cycles) {
    constexpr int N = 27, RH = 16;
    float in[RH + N - 1];
#pragma unroll
    for (int i = 0; i < RH + N - 1; i++) in[i] = (float)(threadIdx.x + i) * 1e-3f;
    float acc[RH];
#pragma unroll
    for (int o = 0; o < RH; o++) acc[o] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < N; j++) {
                float t = taps.f[j];
#pragma unroll
                for (int o = 0; o < RH; o += 2) {
                    asm volatile("{ .reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %4}; mov.b64 c, {%0, %1}; fma.rn.f32x2 c, a, b, c; mov.b64 {%0, %1}, c; }"
                                 : "+f"(acc[o]), "+f"(acc[o + 1]) : "f"(in[o + j]), "f"(in[o + j + 1]), "f"(t));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) {
#pragma unroll
                for (int o = 0; o < RH; o++) acc[o] = __fmaf_rn(in[o + j], MODE == 1 ? kT[j] : taps.f[j], acc[o]);
            }
        }
#pragma unroll
        for (int i = 0; i < RH; i++) in[i] = acc[i] * 1e-3f;  // keep the loop from being hoisted
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int o = 0; o < RH; o++) s += acc[o];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    float *out; long long *cyc;
    const int grid = 148 * 8;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, 8);
    Taps t; for (int i = 0; i < 64; i++) t.f[i] = 0.03f + i * 1e-4f;
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++) {
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<grid, 256>>>(out, t, iters, cyc);
            if (mode == 1) probe<1><<<grid, 256>>>(out, t, iters, cyc);
            if (mode == 2) probe<2><<<grid, 256>>>(out, t, iters, cyc);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms, e0, e1);
        }
        double fma = (double)grid * 256 * 27.0 * 16 * iters;
        printf("mode %d (%s): %.3f ms, %.2f TFMA/s = %.1f FMA/clk/SM at 1.90 GHz (%s)\n", mode,
               mode == 0 ? "register taps" : mode == 1 ? "immediate taps" : "fma.f32x2", ms, fma / ms * 1e-9,
               fma / (ms * 1e-3) / 148 / 1.90e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
