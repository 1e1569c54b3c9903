// Throughput probe for the byte-SAD inner loop of k_match_l1 on sm_100a: candidates for sum |a-b| over 4 packed u8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sad_probe tools/sad_probe.cu && tools/sad_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1 + i); acc[i] = i; }
    uint32_t b = seed ^ threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) acc[i] = __vsadu4(a[i], b) + acc[i];                       // VABSDIFF4.U8.ACC
            if (MODE == 1) acc[i] = __dp4a(__vminu4(a[i], b), 0x01010101u, acc[i]);  // min + dp4a (sum of mins)
            if (MODE == 2) acc[i] = __dp4a(a[i], b, acc[i]);                          // dp4a alone
            if (MODE == 3) acc[i] = __vminu4(a[i], b) + acc[i];                       // packed min alone (+add)
            if (MODE == 4) { uint32_t d = __vabsdiffu4(a[i], b); acc[i] = __dp4a(d, 0x01010101u, acc[i]); }
            if (MODE == 6) { if (i & 1) acc[i] = __vsadu4(a[i], b) + acc[i]; else acc[i] = __dp4a(__vabsdiffu4(a[i], b), 0x01010101u, acc[i]); }
            if (MODE == 7) { if ((i & 3) == 0) acc[i] = __vsadu4(a[i], b) + acc[i]; else acc[i] = __dp4a(__vabsdiffu4(a[i], b), 0x01010101u, acc[i]); }
            if (MODE == 8) { uint32_t d = __vabsdiffu4(a[i], b); acc[i] += (d & 0x00ff00ffu) + ((d >> 8) & 0x00ff00ffu); }  // 16-bit lane sums
            if (MODE == 9) { acc[i] ^= __vabsdiffu4(a[i], b); }  // VABSDIFF4 rate (+ xor)
            if (MODE == 5) { uint32_t d = (a[i] | 0x80808080u) - (b & 0x7f7f7f7fu); acc[i] += d; }  // plain int ops ref
        }
        b = b * 1664525u + 1013904223u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name) {
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(d, 12345u, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 12345u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)iters * 8;  // 4-byte group operations
    printf("%-28s %.3f ms  %.1f G group-ops/s  = %.1f per clk per SM (at 1.965 GHz)\n", name, ms, ops / ms * 1e-6,
           ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    run<0>("vsadu4 (+acc)");
    run<1>("vminu4 + dp4a");
    run<2>("dp4a");
    run<3>("vminu4 + iadd");
    run<4>("vabsdiffu4 + dp4a");
    run<5>("lop/iadd reference (3 ops)");
    run<6>("1 vsadu4 : 1 (vabsdiff+dp4a)");
    run<7>("1 vsadu4 : 3 (vabsdiff+dp4a)");
    run<8>("vabsdiffu4 + 16-bit lane sums");
    run<9>("vabsdiffu4 + xor");
    return 0;
}
