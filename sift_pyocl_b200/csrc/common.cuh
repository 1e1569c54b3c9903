// common.cuh -- shared helpers for the sm_100a kernels of libsiftb200.so
//
// Numerics contract: this translation unit is compiled with -fmad=false, so no
// implicit contraction happens anywhere; the convolution uses explicit __fmaf_rn per tap (the
// OpenCL reference's `sum += in*filter` under the default FP_CONTRACT ON).  OpenCL built-ins with
// device-defined error (exp, atan2, sin, cos, pow, rsqrt) are evaluated in double and rounded once
// to fp32 ("correctly rounded"), which makes the results reproducible against the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SIFTB_M_PI_F 3.14159274101257f
#define SIFTB_M_1_PI_F 0.318309886183791f
#define SIFTB_MAX_TAPS 64

struct Taps {
    float f[SIFTB_MAX_TAPS];  // passed by value: lands in the constant bank, statically indexed when unrolled
};

__device__ __forceinline__ float cr_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float cr_sinf(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cr_cosf(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float cr_exp2f(float x) { return (float)exp2((double)x); }
__device__ __forceinline__ float cr_rsqrtf(float x) { return (float)(1.0 / sqrt((double)x)); }

// reference mirror rule, convolution.cl:41-50: p<0 -> -p-1 ; p>=dim -> 2*dim-1-p
__device__ __forceinline__ int mirror_index(int p, int dim) {
    if (p < 0) p = -p - 1;
    if (p >= dim) p = 2 * dim - 1 - p;
    // positions further than one reflection away only occur for tile padding outside the image
    p = max(0, min(dim - 1, p));
    return p;
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// warp-aggregated append: returns the slot of this lane (or -1 when !pred). All 32 lanes must call.
__device__ __forceinline__ int warp_append(bool pred, int *counter) {
    unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return -1;
    int leader = __ffs(m) - 1;
    int base = 0;
    if ((threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(m & lanemask_lt()) : -1;
}

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned float_to_ordered(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
