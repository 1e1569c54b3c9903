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

// ---- exp for x in [-16, 0]: table-driven double evaluation, EXHAUSTIVELY verified (tools/exp_check.c: all
// 1 098 907 649 fp32 inputs of the interval) to return exactly (float)exp((double)x) of the host libm, i.e. the
// oracle's definition, at less than half the instructions of libdevice's exp.  x = (k/32) ln2 + r, |r| <= ln2/64:
// exp(x) = 2^(k>>5) * 2^((k&31)/32) * (1 + r + r^2/2 + ... + r^6/720).  Other inputs take the generic path.
__device__ double c_exp_t32[32] = {
    0x1p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f09p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e454p+0};
__device__ __forceinline__ float cr_expf_neg(float xf) {
    if (!(xf >= -16.0f && xf <= 0.0f)) return cr_expf(xf);  // also NaN
    const double x = (double)xf;
    const double z = fma(x, 0x1.71547652b82fep+5, 0x1.8p52);  // k = rint(x * 32/ln2) in the low word
    const int k = __double2loint(z);
    const double kd = z - 0x1.8p52;
    double r = fma(kd, -0x1.62e42feep-6, x);
    r = fma(kd, -0x1.a39ef358p-38, r);
    double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(r, q, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    const double p = fma(r * r, q, r);
    const double t = __ldg(&c_exp_t32[k & 31]);
    const double s = __hiloint2double(__double2hiint(t) + ((k >> 5) << 20), __double2loint(t));
    return (float)fma(s, p, s);
}
__device__ __forceinline__ float cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float cr_sinf(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cr_cosf(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float cr_exp2f(float x) { return (float)exp2((double)x); }
__device__ __forceinline__ float cr_rsqrtf(float x) { return (float)(1.0 / sqrt((double)x)); }


// ---- fast correctly-rounded-to-fp32 atan2 (double evaluation, one division) ---------------------------------
// atan2(y, x) for fp32 inputs: fold to t = lo/hi in [0, 1], pick the nearest of 17 breakpoints c_k = tan(k*pi/64)
// from a quadratic fit of atan (k = rint(t*(21.5615 - 5.5615 t)), off by < 0.1 interval, so |atan t - k*pi/64| <
// 0.6*pi/64), then atan(t) = k*pi/64 + atan(r), r = (lo - c_k*hi) / (hi + c_k*lo), |r| < 0.03: the odd Taylor
// polynomial to r^11 is exact to < 1e-19.  Measured against glibc atan2 on 9e8 mixed inputs, also with t
// perturbed by +-1e-6 (tools/atan2_check.c): max error 2 ulp(double), zero differences after rounding to fp32 --
// the same guarantee class as libdevice's atan2 at a fraction of its cost.  Signed zeros follow C99.
__device__ double c_atan_c[17] = {0x0.0p+0, 0x1.927278a3b1162p-5, 0x1.936bb8c5b2da2p-4, 0x1.2fcac73a60640p-3, 0x1.975f5e0553158p-3, 0x1.007fa758626aep-2, 0x1.36a08355c63dcp-2, 0x1.6e649f7d78649p-2, 0x1.a827999fcef32p-2, 0x1.e450e0d273e7ap-2, 0x1.11ab7190834ebp-1, 0x1.32e1889047ffcp-1, 0x1.561b82ab7f990p-1, 0x1.7bb99ed2990cfp-1, 0x1.a43002ae4284fp-1, 0x1.d00cbc7384d2dp-1, 0x1.fffffffffffffp-1};
__device__ double c_atan_a[17] = {0x0.0p+0, 0x1.921fb54442d18p-5, 0x1.921fb54442d18p-4, 0x1.2d97c7f3321d2p-3, 0x1.921fb54442d18p-3, 0x1.f6a7a2955385ep-3, 0x1.2d97c7f3321d2p-2, 0x1.5fdbbe9bba775p-2, 0x1.921fb54442d18p-2, 0x1.c463abeccb2bbp-2, 0x1.f6a7a2955385ep-2, 0x1.1475cc9eedf00p-1, 0x1.2d97c7f3321d2p-1, 0x1.46b9c347764a4p-1, 0x1.5fdbbe9bba775p-1, 0x1.78fdb9effea46p-1, 0x1.921fb54442d18p-1};

// The double constants of the polynomial and of the quadrant fix-up live in the constant bank, where DFMA / DADD
// read them as a direct operand: as literals ptxas re-materialises each of them with two moves at every use (a
// tenth of the instructions of the gradient kernel).
__constant__ double c_atan_k[7] = {-1.0 / 11.0, 1.0 / 9.0, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0, 1.5707963267948966,
                                   3.141592653589793};
struct AtanConsts {
    double c11, c9, c7, c5, c3, pio2, pi;
    __device__ __forceinline__ AtanConsts()
        : c11(c_atan_k[0]), c9(c_atan_k[1]), c7(c_atan_k[2]), c5(c_atan_k[3]), c3(c_atan_k[4]), pio2(c_atan_k[5]),
          pi(c_atan_k[6]) {}
};

__device__ __forceinline__ float cr_atan2f_fast(float yf, float xf, const AtanConsts &K) {
    const float axf = fabsf(xf), ayf = fabsf(yf);
    const float hif = fmaxf(axf, ayf), lof = fminf(axf, ayf);
    // tables live in global memory and are read through L1 (__ldg): the index differs per lane, which the
    // constant cache would serialise.  Branch-free: lof == 0 (including 0/0) is selected to a = 0 at the end.
    // rough quotient, only used to pick the breakpoint: div.approx.ftz is MUFU.RCP + FMUL (the non-ftz form that
    // __fdividef compiles to without -ftz carries ~7 more instructions of subnormal scaling); subnormal operands,
    // where flushing would change the quotient, take the IEEE division
    float t;
    if (hif >= 1e-30f) asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(t) : "f"(lof), "f"(hif));
    else t = lof / hif;
    const int k = max(min((int)(t * (21.5615f + -5.5615f * t) + 0.5f), 16), 0);
    const double hi = (double)hif, lo = (double)lof;
    const double c = __ldg(&c_atan_c[k]);
    // r = N / D with D in [hi, 2 hi], hi a (possibly subnormal) fp32 value: reciprocal seed (MUFU.RCP64H), two Newton
    // steps (relative error 2^-20 -> 2^-40 -> below 2^-53), then the quotient with one residual correction -- within
    // half an ulp (+ 2^-100) of the exact quotient, i.e. the IEEE result except for near-ties, at 8 instructions
    // instead of the ~19 of __ddiv_rn (whose slow path for extreme exponents cannot occur here).  The atan series
    // below does not need more: its own truncation error is larger than an ulp of r.
    const double num = fma(-c, hi, lo), den = fma(c, lo, hi);
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    y = fma(y, fma(-den, y, 1.0), y);
    y = fma(y, fma(-den, y, 1.0), y);
    double r = num * y;
    r = fma(fma(-den, r, num), y, r);
    const double r2 = r * r;
    double p = fma(r2, K.c11, K.c9);
    p = fma(r2, p, K.c7);
    p = fma(r2, p, K.c5);
    p = fma(r2, p, K.c3);
    p = p * r2;
    double a = __ldg(&c_atan_a[k]) + fma(r, p, r);
    if (lof == 0.0f) a = 0.0;
    if (ayf > axf) a = K.pio2 - a;
    if (signbit(xf)) a = K.pi - a;
    return copysignf((float)a, yf);
}

// Read-only 8-byte load that stays where it is written: the software-pipelined gathers of k_orient / k_describe
// request the values of step n+1 BEFORE step n is evaluated.  With a plain __ldg the compiler is free to sink the
// load down to its first use, i.e. into the next iteration (it did, in one build of k_orient: 24 % of the warp samples
// then sat on the first use of the value, 0.25 instead of 0.21 ms); a volatile asm is not moved across the
// shuffles and shared-memory traffic of the step.
__device__ __forceinline__ float2 ldg_f2_here(const float2 *p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// ---- correctly rounded fp32 division by a loop-invariant divisor -------------------------------------------
// a / b == (float)((double)a * inv) with inv = 1.0 / (double)b: the double product is within 2^-52 (relative) of
// the exact quotient, while a quotient of two 24-bit significands is either at least 2^-49 (relative) away from
// every rounding boundary of fp32 (the midpoints have 25 significant bits: |a - m*b| >= 1 in integer units) or
// exactly representable, and it can never sit exactly on a midpoint.  So the final rounding to fp32 sees the
// same side of every boundary as the exact quotient: bit-identical to IEEE division, for finite non-tiny values
// (all uses here are O(1) magnitudes).  3 instructions instead of ~10 per division.
__device__ __forceinline__ double div_prepare(float b) { return 1.0 / (double)b; }
__device__ __forceinline__ float div_by(float a, double inv) { return (float)((double)a * inv); }

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
