// k_blur.cuh -- fused separable Gaussian blur + DoG + decimation.
//
// Replaces, per blur, four reference launches: convolution.cl:16 horizontal_convolution,
// convolution.cl:62 vertical_convolution (via plan.py:571 _gaussian_convolution, with its `tmp`
// plane round trip), algebra.cl:18 combine (DoG[s] = G[s] - G[s+1], plan.py:619-623) and, for
// G[3], preprocess.cl:266 shrink (next-octave base = G[3][::2, ::2], plan.py:739-745).  The first
// blur of an image additionally fuses preprocess.cl:238 normalizes into its loads.
//
// Arithmetic is bit-identical to the oracle: per output pixel, sum = fmaf(in, taps[n-1-j], sum)
// for j = 0..n-1 along x, rounded to fp32, then the same along y (convolution.cl:45-52,84-98).
#pragma once
#include "common.cuh"

#define BLUR_TW 128
#define BLUR_TH 32
#define BLUR_THREADS 256

struct BlurArgs {
    const float *in;   // source plane (G[s], or the raw image for the first blur)
    int in_pitch;
    float *outG;       // G[s+1]
    int out_pitch;
    float *outD;       // DoG[s] = in - outG (nullable)
    float *outHalf;    // outG[::2, ::2] (nullable)
    int half_pitch, half_w, half_h;
    int w, h;
    int ntaps;
    const unsigned *norm_mm;  // non-null: apply 255*(x-min)/(max-min) to every loaded pixel
};

// v1: generic tap count, tile (BLUR_TH+2c) x (BLUR_TW+2c) staged in shared memory with mirrored borders.
__global__ void __launch_bounds__(BLUR_THREADS) k_blur_generic(BlurArgs a, Taps taps) {
    extern __shared__ float smem[];
    const int n = a.ntaps;
    const int c = n >> 1;  // odd sizes only (utils.kernel_size(odd=True))
    const int tw = BLUR_TW + 2 * c;          // staged tile width
    const int tws = tw | 1;                  // odd pitch: conflict-free column walks
    const int th = BLUR_TH + 2 * c;
    float *tile = smem;                      // th x tws
    float *hbuf = smem + th * tws;           // th x BLUR_TW
    const int x0 = blockIdx.x * BLUR_TW, y0 = blockIdx.y * BLUR_TH;

    float mn = 0.f, den = 1.f;
    const bool norm = a.norm_mm != nullptr;
    if (norm) {
        mn = ordered_to_float(a.norm_mm[0]);
        den = ordered_to_float(a.norm_mm[1]) - mn;
    }
    for (int i = threadIdx.x; i < th * tw; i += BLUR_THREADS) {
        int ty = i / tw, tx = i - ty * tw;
        int gy = mirror_index(y0 - c + ty, a.h), gx = mirror_index(x0 - c + tx, a.w);
        float v = __ldg(a.in + (long)gy * a.in_pitch + gx);
        if (norm) v = (255.0f * (v - mn)) / den;  // preprocess.cl:250
        tile[ty * tws + tx] = v;
    }
    __syncthreads();
    // horizontal pass over all th rows
    for (int i = threadIdx.x; i < th * BLUR_TW; i += BLUR_THREADS) {
        int ty = i / BLUR_TW, tx = i - ty * BLUR_TW;
        const float *p = tile + ty * tws + tx;
        float sum = 0.0f;
        for (int j = 0; j < n; j++) sum = __fmaf_rn(p[j], taps.f[n - 1 - j], sum);
        hbuf[ty * BLUR_TW + tx] = sum;
    }
    __syncthreads();
    // vertical pass + epilogue
    for (int i = threadIdx.x; i < BLUR_TH * BLUR_TW; i += BLUR_THREADS) {
        int ty = i / BLUR_TW, tx = i - ty * BLUR_TW;
        int gx = x0 + tx, gy = y0 + ty;
        if (gx >= a.w || gy >= a.h) continue;
        const float *p = hbuf + ty * BLUR_TW + tx;
        float sum = 0.0f;
        for (int j = 0; j < n; j++) sum = __fmaf_rn(p[j * BLUR_TW], taps.f[n - 1 - j], sum);
        a.outG[(long)gy * a.out_pitch + gx] = sum;
        if (a.outD) a.outD[(long)gy * a.out_pitch + gx] = tile[(ty + c) * tws + tx + c] - sum;
        if (a.outHalf && !(gx & 1) && !(gy & 1) && (gx >> 1) < a.half_w && (gy >> 1) < a.half_h)
            a.outHalf[(long)(gy >> 1) * a.half_pitch + (gx >> 1)] = sum;
    }
}

static inline size_t blur_generic_smem(int ntaps) {
    int c = ntaps >> 1;
    int tw = BLUR_TW + 2 * c, tws = tw | 1, th = BLUR_TH + 2 * c;
    return (size_t)(th * tws + th * BLUR_TW) * sizeof(float);
}
