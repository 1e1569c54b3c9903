// k_warp.cuh -- inverse-mapped affine warp of LinearAlign (replaces transform.cl:22 `transform` and :116
// `transform_RGB`, launched from alignment.py:329-346).
//
// Output pixel (y, x) samples the input at (sy, sx) = A * (y, x) + b, bilinear (mode 1) or nearest (mode 0), with
// `fill` wherever the sample or one of its bilinear neighbours falls outside the image, and in the last half
// pixel of the image ("to be coherent with scipy", transform.cl:100-104).  The per-pixel arithmetic (products,
// then the sum, then the offset; weights next - t and t - prev; x-interpolation before y) is the reference's, so the
// result is bit-identical to the oracle; the kernel organisation is not: a thread produces four horizontally
// adjacent output pixels (their source samples lie on a short line segment, so a warp gathers from a few image rows)
// and writes them with one 128-bit store.
#pragma once
#include "common.cuh"

struct WarpMap {
    float ayy, ayx, axy, axx;  // sy = ayy*y + ayx*x + by ; sx = axy*y + axx*x + bx  (matrix rows 0 / 1, alignment.py:324)
    float by, bx;
    int src_w, src_h, dst_w, dst_h;
    float fill;
    int bilinear;
};

// source position of output pixel (y, x): transform.cl:40-45 (dot product, then the offset)
__device__ __forceinline__ void warp_source(const WarpMap &m, int y, int x, float &sy, float &sx) {
    sx = m.axy * (float)y + m.axx * (float)x;
    sy = m.ayy * (float)y + m.ayx * (float)x;
    sx += m.bx;
    sy += m.by;
}

// Fetch is a functor (row, col) -> float so that the grey and the per-channel RGB kernels share the sampling rule.
template <typename Fetch>
__device__ __forceinline__ float warp_sample(const WarpMap &m, float sy, float sx, Fetch fetch) {
    float v = m.fill;
    const bool inside = 0.0f <= sx && sx < (float)m.src_w && 0.0f <= sy && sy < (float)m.src_h;
    if (inside) {
        const int x0 = (int)sx, y0 = (int)sy;
        if (m.bilinear) {
            const int x1 = x0 + 1, y1 = y0 + 1;
            const bool x_out = x1 >= m.src_w, y_out = y1 >= m.src_h;
            const float v00 = fetch(y0, x0);
            const float v01 = x_out ? m.fill : fetch(y0, x1);
            const float v10 = y_out ? m.fill : fetch(y1, x0);
            const float v11 = (x_out || y_out) ? m.fill : fetch(y1, x1);
            const float wx0 = (float)x1 - sx, wx1 = sx - (float)x0;
            const float wy0 = (float)y1 - sy, wy1 = sy - (float)y0;
            const float top = wx0 * v00 + wx1 * v01;
            const float bot = wx0 * v10 + wx1 * v11;
            v = wy0 * top + wy1 * bot;
        } else {
            v = fetch(y0, x0);
        }
    }
    // the last half pixel of the image is filled as well (transform.cl:100-104)
    if (sx >= (float)m.src_w + -0.5f || sy >= (float)m.src_h + -0.5f) v = m.fill;
    return v;
}

#define WARP_THREADS 128
// grid (ceil(dst_w / (4 * WARP_THREADS)), dst_h)
static __global__ void __launch_bounds__(WARP_THREADS) k_warp_f32(const float *__restrict__ src, float *__restrict__ dst,
                                                            WarpMap m) {
    const int x4 = 4 * (blockIdx.x * WARP_THREADS + threadIdx.x), y = blockIdx.y;
    if (x4 >= m.dst_w) return;
    auto fetch = [&](int r, int c) { return __ldg(src + (size_t)r * m.src_w + c); };
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float sy, sx;
        warp_source(m, y, x4 + k, sy, sx);
        v[k] = warp_sample(m, sy, sx, fetch);
    }
    float *row = dst + (size_t)y * m.dst_w;
    if (x4 + 3 < m.dst_w && (m.dst_w & 3) == 0) {
        *reinterpret_cast<float4 *>(row + x4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x4 + k < m.dst_w) row[x4 + k] = v[k];
    }
}

// interleaved uint8 RGB: one thread per output pixel, the three channels share the source position
// (transform.cl:116-203 evaluates it once per byte); the result is stored as (uchar)(int)value like the implicit
// float -> uchar store of the reference
static __global__ void __launch_bounds__(WARP_THREADS) k_warp_rgb8(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                             WarpMap m) {
    const int x = blockIdx.x * WARP_THREADS + threadIdx.x, y = blockIdx.y;
    if (x >= m.dst_w) return;
    float sy, sx;
    warp_source(m, y, x, sy, sx);
    uint8_t *o = dst + 3 * ((size_t)y * m.dst_w + x);
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        auto fetch = [&](int r, int c) { return (float)__ldg(src + 3 * ((size_t)r * m.src_w + c) + ch); };
        o[ch] = (uint8_t)(int)warp_sample(m, sy, sx, fetch);
    }
}

static inline WarpMap make_warp_map(const float matrix[4], const float offset[2], int src_h, int src_w, int dst_h,
                                    int dst_w, float fill, int mode) {
    WarpMap m;
    m.ayy = matrix[0]; m.ayx = matrix[1]; m.axy = matrix[2]; m.axx = matrix[3];
    m.by = offset[0]; m.bx = offset[1];
    m.src_w = src_w; m.src_h = src_h; m.dst_w = dst_w; m.dst_h = dst_h;
    m.fill = fill; m.bilinear = mode == 1;
    return m;
}
static inline cudaError_t launch_warp_f32(cudaStream_t st, const float *src, float *dst, const WarpMap &m) {
    dim3 grid((m.dst_w + 4 * WARP_THREADS - 1) / (4 * WARP_THREADS), m.dst_h);
    k_warp_f32<<<grid, WARP_THREADS, 0, st>>>(src, dst, m);
    return cudaGetLastError();
}
static inline cudaError_t launch_warp_rgb8(cudaStream_t st, const uint8_t *src, uint8_t *dst, const WarpMap &m) {
    dim3 grid((m.dst_w + WARP_THREADS - 1) / WARP_THREADS, m.dst_h);
    k_warp_rgb8<<<grid, WARP_THREADS, 0, st>>>(src, dst, m);
    return cudaGetLastError();
}
