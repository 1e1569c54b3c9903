// k_frontend.cuh -- input conversion, global min/max, normalisation.
// Replaces preprocess.cl:*_to_float / rgb_to_float (:53-223), reductions.cl:62-239 and
// preprocess.cl:238 normalizes (the latter is fused into the first blur's loads, see k_blur.cuh).
#pragma once
#include "common.cuh"

// minmax[0] = ordered(min), minmax[1] = ordered(max); reset with k_minmax_reset before use.
__global__ void k_minmax_reset(unsigned *mm) {
    mm[0] = 0xffffffffu;  // +inf side for min
    mm[1] = 0u;           // -inf side for max
}

__device__ __forceinline__ void block_minmax_commit(float mn, float mx, unsigned *mm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ float s_mn[32], s_mx[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_mn[wid] = mn; s_mx[wid] = mx; }
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        mn = lane < nw ? s_mn[lane] : s_mn[0];
        mx = lane < nw ? s_mx[lane] : s_mx[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            atomicMin(&mm[0], float_to_ordered(mn));
            atomicMax(&mm[1], float_to_ordered(mx));
        }
    }
}

// fp32 dense image: float4 loads when n % 4 == 0 and the pointer is 16-B aligned (checked on host)
__global__ void __launch_bounds__(256) k_minmax_f32(const float *__restrict__ img, long n, int vec4, unsigned *mm) {
    float mn = INFINITY, mx = -INFINITY;
    long stride = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec4) {
        const float4 *p = reinterpret_cast<const float4 *>(img);
        long n4 = n >> 2;
        for (; i < n4; i += stride) {
            float4 v = __ldg(p + i);
            mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
            mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
    } else {
        for (; i < n; i += stride) {
            float v = __ldg(img + i);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
    block_minmax_commit(mn, mx, mm);
}

template <typename T>
__device__ __forceinline__ float px_to_float(const T *p, long i) { return (float)p[i]; }

// integer / f64 / RGB -> dense fp32 plane, min/max fused (one pass over the raw image)
template <typename T, bool RGB>
__global__ void __launch_bounds__(256) k_convert_minmax(const T *__restrict__ raw, float *__restrict__ out, long n,
                                                         unsigned *mm) {
    float mn = INFINITY, mx = -INFINITY;
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v;
        if (RGB) {  // preprocess.cl:221
            float r = 0.299f * (float)raw[3 * i], g = 0.587f * (float)raw[3 * i + 1], b = 0.114f * (float)raw[3 * i + 2];
            v = (r + g) + b;
        } else {
            v = (float)raw[i];
        }
        out[i] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    block_minmax_commit(mn, mx, mm);
}

// stand-alone normalizes (stage hook only; the pipeline fuses it into the first blur)
__global__ void __launch_bounds__(256) k_normalize(const float *__restrict__ in, float *__restrict__ out, long n,
                                                    const unsigned *__restrict__ mm) {
    float mn = ordered_to_float(mm[0]), mx = ordered_to_float(mm[1]);
    float den = mx - mn;
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (255.0f * (in[i] - mn)) / den;
}
