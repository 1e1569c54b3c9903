// k_match.cuh -- brute-force L1 matcher with ratio test, and the affine warps (grey and RGB).
//
// k_match_l1 replaces matching_gpu.cl:52 / matching_cpu.cl:57 `matching` plus memset.cl
// memset_kp/memset_int (match.py:244-255).  k_transform replaces transform.cl:22 `transform`
// (alignment.py:336-346).
#pragma once
#include "common.cuh"

// 144-byte AoS records -> dense 128-byte descriptor rows (16-B aligned for vector loads)
__global__ void __launch_bounds__(256) k_extract_desc(const uint8_t *__restrict__ recs, int n,
                                                       uint32_t *__restrict__ desc) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one 32-bit word each
    if (i >= (long)n * 32) return;
    long row = i >> 5;
    int wd = (int)(i & 31);
    desc[i] = *reinterpret_cast<const uint32_t *>(recs + row * 144 + 16 + 4 * wd);
}

#define MATCH_TILE 64
// one thread per query row of list 1; list 2 streamed through shared memory in tiles
__global__ void __launch_bounds__(128) k_match_l1(const uint32_t *__restrict__ d1, int n1,
                                                   const uint32_t *__restrict__ d2, int n2, float ratio_th,
                                                   int2 *__restrict__ pairs, int cap, int *__restrict__ counter) {
    __shared__ uint4 tile[MATCH_TILE * 8];
    const int gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid0 < n1;
    uint32_t q[32];
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(d1) + (long)(active ? gid0 : 0) * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = __ldg(p + i);
            q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
        }
    }
    float dist1 = 1000000000000.0f, dist2 = 1000000000000.0f;  // matching_cpu.cl:71
    int current_min = 0;
    for (int base = 0; base < n2; base += MATCH_TILE) {
        const int rows = min(MATCH_TILE, n2 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 8; i += blockDim.x)
            tile[i] = __ldg(reinterpret_cast<const uint4 *>(d2) + (long)base * 8 + i);
        __syncthreads();
        for (int r = 0; r < rows; r++) {
            unsigned dist = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 v = tile[r * 8 + i];
                dist = __vsadu4(q[4 * i], v.x) + dist;
                dist = __vsadu4(q[4 * i + 1], v.y) + dist;
                dist = __vsadu4(q[4 * i + 2], v.z) + dist;
                dist = __vsadu4(q[4 * i + 3], v.w) + dist;
            }
            const float fd = (float)(int)dist;
            if (fd < dist1) { dist2 = dist1; dist1 = fd; current_min = base + r; }
            else if (fd < dist2) { dist2 = fd; }
        }
    }
    const bool emit = active && (dist2 != 0.0f) && (dist1 / dist2 < ratio_th);  // matching_cpu.cl:100
    const int slot = warp_append(emit, counter);
    if (emit && slot < cap) pairs[slot] = make_int2(gid0, current_min);
}

// transform.cl:34-106; one thread per output pixel
__global__ void __launch_bounds__(256) k_transform(const float *__restrict__ image, float *__restrict__ output,
                                                    float m0, float m1, float m2, float m3, float off0, float off1,
                                                    int image_width, int image_height, int output_width,
                                                    int output_height, float fill, int mode) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= output_width || y >= output_height) return;
    float tx = m2 * (float)y + m3 * (float)x;  // dot(mat.s23, (y, x))
    float ty = m0 * (float)y + m1 * (float)x;
    tx += off1;
    ty += off0;
    const int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
    float interp = fill;
    if (0.0f <= tx && tx < (float)image_width && 0.0f <= ty && ty < (float)image_height) {
        if (mode == 1) {
            const float image_p = image[(long)ty_prev * image_width + tx_prev];
            const bool xo = tx_next >= image_width, yo = ty_next >= image_height;
            const float image_x = xo ? fill : image[(long)ty_prev * image_width + tx_next];
            const float image_y = yo ? fill : image[(long)ty_next * image_width + tx_prev];
            const float image_n = (xo || yo) ? fill : image[(long)ty_next * image_width + tx_next];
            const float wxn = (float)tx_next - tx, wxp = tx - (float)tx_prev;
            const float wyn = (float)ty_next - ty, wyp = ty - (float)ty_prev;
            const float interp1 = wxn * image_p + wxp * image_x;
            const float interp2 = wxn * image_y + wxp * image_n;
            interp = wyn * interp1 + wyp * interp2;
        } else {
            interp = image[(long)((int)ty) * image_width + ((int)tx)];
        }
    }
    if (tx >= (float)image_width + -0.5f) interp = fill;
    if (ty >= (float)image_height + -0.5f) interp = fill;
    output[(long)y * output_width + x] = interp;
}

// transform.cl:116-203 transform_RGB: same mapping per colour channel of an interleaved uint8 image.
// grid (ceil(3*out_w/256), out_h): one thread per output byte.
__global__ void __launch_bounds__(256) k_transform_rgb(const uint8_t *__restrict__ image, uint8_t *__restrict__ output,
                                                        float m0, float m1, float m2, float m3, float off0, float off1,
                                                        int image_width, int image_height, int output_width,
                                                        int output_height, float fill, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (i >= 3 * output_width || y >= output_height) return;
    const int x = i / 3, color = i - 3 * x;
    float tx = m2 * (float)y + m3 * (float)x;
    float ty = m0 * (float)y + m1 * (float)x;
    tx += off1;
    ty += off0;
    const int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
    float interp = fill;
    if (0.0f <= tx && tx < (float)image_width && 0.0f <= ty && ty < (float)image_height) {
        if (mode == 1) {
            const bool xo = tx_next >= image_width, yo = ty_next >= image_height;
            const float image_p = (float)image[3 * ((long)ty_prev * image_width + tx_prev) + color];
            const float image_x = xo ? fill : (float)image[3 * ((long)ty_prev * image_width + tx_next) + color];
            const float image_y = yo ? fill : (float)image[3 * ((long)ty_next * image_width + tx_prev) + color];
            const float image_n = (xo || yo) ? fill : (float)image[3 * ((long)ty_next * image_width + tx_next) + color];
            const float wxn = (float)tx_next - tx, wxp = tx - (float)tx_prev;
            const float wyn = (float)ty_next - ty, wyp = ty - (float)ty_prev;
            const float interp1 = wxn * image_p + wxp * image_x;
            const float interp2 = wxn * image_y + wxp * image_n;
            interp = wyn * interp1 + wyp * interp2;
        } else {
            interp = (float)image[3 * ((long)((int)ty) * image_width + ((int)tx)) + color];
        }
    }
    if (tx >= (float)image_width + -0.5f) interp = fill;
    if (ty >= (float)image_height + -0.5f) interp = fill;
    output[3 * ((long)y * output_width + x) + color] = (uint8_t)(int)interp;  // implicit float -> uchar store
}
