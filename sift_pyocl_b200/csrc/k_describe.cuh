// k_describe.cuh -- 4x4x8 SIFT descriptor, one warp per keypoint, bit-identical to the sequential
// reference kernel keypoints_cpu.cl:36-160 (CPU-variant semantics).
//
// The reference accumulates `hist[bin] += w` while scanning the (2R+1)^2 window in row-major order;
// fp32 addition is not associative, so every bin must see its contributions in exactly that order.
// Parallel scheme:
//   phase 1  the 32 lanes evaluate 32 consecutive window samples (row-major), the valid ones are
//            compacted IN ORDER into a shared-memory record buffer; each record also registers
//            itself in the ordered list of every cell (r, c) it touches (<= 4 of the 16 cells);
//   phase 2  lane X < 16 owns cell X = r*4+c and walks its own list, adding the two orientation
//            contributions of each record into its 8 bins -- lists are in sample order, bins of
//            different cells are independent, so all 16 lanes run concurrently;
//   phase 3  L2 normalisation / 0.2 clamp / renormalisation / x512 -> uint8; the two
//            sum-of-squares are accumulated sequentially by one lane (order matters), the rest is
//            lane-parallel.
// Also performs the host-side NaN filtering and record assembly of plan.py:546-565.
#pragma once
#include "common.cuh"
#include "k_keypoint.cuh"

#define DESC_WARPS 4          // warps per CTA
#define DESC_BATCH 128        // records buffered between two phase-2 sweeps

struct DescSmem {
    float4 rec_f[DESC_BATCH];            // rw0 = mag*(1-rfrac), rw1 = mag*rfrac, cfrac, ofrac
    uint32_t rec_i[DESC_BATCH];          // (ri+1) | (ci+1)<<8 | o0<<16 | o1<<24
    uint8_t lists[16][DESC_BATCH];       // per cell: record indices in sample order
    uint32_t cellmask[16];
    uint32_t count[16];
    float hist[128];                     // [o*16 + cell]
};

__device__ __forceinline__ void describe_flush(DescSmem &S, int lane) {
    // phase 2: lane X < 16 consumes its ordered list
    if (lane < 16) {
        const int r = lane >> 2, c = lane & 3;
        const int n = (int)S.count[lane];
        for (int e = 0; e < n; e++) {
            const int idx = S.lists[lane][e];
            const float4 f = S.rec_f[idx];
            const uint32_t u = S.rec_i[idx];
            const int ri = (int)(u & 0xff) - 1, ci = (int)((u >> 8) & 0xff) - 1;
            const int o0 = (u >> 16) & 0xff, o1 = (u >> 24) & 0xff;
            const float rweight = (r == ri) ? f.x : f.y;                    // (r == 0) ? 1 - rfrac : rfrac
            const float cweight = rweight * ((c == ci) ? 1.0f - f.z : f.z);  // (c == 0) ? 1 - cfrac : cfrac
            float *h0 = &S.hist[o0 * 16 + lane];
            *h0 += cweight * (1.0f - f.w);
            float *h1 = &S.hist[o1 * 16 + lane];
            *h1 += cweight * f.w;
        }
        S.count[lane] = 0;
    }
    __syncwarp();
}

// One warp computes the descriptor of keypoint k into out128 (128 bytes, 4-byte aligned).
__device__ __forceinline__ void describe_warp(DescSmem &S, const float4 k, const float *__restrict__ grad,
                                              const float *__restrict__ orim, int pitch, int grad_width,
                                              int grad_height, int octsize, uint8_t *out128) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 128; i += 32) S.hist[i] = 0.0f;
    if (lane < 16) { S.cellmask[lane] = 0; S.count[lane] = 0; }
    // keypoints_cpu.cl:55-61
    const float row = k.y / (float)octsize, col = k.x / (float)octsize, angle = k.w;
    const int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    const float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    const float spacing = k.z / (float)octsize * 3.0f;
    const int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    const float drow = row - (float)irow, dcol = col - (float)icol;
    const int side = 2 * iradius + 1;
    const int total = (iradius >= 0 && iradius < 16384) ? side * side : 0;
    int nrec = 0;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        if (nrec + 32 > DESC_BATCH) { describe_flush(S, lane); nrec = 0; }
        const int t = base + lane;
        bool valid = false;
        float rw0 = 0.f, rw1 = 0.f, cfrac = 0.f, ofrac = 0.f;
        int ri = 0, ci = 0, o0 = 0, o1 = 0;
        if (t < total) {
            const int ti = t / side;
            const int i = ti - iradius, j = (t - ti * side) - iradius;
            const float rx = ((cosine * (float)i - sine * (float)j) - drow) / spacing + 1.5f;
            const float cx = ((sine * (float)i + cosine * (float)j) - dcol) / spacing + 1.5f;
            if ((rx > -1.0f && rx < 4.0f && cx > -1.0f && cx < 4.0f && (irow + i) >= 0 && (irow + i) < grad_height &&
                 (icol + j) >= 0 && (icol + j) < grad_width)) {
                const long q = (long)(irow + i) * pitch + (icol + j);
                const float er = rx - 1.5f, ec = cx - 1.5f;
                const float mag = grad[q] * cr_expf(-0.125f * (er * er + ec * ec));
                float ori = orim[q] - angle;
                while (ori > 2.0f * SIFTB_M_PI_F) ori -= 2.0f * SIFTB_M_PI_F;
                while (ori < 0.0f) ori += 2.0f * SIFTB_M_PI_F;
                const float oval = (4.0f * ori) * SIFTB_M_1_PI_F;
                ri = (int)((rx >= 0.0f) ? rx : rx - 1.0f);
                ci = (int)((cx >= 0.0f) ? cx : cx - 1.0f);
                const int oi = (int)((oval >= 0.0f) ? oval : oval - 1.0f);
                const float rfrac = rx - (float)ri;
                cfrac = cx - (float)ci;
                ofrac = oval - (float)oi;
                if ((ri >= -1 && ri < 4 && oi >= 0 && oi <= 8 && rfrac >= 0.0f && rfrac <= 1.0f)) {
                    valid = true;
                    rw0 = mag * (1.0f - rfrac);
                    rw1 = mag * rfrac;
                    o0 = (oi >= 8) ? 0 : oi;          // oindex = oi + orr; if (oindex >= 8) oindex = 0
                    o1 = (oi + 1 >= 8) ? 0 : oi + 1;
                }
            }
        }
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (vm == 0) continue;
        int idx = 0;
        if (valid) {
            idx = nrec + __popc(vm & lanemask_lt());
            S.rec_f[idx] = make_float4(rw0, rw1, cfrac, ofrac);
            S.rec_i[idx] = (uint32_t)(ri + 1) | ((uint32_t)(ci + 1) << 8) | ((uint32_t)o0 << 16) | ((uint32_t)o1 << 24);
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const int rr = ri + (d >> 1), cc = ci + (d & 1);
                if (rr >= 0 && rr < 4 && cc >= 0 && cc < 4) atomicOr(&S.cellmask[rr * 4 + cc], 1u << lane);
            }
        }
        __syncwarp();
        if (valid) {
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const int rr = ri + (d >> 1), cc = ci + (d & 1);
                if (rr >= 0 && rr < 4 && cc >= 0 && cc < 4) {
                    const int X = rr * 4 + cc;
                    S.lists[X][S.count[X] + __popc(S.cellmask[X] & lanemask_lt())] = (uint8_t)idx;
                }
            }
        }
        __syncwarp();
        if (lane < 16) {
            S.count[lane] += __popc(S.cellmask[lane]);
            S.cellmask[lane] = 0;
        }
        __syncwarp();
        nrec += __popc(vm);
    }
    describe_flush(S, lane);
    // phase 3, keypoints_cpu.cl:127-160.  descriptor index i = (r*4+c)*8 + o  <->  hist[o*16 + r*4+c]
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane * 4 + q;
        v[q] = S.hist[(i & 7) * 16 + (i >> 3)];
    }
    __syncwarp();
    float *seq = S.hist;  // reuse as the i-ordered scratch of squares
#pragma unroll
    for (int q = 0; q < 4; q++) seq[lane * 4 + q] = v[q] * v[q];
    __syncwarp();
    float norm = 0.0f;
    if (lane == 0)
        for (int i = 0; i < 128; i++) norm += seq[i];
    norm = cr_rsqrtf(__shfl_sync(0xffffffffu, norm, 0));
    bool changed = false;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        v[q] *= norm;
        if (v[q] > 0.2f) { v[q] = 0.2f; changed = true; }
        seq[lane * 4 + q] = v[q] * v[q];
    }
    changed = __any_sync(0xffffffffu, changed);
    __syncwarp();
    if (changed) {
        norm = 0.0f;
        if (lane == 0)
            for (int i = 0; i < 128; i++) norm += seq[i];
        norm = cr_rsqrtf(__shfl_sync(0xffffffffu, norm, 0));
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] *= norm;
    }
    uint32_t packed = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float x = 512.0f * v[q];
        const int intval = (x != x) ? 0 : (int)x;
        packed |= (uint32_t)min(255, intval) << (8 * q);  // intval >= 0 here (hist >= 0)
    }
    reinterpret_cast<uint32_t *>(out128)[lane] = packed;
    __syncwarp();
}

// Pipeline form: warps grid-stride over the keypoints of the octave; rows with NaN are dropped
// (plan.py:546-550) and survivors appended to the final record array (plan.py:555-565).
__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(GradPlanes G, const float4 *__restrict__ kp,
                                                               const int *__restrict__ kp_scale,
                                                               const int *__restrict__ n_base_p,
                                                               const int *__restrict__ n_extra_p, int cap, int octsize,
                                                               KpRecord *__restrict__ out, int out_cap,
                                                               int *__restrict__ n_out, int *__restrict__ n_out_oct) {
    __shared__ DescSmem smem[DESC_WARPS];
    DescSmem &S = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int n = min(min(*n_base_p, cap) + *n_extra_p, cap);
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int gid0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gid0 < n; gid0 += nwarps) {
        const float4 k = kp[gid0];
        if (!(k.y >= 0.0f)) continue;
        const float s = ((k.x + k.y) + k.z) + k.w;
        if (s != s) continue;
        int slot = 0;
        if (lane == 0) {
            slot = atomicAdd(n_out, 1);
            atomicAdd(n_out_oct, 1);
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= out_cap) continue;
        const int sc = kp_scale[gid0];
        KpRecord *o = out + slot;
        if (lane == 0) { o->x = k.x; o->y = k.y; o->scale = k.z; o->angle = k.w; }
        describe_warp(S, k, G.grad[sc - 1], G.ori[sc - 1], G.pitch, G.w, G.h, octsize, o->desc);
    }
}

// Stage-hook form: desc[i] for every input row (no filtering), single gradient plane
__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe_rows(const float *__restrict__ grad,
                                                                    const float *__restrict__ ori, int pitch, int w,
                                                                    int h, const float4 *__restrict__ kp, int n,
                                                                    int octsize, uint8_t *__restrict__ desc) {
    __shared__ DescSmem smem[DESC_WARPS];
    DescSmem &S = smem[threadIdx.x >> 5];
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int gid0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gid0 < n; gid0 += nwarps) {
        const float4 k = kp[gid0];
        if (!(k.y >= 0.0f)) continue;
        describe_warp(S, k, grad, ori, pitch, w, h, octsize, desc + 128L * gid0);
    }
}
