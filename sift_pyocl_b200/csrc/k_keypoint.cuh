// k_keypoint.cuh -- gradient planes, orientation assignment, 4x4x8 descriptors.
//
// k_gradient  replaces image.cl:47 compute_gradient_orientation (3 launches/octave -> 1).
// k_orient    replaces orientation_cpu.cl:41 orientation_assignment (CPU-variant semantics, SURVEY
//             App. A.7): one warp per keypoint; every lane owns histogram bins and the samples are
//             committed in the reference's row-major order, so the fp32 sums are bit-identical to
//             the sequential kernel.
// k_describe  replaces keypoints_cpu.cl:36 descriptor (CPU-variant semantics, App. A.8) and the
//             host-side NaN filtering / record assembly of plan.py:546-565.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// image.cl:47-81.  grid (ceil(w/256), h, nplanes), block 256
struct GradArgs {
    const float *g[3];
    float *grad[3];
    float *ori[3];
    int pitch, w, h;
};

__global__ void __launch_bounds__(256) k_gradient(GradArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= a.w) return;
    const float *g = a.g[z];
    const long pos = (long)y * a.pitch + x;
    float xgrad, ygrad;
    if (x == 0) xgrad = 2.0f * (g[pos + 1] - g[pos]);
    else if (x == a.w - 1) xgrad = 2.0f * (g[pos] - g[pos - 1]);
    else xgrad = g[pos + 1] - g[pos - 1];
    if (y == 0) ygrad = 2.0f * (g[pos] - g[pos + a.pitch]);
    else if (y == a.h - 1) ygrad = 2.0f * (g[pos - a.pitch] - g[pos]);
    else ygrad = g[pos - a.pitch] - g[pos + a.pitch];
    a.grad[z][pos] = sqrtf(xgrad * xgrad + ygrad * ygrad);
    a.ori[z][pos] = cr_atan2f(-ygrad, xgrad);
}

// ---------------------------------------------------------------------------------------------
struct GradPlanes {
    const float *grad[3];
    const float *ori[3];
    int pitch, w, h;
};

// One warp per keypoint (grid-stride).  kp rows in: (peak, row, col, sigma); out: (x, y, sigma*oct, angle).
// Extra-orientation keypoints are appended at n_base + atomicAdd(n_extra).
__global__ void __launch_bounds__(256) k_orient(GradPlanes G, float4 *__restrict__ kp, int *__restrict__ kp_scale,
                                                 const int *__restrict__ n_base_p, int *__restrict__ n_extra, int cap,
                                                 int octsize, float OriSigma, int *__restrict__ stage) {
    __shared__ float s_hist[8][36];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n_base = min(*n_base_p, cap);
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float *hist = s_hist[wib];
    for (int gid0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gid0 < n_base; gid0 += nwarps) {
        float4 k = kp[gid0];
        const int sc = kp_scale[gid0];
        if (!(k.y >= 0.0f)) continue;  // warp-uniform
        const float *grad = G.grad[sc - 1], *ori = G.ori[sc - 1];
        const int row = (int)((double)k.y + 0.5), col = (int)((double)k.z + 0.5);  // orientation_cpu.cl:67-68
        const float sigma = OriSigma * k.w;
        const int radius = (int)((double)sigma * 3.0);  // :71
        const int rmin = max(0, row - radius), cmin = max(0, col - radius);
        const int rmax = min(row + radius, G.h - 2), cmax = min(col + radius, G.w - 2);
        const float two_s2 = (2.0f * sigma) * sigma;
        const float rad2 = ((float)(radius * radius)) + 0.5f;
        const int ncols = cmax - cmin + 1, nrows = rmax - rmin + 1;
        const int total = (ncols > 0 && nrows > 0) ? ncols * nrows : 0;
        float h0 = 0.0f, h1 = 0.0f;  // bins lane and lane+32
        for (int base = 0; base < total; base += 32) {
            const int idx = base + lane;
            int bin = -1;
            float w = 0.0f;
            if (idx < total) {
                const int rr = idx / ncols;
                const int r = rmin + rr, c = cmin + (idx - rr * ncols);
                const float gval = grad[(long)r * G.pitch + c];
                float dif = ((float)r - k.y);
                float distsq = dif * dif;
                dif = ((float)c - k.z);
                distsq += dif * dif;
                if (gval > 0.0f && distsq < rad2) {
                    const float angle = ori[(long)r * G.pitch + c];
                    int b = (int)(36.0f * ((angle + SIFTB_M_PI_F) + 0.001f) / (2.0f * SIFTB_M_PI_F));
                    if (b >= 0 && b <= 36) {
                        bin = min(b, 35);
                        w = cr_expf(-distsq / two_s2) * gval;
                    }
                }
            }
            // commit in lane order == the reference's row-major sample order
            unsigned m = __ballot_sync(0xffffffffu, bin >= 0);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int b = __shfl_sync(0xffffffffu, bin, src);
                const float wv = __shfl_sync(0xffffffffu, w, src);
                if (b == lane) h0 += wv;
                if (b == lane + 32) h1 += wv;
            }
        }
        hist[lane] = h0;
        if (lane < 4) hist[32 + lane] = h1;
        __syncwarp();
        // orientation_cpu.cl:100-108 -- six in-place smoothing passes.  In place means: bins 0..34 see
        // the OLD neighbours (prev is carried), bin 35 sees the NEW bin 0.  "/ 3.0" is a double division.
        for (int j = 0; j < 6; j++) {
            const float a0 = hist[(lane + 35) % 36], b0 = hist[lane], c0 = hist[lane + 1];
            float a1 = 0.f, b1 = 0.f, c1 = 0.f;
            if (lane < 3) { a1 = hist[31 + lane], b1 = hist[32 + lane], c1 = hist[33 + lane]; }
            float o34 = hist[34], o35 = hist[35];
            __syncwarp();
            const float n0 = (float)((double)((a0 + b0) + c0) / 3.0);
            hist[lane] = n0;
            if (lane < 3) hist[32 + lane] = (float)((double)((a1 + b1) + c1) / 3.0);
            __syncwarp();
            if (lane == 0) hist[35] = (float)((double)((o34 + o35) + n0) / 3.0);
            __syncwarp();
        }
        if (lane == 0) {
            float maxval = 0.0f;
            int argmax = 0;
            for (int i = 0; i < 36; i++)
                if (maxval < hist[i]) { maxval = hist[i]; argmax = i; }
            const int prev = (argmax == 0 ? 35 : argmax - 1), next = (argmax == 35 ? 0 : argmax + 1);
            float hist_prev = hist[prev], hist_next = hist[next];
            if (maxval < 0.0f) { hist_prev = -hist_prev; maxval = -maxval; hist_next = -hist_next; }
            const float interp = 0.5f * (hist_prev - hist_next) / ((hist_prev - 2.0f * maxval) + hist_next);
            const float angle = (2.0f * SIFTB_M_PI_F) * (((float)argmax + 0.5f) + interp) / 36.0f - SIFTB_M_PI_F;
            float4 o;
            o.x = k.z * (float)octsize;
            o.y = k.y * (float)octsize;
            o.z = k.w * (float)octsize;
            o.w = angle;
            kp[gid0] = o;
            int added = 1;
            for (int i = 0; i < 36; i++) {
                const int pv = (i == 0 ? 35 : i - 1), nx = (i == 35 ? 0 : i + 1);
                float hp = hist[pv], hc = hist[i], hn = hist[nx];
                if (hc > hp && hc > hn && hc >= 0.8f * maxval && i != argmax) {
                    if (hc < 0.0f) { hp = -hp; hc = -hc; hn = -hn; }
                    const float itp = 0.5f * (hp - hn) / ((hp - 2.0f * hc) + hn);
                    // orientation_cpu.cl:166: "/36.0" promotes the tail of the expression to double
                    const float a2 = (float)((double)((2.0f * SIFTB_M_PI_F) * (((float)i + 0.5f) + itp)) / 36.0 -
                                             (double)SIFTB_M_PI_F);
                    if (a2 >= -SIFTB_M_PI_F && a2 <= SIFTB_M_PI_F) {
                        const int old = n_base + atomicAdd(n_extra, 1);
                        if (old < cap) {
                            kp[old] = make_float4(o.x, o.y, o.z, a2);
                            kp_scale[old] = sc;
                        }
                        added++;
                    }
                }
            }
            if (stage) atomicAdd(&stage[(sc - 1) * 3 + 2], added);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// keypoints_cpu.cl:49-160 for one keypoint (x, y, sigma*oct, angle) -> 128 bytes.
// v1: one thread per keypoint, literal restatement (histogram in local memory).
__device__ void describe_one(const float4 k, const float *__restrict__ grad, const float *__restrict__ orim,
                             int pitch, int grad_width, int grad_height, int octsize, uint8_t *out) {
    float tmp_descriptors[128];
    for (int i = 0; i < 128; i++) tmp_descriptors[i] = 0.0f;
    const float row = k.y / (float)octsize, col = k.x / (float)octsize, angle = k.w;
    const int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    const float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    const float spacing = k.z / (float)octsize * 3.0f;
    const int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    const float drow = row - (float)irow, dcol = col - (float)icol;
    for (int i = -iradius; i <= iradius; i++) {
        for (int j = -iradius; j <= iradius; j++) {
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
                const int ri = (int)((rx >= 0.0f) ? rx : rx - 1.0f), ci = (int)((cx >= 0.0f) ? cx : cx - 1.0f),
                          oi = (int)((oval >= 0.0f) ? oval : oval - 1.0f);
                const float rfrac = rx - (float)ri, cfrac = cx - (float)ci, ofrac = oval - (float)oi;
                if ((ri >= -1 && ri < 4 && oi >= 0 && oi <= 8 && rfrac >= 0.0f && rfrac <= 1.0f)) {
                    for (int r = 0; r < 2; r++) {
                        const int rindex = ri + r;
                        if ((rindex >= 0 && rindex < 4)) {
                            const float rweight = mag * ((r == 0) ? 1.0f - rfrac : rfrac);
                            for (int c = 0; c < 2; c++) {
                                const int cindex = ci + c;
                                if ((cindex >= 0 && cindex < 4)) {
                                    const float cweight = rweight * ((c == 0) ? 1.0f - cfrac : cfrac);
                                    for (int orr = 0; orr < 2; orr++) {
                                        int oindex = oi + orr;
                                        if (oindex >= 8) oindex = 0;
                                        tmp_descriptors[(rindex * 4 + cindex) * 8 + oindex] +=
                                            cweight * ((orr == 0) ? 1.0f - ofrac : ofrac);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    float norm = 0.0f;
    for (int i = 0; i < 128; i++) norm += tmp_descriptors[i] * tmp_descriptors[i];
    norm = cr_rsqrtf(norm);
    for (int i = 0; i < 128; i++) tmp_descriptors[i] *= norm;
    bool changed = false;
    norm = 0.0f;
    for (int i = 0; i < 128; i++) {
        if (tmp_descriptors[i] > 0.2f) { tmp_descriptors[i] = 0.2f; changed = true; }
        norm += tmp_descriptors[i] * tmp_descriptors[i];
    }
    if (changed) {
        norm = cr_rsqrtf(norm);
        for (int i = 0; i < 128; i++) tmp_descriptors[i] *= norm;
    }
    for (int i = 0; i < 128; i++) {
        const float v = 512.0f * tmp_descriptors[i];  // 512.0 * v in double is exact, == fp32 product
        const int intval = (v != v) ? 0 : (int)v;
        out[i] = (uint8_t)min(255, intval);
    }
}

struct KpRecord {  // == siftb_kp
    float x, y, scale, angle;
    uint8_t desc[128];
};

// Pipeline form: thread per keypoint over [0, n_base + n_extra); rows with a NaN coordinate are dropped
// (plan.py:546-550) and the survivors are appended to the final record array.
__global__ void __launch_bounds__(64) k_describe(GradPlanes G, const float4 *__restrict__ kp,
                                                  const int *__restrict__ kp_scale, const int *__restrict__ n_base_p,
                                                  const int *__restrict__ n_extra_p, int cap, int octsize,
                                                  KpRecord *__restrict__ out, int out_cap, int *__restrict__ n_out,
                                                  int *__restrict__ n_out_oct) {
    const int n = min(min(*n_base_p, cap) + *n_extra_p, cap);
    const int stride = gridDim.x * blockDim.x;
    for (int gid0 = blockIdx.x * blockDim.x + threadIdx.x; gid0 < n; gid0 += stride) {
        const float4 k = kp[gid0];
        if (!(k.y >= 0.0f)) continue;
        const float s = ((k.x + k.y) + k.z) + k.w;
        if (s != s) continue;
        const int slot = atomicAdd(n_out, 1);
        atomicAdd(n_out_oct, 1);
        if (slot >= out_cap) continue;
        const int sc = kp_scale[gid0];
        KpRecord *o = out + slot;
        o->x = k.x; o->y = k.y; o->scale = k.z; o->angle = k.w;
        describe_one(k, G.grad[sc - 1], G.ori[sc - 1], G.pitch, G.w, G.h, octsize, o->desc);
    }
}

// Stage-hook form: desc[i] for every input row (no filtering), single gradient plane
__global__ void __launch_bounds__(64) k_describe_rows(const float *__restrict__ grad, const float *__restrict__ ori,
                                                       int pitch, int w, int h, const float4 *__restrict__ kp, int n,
                                                       int octsize, uint8_t *__restrict__ desc) {
    const int gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid0 >= n) return;
    const float4 k = kp[gid0];
    if (!(k.y >= 0.0f)) return;
    describe_one(k, grad, ori, pitch, w, h, octsize, desc + 128L * gid0);
}
