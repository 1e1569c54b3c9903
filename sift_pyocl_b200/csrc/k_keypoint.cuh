// k_keypoint.cuh -- gradient planes, orientation assignment, 4x4x8 descriptors.
//
// k_gradient  replaces image.cl:47 compute_gradient_orientation (3 launches/octave -> 1).
// k_orient    replaces orientation_cpu.cl:41 orientation_assignment (CPU-variant semantics, SURVEY
//             App. A.7): one warp per keypoint; every lane owns histogram bins and the samples are
//             committed in the reference's row-major order, so the fp32 sums are bit-identical to
//             the sequential kernel.
// (the descriptor kernel lives in k_describe.cuh)
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// image.cl:47-81
// The gradient magnitude and orientation of a pixel are stored INTERLEAVED, one float2 (grad, ori) per pixel:
// orientation assignment and descriptors always read both values of a sample, so one 8-byte gather replaces two
// 4-byte gathers from different planes (half the memory instructions and cache lines of the two hottest kernels).
struct GradArgs {
    const float *g[3];
    float2 *go[3];    // (gradient magnitude, orientation) per pixel, row pitch `pitch` pixels
    int pitch, w, h;
};

// stage hooks only: separate host planes <-> the interleaved device layout
__global__ void __launch_bounds__(256) k_interleave(const float *__restrict__ grad, const float *__restrict__ ori, long n,
                                                     float2 *__restrict__ go) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        go[i] = make_float2(grad[i], ori[i]);
}
__global__ void __launch_bounds__(256) k_deinterleave(const float2 *__restrict__ go, long n, float *__restrict__ grad,
                                                       float *__restrict__ ori) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float2 v = go[i];
        grad[i] = v.x;
        ori[i] = v.y;
    }
}

#define GRAD_ROWS 8
// grid (ceil(w/256), ceil(h/GRAD_ROWS), nplanes), block 256: a thread walks GRAD_ROWS rows of its column, keeping
// the vertical neighbours in registers (one coalesced centre load per pixel; left/right come from L1).
__global__ void __launch_bounds__(256) k_gradient(GradArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.z;
    const int y0 = blockIdx.y * GRAD_ROWS;
    if (x >= a.w) return;
    const float *g = a.g[z];
    float2 *gop = a.go[z];
    const int xm = x == 0 ? 0 : x - 1, xp = x == a.w - 1 ? x : x + 1;
    const float xs = (x == 0 || x == a.w - 1) ? 2.0f : 1.0f;  // one-sided differences are doubled (image.cl:61-66)
    // sliding window of the centre column: up, cur, down
    const AtanConsts K;
    float up = g[(long)(y0 > 0 ? y0 - 1 : 0) * a.pitch + x];
    float cur = g[(long)y0 * a.pitch + x];
#pragma unroll
    for (int r = 0; r < GRAD_ROWS; r++) {
        const int y = y0 + r;
        if (y >= a.h) break;
        const long pos = (long)y * a.pitch + x;
        const float dn = g[(long)(y + 1 < a.h ? y + 1 : y) * a.pitch + x];
        // image.cl:58-66: xgrad = I[x+1]-I[x-1], doubled one-sided at the borders (then I[xm] or I[xp] is the centre)
        const float xgrad = xs * (g[(long)y * a.pitch + xp] - g[(long)y * a.pitch + xm]);
        // image.cl:67-72: ygrad = I[y-1]-I[y+1] ("up minus down"), doubled one-sided at the borders
        float ygrad;
        if (y == 0) ygrad = 2.0f * (cur - dn);
        else if (y == a.h - 1) ygrad = 2.0f * (up - cur);
        else ygrad = up - dn;
        gop[pos] = make_float2(sqrtf(xgrad * xgrad + ygrad * ygrad), cr_atan2f_fast(-ygrad, xgrad, K));
        up = cur;
        cur = dn;
    }
}

// Vector form for planes whose pitch is a multiple of 4 floats and whose base is 16-byte aligned (every plane of
// a SiftPlan): a thread owns 4 adjacent columns and walks GRAD4_ROWS rows.  One 128-bit load per row gives the
// centre values; the horizontal neighbours come from the adjacent lanes (shuffles; the two edge lanes of a warp
// read one extra value), the vertical neighbours are the previous / next row kept in registers, and the row after
// next is requested before the current one is evaluated.  128-bit stores of both result planes.
#define GRAD4_ROWS 16
__device__ __forceinline__ void grad_one(float xgrad, float ygrad, float &gr, float &orv, const AtanConsts &K) {
    gr = sqrtf(xgrad * xgrad + ygrad * ygrad);
    orv = cr_atan2f_fast(-ygrad, xgrad, K);
}
// one plane: g -> gop, block (bx, by) of the grid (ceil(w/512), ceil(h/GRAD4_ROWS))
struct GradPlane {
    const float *g;
    float2 *go;
    int pitch, w, h;
};
__device__ __forceinline__ void gradient4_block(const GradPlane a, int bx, int by) {
    const int x4 = (bx * blockDim.x + threadIdx.x) * 4, lane = threadIdx.x & 31;
    const int y0 = by * GRAD4_ROWS;
    const float *g = a.g;
    float2 *gop = a.go;
    const bool active = x4 < a.w;
    const int xc = active ? x4 : 0;  // idle threads of a partly filled warp still take part in the shuffles
    auto ldrow = [&](int y) {
        y = max(0, min(y, a.h - 1));
        return *reinterpret_cast<const float4 *>(g + (long)y * a.pitch + xc);
    };
    // the two edge lanes of a warp need one value from outside the warp's 128 columns: lane 0 the pixel left of its
    // group, lane 31 the pixel right of it.  Like the rows, it is requested one iteration ahead (a load issued and
    // consumed in the same iteration stalls the whole warp for an L2 round trip: 25 % of this kernel's samples).
    const bool has_left = active && x4 > 0, has_right = x4 + 4 < a.w;
    // (one load per lane: two predicated loads into the same register wait for each other)
    const int edge_off = lane == 0 ? -1 : 4;
    const bool edge_on = (lane == 0 && has_left) || (lane == 31 && has_right);
    auto ldedge = [&](int y) {
        y = max(0, min(y, a.h - 1));
        float v = 0.0f;
        if (edge_on) v = g[(long)y * a.pitch + xc + edge_off];
        return v;
    };
    const AtanConsts K;
    float4 up = ldrow(y0 - 1), cur = ldrow(y0), dn = ldrow(y0 + 1);
    float edge = ldedge(y0);
    const int y_end = min(y0 + GRAD4_ROWS, a.h);
    for (int y = y0; y < y_end; y++) {
        const float4 nxt = ldrow(y + 2);
        const float edge_nxt = ldedge(y + 1);
        // horizontal neighbours of the 4-column group
        float left = __shfl_up_sync(0xffffffffu, cur.w, 1), right = __shfl_down_sync(0xffffffffu, cur.x, 1);
        if (lane == 0) left = has_left ? edge : cur.x;
        if (lane == 31) right = has_right ? edge : cur.w;
        // image.cl:58-66: xgrad = I[x+1]-I[x-1]; at the two image borders the one-sided difference, doubled
        const int last = a.w - 1 - x4;  // element index of the last image column inside this group (or >= 4)
        const float l0 = x4 == 0 ? cur.x : left, s0 = (x4 == 0 || last == 0) ? 2.0f : 1.0f;
        const float xg0 = s0 * ((last == 0 ? cur.x : cur.y) - l0);
        const float xg1 = (last == 1 ? 2.0f : 1.0f) * ((last == 1 ? cur.y : cur.z) - cur.x);
        const float xg2 = (last == 2 ? 2.0f : 1.0f) * ((last == 2 ? cur.z : cur.w) - cur.y);
        const float xg3 = (last == 3 ? 2.0f : 1.0f) * ((last == 3 ? cur.w : right) - cur.z);
        // image.cl:67-72: ygrad = I[y-1]-I[y+1] ("up minus down"), doubled one-sided at the borders
        float4 yg;
        if (y == 0) yg = make_float4(2.0f * (cur.x - dn.x), 2.0f * (cur.y - dn.y), 2.0f * (cur.z - dn.z), 2.0f * (cur.w - dn.w));
        else if (y == a.h - 1) yg = make_float4(2.0f * (up.x - cur.x), 2.0f * (up.y - cur.y), 2.0f * (up.z - cur.z), 2.0f * (up.w - cur.w));
        else yg = make_float4(up.x - dn.x, up.y - dn.y, up.z - dn.z, up.w - dn.w);
        float4 gr, orv;
        grad_one(xg0, yg.x, gr.x, orv.x, K);
        grad_one(xg1, yg.y, gr.y, orv.y, K);
        grad_one(xg2, yg.z, gr.z, orv.z, K);
        grad_one(xg3, yg.w, gr.w, orv.w, K);
        if (active) {
            const long pos = (long)y * a.pitch + x4;
            if (last >= 3) {  // 32 contiguous bytes per thread
                float4 *dst = reinterpret_cast<float4 *>(gop + pos);
                dst[0] = make_float4(gr.x, orv.x, gr.y, orv.y);
                dst[1] = make_float4(gr.z, orv.z, gr.w, orv.w);
            } else {  // ragged right edge: only the columns inside the image
                gop[pos] = make_float2(gr.x, orv.x);
                if (last >= 1) gop[pos + 1] = make_float2(gr.y, orv.y);
                if (last >= 2) gop[pos + 2] = make_float2(gr.z, orv.z);
            }
        }
        up = cur;
        cur = dn;
        dn = nxt;
        edge = edge_nxt;
    }
}

__global__ void __launch_bounds__(128, 8) k_gradient4(GradArgs a) {
    GradPlane pl;
    pl.g = a.g[blockIdx.z]; pl.go = a.go[blockIdx.z]; pl.pitch = a.pitch; pl.w = a.w; pl.h = a.h;
    gradient4_block(pl, blockIdx.x, blockIdx.y);
}

// The three planes of EVERY octave in one launch (the Gaussian planes of all octaves are kept): 1-D grid, the table
// (device memory, written once per plan) gives the first block of each (octave, plane).
#define GRAD_MAXPLANES 48
struct GradTable {
    GradPlane plane[GRAD_MAXPLANES];
    int bx[GRAD_MAXPLANES], start[GRAD_MAXPLANES + 1];
    int n_planes;
};
__global__ void __launch_bounds__(128, 8) k_gradient4_all(const GradTable *__restrict__ T) {
    int pidx = 0;
    const int n = T->n_planes;
    while (pidx + 1 < n && (int)blockIdx.x >= T->start[pidx + 1]) pidx++;
    const int local = blockIdx.x - T->start[pidx], bxn = T->bx[pidx];
    gradient4_block(T->plane[pidx], local % bxn, local / bxn);
}

// ---------------------------------------------------------------------------------------------
// Gradient / orientation planes of EVERY octave of an image: orientation assignment and descriptors run once
// per image over the keypoints of all octaves (a keypoint carries tag = octave << 8 | scale), so that the
// small octaves do not each pay the latency of a nearly empty launch.
#ifndef SIFTB_KOCT
#define SIFTB_KOCT 16
#endif
struct OctTable {
    const float2 *go[SIFTB_KOCT][3];  // (gradient magnitude, orientation) planes, see GradArgs
    int pitch[SIFTB_KOCT], w[SIFTB_KOCT], h[SIFTB_KOCT];
    int octsize[SIFTB_KOCT];
};

// Size class of a keypoint's descriptor window = its radius in pixels (keypoints_cpu.cl:60-61), clamped to 63.
// k_describe processes the keypoints in descending class order so that the four keypoints sharing a warp have
// windows of the same size (the warp runs as long as its largest window).
#define DESC_CLASSES 64
__device__ __forceinline__ int desc_size_class(float sigma_oct, int octsize) {
    const float spacing = sigma_oct / (float)octsize * 3.0f;
    const int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    return min(max(iradius, 0), DESC_CLASSES - 1);
}

#define ORI_MAXROWS 96  // window rows handled by the chord table (radius <= 47; the default sigmas need 41)
// One warp per keypoint (grid-stride).  kp rows in: (peak, row, col, sigma); out: (x, y, sigma*oct, angle).
// Extra-orientation keypoints are appended at n_base + atomicAdd(n_extra).
// stage: [octave][3 scales][3] counters (may be null); oct_valid[o]: records octave o will emit (non-NaN rows).
// GPUVAR: the semantics of orientation_gpu.cl instead of orientation_cpu.cl (devicetype "GPU" in the
// reference): bin = (int)(18 (ori + pi) / pi) with a +-36 wrap, at most 128 columns per window row, smoothing
// multiplies by (1.0f / 3.0f), the maximum comes from the reference's tree reduction (its tie rules), the angle is
// (argmax + 0.5 + interp) / 18 wrapped into [0, 2] then (a - 1) pi, extra peaks are not range-filtered.  The samples
// are accumulated in the same row-major order in both variants (lane 0 of the reference's work-group adds a row's
// values in column order, orientation_gpu.cl:141-145).
template <bool GPUVAR>
__global__ void __launch_bounds__(256, 5) k_orient(OctTable T, float4 *__restrict__ kp, int *__restrict__ kp_tag,
                                                 const int *__restrict__ n_base_p, int *__restrict__ n_extra, int cap,
                                                 float OriSigma, int *__restrict__ stage, int *__restrict__ oct_valid,
                                                 int *__restrict__ size_hist, int *__restrict__ q_head) {
    __shared__ float s_hist[8][36];
    // per window row: x = packed index of the row's first candidate, y = (first column) - x; two sentinel rows
    __shared__ int2 s_rows[8][ORI_MAXROWS + 3];
    // per-CTA partial counters, added to the global ones once at the end (one keypoint = three increments on a handful
    // of hot addresses otherwise: ~2e5 same-address atomics per image)
    __shared__ int s_stage[SIFTB_KOCT * 9], s_valid[SIFTB_KOCT], s_size[DESC_CLASSES];
    for (int i = threadIdx.x; i < SIFTB_KOCT * 9; i += blockDim.x) s_stage[i] = 0;
    for (int i = threadIdx.x; i < SIFTB_KOCT; i += blockDim.x) s_valid[i] = 0;
    for (int i = threadIdx.x; i < DESC_CLASSES; i += blockDim.x) s_size[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int2 *rows = s_rows[wib];
    const int n_base = min(*n_base_p, cap);
    float *hist = s_hist[wib];
    // Keypoints are handed out one at a time from a queue (*q_head, zero at launch): window sizes differ by 4x, and with
    // a fixed keypoint -> warp assignment the warps finished spread over the last quarter of the kernel (ncu: 14 % of
    // the warp samples sat at the final barrier).  The next index is requested before the current keypoint is processed
    // and only read (shuffle) after it.
    int ticket = 0;  // lane 0: the next index, broadcast only when the current keypoint is done
    if (lane == 0) ticket = atomicAdd(q_head, 1);
    for (int gid0 = __shfl_sync(0xffffffffu, ticket, 0); gid0 < n_base; gid0 = __shfl_sync(0xffffffffu, ticket, 0)) {
        if (lane == 0) ticket = atomicAdd(q_head, 1);
        float4 k = kp[gid0];
        const int tag = kp_tag[gid0];
        const int sc = tag & 0xff, oct = tag >> 8;
        if (!(k.y >= 0.0f)) continue;  // warp-uniform
        const float2 *go = T.go[oct][sc - 1];
        const int Gpitch = T.pitch[oct], Gw = T.w[oct], Gh = T.h[oct], octsize = T.octsize[oct];
        const int row = (int)((double)k.y + 0.5), col = (int)((double)k.z + 0.5);  // orientation_cpu.cl:67-68
        const float sigma = OriSigma * k.w;
        const int radius = (int)((double)sigma * 3.0);  // :71
        const int rmin = max(0, row - radius), cmin = max(0, col - radius);
        const int rmax = min(row + radius, Gh - 2);
        int cmax = min(col + radius, Gw - 2);
        if (GPUVAR) cmax = min(cmax, cmin + 127);  // c = cmin + lid0, lid0 < WORKGROUP_SIZE (orientation_gpu.cl:126-129)
        const float two_s2 = (2.0f * sigma) * sigma;
        const double inv_two_s2 = div_prepare(two_s2), inv_two_pi = div_prepare(2.0f * SIFTB_M_PI_F);
        const float rad2 = ((float)(radius * radius)) + 0.5f;
        const int ncols = cmax - cmin + 1, nrows = rmax - rmin + 1;
        hist[lane] = 0.0f;
        if (lane < 4) hist[32 + lane] = 0.0f;
        // Candidate columns of every window row: the reference scans the square [rmin, rmax] x [cmin, cmax] and
        // keeps the samples with distsq < rad2 (orientation_cpu.cl:78-85); per row those lie on the chord
        // |c - k.z| < sqrt(rad2 - dr^2).  The chord is computed in double with margins (1e-5 relative, 1e-3
        // absolute) far above the fp32 evaluation error of distsq (< 3e-7 relative), so it is a superset; every
        // candidate still goes through the reference's exact fp32 test.  Windows with more than ORI_MAXROWS rows
        // (never with the default sigmas) scan the full square.
        const bool chord = nrows <= ORI_MAXROWS;  // warp-uniform
        int total = (ncols > 0 && nrows > 0) ? ncols * nrows : 0;
        if (chord && total > 0) {
            int base = 0;
            for (int r0 = 0; r0 < nrows; r0 += 32) {  // warp-uniform trip count
                const int rr = r0 + lane;
                int lo = 0, len = 0;
                if (rr < nrows) {
                    const double drd = (double)(rmin + rr) - (double)k.y;
                    const double hw2 = (double)rad2 * 1.00001 - drd * drd * 0.99999 + 1e-3;
                    if (hw2 > 0.0) {
                        const double hw = sqrt(hw2);
                        lo = max(cmin, (int)ceil((double)k.z - hw));
                        len = max(0, min(cmax, (int)floor((double)k.z + hw)) - lo + 1);
                    }
                }
                int incl = len;  // inclusive prefix sum over the 32 rows of this chunk
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                const int start = base + incl - len;
                if (rr < nrows) rows[rr] = make_int2(start, lo - start);
                base += __shfl_sync(0xffffffffu, incl, 31);
            }
            total = base;
            if (lane < 3) rows[nrows + lane] = make_int2(lane == 0 ? total : 0x7fffffff, 0);
        }
        __syncwarp();
        // Per-lane cursor over the candidates in row-major order: lane l takes candidates l, l + 32, ... (packed index
        // `cand`); its row is found by walking the row table forward, two rows per probe.  The gradient / orientation
        // values of chunk c+1 are requested before chunk c is evaluated and committed (L2 / DRAM gathers: ncu showed
        // the warps mostly waiting on these loads).
        // The loop below is unrolled twice by hand over two sets of these values (A, B): with one set the compiler
        // copied "next" into "current" at the END of the iteration, and that copy waited for the load issued at its
        // beginning (24 % of the warp samples, 0.25 instead of 0.21 ms).
        struct Sample {
            float gval, ang;
            int r, c;
            bool ok;
        };
        int rcur = 0, cur_off = 0, cand = lane;
        if (chord && total > 0) cur_off = rows[0].y;
        auto locate = [&](Sample &n) {
            n.ok = cand < total;
            n.gval = 0.0f; n.ang = 0.0f; n.r = 0; n.c = 0;
            if (chord) {
                if (total > 0) {
                    int2 e1 = rows[rcur + 1], e2 = rows[rcur + 2];
                    while (cand >= e2.x) {
                        rcur += 2;
                        cur_off = e2.y;
                        e1 = rows[rcur + 1];
                        e2 = rows[rcur + 2];
                    }
                    if (cand >= e1.x) {
                        rcur++;
                        cur_off = e1.y;
                    }
                }
                n.r = rmin + rcur;
                n.c = cand + cur_off;
            } else if (n.ok) {
                const int rr = cand / ncols;
                n.r = rmin + rr;
                n.c = cmin + (cand - rr * ncols);
            }
            if (n.ok) {
                const float2 v = ldg_f2_here(go + ((long)n.r * Gpitch + n.c));
                n.gval = v.x;
                n.ang = v.y;
            }
            cand += 32;
        };
        auto process = [&](const Sample &cur) {
            int bin = -1;
            float w = 0.0f;
            if (cur.ok) {
                float dif = ((float)cur.r - k.y);
                float distsq = dif * dif;
                dif = ((float)cur.c - k.z);
                distsq += dif * dif;
                if (cur.gval > 0.0f && distsq < rad2) {
                    if (GPUVAR) {  // orientation_gpu.cl:135-139
                        int b = (int)((18.0f * (cur.ang + SIFTB_M_PI_F)) * SIFTB_M_1_PI_F);
                        if (b < 0) b += 36;
                        if (b > 35) b -= 36;
                        bin = max(0, min(b, 35));  // (the clamp only guards shared memory against a NaN plane)
                        w = cr_expf_neg(div_by(-distsq, inv_two_s2)) * cur.gval;
                    } else {
                        int b = (int)div_by(36.0f * ((cur.ang + SIFTB_M_PI_F) + 0.001f), inv_two_pi);
                        if (b >= 0 && b <= 36) {
                            bin = min(b, 35);
                            w = cr_expf_neg(div_by(-distsq, inv_two_s2)) * cur.gval;
                        }
                    }
                }
            }
            // commit: bins are independent chains; lanes that hit the same bin add in lane order (== the reference's
            // row-major sample order), different bins add concurrently.  Neighbouring pixels have similar
            // orientations, several lanes per bin are typical: the chains run in registers -- a lane gets the running
            // sum from the previous lane of its bin by shuffle -- and shared memory sees one read by the first and
            // one write by the last lane of every bin.
            if (__any_sync(0xffffffffu, bin >= 0)) {
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                const unsigned before = peers & lanemask_lt();
                const int rank = bin >= 0 ? __popc(before) : 0;
                const int prev_lane = before ? 31 - __clz(before) : lane;
                const int rounds = __reduce_max_sync(0xffffffffu, rank);
                float run = 0.0f;  // the bin's value after this lane's term
                if (bin >= 0 && rank == 0) run = hist[bin] + w;
                for (int rd = 1; rd <= rounds; rd++) {
                    const float upto = __shfl_sync(0xffffffffu, run, prev_lane);
                    if (rank == rd) run = upto + w;
                }
                // (the write below is ordered after the first lane's read of the same bin by the data flow through
                // the shuffles; the barrier states it for compute-sanitizer's racecheck and costs no instruction)
                __syncwarp();
                if (bin >= 0 && (peers >> lane) == 1u) hist[bin] = run;  // last lane of its bin
                __syncwarp();  // the next step's first lane of this bin may be another lane
            }
        };
        Sample sa, sb;
        locate(sa);
        for (int base = 0; base < total; base += 64) {
            locate(sb);
            process(sa);
            if (base + 32 >= total) break;  // warp-uniform
            locate(sa);
            process(sb);
        }
        // orientation_cpu.cl:100-108 -- six in-place smoothing passes.  In place means: bins 0..34 see
        // the OLD neighbours (prev is carried), bin 35 sees the NEW bin 0.  "/ 3.0" is a double division.
        // x / 3.0 in double, rounded to fp32: q = x*(1/3) corrected once through the exact residual gives the same
        // fp32 result as the IEEE division for EVERY fp32 x (exhaustive check, tools/div3_check.c)
        auto third = [](float x) {
            const double xd = (double)x, q0 = xd * (1.0 / 3.0);
            return (float)fma(fma(-3.0, q0, xd), 1.0 / 3.0, q0);
        };
        for (int j = 0; j < 6; j++) {
            const float a0 = hist[(lane + 35) % 36], b0 = hist[lane], c0 = hist[lane + 1];
            float a1 = 0.f, b1 = 0.f, c1 = 0.f;
            if (lane < 3) { a1 = hist[31 + lane], b1 = hist[32 + lane], c1 = hist[33 + lane]; }
            float o34 = hist[34], o35 = hist[35];
            __syncwarp();
            // GPU variant: "* ONE_3" in fp32 (orientation_gpu.cl:160-170); same data flow (old neighbours, bin 35 sees
            // the new bin 0: what lock-step execution makes of the reference's racy update)
            const float ONE_3 = 1.0f / 3.0f;
            const float n0 = GPUVAR ? ((a0 + b0) + c0) * ONE_3 : third((a0 + b0) + c0);
            hist[lane] = n0;
            if (lane < 3) hist[32 + lane] = GPUVAR ? ((a1 + b1) + c1) * ONE_3 : third((a1 + b1) + c1);
            __syncwarp();
            if (lane == 0) hist[35] = GPUVAR ? ((o34 + o35) + n0) * ONE_3 : third((o34 + o35) + n0);
            __syncwarp();
        }
        // orientation_cpu.cl:110-121 -- argmax = first bin holding the largest value > 0 (0 if there is none;
        // NaN bins never win a comparison).  Lane i looks at bin i, lanes 0..3 also at bins 32..35.
        const float h0 = hist[lane], h1 = lane < 4 ? hist[32 + lane] : 0.0f;
        const float m0 = h0 > 0.0f ? h0 : 0.0f, m1 = h1 > 0.0f ? h1 : 0.0f;
        float maxval = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(m0, m1))));
        int argmax = 0;
        if (GPUVAR) {
            // orientation_gpu.cl:187-236: bins 32..35 are folded into lanes 0..3 (a tie keeps the HIGHER bin), then
            // the halving steps 16, 8, 4, 2, 1 (a tie keeps the LOWER lane)
            float v = h0;
            int pbin = lane;
            if (lane < 4 && !(h0 > h1)) { v = h1; pbin = lane + 32; }
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const float uv = __shfl_down_sync(0xffffffffu, v, step);
                const int up = __shfl_down_sync(0xffffffffu, pbin, step);
                if (lane < step && uv > v) { v = uv; pbin = up; }
            }
            maxval = __shfl_sync(0xffffffffu, v, 0);
            argmax = __shfl_sync(0xffffffffu, pbin, 0);
        } else if (maxval > 0.0f) {
            const unsigned e0 = __ballot_sync(0xffffffffu, h0 == maxval);
            const unsigned e1 = __ballot_sync(0xffffffffu, lane < 4 && h1 == maxval);
            argmax = e0 ? __ffs(e0) - 1 : 31 + __ffs(e1);
        }
        const float ONE_18 = 1.0f / 18.0f;
        auto gpu_angle = [&](int i, float itp) {  // orientation_gpu.cl:255-264, 303-306
            float a = (((float)i + 0.5f) + itp) * ONE_18;
            if (a < 0.0f) a += 2.0f;
            else if (a > 2.0f) a -= 2.0f;
            return (a - 1.0f) * SIFTB_M_PI_F;
        };
        float angle;
        float4 o;
        {
            const int prev = (argmax == 0 ? 35 : argmax - 1), next = (argmax == 35 ? 0 : argmax + 1);
            float hist_prev = hist[prev], hist_next = hist[next];
            if (!GPUVAR && maxval < 0.0f) { hist_prev = -hist_prev; maxval = -maxval; hist_next = -hist_next; }
            const float interp = 0.5f * (hist_prev - hist_next) / ((hist_prev - 2.0f * maxval) + hist_next);
            angle = GPUVAR ? gpu_angle(argmax, interp)
                           : (2.0f * SIFTB_M_PI_F) * (((float)argmax + 0.5f) + interp) / 36.0f - SIFTB_M_PI_F;
            o.x = k.z * (float)octsize;
            o.y = k.y * (float)octsize;
            o.z = k.w * (float)octsize;
            o.w = angle;
            if (lane == 0) kp[gid0] = o;
        }
        // orientation_cpu.cl:131-172 -- every other local peak >= 0.8 max gives an extra keypoint
        auto peak_angle = [&](int i, float &a2) {
            const int pv = (i == 0 ? 35 : i - 1), nx = (i == 35 ? 0 : i + 1);
            float hp = hist[pv], hc = hist[i], hn = hist[nx];
            if (!(hc > hp && hc > hn && hc >= 0.8f * maxval && i != argmax)) return false;
            if (!GPUVAR && hc < 0.0f) { hp = -hp; hc = -hc; hn = -hn; }
            const float itp = 0.5f * (hp - hn) / ((hp - 2.0f * hc) + hn);
            if (GPUVAR) {  // every such peak becomes a keypoint (no range filter, orientation_gpu.cl:286-311)
                a2 = gpu_angle(i, itp);
                return true;
            }
            // orientation_cpu.cl:166: "/36.0" promotes the tail of the expression to double
            a2 = (float)((double)((2.0f * SIFTB_M_PI_F) * (((float)i + 0.5f) + itp)) / 36.0 - (double)SIFTB_M_PI_F);
            return a2 >= -SIFTB_M_PI_F && a2 <= SIFTB_M_PI_F;
        };
        float a2_0 = 0.0f, a2_1 = 0.0f;
        const bool p0 = peak_angle(lane, a2_0);
        const bool p1 = lane < 4 && peak_angle(32 + lane, a2_1);
        const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
        const int n_extra_here = __popc(b0) + __popc(b1);
        int n_stored = 0;  // extra rows that fit into the list: only those are counted below (the reference's
                           // counter runs past the buffer, plan.py:771 only warns; here overflow truncates cleanly)
        if (n_extra_here) {
            int slot0 = 0;
            if (lane == 0) slot0 = n_base + atomicAdd(n_extra, n_extra_here);
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            n_stored = max(0, min(n_extra_here, cap - slot0));
            if (p0) {
                const int old = slot0 + __popc(b0 & lanemask_lt());
                if (old < cap) { kp[old] = make_float4(o.x, o.y, o.z, a2_0); kp_tag[old] = tag; }
            }
            if (p1) {
                const int old = slot0 + __popc(b0) + __popc(b1 & lanemask_lt());
                if (old < cap) { kp[old] = make_float4(o.x, o.y, o.z, a2_1); kp_tag[old] = tag; }
            }
        }
        if (lane == 0) {
            const int added = 1 + n_stored;
            atomicAdd(&s_stage[oct * 9 + (sc - 1) * 3 + 2], added);
            // rows whose angle is NaN (flat histogram) are dropped on output (plan.py:546-550)
            atomicAdd(&s_valid[oct], added - ((angle != angle) ? 1 : 0));
            // descriptor-window size class of this keypoint and of its extra orientations (same sigma)
            atomicAdd(&s_size[desc_size_class(o.z, octsize)], added);
        }
        __syncwarp();
    }
    __syncthreads();
    if (stage)
        for (int i = threadIdx.x; i < SIFTB_KOCT * 9; i += blockDim.x)
            if (s_stage[i]) atomicAdd(&stage[i], s_stage[i]);
    if (oct_valid)
        for (int i = threadIdx.x; i < SIFTB_KOCT; i += blockDim.x)
            if (s_valid[i]) atomicAdd(&oct_valid[i], s_valid[i]);
    if (size_hist)
        for (int i = threadIdx.x; i < DESC_CLASSES; i += blockDim.x)
            if (s_size[i]) atomicAdd(&size_hist[i], s_size[i]);
}

struct KpRecord {  // == siftb_kp
    float x, y, scale, angle;
    uint8_t desc[128];
};
