// k_extrema.cuh -- 3x3x3 DoG extrema detection and sub-pixel refinement.
//
// k_extrema replaces image.cl:119 local_maxmin (one launch per scale in the reference,
// plan.py:626-638) plus memset.cl memset_float/_int (plan.py:797): all three scales of an octave in
// one launch, warp-aggregated append.  k_refine replaces image.cl:235 interp_keypoint AND
// algebra.cl:57 compact (+ the host round trips of plan.py:642,758-795): rejected candidates are
// simply not written, so there are no holes to compact.
#pragma once
#include "common.cuh"

struct DogStack {
    const float *d[5];
    int pitch, w, h;
};

// image.cl:141-212 for one pixel and one scale, split in two steps with the same decisions as the reference's
// flag loops (strict comparisons, plateaus fire):
//  * maxmin_gate: contrast gate and the two neighbouring scales at the same pixel, from the five DoG values
//    c[0..4] already in registers.  gate: smallest fp32 strictly above 0.8*peak_thresh evaluated in double
//    (image.cl:152 compares in double: fabs(val) > 0.8*peak_thresh <=> fabsf(val) >= gate for fp32 val); computed
//    on the host.  2-3.5 % of the pixels survive it.
//  * maxmin_rest: the other 24 neighbours, ordered so that most pixels leave after a few compares, then the edge
//    test (image.cl:180-186: H00/H11 in double (literal 2.0), H01 float differences then /4.0).
__device__ __forceinline__ bool maxmin_gate(const float c[5], int scale, float gate) {
    // branch-free (12 of these run per thread and row): for val > 0 the neighbours must not exceed it, otherwise
    // they must not lie below it -- the reference's sign-flipped strict comparisons, NaN compares false either way
    const float val = c[scale], a = c[scale - 1], b = c[scale + 1];
    const bool strong = fabsf(val) >= gate;
    const bool above = (a > val) | (b > val), below = (a < val) | (b < val);
    return strong & !((val > 0.0f) ? above : below);
}
__device__ __forceinline__ bool maxmin_rest(const DogStack &D, long pos, int scale, float edthresh) {
    // Two batches of independent loads instead of 24 loads behind 24 early-outs (each a dependent L2 round trip):
    // the 8 in-plane neighbours first -- they decide most candidates and also feed the edge test -- then the 16
    // remaining neighbours in the two adjacent scales.
    const float *dc = D.d[scale];
    const long up = pos - D.pitch, dn_ = pos + D.pitch;
    const float val = dc[pos];
    const float n_l = dc[pos - 1], n_r = dc[pos + 1];
    const float u_l = dc[up - 1], u_c = dc[up], u_r = dc[up + 1];
    const float d_l = dc[dn_ - 1], d_c = dc[dn_], d_r = dc[dn_ + 1];
    const float sgn = (val > 0.0f) ? 1.0f : -1.0f;
    const float sval = sgn * val;
    if (sgn * n_l > sval || sgn * n_r > sval || sgn * u_l > sval || sgn * u_c > sval || sgn * u_r > sval ||
        sgn * d_l > sval || sgn * d_c > sval || sgn * d_r > sval)
        return false;
    float o[16];
#pragma unroll
    for (int dpl = 0; dpl < 2; dpl++) {
        const float *pl = D.d[scale + 2 * dpl - 1] + pos;
        o[8 * dpl + 0] = pl[-1];
        o[8 * dpl + 1] = pl[1];
#pragma unroll
        for (int dr = 0; dr < 2; dr++) {
            const float *rowp = pl + (long)(2 * dr - 1) * D.pitch;
            o[8 * dpl + 2 + 3 * dr] = rowp[-1];
            o[8 * dpl + 3 + 3 * dr] = rowp[0];
            o[8 * dpl + 4 + 3 * dr] = rowp[1];
        }
    }
    bool beaten = false;
#pragma unroll
    for (int i = 0; i < 16; i++) beaten = beaten || (sgn * o[i] > sval);
    if (beaten) return false;
    float H00 = (float)(((double)u_c - 2.0 * (double)val) + (double)d_c);
    float H11 = (float)(((double)n_l - 2.0 * (double)val) + (double)n_r);
    float d1 = d_r - d_l;
    float d2 = u_r - u_l;
    float H01 = (float)((double)(d1 - d2) / 4.0);
    float det = H00 * H11 - H01 * H01, trace = H00 + H11;  // -fmad=false: no contraction
    float tt = edthresh * trace;
    tt = tt * trace;
    if (det < tt) return false;
    return val != 0.0f;
}

#define EXT_ROWS 16
// Vector form (every plane of a SiftPlan: pitch % 4 == 0, 16-byte aligned base).
// grid: (ceil(w/512), ceil((h-2*border)/EXT_ROWS)), block 128 threads; a thread owns FOUR adjacent columns and walks
// EXT_ROWS rows: one 128-bit load per DoG plane and row gives the five values of its four pixels (the row after next
// is already requested), maxmin_gate runs on all 12 (pixel, scale) pairs from registers and leaves a 12-bit survivor
// mask per thread.  The survivors of a warp (2-3.5 % of the pairs) are queued in shared memory -- slots come from a
// warp prefix sum of the masks' popcounts -- and finished 32 at a time, one per lane (maxmin_rest is long and would
// otherwise run for one or two lanes of a warp at a time).
// cand rows: (val, row, col, scale).  n_cand = total candidates, stage[(s-1)*3] = per-scale count.
#define EXT_QSIZE 512  // ring entries per warp: <= 31 left over + 32 lanes x 12 new survivors per row
// (bx, by): the block's position in the grid described above
__device__ __forceinline__ void extrema_block(const DogStack &D, int border, float gate, float edthresh,
                                              float4 *__restrict__ cand, int cap, int *__restrict__ n_cand,
                                              int *__restrict__ stage /* [3][3] or null */, int scale_lo, int nscales,
                                              int bx, int by) {
    __shared__ unsigned short s_q[4][EXT_QSIZE];  // (row offset << 9) | (scale index << 7) | (lane << 2) | column in group
    __shared__ float4 s_hits[4][64];
    const int lane = threadIdx.x & 31;
    const int x4 = 4 * (bx * blockDim.x + threadIdx.x);
    const int col0 = x4 - 4 * lane;                // first column of the warp
    const int row0 = border + by * EXT_ROWS;
    unsigned short *q = s_q[threadIdx.x >> 5];
    const bool active = x4 < D.pitch;              // the 128-bit loads stay inside the (padded) row
    unsigned colmask = 0;                          // columns of my group inside [border, w - border)
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (x4 + k >= border && x4 + k < D.w - border) colmask |= 1u << k;
    if (!active) colmask = 0;
    colmask |= (colmask << 4) | (colmask << 8);    // the same columns for the three scales
    unsigned scalemask = 0;
#pragma unroll
    for (int si = 0; si < 3; si++)
        if (1 + si >= scale_lo && 1 + si < scale_lo + nscales) scalemask |= 0xfu << (4 * si);
    colmask &= scalemask;
    int head = 0, qn = 0;  // warp-uniform ring state
    // Extrema found by this warp are collected in shared memory and appended to the candidate list in batches: one
    // atomicAdd round trip per ~32 candidates instead of one per group of survivors (the warp waits for its result).
    float4 *hits = s_hits[threadIdx.x >> 5];
    int n_hits = 0, cnt_s1 = 0, cnt_s2 = 0, cnt_s3 = 0;  // warp-uniform
    auto flush = [&]() {
        if (n_hits == 0) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(n_cand, n_hits);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < n_hits; i += 32)
            if (base + i < cap) cand[base + i] = hits[i];  // image.cl:202-208
        n_hits = 0;
        __syncwarp();
    };
    auto finish = [&](int nb) {  // the first nb (<= 32) queued survivors, one per lane
        bool hit = false;
        int gid1 = 0, gcol = 0, scale = 1;
        long pos = 0;
        if (lane < nb) {
            const unsigned e = q[(head + lane) & (EXT_QSIZE - 1)];
            gid1 = row0 + (int)(e >> 9);
            gcol = col0 + (int)(e & 127u);
            scale = 1 + (int)((e >> 7) & 3u);
            pos = (long)gid1 * D.pitch + gcol;
            hit = maxmin_rest(D, pos, scale, edthresh);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            if (hit) hits[n_hits + __popc(m & lanemask_lt())] = make_float4(D.d[scale][pos], (float)gid1, (float)gcol, (float)scale);
            n_hits += __popc(m);
            cnt_s1 += __popc(__ballot_sync(0xffffffffu, hit && scale == 1));
            cnt_s2 += __popc(__ballot_sync(0xffffffffu, hit && scale == 2));
            cnt_s3 += __popc(__ballot_sync(0xffffffffu, hit && scale == 3));
            __syncwarp();
            if (n_hits > 32) flush();  // room for the next 32
        }
        head = (head + nb) & (EXT_QSIZE - 1);
        qn -= nb;
    };
    const int r_end = min(EXT_ROWS, D.h - border - row0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld5 = [&](int r, float4 v[5]) {
#pragma unroll
        for (int i = 0; i < 5; i++)
            v[i] = (active && r < r_end) ? __ldg(reinterpret_cast<const float4 *>(D.d[i] + (long)(row0 + r) * D.pitch + x4)) : zero4;
    };
    float4 c1[5], c2[5];
    ld5(0, c1);
    ld5(1, c2);
    for (int r = 0; r < r_end; r++) {
        float4 c[5];
#pragma unroll
        for (int i = 0; i < 5; i++) { c[i] = c1[i]; c1[i] = c2[i]; }
        ld5(r + 2, c2);
        unsigned mask = 0;
#pragma unroll
        for (int si = 0; si < 3; si++) {
            const float px[4][5] = {{c[0].x, c[1].x, c[2].x, c[3].x, c[4].x}, {c[0].y, c[1].y, c[2].y, c[3].y, c[4].y},
                                    {c[0].z, c[1].z, c[2].z, c[3].z, c[4].z}, {c[0].w, c[1].w, c[2].w, c[3].w, c[4].w}};
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (maxmin_gate(px[k], 1 + si, gate)) mask |= 1u << (4 * si + k);
        }
        mask &= colmask;
        if (__any_sync(0xffffffffu, mask != 0)) {
            // queue slots: exclusive prefix sum of the popcounts over the lanes
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int slot = head + qn + incl - cnt;
            while (mask) {
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                q[slot++ & (EXT_QSIZE - 1)] = (unsigned short)((r << 9) | ((bit >> 2) << 7) | (lane << 2) | (bit & 3));
            }
            qn += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        while (qn >= 32) { finish(32); __syncwarp(); }
    }
    while (qn > 0) { finish(min(qn, 32)); __syncwarp(); }
    flush();
    if (stage && lane == 0) {
        if (cnt_s1) atomicAdd(&stage[0], cnt_s1);
        if (cnt_s2) atomicAdd(&stage[3], cnt_s2);
        if (cnt_s3) atomicAdd(&stage[6], cnt_s3);
    }
}

__global__ void __launch_bounds__(128) k_extrema(DogStack D, int border, float gate, float edthresh,
                                                  float4 *__restrict__ cand, int cap, int *__restrict__ n_cand,
                                                  int *__restrict__ stage, int scale_lo, int nscales) {
    extrema_block(D, border, gate, edthresh, cand, cap, n_cand, stage, scale_lo, nscales, blockIdx.x, blockIdx.y);
}

// All octaves of an image in ONE launch (the DoG planes of every octave are kept): the small octaves -- a few dozen
// blocks each, pure launch latency on their own -- ride in the tail of octave 0's grid.  The table lives in device
// memory (written once per plan and slot); 1-D grid, block b belongs to the octave whose [start, start + blocks)
// range holds it.
#define SIFTB_KOCT 16
struct PyrTable {
    DogStack ds[SIFTB_KOCT];
    float4 *cand[SIFTB_KOCT];
    int *n_cand[SIFTB_KOCT];      // per octave: candidates found (k_extrema) ...
    int *n_kp_oct[SIFTB_KOCT];    // ... and kept by k_refine
    int *stage[SIFTB_KOCT];       // [3 scales][3] counters
    int cap[SIFTB_KOCT];
    float edthresh[SIFTB_KOCT];
    int ext_bx[SIFTB_KOCT], ext_start[SIFTB_KOCT + 1];   // k_extrema_all: blocks per row of octave o, first block of octave o
    int n_oct;
};
__global__ void __launch_bounds__(128, 5) k_extrema_all(const PyrTable *__restrict__ T, int border, float gate) {
    int o = 0;
    const int n_oct = T->n_oct;
    while (o + 1 < n_oct && (int)blockIdx.x >= T->ext_start[o + 1]) o++;
    const int local = blockIdx.x - T->ext_start[o], bxn = T->ext_bx[o];
    const DogStack D = T->ds[o];
    extrema_block(D, border, gate, T->edthresh[o], T->cand[o], T->cap[o], T->n_cand[o], T->stage[o], 1, 3, local % bxn,
                  local / bxn);
}

// Scalar form for planes whose pitch is not a multiple of 4 floats (stage hook on dense host planes).
// grid: (ceil(w/128), ceil((h-2*border)/EXT_ROWS)), block 128 threads along x; every thread walks EXT_ROWS rows
// of its column and runs maxmin_gate on nscales scales per pixel from the five DoG values loaded together.  The
// survivors of a warp are queued in shared memory and finished 32 at a time, one per lane (maxmin_rest is long and
// would otherwise run for one or two lanes of a warp at a time).
// cand rows: (val, row, col, scale).  n_cand = total candidates, stage[(s-1)*3] = per-scale count.
__global__ void __launch_bounds__(128) k_extrema_scalar(DogStack D, int border, float gate, float edthresh,
                                                  float4 *__restrict__ cand, int cap, int *__restrict__ n_cand,
                                                  int *__restrict__ stage /* [3][3] or null */, int scale_lo,
                                                  int nscales) {
    __shared__ unsigned short s_q[4][128];  // per warp: (row offset << 7) | (scale index << 5) | lane
    const int gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int row0 = border + blockIdx.y * EXT_ROWS;
    const int lane = threadIdx.x & 31, col0 = gid0 - lane;
    unsigned short *q = s_q[threadIdx.x >> 5];
    const bool col_ok = gid0 >= border && gid0 < D.w - border;
    int head = 0, qn = 0;  // warp-uniform ring state
    auto finish = [&](int nb) {  // the first nb (<= 32) queued survivors, one per lane
        bool hit = false;
        int gid1 = 0, gcol = 0, scale = 1;
        long pos = 0;
        if (lane < nb) {
            const unsigned e = q[(head + lane) & 127];
            gid1 = row0 + (int)(e >> 7);
            gcol = col0 + (int)(e & 31u);
            scale = 1 + (int)((e >> 5) & 3u);
            pos = (long)gid1 * D.pitch + gcol;
            hit = maxmin_rest(D, pos, scale, edthresh);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(n_cand, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (hit) {
                const int slot = base + __popc(m & lanemask_lt());
                if (slot < cap) cand[slot] = make_float4(D.d[scale][pos], (float)gid1, (float)gcol, (float)scale);  // image.cl:202-208
            }
            if (stage) {
#pragma unroll
                for (int sc = 1; sc <= 3; sc++) {
                    const unsigned msc = __ballot_sync(0xffffffffu, hit && scale == sc);
                    if (lane == 0 && msc) atomicAdd(&stage[(sc - 1) * 3 + 0], __popc(msc));
                }
            }
        }
        head = (head + nb) & 127;
        qn -= nb;
    };
    const int r_end = min(EXT_ROWS, D.h - border - row0);
    // the five values of the next two rows are requested before the current row is tested (bytes in flight)
    float c1[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, c2[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (col_ok) {
#pragma unroll
        for (int i = 0; i < 5; i++) {
            if (0 < r_end) c1[i] = D.d[i][(long)row0 * D.pitch + gid0];
            if (1 < r_end) c2[i] = D.d[i][(long)(row0 + 1) * D.pitch + gid0];
        }
    }
    for (int r = 0; r < r_end; r++) {
        float c[5];
#pragma unroll
        for (int i = 0; i < 5; i++) { c[i] = c1[i]; c1[i] = c2[i]; }
        if (col_ok && r + 2 < r_end) {
#pragma unroll
            for (int i = 0; i < 5; i++) c2[i] = D.d[i][(long)(row0 + r + 2) * D.pitch + gid0];
        }
#pragma unroll
        for (int si = 0; si < 3; si++) {
            const int scale = 1 + si;
            if (scale < scale_lo || scale >= scale_lo + nscales) continue;  // block-uniform
            const bool surv = col_ok && maxmin_gate(c, scale, gate);
            const unsigned m = __ballot_sync(0xffffffffu, surv);
            if (surv) q[(head + qn + __popc(m & lanemask_lt())) & 127] = (unsigned short)((r << 7) | (si << 5) | lane);
            qn += __popc(m);
        }
        __syncwarp();
        while (qn >= 32) { finish(32); __syncwarp(); }
    }
    while (qn > 0) { finish(min(qn, 32)); __syncwarp(); }
}

// image.cl:249-366 for one candidate; returns true if kept, result (peak, row, col, sigma)
__device__ __forceinline__ bool interp_one(const DogStack &D, float4 k, float peak_thresh, float InitSigma,
                                           float4 *out) {
    int r = (int)k.y, c = (int)k.z;
    const int scale = (int)k.w;
    const float *Dp = D.d[scale - 1], *Dc = D.d[scale], *Dn = D.d[scale + 1];
    const int width = D.w, height = D.h, P = D.pitch;
    float solution0 = 0.f, solution1 = 0.f, solution2 = 0.f, peakval = 0.f;
    int loop = 1, movesRemain = 5, newr = r, newc = c;
    while (loop == 1) {
        r = newr, c = newc;
        const long pos = (long)newr * P + newc, up = pos - P, dn = pos + P;
        float g0 = (Dn[pos] - Dp[pos]) / 2.0f;
        float g1 = (Dc[dn] - Dc[up]) / 2.0f;
        float g2 = (Dc[pos + 1] - Dc[pos - 1]) / 2.0f;
        float t2 = 2.0f * Dc[pos];
        float H00 = (Dp[pos] - t2) + Dn[pos];
        float H11 = (Dc[up] - t2) + Dc[dn];
        float H22 = (Dc[pos - 1] - t2) + Dc[pos + 1];
        float H01 = ((Dn[dn] - Dn[up]) - (Dp[dn] - Dp[up])) / 4.0f;
        float H02 = ((Dn[pos + 1] - Dn[pos - 1]) - (Dp[pos + 1] - Dp[pos - 1])) / 4.0f;
        float H12 = ((Dc[dn + 1] - Dc[dn - 1]) - (Dc[up + 1] - Dc[up - 1])) / 4.0f;
        float H10 = H01, H20 = H02, H21 = H12;
        // image.cl:300 (left-to-right, no contraction)
        float t1 = (H02 * H11) * H20, t2b = (H01 * H12) * H20, t3 = (H02 * H10) * H21;
        float t4 = (H00 * H12) * H21, t5 = (H01 * H10) * H22, t6 = (H00 * H11) * H22;
        float det = ((((-t1 + t2b) + t3) - t4) - t5) + t6;
        float K00 = H11 * H22 - H12 * H21;
        float K01 = H02 * H21 - H01 * H22;
        float K02 = H01 * H12 - H02 * H11;
        float K10 = H12 * H20 - H10 * H22;
        float K11 = H00 * H22 - H02 * H20;
        float K12 = H02 * H10 - H00 * H12;
        float K20 = H10 * H21 - H11 * H20;
        float K21 = H01 * H20 - H00 * H21;
        float K22 = H00 * H11 - H01 * H10;
        solution0 = -((g0 * K00 + g1 * K01) + g2 * K02) / det;
        solution1 = -((g0 * K10 + g1 * K11) + g2 * K12) / det;
        solution2 = -((g0 * K20 + g1 * K21) + g2 * K22) / det;
        peakval = Dc[pos] + 0.5f * ((solution0 * g0 + solution1 * g1) + solution2 * g2);
        if (solution1 > 0.6f && newr < height - 3) newr++;
        else if (solution1 < -0.6f && newr > 3) newr--;
        if (solution2 > 0.6f && newc < width - 3) newc++;
        else if (solution2 < -0.6f && newc > 3) newc--;
        if (movesRemain > 0 && (newr != r || newc != c)) movesRemain--;
        else loop = 0;
    }
    if (fabsf(solution0) <= 1.5f && fabsf(solution1) <= 1.5f && fabsf(solution2) <= 1.5f &&
        fabsf(peakval) >= peak_thresh) {
        out->x = peakval;
        out->y = (float)r + solution1;
        out->z = (float)c + solution2;
        out->w = InitSigma * cr_exp2f((((float)scale) + solution0) / 3.0f);  // pow(2.0f, .), image.cl:355
        return true;
    }
    return false;
}

// All octaves in one launch: thread t takes candidate t of the concatenation of the per-octave candidate lists
// (grid-stride over the device-resident counts); survivors are appended to the image-wide keypoint list.
__global__ void __launch_bounds__(128) k_refine_all(const PyrTable *__restrict__ T, float peak_thresh, float InitSigma,
                                                     float4 *__restrict__ kp, int *__restrict__ kp_tag, int kp_cap,
                                                     int *__restrict__ n_kp) {
    __shared__ int s_first[SIFTB_KOCT + 1];  // first[o] = candidates of the octaves before o
    const int n_oct = T->n_oct;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int o = 0; o < n_oct; o++) {
            s_first[o] = acc;
            acc += min(*T->n_cand[o], T->cap[o]);
        }
        for (int o = n_oct; o <= SIFTB_KOCT; o++) s_first[o] = acc;
    }
    __syncthreads();
    const int n = s_first[SIFTB_KOCT];
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (n + stride - 1) / stride;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < rounds; it++, i += stride) {
        bool keep = false;
        float4 res = make_float4(-1.f, -1.f, -1.f, -1.f);
        int scale = 0, oct = 0;
        if (i < n) {
            while (oct + 1 < n_oct && i >= s_first[oct + 1]) oct++;
            const float4 k = T->cand[oct][i - s_first[oct]];
            scale = (int)k.w;
            if ((int)k.y != -1) keep = interp_one(T->ds[oct], k, peak_thresh, InitSigma, &res);
        }
        int slot = warp_append(keep, n_kp);
        if (keep && slot < kp_cap) { kp[slot] = res; kp_tag[slot] = (oct << 8) | scale; }
        // counters per (octave, scale): the lanes of a warp that kept a candidate of the same (octave, scale) share
        // one atomicAdd (tens of thousands of single increments on a handful of addresses serialise in L2)
        const int key = keep ? ((oct << 2) | scale) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (keep && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
            atomicAdd(&T->stage[oct][(scale - 1) * 3 + 1], __popc(peers));
            atomicAdd(T->n_kp_oct[oct], __popc(peers));
        }
    }
}

// grid-stride over the device-resident candidate count; survivors appended to kp / kp_scale
__global__ void __launch_bounds__(128) k_refine(DogStack D, const float4 *__restrict__ cand,
                                                 const int *__restrict__ n_cand, int cap, float peak_thresh,
                                                 float InitSigma, float4 *__restrict__ kp, int *__restrict__ kp_tag,
                                                 int kp_cap, int *__restrict__ n_kp, int *__restrict__ stage,
                                                 int octave, int *__restrict__ n_kp_oct) {
    const int n = min(*n_cand, cap);
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (n + stride - 1) / stride;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < rounds; it++, i += stride) {
        bool keep = false;
        float4 res = make_float4(-1.f, -1.f, -1.f, -1.f);
        int scale = 0;
        if (i < n) {
            float4 k = cand[i];
            scale = (int)k.w;
            if ((int)k.y != -1) keep = interp_one(D, k, peak_thresh, InitSigma, &res);
        }
        int slot = warp_append(keep, n_kp);
        if (keep) {
            if (slot < kp_cap) { kp[slot] = res; kp_tag[slot] = (octave << 8) | scale; }
            if (stage) atomicAdd(&stage[(scale - 1) * 3 + 1], 1);
            if (n_kp_oct) atomicAdd(n_kp_oct, 1);
        }
    }
}
