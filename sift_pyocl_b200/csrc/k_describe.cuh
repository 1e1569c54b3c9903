// k_describe.cuh -- 4x4x8 SIFT descriptor, bit-identical to the sequential reference kernel
// keypoints_cpu.cl:36-160 (CPU-variant semantics).
//
// The reference accumulates `hist[bin] += w` while scanning the (2R+1)^2 window in row-major order;
// fp32 addition is not associative, so every bin must see its contributions in exactly that order.
//
// Parallel scheme: 8 lanes (an "octet") per keypoint, 4 keypoints per warp.
//   * candidates: per window row, the j-interval that can pass the reference's (rx, cx) test is computed in
//     double with a rigorous bound on the fp32 evaluation error folded in; the candidates of all rows are
//     packed back to back and lane l of the octet takes candidates l, l + 8, ... (8 per pass);
//   * evaluation: each lane evaluates its candidate with the reference's exact fp32 expressions and stages, for
//     each of the 8 parity classes (row&1, col&1, ori&1), ONE (bin address, value) term: a sample feeds up to 8
//     bins (2 rows x 2 columns x 2 orientations of the trilinear interpolation) and those always differ in all
//     three parities.  Missing terms are +0.0 on the class's home bin (an exact no-op: every term is >= +0);
//   * commit: lane p owns parity class p and performs 8 unconditional `hist[addr] += value` steps per pass, in
//     sample order: all lanes busy, bins of one sample never collide, and each bin receives its contributions
//     in the reference's order.  (ori == 2*pi exactly puts both orientation terms into bin 0; the second one is
//     cweight * 0 = +0 there, so one add suffices.)
//   * finish: L2 normalisation / 0.2 clamp / renormalisation / x512 -> uint8; the two
//     sums of squares are accumulated sequentially by one lane of the octet (order matters).
// Also performs the host-side NaN filtering and record assembly of plan.py:546-565.
#pragma once
#include "common.cuh"
#include "k_keypoint.cuh"

#define DESC_WARPS 4  // warps per CTA -> 16 keypoints per CTA

#define DESC_MAXROWS 100  // window rows per keypoint handled by the interval table (iradius <= 49; default sigmas need 97)

struct DescRows {          // per-octet table of the non-empty window rows: only the j-interval that can be valid
    signed char i[DESC_MAXROWS];   // window row i (-iradius..iradius; |.| <= 49 for tabled windows)
    signed char jlo[DESC_MAXROWS]; // first candidate j of the row
    signed char jhi[DESC_MAXROWS]; // last candidate j of the row
};

// Staging area of one octet for one pass (8 samples): the evaluating lane s writes, for each of the 8 parity
// classes p = (row&1)<<2 | (col&1)<<1 | (ori&1), the ONE contribution of sample s to a bin of that class as
// (byte offset of the bin inside the octet's histogram, value).  A sample feeds at most 8 bins (2 rows x 2
// columns x 2 orientations of the trilinear interpolation) and those always differ in all three parities, so
// every class receives exactly one (possibly null) term per sample.  Null terms (neighbour cell outside the 4x4
// grid, sample rejected by the reference's tests) are stored as value +0.0 on the class's home bin: every
// histogram term is >= +0, so adding +0.0 is an exact no-op.
// Rows are 10 float2 apart (80 B) and octets 704 B apart: the 8 lanes of an octet then store their 16-byte
// chunks to 8 different bank quads, and the per-class 8-byte loads of two neighbouring octets do not collide.
struct __align__(16) DescStage {
    float2 e[8][10];
    float2 pad[8];
};

// One warp, 4 keypoints (octet g handles kp[g] when act is true for that octet).
// whist: the warp's 16 x 32 floats of histogram storage.  Octet g owns the bank group 8g..8g+7; bin (r, c, o)
// lives in row (o>>1) + 4*(c>>1) + 8*(r>>1), bank 8g + 4*(r&1) + 2*(c&1) + (o&1): the 8 lanes of an octet (one
// per parity class) and the 4 octets of the warp always hit 32 different banks -- every histogram access of the
// commit loop is a single conflict-free wavefront.
__device__ __forceinline__ void describe_octets(float *__restrict__ whist, DescRows &rows, DescStage &stage,
                                                bool act, const float4 k,
                                                const float *__restrict__ grad, const float *__restrict__ orim,
                                                int pitch, int grad_width, int grad_height, int octsize,
                                                uint8_t *out128) {
    const int lane = threadIdx.x & 31, l8 = lane & 7, obase = lane & 24;
    const unsigned omask = 0xffu << obase;  // lanes of my octet
    float *hist = whist + obase;            // bank group of my octet
#pragma unroll
    for (int r = 0; r < 16; r++) hist[32 * r + l8] = 0.0f;
    // keypoints_cpu.cl:55-61
    const float row = k.y / (float)octsize, col = k.x / (float)octsize, angle = k.w;
    const int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    const float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    const float spacing = k.z / (float)octsize * 3.0f;
    const double inv_spacing = div_prepare(spacing);  // x / spacing == div_by(x, inv_spacing), see common.cuh
    int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    const float drow = row - (float)irow, dcol = col - (float)icol;
    if (!(act && iradius >= 0)) iradius = -1;
    const int nrows = 2 * iradius + 1;  // 0 rows for an inactive octet
    const bool tabled = nrows <= DESC_MAXROWS;
    // ---- row table: the reference scans j = -R..R of every row and keeps the samples with rx, cx in (-1, 4)
    // whose pixel is inside the image (keypoints_cpu.cl:66-67); those form one j-interval per row.  A
    // conservative superset of it is computed here (real-valued bounds widened by a full sample on each side) so
    // that only candidate samples are evaluated; every evaluated sample still goes through the reference's exact
    // fp32 test.  Rows with an empty interval are dropped.
    int my_passes = 0, nrc = 0;  // passes (8 samples each) of this octet, number of table rows
    // untabled (enormous window, custom init_sigma): every row of the square that lies inside the image, full width
    const int u_r0 = max(0, iradius - irow), u_r1 = min(nrows - 1, iradius + grad_height - 1 - irow);
    const int u_jlo = max(-iradius, -icol), u_jhi = min(iradius, grad_width - 1 - icol);
    if (tabled) {
        // rx in (-1, 4)  <=>  |cs*i - sn*j - drow| < L = 2.5*spacing in exact arithmetic; the fp32 evaluation of
        // the reference (five roundings on magnitudes <= R + L) moves the left side by less than E.  Candidates
        // therefore satisfy |A - sn*j| < L + E and |B + cs*j| < L + E: exact j-intervals, computed in double.
        const double L = 2.5 * (double)spacing, sn = (double)sine, cs = (double)cosine;
        const double LE = L + 2e-6 * ((double)iradius + L + 2.0);
        const bool use_sn = fabs(sn) > 1e-9, use_cs = fabs(cs) > 1e-9;
        const double isn = 1.0 / sn, ics = 1.0 / cs;
        for (int r = l8; r < nrows; r += 8) {
            const int i = r - iradius;
            double lo = -(double)iradius, hi = (double)iradius;
            if (irow + i < 0 || irow + i >= grad_height) hi = lo - 1.0;  // row outside the image
            const double A = cs * (double)i - (double)drow;
            if (use_sn) {
                const double a = (A - LE) * isn, b = (A + LE) * isn;
                lo = fmax(lo, fmin(a, b) - 1e-6);
                hi = fmin(hi, fmax(a, b) + 1e-6);
            }
            const double B = sn * (double)i - (double)dcol;
            if (use_cs) {
                const double a = (-LE - B) * ics, b = (LE - B) * ics;
                lo = fmax(lo, fmin(a, b) - 1e-6);
                hi = fmin(hi, fmax(a, b) + 1e-6);
            }
            int jlo = (int)ceil(lo), jhi = (int)floor(hi);
            jlo = max(jlo, max(-iradius, -icol));
            jhi = min(jhi, min(iradius, grad_width - 1 - icol));
            if (jhi < jlo) { jlo = 1; jhi = 0; }  // empty (the real-valued bounds may not fit the table's type)
            rows.jlo[r] = (signed char)jlo;
            rows.jhi[r] = (signed char)jhi;
        }
        __syncwarp();
        if (l8 == 0) {  // drop the empty rows (in place: the write index never overtakes the read index)
            int n = 0, acc = 0;
            for (int r = 0; r < nrows; r++) {
                const int jlo = rows.jlo[r], jhi = rows.jhi[r];
                if (jhi >= jlo) {
                    rows.i[n] = (signed char)(r - iradius);
                    rows.jlo[n] = (signed char)jlo;
                    rows.jhi[n] = (signed char)jhi;
                    n++;
                    acc += jhi - jlo + 1;
                }
            }
            nrc = n;
            my_passes = (acc + 7) >> 3;  // the candidates of all rows are packed back to back, 8 per pass
        }
        __syncwarp();
    } else if (u_r1 >= u_r0 && u_jhi >= u_jlo) {
        nrc = u_r1 - u_r0 + 1;
        my_passes = (nrc * (u_jhi - u_jlo + 1) + 7) >> 3;
    }
    nrc = __shfl_sync(0xffffffffu, nrc, obase);
    my_passes = __shfl_sync(0xffffffffu, my_passes, obase);
    // warp-uniform trip count: the longest of the 4 octets
    const int passes_max = __reduce_max_sync(0xffffffffu, my_passes);
    // Per-lane cursor over the candidate samples in row-major order: lane l8 takes candidates l8, l8 + 8, ... of
    // the octet (table row rcur, column jcur, last column of that row jend, pixel offset of (row, j = 0)).  The
    // gradient / orientation values of the lane's sample of pass p+1 are requested before pass p is evaluated
    // and committed, so the L2 / DRAM gather latency overlaps the commit loop.
    int rcur = -1, jcur = l8, jend = -1, n_i = 0, n_j = 0;
    long rowoff = 0;
    bool n_in = false;
    float n_g = 0.0f, n_o = 0.0f;
    auto fetch = [&]() {
        while (jcur > jend && rcur < nrc) {  // into the next table row(s)
            const int over = jcur - jend - 1;
            rcur++;
            if (rcur < nrc) {
                if (tabled) {
                    n_i = rows.i[rcur];
                    jcur = rows.jlo[rcur] + over;
                    jend = rows.jhi[rcur];
                } else {
                    n_i = u_r0 + rcur - iradius;
                    jcur = u_jlo + over;
                    jend = u_jhi;
                }
                rowoff = (long)(irow + n_i) * pitch + icol;
            }
        }
        n_j = jcur;
        n_in = rcur < nrc;
        if (n_in) {
            n_g = grad[rowoff + n_j];
            n_o = orim[rowoff + n_j];
        }
        jcur += 8;
    };
    fetch();
    for (int p = 0; p < passes_max; p++) {
        const int i = n_i, j = n_j;
        const bool in_image = n_in;
        const float g_val = n_g, o_val = n_o;
        fetch();
        // terms of this lane's sample, indexed by the parity of the row / column / orientation bin they go to;
        // the defaults describe a null sample
        float rw_e = 0.0f, rw_o = 0.0f, cf_e = 0.0f, cf_o = 0.0f, ow_e = 0.0f, ow_o = 0.0f;
        int ra_e = 0, ra_o = 16, ca_e = 0, ca_o = 8, oa_e = 0, oa_o = 4;  // byte offsets in the bank group, home cell
        if (in_image) {
            const float rx = div_by((cosine * (float)i - sine * (float)j) - drow, inv_spacing) + 1.5f;
            const float cx = div_by((sine * (float)i + cosine * (float)j) - dcol, inv_spacing) + 1.5f;
            if (rx > -1.0f && rx < 4.0f && cx > -1.0f && cx < 4.0f) {
                const float er = rx - 1.5f, ec = cx - 1.5f;
                const float mag = g_val * cr_expf_neg(-0.125f * (er * er + ec * ec));
                float ori = o_val - angle;
                while (ori > 2.0f * SIFTB_M_PI_F) ori -= 2.0f * SIFTB_M_PI_F;
                while (ori < 0.0f) ori += 2.0f * SIFTB_M_PI_F;
                const float oval = (4.0f * ori) * SIFTB_M_1_PI_F;
                // keypoints_cpu.cl:77-79 `(int)((v >= 0.0f) ? v : v - 1.0f)`: for v in (-1, 4) that is floor(v)
                // (v - 1 lies in (-2, -1) for negative v, so the truncation gives -1); oval >= 0 always
                const int ri = __float2int_rd(rx);
                const int ci = __float2int_rd(cx);
                const int oi = (int)((oval >= 0.0f) ? oval : oval - 1.0f);
                const float rfrac = rx - (float)ri, cfrac = cx - (float)ci, ofrac = oval - (float)oi;
                if ((ri >= -1 && ri < 4 && oi >= 0 && oi <= 8 && rfrac >= 0.0f && rfrac <= 1.0f)) {
                    // rows ri (weight mag*(1-rfrac)) and ri+1 (mag*rfrac): the even one and the odd one
                    const float rw0 = mag * (1.0f - rfrac), rw1 = mag * rfrac;
                    const int rodd = ri & 1, r_e = ri + rodd, r_o = ri + 1 - rodd;
                    if ((unsigned)r_e < 4u) { rw_e = rodd ? rw1 : rw0; ra_e = 512 * r_e; }
                    if ((unsigned)r_o < 4u) { rw_o = rodd ? rw0 : rw1; ra_o = 512 * r_o - 496; }
                    const float cf0 = 1.0f - cfrac;
                    const int codd = ci & 1, c_e = ci + codd, c_o = ci + 1 - codd;
                    if ((unsigned)c_e < 4u) { cf_e = codd ? cfrac : cf0; ca_e = 256 * c_e; }
                    if ((unsigned)c_o < 4u) { cf_o = codd ? cf0 : cfrac; ca_o = 256 * c_o - 248; }
                    const float of0 = 1.0f - ofrac;
                    if (oi < 8) {
                        const int o1 = (oi + 1) & 7;  // oindex = oi + orr; if (oindex >= 8) oindex = 0
                        const bool oodd = oi & 1;
                        ow_e = oodd ? ofrac : of0;
                        ow_o = oodd ? of0 : ofrac;
                        oa_e = 64 * (oodd ? o1 : oi);
                        oa_o = 64 * (oodd ? oi : o1) - 60;
                    } else {
                        // oi == 8 <=> ori == 2*pi exactly (then oval == 8.0f and ofrac == 0): both terms go to
                        // bin 0, cweight*(1-ofrac) then cweight*ofrac = +0 -- the second add is a no-op
                        ow_e = of0;
                        oa_e = 0;
                    }
                }
            }
        }
        {
            // (rweight * c-factor) * o-factor, in the reference's multiplication order (keypoints_cpu.cl:93-104)
            const float cw_ee = rw_e * cf_e, cw_eo = rw_e * cf_o, cw_oe = rw_o * cf_e, cw_oo = rw_o * cf_o;
            float4 *dst = reinterpret_cast<float4 *>(&stage.e[l8][0]);
            dst[0] = make_float4(__int_as_float(ra_e + ca_e + oa_e), cw_ee * ow_e,
                                 __int_as_float(ra_e + ca_e + oa_o), cw_ee * ow_o);
            dst[1] = make_float4(__int_as_float(ra_e + ca_o + oa_e), cw_eo * ow_e,
                                 __int_as_float(ra_e + ca_o + oa_o), cw_eo * ow_o);
            dst[2] = make_float4(__int_as_float(ra_o + ca_e + oa_e), cw_oe * ow_e,
                                 __int_as_float(ra_o + ca_e + oa_o), cw_oe * ow_o);
            dst[3] = make_float4(__int_as_float(ra_o + ca_o + oa_e), cw_oo * ow_e,
                                 __int_as_float(ra_o + ca_o + oa_o), cw_oo * ow_o);
        }
        __syncwarp();
        // commit the 8 samples of the pass in sample order: lane (pr, pc, po) adds the one term of its class
#pragma unroll
        for (int sidx = 0; sidx < 8; sidx++) {
            const float2 t = stage.e[sidx][l8];
            float *bin = reinterpret_cast<float *>(reinterpret_cast<char *>(hist) + __float_as_int(t.x));
            *bin += t.y;
        }
        __syncwarp();
    }
    __syncwarp();
    // finish, keypoints_cpu.cl:127-160: each lane of the octet owns 16 consecutive descriptor entries
    // i = 16*l8 + q, i.e. r = l8>>1, c = 2*(l8&1) + (q>>3), o = q&7
    float *mine = hist + 32 * (4 * (l8 & 1) + 8 * (l8 >> 2)) + 4 * ((l8 >> 1) & 1);
#define DESC_QOFF(q) (32 * (((q) & 7) >> 1) + 2 * ((q) >> 3) + ((q) & 1))
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) v[q] = mine[DESC_QOFF(q)];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) mine[DESC_QOFF(q)] = v[q] * v[q];
    __syncwarp();
    // the sums of squares are sequential over i = 0..127 in the reference: one lane adds them in that order
    auto ordered_sum = [&]() {
        float acc = 0.0f;
        for (int rc = 0; rc < 16; rc++) {
            const float *cell = hist + 32 * (4 * ((rc & 3) >> 1) + 8 * (rc >> 3)) + 4 * ((rc >> 2) & 1) + 2 * (rc & 1);
#pragma unroll
            for (int o = 0; o < 8; o++) acc += cell[32 * (o >> 1) + (o & 1)];
        }
        return acc;
    };
    float norm = 0.0f;
    if (l8 == 0) norm = ordered_sum();
    norm = cr_rsqrtf(__shfl_sync(0xffffffffu, norm, obase));
    bool changed = false;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) {
        v[q] *= norm;
        if (v[q] > 0.2f) { v[q] = 0.2f; changed = true; }
        mine[DESC_QOFF(q)] = v[q] * v[q];
    }
    changed = (__ballot_sync(0xffffffffu, changed) & omask) != 0;
    __syncwarp();
    float norm2 = 0.0f;
    if (l8 == 0 && changed) norm2 = ordered_sum();
    norm2 = cr_rsqrtf(__shfl_sync(0xffffffffu, norm2, obase));
    if (changed) {
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] *= norm2;
    }
    if (act) {
        uint32_t w4[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            uint32_t pk = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float x = 512.0f * v[q4 * 4 + q];
                const int intval = (x != x) ? 0 : (int)x;
                pk |= (uint32_t)min(255, intval) << (8 * q);  // intval >= 0 here (hist >= 0)
            }
            w4[q4] = pk;
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(out128) + l8 * 4;  // 16-byte records offset: 4-B aligned
        dst[0] = w4[0]; dst[1] = w4[1]; dst[2] = w4[2]; dst[3] = w4[3];
    }
    __syncwarp();
}

// prefix of the per-octave record counts -> first record slot of every octave, and the total
__global__ void k_octave_offsets(const int *__restrict__ oct_valid, int n_oct, int *__restrict__ oct_offset,
                                 int *__restrict__ n_out, int *__restrict__ n_out_oct /* stride 4 */,
                                 const int *__restrict__ size_hist, int *__restrict__ size_start,
                                 int *__restrict__ n_order) {
    int acc = 0;
    for (int o = 0; o < n_oct; o++) {
        oct_offset[o] = acc;
        acc += oct_valid[o];
        n_out_oct[4 * o] = oct_valid[o];
    }
    *n_out = acc;
    acc = 0;  // processing order: largest windows first
    for (int b = DESC_CLASSES - 1; b >= 0; b--) {
        size_start[b] = acc;
        acc += size_hist[b];
    }
    *n_order = acc;  // keypoints k_orient accepted and stored (== entries k_size_order writes)
}

// order[] = keypoint indices sorted by descending descriptor-window size class (counting sort, unstable)
__global__ void __launch_bounds__(256) k_size_order(const float4 *__restrict__ kp, const int *__restrict__ kp_tag,
                                                     const int *__restrict__ n_base_p, const int *__restrict__ n_extra_p,
                                                     int cap, const int *__restrict__ size_start,
                                                     int *__restrict__ size_fill, int *__restrict__ order) {
    const int n = min(min(*n_base_p, cap) + *n_extra_p, cap);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 k = kp[i];
        if (!(k.y >= 0.0f)) continue;  // rows k_orient skipped (never counted in size_hist)
        const int b = desc_size_class(k.z, 1 << (kp_tag[i] >> 8));
        const int pos = size_start[b] + atomicAdd(&size_fill[b], 1);
        if (pos < cap) order[pos] = i;
    }
}

// Pipeline form: octets fetch keypoints of ALL octaves from a work queue; rows with NaN are dropped
// (plan.py:546-550); the survivors of octave o go to out[oct_offset[o] + ...], i.e. the output is grouped by
// octave in octave order like the reference's concatenation (plan.py:555-565).
__global__ void __launch_bounds__(DESC_WARPS * 32, 7) k_describe(OctTable T, const float4 *__restrict__ kp,
                                                               const int *__restrict__ kp_tag,
                                                               const int *__restrict__ n_order_p, int cap,
                                                               KpRecord *__restrict__ out, int out_cap,
                                                               const int *__restrict__ oct_offset,
                                                               int *__restrict__ oct_fill, int *__restrict__ queue,
                                                               const int *__restrict__ order) {
    __shared__ float s_hist[DESC_WARPS][16 * 32];
    __shared__ DescRows s_rows[DESC_WARPS * 4];
    __shared__ DescStage s_stage[DESC_WARPS * 4];
    const int lane = threadIdx.x & 31, l8 = lane & 7, obase = lane & 24;
    float *hist = s_hist[threadIdx.x >> 5];
    const int n = min(*n_order_p, cap);
    // dynamic work queue: every warp fetches 4 keypoints at a time, so warps with small windows simply fetch
    // more often and the last wave is not quantised to the grid size
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(queue, 4);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        bool act = base + (lane >> 3) < n;
        const int gid0 = act ? order[base + (lane >> 3)] : 0;
        float4 k = make_float4(0.f, 0.f, 1.f, 0.f);
        int sc = 1, oct = 0;
        if (act) {
            k = kp[gid0];
            const int tag = kp_tag[gid0];
            sc = tag & 0xff;
            oct = tag >> 8;
            const float s = ((k.x + k.y) + k.z) + k.w;
            act = (k.y >= 0.0f) && !(s != s);
        }
        int slot = 0;
        if (act && l8 == 0) slot = oct_offset[oct] + atomicAdd(&oct_fill[oct], 1);
        slot = __shfl_sync(0xffffffffu, slot, obase);
        if (slot >= out_cap) act = false;
        KpRecord *o = out + (act ? slot : 0);
        if (act && l8 == 0) { o->x = k.x; o->y = k.y; o->scale = k.z; o->angle = k.w; }
        describe_octets(hist, s_rows[threadIdx.x >> 3], s_stage[threadIdx.x >> 3], act, k, T.grad[oct][sc - 1],
                        T.ori[oct][sc - 1], T.pitch[oct], T.w[oct], T.h[oct], T.octsize[oct], o->desc);
    }
}

// Stage-hook form: desc[i] for every input row (no filtering), single gradient plane
__global__ void __launch_bounds__(DESC_WARPS * 32, 8) k_describe_rows(const float *__restrict__ grad,
                                                                    const float *__restrict__ ori, int pitch, int w,
                                                                    int h, const float4 *__restrict__ kp, int n,
                                                                    int octsize, uint8_t *__restrict__ desc) {
    __shared__ float s_hist[DESC_WARPS][16 * 32];
    __shared__ DescRows s_rows[DESC_WARPS * 4];
    __shared__ DescStage s_stage[DESC_WARPS * 4];
    float *hist = s_hist[threadIdx.x >> 5];
    const int noct = (gridDim.x * blockDim.x) >> 3;
    const int rounds = (n + noct - 1) / noct;
    int gid0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    for (int it = 0; it < rounds; it++, gid0 += noct) {
        bool act = gid0 < n;
        float4 k = make_float4(0.f, 0.f, 1.f, 0.f);
        if (act) {
            k = kp[gid0];
            act = k.y >= 0.0f;
        }
        describe_octets(hist, s_rows[threadIdx.x >> 3], s_stage[threadIdx.x >> 3], act, k, grad, ori, pitch, w, h, octsize,
                        desc + 128L * (act ? gid0 : 0));
    }
}
