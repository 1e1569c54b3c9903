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
//   * evaluation: each lane evaluates its candidate with the reference's exact fp32 expressions -- straight-line
//     code: a rejected sample is carried along with magnitude zero, and the histogram has a guard ring of cells so
//     that the trilinear neighbours need no range test -- and stages, for each of the 8 parity classes (row&1,
//     col&1, ori&1), ONE (bin address, value) term: a sample feeds 8 bins (2 rows x 2 columns x 2 orientations of
//     the trilinear interpolation) and those always differ in all three parities.  Terms of rejected samples are
//     +-0.0 (an exact no-op: every histogram term is >= +0);
//   * commit: lane p owns parity class p and performs 8 unconditional `hist[addr] += value` steps per pass, in
//     sample order: all lanes busy, bins of one sample never collide, and each bin receives its contributions
//     in the reference's order.  (ori == 2*pi exactly puts both orientation terms into bin 0; the second one is
//     cweight * 0 = +0 there, so one add suffices.)
//   * finish: L2 normalisation / 0.2 clamp / renormalisation / x512 -> uint8; the two
//     sums of squares are accumulated sequentially by one lane of the octet (order matters).
// The same code, compiled with GPUVAR, reproduces keypoints_gpu2.cl instead (fixed [-64, 64)^2 window, integer
// accumulation of (uint)(100000 * term), halving-tree norms, uchar wrap): see describe_octets.
// Also performs the host-side NaN filtering and record assembly of plan.py:546-565.
#pragma once
#include "common.cuh"
#include "k_keypoint.cuh"

#define DESC_WARPS 4  // warps per CTA -> 16 keypoints per CTA

#define DESC_MAXROWS 101  // window rows per keypoint handled by the interval table (default sigmas: 97, GPU variant 101)

// Per-octet table of the non-empty window rows: only the j-interval that can be valid.  Three signed bytes per row
// (window row i, -66..63 for tabled windows; first and last candidate j): 306 bytes per octet -- together with the
// swizzled stage below this lets SEVEN CTAs share an SM's shared memory.
struct DescRows {
    signed char i[DESC_MAXROWS], jlo[DESC_MAXROWS], jhi[DESC_MAXROWS];
};

// Staging area of one octet for one pass (8 samples): the evaluating lane s writes, for each of the 8 parity
// classes p = (row&1)<<2 | (col&1)<<1 | (ori&1), the ONE contribution of sample s to a bin of that class as
// (shared-memory byte address of the bin, value).  A sample feeds 8 bins (2 rows x 2 columns x 2 orientations of
// the trilinear interpolation) and those always differ in all three parities, so every class receives exactly one
// term per sample.  Rejected samples stage value +-0.0 (every histogram term is >= +0, so adding a zero of either
// sign is an exact no-op).
// Layout: row s = 64 bytes = four 16-byte chunks (chunk c = classes 2c, 2c + 1), stored at chunk position
// c ^ ((s >> 1) & 3): the 8 lanes of an octet then write their chunk c to 8 different bank quads (one wavefront per
// octet) although the rows are only 64 bytes apart.  The four stages of a warp sit at byte offsets 0, 576, 1088, 1664
// of a 2176-byte block: octets 0 / 1 and 2 / 3 -- the pairs that share a half-warp, the unit of a 64-bit shared load
// -- are 64 bytes apart modulo 128, so the 64-byte rows they read together cover 32 different banks.
struct __align__(16) DescStage {
    float2 e[8][8];
};
#define DESC_STAGE_WARP_BYTES 2176
__device__ __forceinline__ int desc_stage_offset(int octet_in_warp) {  // 0, 576, 1088, 1664
    return 544 * octet_in_warp + 32 * (octet_in_warp & 1);
}
// one raw shared buffer per CTA, carved by hand (no alignment padding between the arrays: with 32 288 bytes seven
// CTAs fit into an SM's 228 KB, each CTA also reserving 1 KB)
#define DESC_SMEM_HIST (DESC_WARPS * DESC_HROWS * 32 * 4)
#define DESC_SMEM_STAGE (DESC_WARPS * DESC_STAGE_WARP_BYTES)
#define DESC_SMEM_EXP 256
#define DESC_SMEM_ROWS (DESC_WARPS * 4 * 3 * DESC_MAXROWS)
#define DESC_SMEM_BYTES (DESC_SMEM_HIST + DESC_SMEM_STAGE + DESC_SMEM_EXP + DESC_SMEM_ROWS)
__device__ __forceinline__ int desc_stage_chunk(int s, int c) { return 4 * s + (c ^ ((s >> 1) & 3)); }  // float4 index

// Histogram storage of one warp: the 4 x 4 x 8 bins of the reference plus a GUARD RING -- cells r, c in -1..4,
// stored as r0 = r + 1, c0 = c + 1 in 0..5 -- so that the trilinear neighbours of a sample never need a range test
// (the reference's `if (rindex >= 0 && rindex < 4)` etc., keypoints_cpu.cl:91-101, become writes to cells nobody
// reads).  Octet g owns the bank group 8g..8g+7; bin (r0, c0, o) lives in row (o>>1) + 4*(c0>>1) + 12*(r0>>1),
// bank 8g + 4*(r0&1) + 2*(c0&1) + (o&1): the 8 lanes of an octet (one per parity class) and the 4 octets of the
// warp always hit 32 different banks -- every histogram access of the commit loop is a single conflict-free
// wavefront.
#define DESC_HROWS 36
__device__ __forceinline__ int desc_bin(int r0, int c0, int o) {  // float index inside the octet's bank group
    return 32 * ((o >> 1) + 4 * (c0 >> 1) + 12 * (r0 >> 1)) + 4 * (r0 & 1) + 2 * (c0 & 1) + (o & 1);
}

// exp(x) for x in [-16, 0] without the range check of cr_expf_neg (callers clamp): the same arithmetic, hence the
// same result (exhaustively verified, tools/exp_check.c).  The double constants are read from the constant bank as
// direct DFMA operands (as literals ptxas re-materialises each with two moves per use) and the 32-entry table from
// shared memory (32-bit address, no 64-bit pointer arithmetic per sample).
__constant__ double c_exp_k[8] = {0x1.71547652b82fep+5, 0x1.8p52, -0x1.62e42feep-6, -0x1.a39ef358p-38,
                                  1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};
__device__ __forceinline__ float cr_expf_neg_inrange(float xf, const double *__restrict__ s_tab) {
    const double x = (double)xf;
    const double z = fma(x, c_exp_k[0], c_exp_k[1]);  // k = rint(x * 32/ln2) in the low word
    const int k = __double2loint(z);
    const double kd = z - c_exp_k[1];
    double r = fma(kd, c_exp_k[2], x);
    r = fma(kd, c_exp_k[3], r);
    double q = fma(r, c_exp_k[4], c_exp_k[5]);
    q = fma(r, q, c_exp_k[6]);
    q = fma(r, q, c_exp_k[7]);
    q = fma(r, q, 0.5);
    const double p = fma(r * r, q, r);
    const double t = s_tab[k & 31];
    const double s = __hiloint2double(__double2hiint(t) + ((k >> 5) << 20), __double2loint(t));
    return (float)fma(s, p, s);
}

// One warp, 4 keypoints (octet g handles kp[g] when act is true for that octet).
// whist: the warp's DESC_HROWS x 32 floats of histogram storage.
// ANY_ANGLE: the keypoint angle may lie anywhere (stage hook fed by a caller); in the pipeline k_orient guarantees
// [-pi, pi] up to one rounding.  s_exp: the CTA's shared copy of c_exp_t32.
// GPUVAR: the semantics of keypoints_gpu2.cl instead of keypoints_cpu.cl (devicetype "GPU" in the
// reference): the window is [-64, 64)^2 whatever the keypoint's size, every trilinear term is accumulated as
// (uint)(100000 * term) by integer adds (order-free, wraps modulo 2^32), histogram = (float)sum * 0.00001f, the sums
// of squares come from the reference's 128 -> 2 halving tree, and the result is cast to uchar BEFORE MIN(255, .).
// The sample evaluation, the candidate table, the parity-class staging and the commit loop are shared.
template <bool ANY_ANGLE, bool GPUVAR>
__device__ __forceinline__ void describe_octets(float *__restrict__ whist, DescRows &rows, DescStage &stage,
                                                const double *__restrict__ s_exp, bool act, const float4 k,
                                                const float2 *__restrict__ go, int pitch, int grad_width, int grad_height, int octsize,
                                                uint8_t *out128) {
    const int lane = threadIdx.x & 31, l8 = lane & 7, obase = lane & 24;
    const unsigned omask = 0xffu << obase;  // lanes of my octet
    float *hist = whist + obase;            // bank group of my octet
#pragma unroll
    for (int r = 0; r < DESC_HROWS; r++) hist[32 * r + l8] = 0.0f;
    // keypoints_cpu.cl:55-61
    const float row = k.y / (float)octsize, col = k.x / (float)octsize, angle = k.w;
    const int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    const float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    const float spacing = k.z / (float)octsize * 3.0f;
    const double inv_spacing = div_prepare(spacing);  // x / spacing == div_by(x, inv_spacing), see common.cuh
    int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    const float drow = row - (float)irow, dcol = col - (float)icol;
    if (!(act && iradius >= 0)) iradius = -1;
    // window rows / columns that are scanned: [-iradius, iradius] (keypoints_cpu.cl:63-64); the GPU variant scans
    // [-64, 64) -- only the rows and columns within iradius + 2 can hold a sample that passes the (rx, cx) test
    // (the test admits |.| < 2.5 * sqrt(2) * spacing < iradius + 1), so the table covers those
    const int wlo = GPUVAR ? -min(iradius + 2, 64) : -iradius;
    const int whi = GPUVAR ? min(iradius + 2, 63) : iradius;
    const int nrows = iradius < 0 ? 0 : whi - wlo + 1;  // 0 rows for an inactive octet
    const bool tabled = nrows <= DESC_MAXROWS;
    // ---- row table: the reference scans j = -R..R of every row and keeps the samples with rx, cx in (-1, 4)
    // whose pixel is inside the image (keypoints_cpu.cl:66-67); those form one j-interval per row.  A
    // conservative superset of it is computed here (real-valued bounds widened by a rigorous bound on the fp32
    // evaluation error) so that only candidate samples are evaluated; every evaluated sample still goes through
    // the reference's exact fp32 test.  Rows with an empty interval are dropped.
    int my_passes = 0, nrc = 0;  // passes (8 samples each) of this octet, number of table rows
    // untabled (enormous window, custom init_sigma): every row of the square that lies inside the image, full width
    const int u_r0 = max(0, -wlo - irow), u_r1 = min(nrows - 1, -wlo + grad_height - 1 - irow);
    const int u_jlo = max(wlo, -icol), u_jhi = min(whi, grad_width - 1 - icol);
    if (tabled) {
        // rx in (-1, 4)  <=>  |cs*i - sn*j - drow| < L = 2.5*spacing in exact arithmetic; the fp32 evaluation of
        // the reference (five roundings on magnitudes <= R + L) moves the left side by less than E.  Candidates
        // therefore satisfy |A - sn*j| < L + E and |B + cs*j| < L + E: exact j-intervals, computed in double.
        const double L = 2.5 * (double)spacing, sn = (double)sine, cs = (double)cosine;
        const double LE = L + 2e-6 * ((double)iradius + L + 2.0);
        const bool use_sn = fabs(sn) > 1e-9, use_cs = fabs(cs) > 1e-9;
        const double isn = 1.0 / sn, ics = 1.0 / cs;
        for (int r = l8; r < nrows; r += 8) {
            const int i = r + wlo;
            double lo = (double)wlo, hi = (double)whi;
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
            jlo = max(jlo, max(wlo, -icol));
            jhi = min(jhi, min(whi, grad_width - 1 - icol));
            if (jhi < jlo) { jlo = 1; jhi = 0; }  // empty (the real-valued bounds may not fit the table's type)
            rows.i[r] = (signed char)i;
            rows.jlo[r] = (signed char)jlo;
            rows.jhi[r] = (signed char)jhi;
        }
    }
    // (the warp-level barriers sit outside the per-octet branches: the four octets of a warp may take different ones)
    __syncwarp();
    if (tabled) {
        if (l8 == 0) {  // drop the empty rows (in place: the write index never overtakes the read index)
            int n = 0, acc = 0;
            for (int r = 0; r < nrows; r++) {
                const int jlo = rows.jlo[r], jhi = rows.jhi[r];
                if (jhi >= jlo) {
                    rows.i[n] = rows.i[r];
                    rows.jlo[n] = (signed char)jlo;
                    rows.jhi[n] = (signed char)jhi;
                    n++;
                    acc += jhi - jlo + 1;
                }
            }
            nrc = n;
            my_passes = (acc + 7) >> 3;  // the candidates of all rows are packed back to back, 8 per pass
        }
    } else if (u_r1 >= u_r0 && u_jhi >= u_jlo) {
        nrc = u_r1 - u_r0 + 1;
        my_passes = (nrc * (u_jhi - u_jlo + 1) + 7) >> 3;
    }
    __syncwarp();
    nrc = __shfl_sync(0xffffffffu, nrc, obase);
    my_passes = __shfl_sync(0xffffffffu, my_passes, obase);
    // warp-uniform trip count: the longest of the 4 octets
    const int passes_max = __reduce_max_sync(0xffffffffu, my_passes);
    // Per-lane cursor over the candidate samples in row-major order: lane l8 takes candidates l8, l8 + 8, ... of
    // the octet (table row rcur, column jcur, last column of that row jend).  Per row the cursor keeps the pixel
    // offset of (row, j = 0) and the two row-constant products of the rotation, cosine*i and sine*i
    // (keypoints_cpu.cl:63-65 evaluates them per sample; same operands, same fp32 products).  The gradient /
    // orientation values of the lane's sample of pass p+1 are requested before pass p is evaluated and committed,
    // so the L2 / DRAM gather latency overlaps the commit loop.
    // (two sets of "next sample" values, A and B, used alternately by the pass loop unrolled twice by hand: with one
    // set the compiler copies next -> current at the end of the iteration and that copy waits for the load)
    struct Sample {
        int i, j;
        float g, o;
    };
    int rcur = -1, jcur = l8, jend = -1, n_i = 0, rowoff = 0;
    auto fetch = [&](Sample &n) {
        while (jcur > jend && rcur < nrc) {  // into the next table row(s)
            const int over = jcur - jend - 1;
            rcur++;
            if (rcur < nrc) {
                if (tabled) {
                    n_i = rows.i[rcur];
                    jcur = rows.jlo[rcur] + over;
                    jend = rows.jhi[rcur];
                } else {
                    n_i = u_r0 + rcur + wlo;
                    jcur = u_jlo + over;
                    jend = u_jhi;
                }
            }
        }
        n.i = n_i;
        n.j = jcur;
        n.g = 0.0f;  // past the end of the window: a sample of zero gradient, i.e. eight +-0 terms
        n.o = 0.0f;
        if (rcur < nrc) {
            rowoff = (irow + n_i) * pitch + icol;
            const float2 v = ldg_f2_here(go + (rowoff + n.j));
            n.g = v.x;
            n.o = v.y;
        }
        jcur += 8;
    };
    // shared-memory byte address of bin (0, 0, 0) of my octet: the stage carries absolute addresses
    const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist);
    auto pass = [&](const Sample &cur) {
        const float fi = (float)cur.i, fj = (float)cur.j;
        const float g_val = cur.g, o_val = cur.o;
        // ---- evaluate (straight-line code: a rejected sample is carried along with magnitude zero) -------------
        // keypoints_cpu.cl:63-67
        const float rx = div_by((cosine * fi - sine * fj) - drow, inv_spacing) + 1.5f;
        const float cx = div_by((sine * fi + cosine * fj) - dcol, inv_spacing) + 1.5f;
        const bool ok = rx > -1.0f && rx < 4.0f && cx > -1.0f && cx < 4.0f;
        // :69-70 mag = grad * exp(-0.125 * ((rx-1.5)^2 + (cx-1.5)^2)); the argument of an accepted sample lies in
        // [-1.5625, 0]; the clamp only keeps rejected samples inside the fast exp's range
        const float er = rx - 1.5f, ec = cx - 1.5f;
        const float earg = fmaxf(-0.125f * (er * er + ec * ec), -16.0f);
        const float mag = (ok ? g_val : 0.0f) * cr_expf_neg_inrange(earg, s_exp);
        // :71-75 orientation relative to the keypoint, wrapped into [0, 2 pi] (`>` : exactly 2 pi stays).  The
        // planes hold atan2 values in [-pi, pi] and the keypoint angle lies in [-pi, pi] up to one rounding, so
        // the reference's while loops run at most once / twice
        float ori = o_val - angle;
        if (ori > 2.0f * SIFTB_M_PI_F) ori -= 2.0f * SIFTB_M_PI_F;
        if (ori < 0.0f) ori += 2.0f * SIFTB_M_PI_F;
        if (ori < 0.0f) ori += 2.0f * SIFTB_M_PI_F;
        if (ANY_ANGLE && (ori > 2.0f * SIFTB_M_PI_F || ori < 0.0f)) {  // caller-supplied angle far outside [-pi, pi]
            while (ori > 2.0f * SIFTB_M_PI_F) ori -= 2.0f * SIFTB_M_PI_F;
            while (ori < 0.0f) ori += 2.0f * SIFTB_M_PI_F;
        }
        const float oval = (4.0f * ori) * SIFTB_M_1_PI_F;
        // :77-85 integer cells and fractions.  `(int)((v >= 0) ? v : v - 1)` is floor(v) for v in (-1, 4) (v - 1
        // lies in (-2, -1) for negative v, so the truncation gives -1); oval >= 0.  The reference's guards
        // (ri in [-1, 4), oi in [0, 8], rfrac in [0, 1]) always hold for an accepted sample; a rejected one is
        // evaluated at rx = cx = 0 so that its (zero) terms land inside the histogram.
        const float rxs = ok ? rx : 0.0f, cxs = ok ? cx : 0.0f;
        const int ri = __float2int_rd(rxs), ci = __float2int_rd(cxs), oi = (int)oval;
        const float rfrac = rxs - (float)ri, cfrac = cxs - (float)ci, ofrac = oval - (float)oi;
        // :87-104 the 2 x 2 x 2 trilinear terms, sorted by the parity of the cell they go to.  Rows r0 = ri + 1
        // and r0 + 1 (guard-ring coordinates): the even one is (r0 + 1) & ~1, the odd one r0 | 1; if r0 is odd
        // the even row is the upper neighbour and gets weight rfrac.  Same for the columns.  Orientation bins oi
        // and oi + 1 wrap modulo 8 (`if (oindex >= 8) oindex = 0`); oi == 8 means ori == 2 pi exactly, i.e.
        // ofrac == 0: its second term is cweight * 0 and may go to any bin.
        const float rw0 = mag * (1.0f - rfrac), rw1 = mag * rfrac;
        const bool rodd = !(ri & 1), codd = !(ci & 1), oodd = oi & 1;  // r0 = ri + 1 is odd when ri is even
        const float rw_e = rodd ? rw1 : rw0, rw_o = rodd ? rw0 : rw1;
        const float cf0 = 1.0f - cfrac;
        const float cf_e = codd ? cfrac : cf0, cf_o = codd ? cf0 : cfrac;
        const float of0 = 1.0f - ofrac;
        const float ow_e = oodd ? ofrac : of0, ow_o = oodd ? of0 : ofrac;
        // byte offsets of the cells: row index * 128 + bank * 4 (desc_bin), split per dimension
        const unsigned ra_e = hist_sa + 768u * (unsigned)((ri + 2) & ~1), ra_o = hist_sa + 16u + 768u * (unsigned)((ri + 1) & ~1);
        const unsigned ca_e = 256u * (unsigned)((ci + 2) & ~1), ca_o = 8u + 256u * (unsigned)((ci + 1) & ~1);
        const unsigned oa_e = 64u * (unsigned)((oi + 1) & 6), oa_o = 64u * (unsigned)((oi | 1) & 7) - 60u;
        {
            // (rweight * c-factor) * o-factor, in the reference's multiplication order (keypoints_cpu.cl:93-104)
            const float cw_ee = rw_e * cf_e, cw_eo = rw_e * cf_o, cw_oe = rw_o * cf_e, cw_oo = rw_o * cf_o;
            // GPU variant: the term enters the histogram as (uint)(100000 * term) (keypoints_gpu2.cl:189-191); its
            // bit pattern travels through the stage in place of the float
            auto term = [](float t) { return GPUVAR ? __uint_as_float((unsigned)(100000.0f * t)) : t; };
            float4 *dst = reinterpret_cast<float4 *>(&stage.e[0][0]);
            dst[desc_stage_chunk(l8, 0)] = make_float4(__uint_as_float(ra_e + ca_e + oa_e), term(cw_ee * ow_e),
                                 __uint_as_float(ra_e + ca_e + oa_o), term(cw_ee * ow_o));
            dst[desc_stage_chunk(l8, 1)] = make_float4(__uint_as_float(ra_e + ca_o + oa_e), term(cw_eo * ow_e),
                                 __uint_as_float(ra_e + ca_o + oa_o), term(cw_eo * ow_o));
            dst[desc_stage_chunk(l8, 2)] = make_float4(__uint_as_float(ra_o + ca_e + oa_e), term(cw_oe * ow_e),
                                 __uint_as_float(ra_o + ca_e + oa_o), term(cw_oe * ow_o));
            dst[desc_stage_chunk(l8, 3)] = make_float4(__uint_as_float(ra_o + ca_o + oa_e), term(cw_oo * ow_e),
                                 __uint_as_float(ra_o + ca_o + oa_o), term(cw_oo * ow_o));
        }
        __syncwarp();
        // commit the 8 samples of the pass in sample order: lane (pr, pc, po) adds the one term of its class
        float2 term[8];
#pragma unroll
        for (int sidx = 0; sidx < 8; sidx++)  // class l8 of sample sidx: chunk l8 >> 1 (swizzled), half l8 & 1
            term[sidx] = reinterpret_cast<const float2 *>(&stage.e[0][0])[2 * desc_stage_chunk(sidx, l8 >> 1) + (l8 & 1)];
#pragma unroll
        for (int sidx = 0; sidx < 8; sidx++) {
            const unsigned sa = __float_as_uint(term[sidx].x);
            if (GPUVAR) {
                unsigned u;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(sa));
                u += __float_as_uint(term[sidx].y);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa), "r"(u) : "memory");
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa));
                v += term[sidx].y;
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa), "f"(v) : "memory");
            }
        }
        __syncwarp();
    };
    Sample sa, sb;
    fetch(sa);
    for (int p = 0; p < passes_max; p += 2) {
        fetch(sb);
        pass(sa);
        if (p + 1 >= passes_max) break;  // warp-uniform
        fetch(sa);
        pass(sb);
    }
    __syncwarp();
    // finish, keypoints_cpu.cl:127-160: each lane of the octet owns 16 consecutive descriptor entries
    // i = 16*l8 + q, i.e. r = l8>>1, c = 2*(l8&1) + (q>>3), o = q&7  (guard-ring coordinates r + 1, c + 1)
    const int fr0 = (l8 >> 1) + 1, fc0 = 2 * (l8 & 1) + 1;
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const float raw = hist[desc_bin(fr0, fc0 + (q >> 3), q & 7)];
        v[q] = GPUVAR ? (float)__float_as_uint(raw) * 0.00001f : raw;  // keypoints_gpu2.cl:205-209
    }
    __syncwarp();
    bool changed = false;
    if (GPUVAR) {
        // keypoints_gpu2.cl:211-272: hist2[l] = h[l]^2, then hist2[l] += hist2[l + half] for half = 64 .. 2 and
        // rsqrt(hist2[1] + hist2[0]).  The octet's bank group serves as the linear scratch array hist2[128].
        auto scratch = [&](int i) -> float & { return hist[32 * (i >> 3) + (i & 7)]; };
        auto tree_norm = [&]() {
#pragma unroll
            for (int q = 0; q < 16; q++) scratch(16 * l8 + q) = v[q] * v[q];
            __syncwarp();
#pragma unroll
            for (int per = 8; per >= 1; per >>= 1) {  // halves 64, 32, 16, 8: `per` sums per lane
                const int half = 8 * per;
#pragma unroll
                for (int t = 0; t < per; t++) scratch(l8 * per + t) += scratch(l8 * per + t + half);
                __syncwarp();
            }
            if (l8 < 4) scratch(l8) += scratch(l8 + 4);
            __syncwarp();
            if (l8 < 2) scratch(l8) += scratch(l8 + 2);
            __syncwarp();
            const float nrm = cr_rsqrtf(scratch(1) + scratch(0));
            __syncwarp();
            return nrm;
        };
        const float norm = tree_norm();
#pragma unroll
        for (int q = 0; q < 16; q++) {
            v[q] *= norm;
            if (v[q] > 0.2f) { v[q] = 0.2f; changed = true; }
        }
        changed = (__ballot_sync(0xffffffffu, changed) & omask) != 0;
        // all 32 lanes run the second tree (it contains warp-level barriers); only octets that clamped use it
        const float norm2 = tree_norm();
        if (changed) {
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] *= norm2;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) hist[desc_bin(fr0, fc0 + (q >> 3), q & 7)] = v[q] * v[q];
        __syncwarp();
        // the sums of squares are sequential over i = 0..127 in the reference: one lane adds them in that order
        auto ordered_sum = [&]() {
            float acc = 0.0f;
            for (int rc = 0; rc < 16; rc++) {
                const float *cell = hist + desc_bin((rc >> 2) + 1, (rc & 3) + 1, 0);
#pragma unroll
                for (int o = 0; o < 8; o++) acc += cell[32 * (o >> 1) + (o & 1)];
            }
            return acc;
        };
        float norm = 0.0f;
        if (l8 == 0) norm = ordered_sum();
        norm = cr_rsqrtf(__shfl_sync(0xffffffffu, norm, obase));
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) {
            v[q] *= norm;
            if (v[q] > 0.2f) { v[q] = 0.2f; changed = true; }
            hist[desc_bin(fr0, fc0 + (q >> 3), q & 7)] = v[q] * v[q];
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
                // keypoints_cpu.cl:157-159 clamps; keypoints_gpu2.cl:281 casts to uchar first (values >= 256 wrap)
                pk |= (uint32_t)(GPUVAR ? (intval & 0xff) : min(255, intval)) << (8 * q);  // intval >= 0 (hist >= 0)
            }
            w4[q4] = pk;
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(out128) + l8 * 4;  // 16-byte records offset: 4-B aligned
        dst[0] = w4[0]; dst[1] = w4[1]; dst[2] = w4[2]; dst[3] = w4[3];
    }
    __syncwarp();
}

// order[] = keypoint indices sorted by descending descriptor-window size class (counting sort, unstable), plus --
// by block 0 -- the prefix of the per-octave record counts: first record slot of every octave, and the totals.
// Every block derives the class offsets from the histogram k_orient filled (64 adds; cheaper than a launch of its
// own); lanes of a warp that hold the same class share one atomicAdd.
__global__ void __launch_bounds__(256) k_size_order(const float4 *__restrict__ kp, const int *__restrict__ kp_tag,
                                                     const int *__restrict__ n_base_p, const int *__restrict__ n_extra_p,
                                                     int cap, const int *__restrict__ size_hist,
                                                     int *__restrict__ size_fill, int *__restrict__ order,
                                                     int *__restrict__ n_order, const int *__restrict__ oct_valid,
                                                     int n_oct, int *__restrict__ oct_offset, int *__restrict__ n_out,
                                                     int *__restrict__ n_out_oct /* stride 4 */) {
    __shared__ int s_start[DESC_CLASSES];
    if (threadIdx.x == 0) {
        int acc = 0;  // processing order: largest windows first
        for (int b = DESC_CLASSES - 1; b >= 0; b--) {
            s_start[b] = acc;
            acc += size_hist[b];
        }
        if (blockIdx.x == 0) {
            *n_order = acc;  // keypoints k_orient accepted and stored (== entries written below)
            acc = 0;
            for (int o = 0; o < n_oct; o++) {
                oct_offset[o] = acc;
                acc += oct_valid[o];
                n_out_oct[4 * o] = oct_valid[o];
            }
            *n_out = acc;
        }
    }
    __syncthreads();
    const int n = min(min(*n_base_p, cap) + *n_extra_p, cap);
    const int stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {  // warp-uniform trip count
        const int i = i0 + lane;
        int b = -1;
        if (i < n) {
            const float4 k = kp[i];
            if (k.y >= 0.0f) b = desc_size_class(k.z, 1 << (kp_tag[i] >> 8));  // rows k_orient skipped are not ordered
        }
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (lane == leader && b >= 0) pos = atomicAdd(&size_fill[b], __popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (b >= 0) {
            pos += s_start[b] + __popc(peers & lanemask_lt());
            if (pos < cap) order[pos] = i;
        }
    }
}

// Pipeline form: octets fetch keypoints of ALL octaves from a work queue; rows with NaN are dropped
// (plan.py:546-550); the survivors of octave o go to out[oct_offset[o] + ...], i.e. the output is grouped by
// octave in octave order like the reference's concatenation (plan.py:555-565).
template <bool GPUVAR>
__global__ void __launch_bounds__(DESC_WARPS * 32, 7) k_describe(OctTable T, const float4 *__restrict__ kp,
                                                               const int *__restrict__ kp_tag,
                                                               const int *__restrict__ n_order_p, int cap,
                                                               KpRecord *__restrict__ out, int out_cap,
                                                               const int *__restrict__ oct_offset,
                                                               int *__restrict__ oct_fill, int *__restrict__ queue,
                                                               const int *__restrict__ order) {
    __shared__ __align__(16) unsigned char s_raw[DESC_SMEM_BYTES];
    float *hist = reinterpret_cast<float *>(s_raw) + (threadIdx.x >> 5) * (DESC_HROWS * 32);
    DescStage &my_stage = *reinterpret_cast<DescStage *>(s_raw + DESC_SMEM_HIST + (threadIdx.x >> 5) * DESC_STAGE_WARP_BYTES +
                                                         desc_stage_offset((threadIdx.x >> 3) & 3));
    double *s_exp = reinterpret_cast<double *>(s_raw + DESC_SMEM_HIST + DESC_SMEM_STAGE);
    DescRows &my_rows = reinterpret_cast<DescRows *>(s_raw + DESC_SMEM_HIST + DESC_SMEM_STAGE + DESC_SMEM_EXP)[threadIdx.x >> 3];
    if (threadIdx.x < 32) s_exp[threadIdx.x] = c_exp_t32[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, l8 = lane & 7, obase = lane & 24;
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
        describe_octets<false, GPUVAR>(hist, my_rows, my_stage, s_exp, act, k, T.go[oct][sc - 1],
                        T.pitch[oct], T.w[oct], T.h[oct], T.octsize[oct], o->desc);
    }
}

// Stage-hook form: desc[i] for every input row (no filtering), single gradient plane
template <bool GPUVAR>
__global__ void __launch_bounds__(DESC_WARPS * 32, 7) k_describe_rows(const float2 *__restrict__ go, int pitch, int w,
                                                                    int h, const float4 *__restrict__ kp, int n,
                                                                    int octsize, uint8_t *__restrict__ desc) {
    __shared__ __align__(16) unsigned char s_raw[DESC_SMEM_BYTES];
    float *hist = reinterpret_cast<float *>(s_raw) + (threadIdx.x >> 5) * (DESC_HROWS * 32);
    DescStage &my_stage = *reinterpret_cast<DescStage *>(s_raw + DESC_SMEM_HIST + (threadIdx.x >> 5) * DESC_STAGE_WARP_BYTES +
                                                         desc_stage_offset((threadIdx.x >> 3) & 3));
    double *s_exp = reinterpret_cast<double *>(s_raw + DESC_SMEM_HIST + DESC_SMEM_STAGE);
    DescRows &my_rows = reinterpret_cast<DescRows *>(s_raw + DESC_SMEM_HIST + DESC_SMEM_STAGE + DESC_SMEM_EXP)[threadIdx.x >> 3];
    if (threadIdx.x < 32) s_exp[threadIdx.x] = c_exp_t32[threadIdx.x];
    __syncthreads();

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
        describe_octets<true, GPUVAR>(hist, my_rows, my_stage, s_exp, act, k, go, pitch, w, h, octsize,
                        desc + 128L * (act ? gid0 : 0));
    }
}
