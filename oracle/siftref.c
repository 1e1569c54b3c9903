/*
 * siftref.c -- CPU ORACLE for the sift_pyocl keypoint path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's OpenCL kernels (CPU variant) and of the
 * host control flow of sift-src/plan.py / match.py.  It is the parity checker and the timed CPU
 * baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (sift_pyocl_b200) never calls into this library.
 *
 * Parity pin: the reference cannot run here (no PyOpenCL / OpenCL ICD) and ships no golden
 * vectors.  The oracle is pinned against (a) the known-answer relations the reference's own tests
 * assert (scipy convolve1d "reflect", numpy.gradient, numpy taps formula, img[::2, ::2] ...) and
 * (b) outputs of the reference's own pure-python restatement test/test_image_functions.py executed
 * in the build container (tests/golden/make_golden.py -> tests/golden/ npz files).  Descriptor values,
 * matching and alignment are not asserted by any reference test ("parity unpinned" there beyond
 * the python restatement).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Floating-point policy (see DESIGN.md "Numerics contract"):
 *   - all arithmetic is IEEE fp32 exactly as written in the .cl source; unsuffixed literals
 *     (0.5, 3.0, 36.0, 2.0, 4.0, 0.8, 512.0) are double as OpenCL C defines on fp64 devices;
 *   - NO floating-point contraction anywhere (compile with -ffp-contract=off) EXCEPT the
 *     convolution multiply-accumulate `sum += in*filter`, which is evaluated as one fused
 *     fmaf() per tap in tap order: OpenCL C defaults to FP_CONTRACT ON, so every FMA-capable
 *     device the reference ran on (Fermi+/Haswell+) fuses exactly that statement;
 *   - OpenCL built-ins with device-defined ulp error (exp, atan2, sin, cos, pow, rsqrt) are
 *     evaluated correctly rounded: computed in double and rounded once to fp32.  That is the
 *     centre of every conformant device's error interval and is libm-version independent.
 *   - append order of atomics is made deterministic (scan order); parity checks sort anyway.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))
#define MIN(i, j) ((i) < (j) ? (i) : (j))
#define MAX(i, j) ((i) < (j) ? (j) : (i))

/* OpenCL float constants */
#define M_PI_F 3.14159274101257f
#define M_1_PI_F 0.318309886183791f

/* correctly rounded fp32 built-ins (double evaluation, one rounding) */
static inline float cr_expf(float x) { return (float)exp((double)x); }
static inline float cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
static inline float cr_sinf(float x) { return (float)sin((double)x); }
static inline float cr_cosf(float x) { return (float)cos((double)x); }
static inline float cr_exp2f(float x) { return (float)exp2((double)x); } /* pow(2.0f, x) */
static inline float cr_rsqrtf(float x) { return (float)(1.0 / sqrt((double)x)); }

typedef struct {
    float x, y, scale, angle;
    uint8_t desc[128];
} siftref_kp; /* 144 bytes == numpy dtype_kp, plan.py:110-115 */

API int siftref_version(void) { return 1; }

API int siftref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

API void siftref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* utils.py:54-64  kernel_size(sigma, odd=True, cutoff=4)                                      */
API int siftref_kernel_size(double sigma, int odd) {
    int size = (int)ceil(2 * 4 * sigma + 1);
    if (odd && size % 2 == 0) size += 1;
    return size;
}

/* numpy's pairwise float32 add.reduce for n < 128 (what gaussian.sum(dtype=float32) does) */
static float np_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

/* plan.py:315-317 / 336-339 (the documented numpy equivalent of gaussian.cl:56):
 *   x = arange(size) - (size-1)/2 ; g = exp(-(x/sigma)**2/2).astype(f32) ; g /= g.sum(dtype=f32) */
API void siftref_gaussian_taps(double sigma, int size, float *out) {
    for (int i = 0; i < size; i++) {
        double x = (double)i - ((double)size - 1.0) / 2.0;
        double q = x / sigma;
        out[i] = (float)exp(-(q * q) / 2.0);
    }
    float s = np_sum_f32(out, size);
    for (int i = 0; i < size; i++) out[i] = out[i] / s;
}

/* plan.py:213-224  _calc_scales: number of octaves for a (h, w) image */
API int siftref_num_octaves(int h, int w) {
    int n = 1, min_size = 2 * 5 + 2;
    while (MIN(h, w) > min_size) {
        h /= 2;
        w /= 2;
        n++;
    }
    return n - 1;
}

/* ------------------------------------------------------------------------------------------ */
/* preprocess.cl:53-223  integer / RGB -> float conversions                                    */
/* dtype codes: 0=f32 1=u8 2=u16 3=u32 4=u64 5=i32 6=i64 7=f64 8=rgb-u8                          */
API int siftref_to_float(const void *src, int dtype, long n, float *dst) {
    long i;
    switch (dtype) {
    case 0: memcpy(dst, src, n * sizeof(float)); break;
    case 1: for (i = 0; i < n; i++) dst[i] = (float)((const uint8_t *)src)[i]; break;
    case 2: for (i = 0; i < n; i++) dst[i] = (float)((const uint16_t *)src)[i]; break;
    case 3: for (i = 0; i < n; i++) dst[i] = (float)((const uint32_t *)src)[i]; break;
    case 4: for (i = 0; i < n; i++) dst[i] = (float)((const uint64_t *)src)[i]; break;
    case 5: for (i = 0; i < n; i++) dst[i] = (float)((const int32_t *)src)[i]; break;
    case 6: for (i = 0; i < n; i++) dst[i] = (float)((const int64_t *)src)[i]; break;
    case 7: for (i = 0; i < n; i++) dst[i] = (float)((const double *)src)[i]; break; /* plan.py:461 */
    case 8: { /* preprocess.cl:221 */
        const uint8_t *p = (const uint8_t *)src;
        for (i = 0; i < n; i++) {
            float r = 0.299f * p[3 * i];
            float g = 0.587f * p[3 * i + 1];
            float b = 0.114f * p[3 * i + 2];
            dst[i] = (r + g) + b;
        }
        break;
    }
    default: return -1;
    }
    return 0;
}

/* reductions.cl:217-239  max_min_serial (the exact form; the tree version gives the same) */
API void siftref_minmax(const float *data, long n, float *minimum, float *maximum) {
    float mini = data[0], maxi = data[0];
#pragma omp parallel for reduction(min : mini) reduction(max : maxi) schedule(static)
    for (long i = 1; i < n; i++) {
        float v = data[i];
        if (v > maxi) maxi = v;
        if (v < mini) mini = v;
    }
    *minimum = mini;
    *maximum = maxi;
}

/* preprocess.cl:238-252  normalizes: image = max_out*(image-min)/(max-min), in place */
API void siftref_normalize(float *image, long n, float min_in, float max_in, float max_out) {
    float den = max_in - min_in;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) image[i] = (max_out * (image[i] - min_in)) / den;
}

/* preprocess.cl:266-285  shrink: out[y,x] = in[y*sh, x*sw] */
API void siftref_shrink(const float *in, float *out, int scale_w, int scale_h, int large_w, int large_h,
                        int small_w, int small_h) {
    (void)large_h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < small_h; y++)
        for (int x = 0; x < small_w; x++)
            out[x + small_w * y] = in[(long)x * scale_w + (long)y * scale_h * large_w];
}

/* ------------------------------------------------------------------------------------------ */
/* convolution.cl:16-55  horizontal_convolution                                                */
/* sum += input[idx]*filter[hlen-1-jx] for jx = 0..hlen-1; mirrored borders (-1->0, W->W-1).    */
static inline int conv_index(int g, int c, int j, int dim) {
    /* convolution.cl:41-50 : idx = g-c+j ; if (j < c-g) idx = c-g-j-1 ; if (j > dim-1-g+c) idx = dim-(j-(dim-1-g+c)) */
    int j1 = c - g, j2 = dim - 1 - g + c;
    int idx = g - c + j;
    if (j < j1) idx = j1 - j - 1;
    if (j > j2) idx = dim - (j - j2);
    return idx;
}

API void siftref_convolve_h(const float *input, float *output, const float *filter, int hlen, int width,
                            int height) {
    int c = (hlen & 1) ? hlen / 2 : hlen / 2 - 1; /* convolution.cl:31-40 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++) {
        const float *row = input + (long)y * width;
        float *orow = output + (long)y * width;
        for (int x = 0; x < width; x++) orow[x] = 0.0f;
        for (int j = 0; j < hlen; j++) {
            float f = filter[hlen - 1 - j];
            int lo = MAX(0, c - j), hi = MIN(width, width + c - j); /* x with 0 <= x-c+j < width */
            for (int x = 0; x < MIN(lo, width); x++) orow[x] = fmaf(row[conv_index(x, c, j, width)], f, orow[x]);
            for (int x = lo; x < hi; x++) orow[x] = fmaf(row[x - c + j], f, orow[x]);
            for (int x = MAX(hi, lo); x < width; x++) orow[x] = fmaf(row[conv_index(x, c, j, width)], f, orow[x]);
        }
    }
}

/* convolution.cl:62-101  vertical_convolution */
API void siftref_convolve_v(const float *input, float *output, const float *filter, int hlen, int width,
                            int height) {
    int c = (hlen & 1) ? hlen / 2 : hlen / 2 - 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++) {
        float *orow = output + (long)y * width;
        for (int x = 0; x < width; x++) orow[x] = 0.0f;
        for (int j = 0; j < hlen; j++) {
            float f = filter[hlen - 1 - j];
            const float *row = input + (long)conv_index(y, c, j, height) * width;
            for (int x = 0; x < width; x++) orow[x] = fmaf(row[x], f, orow[x]);
        }
    }
}

/* plan.py:571-594  _gaussian_convolution: horizontal into tmp, vertical into output */
API void siftref_blur(const float *input, float *output, float *tmp, const float *filter, int hlen, int width,
                      int height) {
    siftref_convolve_h(input, tmp, filter, hlen, width, height);
    siftref_convolve_v(tmp, output, filter, hlen, width, height);
}

/* algebra.cl:18-38  combine: w = a*u + b*v */
API void siftref_combine(const float *u, float a, const float *v, float b, float *w, long n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        float p = a * u[i];
        float q = b * v[i];
        w[i] = p + q;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* image.cl:47-81  compute_gradient_orientation                                                */
API void siftref_gradient(const float *igray, float *grad, float *ori, int width, int height) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            long pos = (long)y * width + x;
            float xgrad, ygrad;
            if (x == 0)
                xgrad = 2.0f * (igray[pos + 1] - igray[pos]);
            else if (x == width - 1)
                xgrad = 2.0f * (igray[pos] - igray[pos - 1]);
            else
                xgrad = igray[pos + 1] - igray[pos - 1];
            if (y == 0)
                ygrad = 2.0f * (igray[pos] - igray[pos + width]);
            else if (y == height - 1)
                ygrad = 2.0f * (igray[pos - width] - igray[pos]);
            else
                ygrad = igray[pos - width] - igray[pos + width];
            float xx = xgrad * xgrad;
            float yy = ygrad * ygrad;
            grad[pos] = sqrtf(xx + yy);
            ori[pos] = cr_atan2f(-ygrad, xgrad);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* image.cl:119-214  local_maxmin for ONE scale; appends (val, row, col, scale) rows.          */
/* Returns the value of the counter after the call (the counter keeps running past the         */
/* capacity like the reference's atomic_inc; stores are guarded, image.cl:202-208).            */
static int maxmin_pixel(const float *DOGS, int gid0, int gid1, float peak_thresh, int octsize, float EdgeThresh0,
                        float EdgeThresh, int scale, int width, int height) {
    long plane = (long)width * height;
    long index_dog_prev = (scale - 1) * plane, index_dog = scale * plane, index_dog_next = (scale + 1) * plane;
    float res = 0.0f;
    float val = DOGS[index_dog + gid0 + (long)width * gid1];
    /* image.cl:152: fabs(val) > (0.8 * peak_thresh) is a double comparison */
    if (!(fabs((double)val) > (0.8 * (double)peak_thresh))) return 0;
    int ismax = 0, ismin = 0;
    if ((double)val > 0.0) ismax = 1;
    else ismin = 1;
    for (int r = gid1 - 1; r <= gid1 + 1; r++) {
        for (int c = gid0 - 1; c <= gid0 + 1; c++) {
            long pos = (long)r * width + c;
            if (ismax == 1)
                if (DOGS[index_dog_prev + pos] > val || DOGS[index_dog + pos] > val || DOGS[index_dog_next + pos] > val)
                    ismax = 0;
            if (ismin == 1)
                if (DOGS[index_dog_prev + pos] < val || DOGS[index_dog + pos] < val || DOGS[index_dog_next + pos] < val)
                    ismin = 0;
        }
    }
    if (ismax == 1 || ismin == 1) res = val;
    long pos = (long)gid1 * width + gid0;
    const float *D = DOGS + index_dog;
    /* image.cl:180-184: "2.0" and "4.0" are double literals -> double evaluation, one rounding on store */
    float H00 = (float)(((double)D[(long)(gid1 - 1) * width + gid0] - 2.0 * (double)D[pos]) +
                        (double)D[(long)(gid1 + 1) * width + gid0]);
    float H11 = (float)(((double)D[pos - 1] - 2.0 * (double)D[pos]) + (double)D[pos + 1]);
    float d1 = D[(long)(gid1 + 1) * width + gid0 + 1] - D[(long)(gid1 + 1) * width + gid0 - 1];
    float d2 = D[(long)(gid1 - 1) * width + gid0 + 1] - D[(long)(gid1 - 1) * width + gid0 - 1];
    float H01 = (float)((double)(d1 - d2) / 4.0);
    float p = H00 * H11;
    float q = H01 * H01;
    float det = p - q, trace = H00 + H11;
    float edthresh = (octsize <= 1 ? EdgeThresh0 : EdgeThresh);
    float tt = edthresh * trace;
    tt = tt * trace;
    if (det < tt) res = 0.0f;
    return res != 0.0f;
}

API int siftref_local_maxmin(const float *DOGS, float *output, int border_dist, float peak_thresh, int octsize,
                             float EdgeThresh0, float EdgeThresh, int *counter, int nb_keypoints, int scale,
                             int width, int height) {
    int rows = height - 2 * border_dist;
    if (rows <= 0 || width - 2 * border_dist <= 0) return *counter;
    /* per-row hit lists so that the append order is the deterministic row-major scan order */
    int *row_cnt = (int *)calloc(rows, sizeof(int));
    int **row_hits = (int **)calloc(rows, sizeof(int *));
#pragma omp parallel for schedule(dynamic, 8)
    for (int gid1 = border_dist; gid1 < height - border_dist; gid1++) {
        int cap = 0, n = 0, *hits = NULL;
        for (int gid0 = border_dist; gid0 < width - border_dist; gid0++) {
            if (maxmin_pixel(DOGS, gid0, gid1, peak_thresh, octsize, EdgeThresh0, EdgeThresh, scale, width, height)) {
                if (n == cap) {
                    cap = cap ? 2 * cap : 16;
                    hits = (int *)realloc(hits, cap * sizeof(int));
                }
                hits[n++] = gid0;
            }
        }
        row_cnt[gid1 - border_dist] = n;
        row_hits[gid1 - border_dist] = hits;
    }
    long plane = (long)width * height;
    for (int i = 0; i < rows; i++) {
        int gid1 = i + border_dist;
        for (int k = 0; k < row_cnt[i]; k++) {
            int gid0 = row_hits[i][k];
            int old = (*counter)++;
            if (old < nb_keypoints) {
                output[4 * (long)old + 0] = DOGS[scale * plane + gid0 + (long)width * gid1];
                output[4 * (long)old + 1] = (float)gid1;
                output[4 * (long)old + 2] = (float)gid0;
                output[4 * (long)old + 3] = (float)scale;
            }
        }
        free(row_hits[i]);
    }
    free(row_hits);
    free(row_cnt);
    return *counter;
}

/* ------------------------------------------------------------------------------------------ */
/* image.cl:235-366  interp_keypoint, rows [start, end) in place                                */
API void siftref_interp_keypoint(const float *DOGS, float *keypoints, int start_keypoints, int end_keypoints,
                                 float peak_thresh, float InitSigma, int width, int height) {
#pragma omp parallel for schedule(static)
    for (int gid0 = start_keypoints; gid0 < end_keypoints; gid0++) {
        float *k = keypoints + 4 * (long)gid0;
        int r = (int)k[1];
        int c = (int)k[2];
        int scale = (int)k[3];
        if (r == -1) continue;
        long plane = (long)width * height;
        const float *Dp = DOGS + (scale - 1) * plane, *D = DOGS + scale * plane, *Dn = DOGS + (scale + 1) * plane;
        float g0, g1, g2, H00, H11, H22, H01, H02, H12, H10, H20, H21, K00, K11, K22, K01, K02, K12, K10, K20, K21,
            solution0 = 0, solution1 = 0, solution2 = 0, det, peakval = 0;
        long pos;
        int loop = 1, movesRemain = 5;
        int newr = r, newc = c;
        while (loop == 1) {
            r = newr, c = newc;
            pos = (long)newr * width + newc;
            long up = (long)(newr - 1) * width + newc, dn = (long)(newr + 1) * width + newc;
            g0 = (Dn[pos] - Dp[pos]) / 2.0f;
            g1 = (D[dn] - D[up]) / 2.0f;
            g2 = (D[pos + 1] - D[pos - 1]) / 2.0f;
            { float t = 2.0f * D[pos]; H00 = (Dp[pos] - t) + Dn[pos]; }
            { float t = 2.0f * D[pos]; H11 = (D[up] - t) + D[dn]; }
            { float t = 2.0f * D[pos]; H22 = (D[pos - 1] - t) + D[pos + 1]; }
            H01 = ((Dn[dn] - Dn[up]) - (Dp[dn] - Dp[up])) / 4.0f;
            H02 = ((Dn[pos + 1] - Dn[pos - 1]) - (Dp[pos + 1] - Dp[pos - 1])) / 4.0f;
            H12 = ((D[dn + 1] - D[dn - 1]) - (D[up + 1] - D[up - 1])) / 4.0f;
            H10 = H01; H20 = H02; H21 = H12;
            /* image.cl:300: det = -(H02*H11*H20) + H01*H12*H20 + H02*H10*H21 - H00*H12*H21 - H01*H10*H22 + H00*H11*H22 */
            {
                float t1 = (H02 * H11) * H20, t2 = (H01 * H12) * H20, t3 = (H02 * H10) * H21;
                float t4 = (H00 * H12) * H21, t5 = (H01 * H10) * H22, t6 = (H00 * H11) * H22;
                det = ((((-t1 + t2) + t3) - t4) - t5) + t6;
            }
#define P2(a, b, c2, d) ({ float _p = (a) * (b); float _q = (c2) * (d); _p - _q; })
            K00 = P2(H11, H22, H12, H21);
            K01 = P2(H02, H21, H01, H22);
            K02 = P2(H01, H12, H02, H11);
            K10 = P2(H12, H20, H10, H22);
            K11 = P2(H00, H22, H02, H20);
            K12 = P2(H02, H10, H00, H12);
            K20 = P2(H10, H21, H11, H20);
            K21 = P2(H01, H20, H00, H21);
            K22 = P2(H00, H11, H01, H10);
#undef P2
#define D3(a, A, b, B, c2, C) ({ float _a = (a) * (A); float _b = (b) * (B); float _c = (c2) * (C); (_a + _b) + _c; })
            solution0 = -D3(g0, K00, g1, K01, g2, K02) / det;
            solution1 = -D3(g0, K10, g1, K11, g2, K12) / det;
            solution2 = -D3(g0, K20, g1, K21, g2, K22) / det;
            peakval = D[pos] + 0.5f * D3(solution0, g0, solution1, g1, solution2, g2);
#undef D3
            if (solution1 > 0.6f && newr < height - 3) newr++;
            else if (solution1 < -0.6f && newr > 3) newr--;
            if (solution2 > 0.6f && newc < width - 3) newc++;
            else if (solution2 < -0.6f && newc > 3) newc--;
            if (movesRemain > 0 && (newr != r || newc != c)) movesRemain--;
            else loop = 0;
        }
        if (fabsf(solution0) <= 1.5f && fabsf(solution1) <= 1.5f && fabsf(solution2) <= 1.5f &&
            fabsf(peakval) >= peak_thresh) {
            k[0] = peakval;
            k[1] = r + solution1;
            k[2] = c + solution2;
            k[3] = InitSigma * cr_exp2f((((float)scale) + solution0) / 3.0f);
        } else {
            k[0] = -1.0f; k[1] = -1.0f; k[2] = -1.0f; k[3] = -1.0f;
        }
    }
}

/* algebra.cl:57-83 compact + plan.py:758-795 _compact: rows < start kept, rows in [start,end) with
 * s1 != -1 appended in order; returns the new count; rows beyond are reset to -1 (memset_float). */
API int siftref_compact(float *keypoints, int start, int end, int kpsize) {
    int cnt = start;
    for (int i = start; i < end; i++) {
        if (keypoints[4 * (long)i + 1] != -1) {
            if (cnt != i) memcpy(keypoints + 4 * (long)cnt, keypoints + 4 * (long)i, 4 * sizeof(float));
            cnt++;
        }
    }
    for (long i = 4 * (long)cnt; i < 4 * (long)MIN(end, kpsize); i++) keypoints[i] = -1.0f;
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* orientation_cpu.cl:41-173  orientation_assignment for rows [start, end); extra keypoints      */
/* appended at *counter (deterministic keypoint order).  Returns the new counter.              */
static int orient_one(float *k, const float *grad, const float *ori, int octsize, float OriSigma, int grad_width,
                      int grad_height, float *extra_angles) {
    int bin, prev = 0, next = 0, i, j, r, c, n_extra = 0;
    float distsq, gval, angle, interp = 0.0f;
    float hist_prev, hist_next;
    float hist[36];
    for (i = 0; i < 36; i++) hist[i] = 0.0f;
    int row = (int)((double)k[1] + 0.5), col = (int)((double)k[2] + 0.5); /* orientation_cpu.cl:67-68 */
    float sigma = OriSigma * k[3];
    int radius = (int)((double)sigma * 3.0); /* :71 */
    int rmin = MAX(0, row - radius);
    int cmin = MAX(0, col - radius);
    int rmax = MIN(row + radius, grad_height - 2);
    int cmax = MIN(col + radius, grad_width - 2);
    float two_s2 = (2.0f * sigma) * sigma;
    float rad2 = ((float)(radius * radius)) + 0.5f;
    for (r = rmin; r <= rmax; r++) {
        for (c = cmin; c <= cmax; c++) {
            gval = grad[(long)r * grad_width + c];
            float dif = (r - k[1]);
            distsq = dif * dif;
            dif = (c - k[2]);
            { float d2 = dif * dif; distsq += d2; }
            if (gval > 0.0f && distsq < rad2) {
                angle = ori[(long)r * grad_width + c];
                bin = (int)(36.0f * ((angle + M_PI_F) + 0.001f) / (2.0f * M_PI_F));
                if (bin >= 0 && bin <= 36) {
                    bin = MIN(bin, 35);
                    float w = cr_expf(-distsq / two_s2) * gval;
                    hist[bin] += w;
                }
            }
        }
    }
    /* :100-108 smoothing x6, in place, "/ 3.0" double */
    for (j = 0; j < 6; j++) {
        float prv, temp;
        prv = hist[35];
        for (i = 0; i < 36; i++) {
            temp = hist[i];
            hist[i] = (float)((double)((prv + hist[i]) + hist[(i + 1 == 36) ? 0 : i + 1]) / 3.0);
            prv = temp;
        }
    }
    float maxval = 0.0f;
    int argmax = 0;
    for (i = 0; i < 36; i++) {
        if (maxval < hist[i]) { maxval = hist[i]; argmax = i; }
    }
    prev = (argmax == 0 ? 35 : argmax - 1);
    next = (argmax == 35 ? 0 : argmax + 1);
    hist_prev = hist[prev];
    hist_next = hist[next];
    if (maxval < 0.0f) { hist_prev = -hist_prev; maxval = -maxval; hist_next = -hist_next; }
    interp = 0.5f * (hist_prev - hist_next) / ((hist_prev - 2.0f * maxval) + hist_next);
    angle = (2.0f * M_PI_F) * ((argmax + 0.5f) + interp) / 36.0f - M_PI_F;
    {
        float k0 = k[2] * octsize, k1 = k[1] * octsize, k2 = k[3] * octsize;
        k[0] = k0; k[1] = k1; k[2] = k2; k[3] = angle;
    }
    for (i = 0; i < 36; i++) {
        int pv = (i == 0 ? 35 : i - 1), nx = (i == 35 ? 0 : i + 1);
        float hp = hist[pv], hc = hist[i], hn = hist[nx];
        if (hc > hp && hc > hn && hc >= 0.8f * maxval && i != argmax) {
            if (hc < 0.0f) { hp = -hp; hc = -hc; hn = -hn; }
            float itp = 0.5f * (hp - hn) / ((hp - 2.0f * hc) + hn);
            /* :166 "/36.0" is double: the rest of the expression is evaluated in double */
            float a = (float)((double)((2.0f * M_PI_F) * ((i + 0.5f) + itp)) / 36.0 - (double)M_PI_F);
            if (a >= -M_PI_F && a <= M_PI_F) extra_angles[n_extra++] = a;
        }
    }
    return n_extra;
}

/* orientation_gpu.cl:69-315  the GPU variant of orientation_assignment (SURVEY App. A.7 "GPU variant       */
/* differences"): bin = (int)(18 (ori + pi) / pi) with +-36 wrap, only the first WORKGROUP_SIZE = 128 columns of   */
/* a window row are visited, smoothing multiplies by (1.0f / 3.0f), tree argmax (ties: :187-236), angle wrapped    */
/* into [0, 2] then (a - 1) pi, extra peaks not range-filtered.  The cross-warp race of the smoothing step         */
/* (:157-172) is resolved the way lock-step execution resolves it: every bin sees the OLD values of    */
/* its neighbours, bin 35 the NEW bin 0 (the same data flow as the CPU variant).                                   */
static int orient_one_gpu(float *k, const float *grad, const float *ori, int octsize, float OriSigma, int grad_width,
                          int grad_height, float *extra_angles) {
    const int WG = 128;
    const float ONE_3 = 1.0f / 3.0f, ONE_18 = 1.0f / 18.0f;
    int i, j, r, n_extra = 0;
    float hist[36], hist2[128];
    int pos[128];
    for (i = 0; i < 36; i++) hist[i] = 0.0f;
    int row = (int)((double)k[1] + 0.5), col = (int)((double)k[2] + 0.5);
    float sigma = OriSigma * k[3];
    int radius = (int)((double)sigma * 3.0);
    int rmin = MAX(0, row - radius), cmin = MAX(0, col - radius);
    int rmax = MIN(row + radius, grad_height - 2), cmax = MIN(col + radius, grad_width - 2);
    float two_s2 = (2.0f * sigma) * sigma;
    float rad2 = ((float)(radius * radius)) + 0.5f;
    for (r = rmin; r <= rmax; r++) {
        for (int lid0 = 0; lid0 < WG; lid0++) {   /* lane 0 adds the row's samples in lane order (:141-145) */
            int c = cmin + lid0;
            if (c > cmax) break;
            float gval = grad[(long)r * grad_width + c];
            float dr = (r - k[1]), dc = (c - k[2]);
            float t1 = dr * dr, t2 = dc * dc;
            float distsq = t1 + t2;
            if (gval > 0.0f && distsq < rad2) {
                float angle = ori[(long)r * grad_width + c];
                int bin = (int)((18.0f * (angle + M_PI_F)) * M_1_PI_F);
                if (bin < 0) bin += 36;
                if (bin > 35) bin -= 36;
                hist[bin] += cr_expf(-distsq / two_s2) * gval;
            }
        }
    }
    for (j = 0; j < 6; j++) {   /* :157-172 */
        float old[36];
        for (i = 0; i < 36; i++) old[i] = hist[i];
        hist[0] = ((old[35] + old[0]) + old[1]) * ONE_3;
        for (i = 1; i < 35; i++) hist[i] = ((old[i - 1] + old[i]) + old[i + 1]) * ONE_3;
        hist[35] = ((old[34] + old[35]) + hist[0]) * ONE_3;
    }
    /* :187-236 tree reduction for the maximum */
    for (i = 0; i < 32; i++) {
        if (i + 32 < 36) {
            if (hist[i] > hist[i + 32]) { hist2[i] = hist[i]; pos[i] = i; }
            else { hist2[i] = hist[i + 32]; pos[i] = i + 32; }
        } else { hist2[i] = hist[i]; pos[i] = i; }
    }
    for (int step = 16; step >= 1; step >>= 1)
        for (i = 0; i < step; i++)
            if (hist2[i + step] > hist2[i]) { hist2[i] = hist2[i + step]; pos[i] = pos[i + step]; }
    int argmax = pos[0];
    float maxval = hist2[0];
    int prev = (argmax == 0 ? 35 : argmax - 1), next = (argmax == 35 ? 0 : argmax + 1);
    float hist_prev = hist[prev], hist_next = hist[next];
    float interp = 0.5f * (hist_prev - hist_next) / ((hist_prev - 2.0f * maxval) + hist_next);
    float angle = ((argmax + 0.5f) + interp) * ONE_18;
    if (angle < 0.0f) angle += 2.0f;
    else if (angle > 2.0f) angle -= 2.0f;
    {
        float k0 = k[2] * octsize, k1 = k[1] * octsize, k2 = k[3] * octsize;
        k[0] = k0; k[1] = k1; k[2] = k2; k[3] = (angle - 1.0f) * M_PI_F;
    }
    for (i = 0; i < 36; i++) {   /* :286-311 */
        if (i == argmax) continue;
        int pv = (i == 0 ? 35 : i - 1), nx = (i == 35 ? 0 : i + 1);
        float hp = hist[pv], hc = hist[i], hn = hist[nx];
        if (hc > hp && hc > hn && hc >= 0.8f * maxval) {
            float itp = 0.5f * (hp - hn) / ((hp - 2.0f * hc) + hn);
            float a = ((i + 0.5f) + itp) * ONE_18;
            if (a < 0.0f) a += 2.0f;
            else if (a > 2.0f) a -= 2.0f;
            extra_angles[n_extra++] = (a - 1.0f) * M_PI_F;
        }
    }
    return n_extra;
}

/* variant: 0 = orientation_cpu.cl, 1 = orientation_gpu.cl */
API int siftref_orientation_v(float *keypoints, const float *grad, const float *ori, int *counter, int octsize,
                              float OriSigma, int nb_keypoints, int keypoints_start, int keypoints_end, int grad_width,
                              int grad_height, int variant) {
    int n = keypoints_end - keypoints_start;
    if (n <= 0) return *counter;
    float *extras = (float *)malloc((size_t)n * 36 * sizeof(float));
    int *n_extras = (int *)calloc(n, sizeof(int));
#pragma omp parallel for schedule(dynamic, 16)
    for (int gid0 = keypoints_start; gid0 < keypoints_end; gid0++) {
        float *k = keypoints + 4 * (long)gid0;
        if (!(k[1] >= 0.0f)) continue;
        n_extras[gid0 - keypoints_start] =
            (variant ? orient_one_gpu : orient_one)(k, grad, ori, octsize, OriSigma, grad_width, grad_height,
                                                    extras + 36 * (long)(gid0 - keypoints_start));
    }
    for (int i = 0; i < n; i++) {
        const float *k = keypoints + 4 * (long)(keypoints_start + i);
        for (int e = 0; e < n_extras[i]; e++) {
            int old = (*counter)++;
            if (old < nb_keypoints) {
                float *o = keypoints + 4 * (long)old;
                o[0] = k[0]; o[1] = k[1]; o[2] = k[2];
                o[3] = extras[36 * (long)i + e];
            }
        }
    }
    free(extras);
    free(n_extras);
    return *counter;
}
API int siftref_orientation(float *keypoints, const float *grad, const float *ori, int *counter, int octsize,
                            float OriSigma, int nb_keypoints, int keypoints_start, int keypoints_end, int grad_width,
                            int grad_height) {
    return siftref_orientation_v(keypoints, grad, ori, counter, octsize, OriSigma, nb_keypoints, keypoints_start,
                                 keypoints_end, grad_width, grad_height, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* keypoints_cpu.cl:36-160  descriptor for rows [start, end)                                    */
static void describe_one(const float *k, uint8_t *out, const float *grad, const float *orim, int octsize,
                         int grad_width, int grad_height) {
    int i, j;
    float tmp_descriptors[128];
    for (i = 0; i < 128; i++) tmp_descriptors[i] = 0.0f;
    float rx, cx;
    float row = k[1] / octsize, col = k[0] / octsize, angle = k[3];
    int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    float spacing = k[2] / octsize * 3.0f;
    int iradius = (int)(((1.414f * spacing) * 2.5f) + 0.5f);
    float drow = row - irow, dcol = col - icol;
    for (i = -iradius; i <= iradius; i++) {
        for (j = -iradius; j <= iradius; j++) {
            { float a = cosine * i, b = sine * j; rx = ((a - b) - drow) / spacing + 1.5f; }
            { float a = sine * i, b = cosine * j; cx = ((a + b) - dcol) / spacing + 1.5f; }
            if ((rx > -1.0f && rx < 4.0f && cx > -1.0f && cx < 4.0f && (irow + i) >= 0 && (irow + i) < grad_height &&
                 (icol + j) >= 0 && (icol + j) < grad_width)) {
                float er = rx - 1.5f, ec = cx - 1.5f;
                float e1 = er * er, e2 = ec * ec;
                float mag = grad[(icol + j) + (long)(irow + i) * grad_width] * cr_expf(-0.125f * (e1 + e2));
                float ori = orim[(icol + j) + (long)(irow + i) * grad_width] - angle;
                while (ori > 2.0f * M_PI_F) ori -= 2.0f * M_PI_F;
                while (ori < 0.0f) ori += 2.0f * M_PI_F;
                int orr, rindex, cindex, oindex;
                float cweight;
                float oval = (4.0f * ori) * M_1_PI_F;
                int ri = (int)((rx >= 0.0f) ? rx : rx - 1.0f), ci = (int)((cx >= 0.0f) ? cx : cx - 1.0f),
                    oi = (int)((oval >= 0.0f) ? oval : oval - 1.0f);
                float rfrac = rx - ri, cfrac = cx - ci, ofrac = oval - oi;
                if ((ri >= -1 && ri < 4 && oi >= 0 && oi <= 8 && rfrac >= 0.0f && rfrac <= 1.0f)) {
                    for (int r = 0; r < 2; r++) {
                        rindex = ri + r;
                        if ((rindex >= 0 && rindex < 4)) {
                            float rweight = (float)(mag * (float)((r == 0) ? 1.0f - rfrac : rfrac));
                            for (int c = 0; c < 2; c++) {
                                cindex = ci + c;
                                if ((cindex >= 0 && cindex < 4)) {
                                    cweight = rweight * ((c == 0) ? 1.0f - cfrac : cfrac);
                                    for (orr = 0; orr < 2; orr++) {
                                        oindex = oi + orr;
                                        if (oindex >= 8) oindex = 0;
                                        float t = cweight * ((orr == 0) ? 1.0f - ofrac : ofrac);
                                        tmp_descriptors[(rindex * 4 + cindex) * 8 + oindex] += t;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    /* keypoints_cpu.cl:127-160 normalise, clamp 0.2, renormalise if clamped, x512 -> u8 */
    float norm = 0;
    for (i = 0; i < 128; i++) { float t = tmp_descriptors[i] * tmp_descriptors[i]; norm += t; }
    norm = cr_rsqrtf(norm);
    for (i = 0; i < 128; i++) tmp_descriptors[i] *= norm;
    int changed = 0;
    norm = 0;
    for (i = 0; i < 128; i++) {
        if (tmp_descriptors[i] > 0.2f) { tmp_descriptors[i] = 0.2f; changed = 1; }
        float t = tmp_descriptors[i] * tmp_descriptors[i];
        norm += t;
    }
    if (changed) {
        norm = cr_rsqrtf(norm);
        for (i = 0; i < 128; i++) tmp_descriptors[i] *= norm;
    }
    for (i = 0; i < 128; i++) {
        double v = 512.0 * (double)tmp_descriptors[i];
        int intval = (v != v) ? 0 : (int)v; /* (int)NaN: defined here as 0 (x86 cvtt gives INT_MIN -> uchar 0) */
        out[i] = (uint8_t)MIN(255, intval);
    }
}

/* keypoints_gpu2.cl:68-284  the GPU variant of descriptor (SURVEY App. A.8 "GPU-variant differences"): fixed   */
/* [-64, 64)^2 window, every trilinear term is accumulated as (uint)(100000 * term) with integer atomics (so the   */
/* order of the samples does not matter; uint32 arithmetic wraps), histogram = (float)sum * 0.00001f, sums of      */
/* squares by the 128 -> 2 tree of :214-241, and the final value is cast to uchar BEFORE MIN(255, .) (wraps >= 256) */
static void describe_one_gpu(const float *k, uint8_t *out, const float *grad, const float *orim, int octsize,
                             int grad_width, int grad_height) {
    int i, j;
    uint32_t acc[128];
    float histogram[128], hist2[128];
    for (i = 0; i < 128; i++) acc[i] = 0;
    float one_octsize = 1.0f / octsize;
    float row = k[1] * one_octsize, col = k[0] * one_octsize, angle = k[3];
    int irow = (int)(row + 0.5f), icol = (int)(col + 0.5f);
    float sine = cr_sinf(angle), cosine = cr_cosf(angle);
    float spacing = k[2] * one_octsize * 3.0f;
    float drow = row - irow, dcol = col - icol;
    for (i = -64; i < 64; i++) {
        for (j = -64; j < 64; j++) {
            float rx, cx;
            { float a = cosine * i, b = sine * j; rx = ((a - b) - drow) / spacing + 1.5f; }
            { float a = sine * i, b = cosine * j; cx = ((a + b) - dcol) / spacing + 1.5f; }
            if (!(rx > -1.0f && rx < 4.0f && cx > -1.0f && cx < 4.0f && (irow + i) >= 0 && (irow + i) < grad_height &&
                  (icol + j) >= 0 && (icol + j) < grad_width))
                continue;
            float er = rx - 1.5f, ec = cx - 1.5f;
            float e1 = er * er, e2 = ec * ec;
            float mag = grad[(icol + j) + (long)(irow + i) * grad_width] * cr_expf(-0.125f * (e1 + e2));
            float ori = orim[(icol + j) + (long)(irow + i) * grad_width] - angle;
            while (ori > 2.0f * M_PI_F) ori -= 2.0f * M_PI_F;
            while (ori < 0.0f) ori += 2.0f * M_PI_F;
            float oval = (4.0f * ori) * M_1_PI_F;
            int ri = (int)((rx >= 0.0f) ? rx : rx - 1.0f), ci = (int)((cx >= 0.0f) ? cx : cx - 1.0f),
                oi = (int)((oval >= 0.0f) ? oval : oval - 1.0f);
            float rfrac = rx - ri, cfrac = cx - ci, ofrac = oval - oi;
            if (!(ri >= -1 && ri < 4 && oi >= 0 && oi <= 8 && rfrac >= 0.0f && rfrac <= 1.0f)) continue;
            for (int r = 0; r < 2; r++) {
                int rindex = ri + r;
                if (!(rindex >= 0 && rindex < 4)) continue;
                float rweight = mag * ((r == 0) ? 1.0f - rfrac : rfrac);
                for (int c = 0; c < 2; c++) {
                    int cindex = ci + c;
                    if (!(cindex >= 0 && cindex < 4)) continue;
                    float cweight = rweight * ((c == 0) ? 1.0f - cfrac : cfrac);
                    for (int orr = 0; orr < 2; orr++) {
                        int oindex = oi + orr;
                        if (oindex >= 8) oindex = 0;
                        float t = cweight * ((orr == 0) ? 1.0f - ofrac : ofrac);
                        float scaled = 100000.0f * t;                       /* :190 */
                        acc[(rindex * 4 + cindex) * 8 + oindex] += (scaled != scaled) ? 0u : (uint32_t)scaled;
                    }
                }
            }
        }
    }
    for (i = 0; i < 128; i++) histogram[i] = (float)acc[i] * 0.00001f;    /* :205-209 */
    for (int pass = 0; pass < 2; pass++) {
        for (i = 0; i < 128; i++) hist2[i] = histogram[i] * histogram[i];
        for (int half = 64; half >= 2; half >>= 1)
            for (i = 0; i < half; i++) hist2[i] += hist2[i + half];
        float norm = cr_rsqrtf(hist2[1] + hist2[0]);
        for (i = 0; i < 128; i++) histogram[i] *= norm;
        if (pass == 1) break;
        int changed = 0;
        for (i = 0; i < 128; i++)
            if (histogram[i] > 0.2f) { histogram[i] = 0.2f; changed = 1; }
        if (!changed) break;
    }
    for (i = 0; i < 128; i++) {
        float v = 512.0f * histogram[i];
        int intval = (v != v) ? 0 : (int)v;
        int wrapped = intval & 0xff;                                        /* (unsigned char) cast, :281 */
        out[i] = (uint8_t)MIN(255, wrapped);
    }
}

/* variant: 0 = keypoints_cpu.cl, 1 = keypoints_gpu2.cl */
API void siftref_descriptor_v(const float *keypoints, uint8_t *descriptors, const float *grad, const float *orim,
                              int octsize, int keypoints_start, int keypoints_end, int grad_width, int grad_height,
                              int variant) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int gid0 = keypoints_start; gid0 < keypoints_end; gid0++) {
        const float *k = keypoints + 4 * (long)gid0;
        if (!(k[1] >= 0.0f)) continue;
        (variant ? describe_one_gpu : describe_one)(k, descriptors + 128 * (long)gid0, grad, orim, octsize, grad_width,
                                                    grad_height);
    }
}
API void siftref_descriptor(const float *keypoints, uint8_t *descriptors, const float *grad, const float *orim,
                            int octsize, int keypoints_start, int keypoints_end, int grad_width, int grad_height) {
    siftref_descriptor_v(keypoints, descriptors, grad, orim, octsize, keypoints_start, keypoints_end, grad_width,
                         grad_height, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* plan.py:432-567 keypoints() + :596-756 _one_octave : the whole path                          */
/* stage_counts (optional, may be NULL): int[octaves][3 scales][3] = {extrema, after interp, after orientation} */
/* variant: 0 = the *_cpu.cl kernels (devicetype "CPU"), 1 = orientation_gpu.cl + keypoints_gpu2.cl ("GPU")       */
API int siftref_keypoints_v(const float *image, int height, int width, double init_sigma, int octave_limit,
                            int pix_per_kp, siftref_kp *out, int out_cap, int *n_per_octave, float *minmax,
                            int *stage_counts, int variant) {
    const int Scales = 3, BorderDist = 5;
    const float PeakThresh = (float)(255.0 * 0.04 / 3.0), EdgeThresh = 0.06f, EdgeThresh1 = 0.08f, OriSigma = 1.5f;
    long N = (long)height * width;
    int octave_max = siftref_num_octaves(height, width);
    if (octave_limit > 0 && octave_limit < octave_max) octave_max = octave_limit; /* par.OctaveMax, SURVEY B5 */
    int kpsize = (int)(N / pix_per_kp); /* plan.py:243 */
    float *G[6];
    for (int i = 0; i < 6; i++) G[i] = (float *)malloc(N * sizeof(float));
    float *tmp = (float *)malloc(N * sizeof(float)), *ori = (float *)malloc(N * sizeof(float));
    float *DoGs = (float *)malloc(5 * N * sizeof(float));
    float *Kp = (float *)malloc((size_t)kpsize * 4 * sizeof(float));
    uint8_t *desc = (uint8_t *)malloc((size_t)kpsize * 128);
    float taps[6][64];
    int ntaps[6];
    double sigmaRatio = pow(2.0, 1.0 / Scales); /* plan.py:108 */
    /* plan.py:297-306 */
    double curSigma = 0.5;
    int has_init = 0;
    if (init_sigma > curSigma) {
        double s = sqrt(init_sigma * init_sigma - curSigma * curSigma);
        ntaps[5] = siftref_kernel_size(s, 1);
        siftref_gaussian_taps(s, ntaps[5], taps[5]);
        has_init = 1;
    }
    {
        double prevSigma = init_sigma;  /* python double, plan.py:123-126 */
        for (int i = 0; i < Scales + 2; i++) {
            double increase = prevSigma * sqrt(sigmaRatio * sigmaRatio - 1.0);
            ntaps[i] = siftref_kernel_size(increase, 1);
            siftref_gaussian_taps(increase, ntaps[i], taps[i]);
            prevSigma *= sigmaRatio;
        }
    }
    memcpy(G[0], image, N * sizeof(float));
    float mn, mx;
    siftref_minmax(G[0], N, &mn, &mx);                  /* plan.py:490-523 */
    if (minmax) { minmax[0] = mn; minmax[1] = mx; }
    siftref_normalize(G[0], N, mn, mx, 255.0f);         /* plan.py:525-530 */
    if (has_init) siftref_blur(G[0], G[0], tmp, taps[5], ntaps[5], width, height); /* plan.py:534-539 */

    int total = 0, w = width, h = height;
    for (int octave = 0; octave < octave_max; octave++) {
        long No = (long)w * h;
        int octsize = 1 << octave;
        int cnt = 0, last_start = 0;
        for (long i = 0; i < 4L * kpsize; i++) Kp[i] = -1.0f;     /* _reset_keypoints plan.py:797 */
        for (int s = 0; s < Scales + 2; s++) {                     /* plan.py:609-625 */
            siftref_blur(G[s], G[s + 1], tmp, taps[s], ntaps[s], w, h);
            siftref_combine(G[s + 1], -1.0f, G[s], +1.0f, DoGs + s * No, No);
        }
        for (int s = 1; s < Scales + 1; s++) {                     /* plan.py:626-733 */
            siftref_local_maxmin(DoGs, Kp, BorderDist, PeakThresh, octsize, EdgeThresh1, EdgeThresh, &cnt, kpsize, s,
                                 w, h);
            if (cnt > kpsize) cnt = kpsize; /* reference would run past the buffer; clamp */
            int n_ext = cnt - last_start;
            siftref_interp_keypoint(DoGs, Kp, last_start, cnt, PeakThresh, (float)init_sigma, w, h); /* numpy.float32(self._init_sigma), plan.py:650 */
            int newcnt = siftref_compact(Kp, last_start, cnt, kpsize);
            int n_int = newcnt - last_start;
            cnt = newcnt;
            siftref_gradient(G[s], tmp, ori, w, h);
            if (newcnt && newcnt > last_start) {
                siftref_orientation_v(Kp, tmp, ori, &cnt, octsize, OriSigma, kpsize, last_start, newcnt, w, h, variant);
                if (cnt > kpsize) cnt = kpsize;
                siftref_descriptor_v(Kp, desc, tmp, ori, octsize, last_start, cnt, w, h, variant);
            }
            if (stage_counts) {
                int *sc = stage_counts + (octave * 3 + (s - 1)) * 3;
                sc[0] = n_ext; sc[1] = n_int; sc[2] = cnt - last_start;
            }
            last_start = cnt;
        }
        if (octave < octave_max - 1) {                             /* plan.py:739-745 */
            siftref_shrink(G[Scales], G[0], 2, 2, w, h, w / 2, h / 2);
        }
        /* plan.py:546-565: drop rows containing NaN, append to the output */
        int kept = 0;
        for (int i = 0; i < last_start; i++) {
            const float *k = Kp + 4 * (long)i;
            float sum = ((k[0] + k[1]) + k[2]) + k[3];
            if (sum != sum) continue;
            if (total < out_cap) {
                out[total].x = k[0]; out[total].y = k[1]; out[total].scale = k[2]; out[total].angle = k[3];
                memcpy(out[total].desc, desc + 128 * (long)i, 128);
            }
            total++;
            kept++;
        }
        if (n_per_octave) n_per_octave[octave] = kept;
        w /= 2;
        h /= 2;
    }
    for (int i = 0; i < 6; i++) free(G[i]);
    free(tmp); free(ori); free(DoGs); free(Kp); free(desc);
    return total;
}
API int siftref_keypoints(const float *image, int height, int width, double init_sigma, int octave_limit,
                          int pix_per_kp, siftref_kp *out, int out_cap, int *n_per_octave, float *minmax,
                          int *stage_counts) {
    return siftref_keypoints_v(image, height, width, init_sigma, octave_limit, pix_per_kp, out, out_cap, n_per_octave,
                               minmax, stage_counts, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* matching_cpu.cl:57-109  matching (L1 on uint8, ratio test); output pairs in kp1 order        */
/* metric: 0 = the reference's L1 distance; 1 = squared L2 distance (not in the reference: the "L2" wording of     */
/* BASELINE config 4), same scan, same tie rules, same ratio threshold applied to the squared distances          */
API int siftref_match_metric(const siftref_kp *keypoints1, const siftref_kp *keypoints2, int *matchings,
                             int max_nb_match, float ratio_th, int size1, int size2, int metric) {
    int *best = (int *)malloc((size_t)MAX(size1, 1) * sizeof(int));
#pragma omp parallel for schedule(static)
    for (int gid0 = 0; gid0 < size1; gid0++) {
        float dist1 = 1000000000000.0f, dist2 = 1000000000000.0f;
        int current_min = 0;
        const uint8_t *desc1 = keypoints1[gid0].desc;
        for (int i = 0; i < size2; i++) {
            const uint8_t *desc2 = keypoints2[i].desc;
            int dist = 0;
            if (metric) {
                for (int j = 0; j < 128; j++) {
                    int d = (int)desc1[j] - (int)desc2[j];
                    dist += d * d;
                }
            } else {
                for (int j = 0; j < 128; j++) {
                    int a = desc1[j], b = desc2[j];
                    dist += (a > b) ? (a - b) : (-a + b);
                }
            }
            if (dist < dist1) { dist2 = dist1; dist1 = (float)dist; current_min = i; }
            else if (dist < dist2) { dist2 = (float)dist; }
        }
        best[gid0] = (dist2 != 0 && dist1 / dist2 < ratio_th) ? current_min : -1;
    }
    int counter = 0;
    for (int gid0 = 0; gid0 < size1; gid0++) {
        if (best[gid0] < 0) continue;
        int old = counter++;
        if (old < max_nb_match) { matchings[2 * old] = gid0; matchings[2 * old + 1] = best[gid0]; }
    }
    free(best);
    return counter;
}
API int siftref_match(const siftref_kp *keypoints1, const siftref_kp *keypoints2, int *matchings, int max_nb_match,
                      float ratio_th, int size1, int size2) {
    return siftref_match_metric(keypoints1, keypoints2, matchings, max_nb_match, ratio_th, size1, size2, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* transform.cl:22-108  affine warp, bilinear (mode 1) or nearest (mode 0)                      */
API void siftref_transform(const float *image, float *output, const float *matrix4, const float *offset2,
                           int image_width, int image_height, int output_width, int output_height, float fill,
                           int mode) {
#pragma omp parallel for schedule(static)
    for (int gid1 = 0; gid1 < output_height; gid1++) {
        for (int gid0 = 0; gid0 < output_width; gid0++) {
            int x = gid0, y = gid1;
            /* dot(mat.s23, (y,x)) = s2*y + s3*x */
            float tx, ty;
            { float a = matrix4[2] * (float)y, b = matrix4[3] * (float)x; tx = a + b; }
            { float a = matrix4[0] * (float)y, b = matrix4[1] * (float)x; ty = a + b; }
            tx += offset2[1];
            ty += offset2[0];
            int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
            float interp = fill;
            if (0.0f <= tx && tx < image_width && 0.0f <= ty && ty < image_height) {
                if (mode == 1) {
                    float image_p = image[(long)ty_prev * image_width + tx_prev];
                    float image_x = (tx_next >= image_width) ? fill : image[(long)ty_prev * image_width + tx_next];
                    float image_y = (ty_next >= image_height) ? fill : image[(long)ty_next * image_width + tx_prev];
                    float image_n = (tx_next >= image_width || ty_next >= image_height)
                                        ? fill
                                        : image[(long)ty_next * image_width + tx_next];
                    float wxn = (float)(tx_next - tx), wxp = (float)(tx - tx_prev);
                    float wyn = (float)(ty_next - ty), wyp = (float)(ty - ty_prev);
                    float interp1, interp2;
                    { float a = wxn * image_p, b = wxp * image_x; interp1 = a + b; }
                    { float a = wxn * image_y, b = wxp * image_n; interp2 = a + b; }
                    { float a = wyn * interp1, b = wyp * interp2; interp = a + b; }
                } else {
                    interp = image[(long)((int)ty) * image_width + ((int)tx)];
                }
            }
            float u = -0.5f, v = -0.5f;
            if (tx >= image_width + u) interp = fill;
            if (ty >= image_height + v) interp = fill;
            output[(long)gid1 * output_width + gid0] = interp;
        }
    }
}

/* transform.cl:116-203  transform_RGB: interleaved uint8, one colour channel at a time */
API void siftref_transform_rgb(const uint8_t *image, uint8_t *output, const float *matrix4, const float *offset2,
                               int image_width, int image_height, int output_width, int output_height, float fill,
                               int mode) {
#pragma omp parallel for schedule(static)
    for (int gid1 = 0; gid1 < output_height; gid1++) {
        for (int gid0 = 0; gid0 < output_width; gid0++) {
            for (int color = 0; color < 3; color++) {
                int x = gid0, y = gid1;
                float tx, ty;
                { float a = matrix4[2] * (float)y, b = matrix4[3] * (float)x; tx = a + b; }
                { float a = matrix4[0] * (float)y, b = matrix4[1] * (float)x; ty = a + b; }
                tx += offset2[1];
                ty += offset2[0];
                int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
                float interp = fill;
                if (0.0f <= tx && tx < image_width && 0.0f <= ty && ty < image_height) {
                    if (mode == 1) {
                        int xo = tx_next >= image_width, yo = ty_next >= image_height;
                        float image_p = image[3 * ((long)ty_prev * image_width + tx_prev) + color];
                        float image_x = xo ? fill : image[3 * ((long)ty_prev * image_width + tx_next) + color];
                        float image_y = yo ? fill : image[3 * ((long)ty_next * image_width + tx_prev) + color];
                        float image_n = (xo || yo) ? fill : image[3 * ((long)ty_next * image_width + tx_next) + color];
                        float wxn = (float)(tx_next - tx), wxp = (float)(tx - tx_prev);
                        float wyn = (float)(ty_next - ty), wyp = (float)(ty - ty_prev);
                        float interp1, interp2;
                        { float a = wxn * image_p, b = wxp * image_x; interp1 = a + b; }
                        { float a = wxn * image_y, b = wxp * image_n; interp2 = a + b; }
                        { float a = wyn * interp1, b = wyp * interp2; interp = a + b; }
                    } else {
                        interp = image[3 * ((long)((int)ty) * image_width + ((int)tx)) + color];
                    }
                }
                if (tx >= image_width + -0.5f) interp = fill;
                if (ty >= image_height + -0.5f) interp = fill;
                output[3 * ((long)gid1 * output_width + gid0) + color] = (uint8_t)(int)interp;
            }
        }
    }
}
