// siftb_api.cu -- C ABI of libsiftb200.so (see include/siftb.h), first translation unit: the SiftPlan path --
// the host-side orchestration that replaces sift-src/plan.py:432-756 (keypoints / _one_octave) -- the warp of the
// frame a plan holds (alignment.py:329-349) and the stage-level test hooks.  (siftb_match.cu: MatchPlan, the
// stateless warps, the NCCL helpers.)
//
// Differences from the reference's control flow (results identical):
//   * no host round trips inside an image: every count stays on the device and drives the next
//     kernel through grid-stride loops; one D->H copy of the counters and one of the records at the end;
//   * the whole pyramid (every octave keeps its own planes) is built first; extrema, refinement, gradient planes,
//     orientation and descriptors then run ONCE per image over all octaves and scales, so records of one octave are
//     not grouped by scale (the reference's order inside a scale group is already nondeterministic: atomic_inc);
//   * blur + DoG + decimation are one kernel per scale instead of four.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "host_common.h"
#include "common.cuh"
#include "k_blur.cuh"
#include "k_blur_tma.cuh"
#include "k_extrema.cuh"
#include "k_frontend.cuh"
#include "k_keypoint.cuh"
#include "k_describe.cuh"
#include "k_warp.cuh"

#define SIFTB_VERSION 100
#define MAX_OCT 32
#define AUX_INTS (8 + 3 * SIFTB_KOCT + 3 * DESC_CLASSES)
#define NSLOT 3  // images in flight per plan
#define NLANE 3  // most images being processed concurrently; 2 by default, SIFTB_LANES=1..3 (compute streams + plane sets), see siftb_plan::Lane

// SIFT constants, param.py:52-79
static const int kScales = 3, kBorderDist = 5;
static const float kPeakThresh = (float)(255.0 * 0.04 / 3.0), kEdgeThresh = 0.06f, kEdgeThresh1 = 0.08f,
                   kOriSigma = 1.5f;

// image.cl:152 `fabs(val) > 0.8 * peak_thresh` is a double comparison; for an fp32 val it equals
// fabsf(val) >= (smallest fp32 strictly above the double product)
static float contrast_gate(float peak_thresh) {
    const double thr = 0.8 * (double)peak_thresh;
    float f = (float)thr;                       // nearest fp32
    if ((double)f > thr) f = nextafterf(f, 0.0f);  // largest fp32 <= thr
    return nextafterf(f, INFINITY);
}
// SIFTB_FORCE_GENERIC=1 routes every blur through the generic kernel (A/B testing of the TMA kernel)
static int env_force_generic() {
    const char *e = getenv("SIFTB_FORCE_GENERIC");
    return e && e[0] == '1';
}

// ---------------------------------------------------------------------------------------------
// taps: utils.py:54-64 kernel_size + plan.py:315-317 numpy formula (identical to the oracle)
static int kernel_size(double sigma) {
    int size = (int)ceil(2 * 4 * sigma + 1);
    if (size % 2 == 0) size += 1;
    return size;
}
static float np_sum_f32(const float *a, int n) {  // numpy pairwise add.reduce, n < 128
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
static void gaussian_taps(double sigma, int size, float *out) {
    for (int i = 0; i < size; i++) {
        double x = (double)i - ((double)size - 1.0) / 2.0;
        double q = x / sigma;
        out[i] = (float)exp(-(q * q) / 2.0);
    }
    float s = np_sum_f32(out, size);
    for (int i = 0; i < size; i++) out[i] = out[i] / s;
}

// ---------------------------------------------------------------------------------------------
// the two TMA boxes of a source plane: full (TH + 2C rows) and continuation (TH rows), k_blur_tma.cuh
struct TbMaps {
    CUtensorMap full, cont;
};
static int tb_encode_pair(TbMaps *m, const float *base, int w, int h, int pitch, int C) {
    return tb_encode(&m->full, base, w, h, pitch, C) || tb_encode(&m->cont, base, w, h, pitch, C, TB_TH);
}

struct Event {
    std::string name;
    cudaEvent_t a, b;
};

struct siftb_plan {
    int device = 0, h = 0, w = 0, dtype = 0, pix_per_kp = 10, n_oct = 0, kpsize = 0, octave_limit = 0;
    int out_cap = 0;  // records of ALL octaves: the reference's limit (kpsize) is per octave (plan.py:243,748-752)
    double init_sigma = 1.6;  // python double in the reference (plan.py:123-126); fp32 only as a kernel argument
    int ow[MAX_OCT], oh[MAX_OCT], opitch[MAX_OCT];
    cudaStream_t stream = nullptr;  // the plan's public queue: carries no kernels, only waits for every submitted image
    std::mutex mtx;
    Taps taps[6];
    int ntaps[6];
    bool has_init = false;
    size_t raw_bytes = 0, dev_bytes = 0;
    // NSLOT image slots so that submit(k+1), submit(k+2) (H->D copies on the copy stream) and collect(k) (D->H of
    // the records, host-side handling of the result) overlap the kernels of another image: with three slots the
    // compute streams always have the next image queued, input already resident, when an image finishes.
    void *d_raws[NSLOT] = {};  // staging for host input (plan dtype)
    // A lane = one compute stream + the planes and keypoint lists of one image being processed.  Images in flight
    // at the same time alternate between the lanes, so the kernels of two images run CONCURRENTLY: the latency-bound
    // phases of one (small octaves, refinement, ordering, launch gaps, kernel tails) are filled by the other
    // (measured: -8.5 % per image).  Lane 1 is only allocated when a second image is submitted while one is in flight
    // (never in profiling mode, where the per-stage events want the kernels of one image alone on the device).
    struct Lane {
        bool ready = false;
        cudaStream_t stream = nullptr;
        float *d_img = nullptr;    // dense fp32 plane for converted integer/RGB/f64 input
        // Gaussian and DoG planes of EVERY octave (octave o: h_o rows at pitch align32(w_o)): the extrema, refinement
        // and gradient kernels run once per image over all octaves after the whole pyramid has been built
        float *G[MAX_OCT][6] = {}, *D[MAX_OCT][5] = {};
        float4 *cands[MAX_OCT] = {};      // per-octave candidate lists, capacity cand_cap[o]
        PyrTable *d_pyr[NSLOT] = {};      // device tables of k_extrema_all / k_refine_all (per slot: counter addresses)
        GradTable *d_grad = nullptr;      // device table of k_gradient4_all
        float2 *gop[SIFTB_KOCT][3] = {};  // (gradient, orientation) planes of every octave (k_keypoint.cuh)
        OctTable table;
        float4 *kp = nullptr;
        int *kp_tag = nullptr;    // octave << 8 | scale
        int *kp_order = nullptr;  // keypoint indices by descending descriptor-window size
        TbMaps tmaps[MAX_OCT][5];  // source G[s] of octave o, boxes for taps[s]
        bool tmaps_ok[MAX_OCT][5] = {};
        TbMaps tmap_img;           // first blur from the converted fp32 plane
        bool tmap_img_ok = false;
    };
    Lane lanes[NLANE];
    int max_lanes = 2;      // SIFTB_LANES=1 keeps every image on one compute stream
    int slot_lane[NSLOT] = {};
    int last_lane = 0;      // lane of the most recent submit
    int cand_cap[MAX_OCT] = {};
    int ext_blocks = 0, grad_blocks = 0;
    int kp_cap = 0;         // keypoints of one image over all octaves
    KpRecord *outs[NSLOT] = {};
    // device counters: [0]=n_out, [1..]: per octave {n_cand, n_kp, n_extra, n_out_oct}; then stage[n_oct][3][3]; then mm[2]
    // per slot AUX_INTS ints: [0] describe work-queue head, [1] refined keypoints (all octaves), [2] extra
    // orientations, [3] keypoints in the descriptor processing order, [4] orientation work-queue head, [8..8+KOCT) records per octave, [8+KOCT..) first record slot per octave, [8+2KOCT..) fill
    int *d_queue = nullptr;
    int *d_cnts[NSLOT] = {};
    int *h_cnts[NSLOT] = {};  // pinned mirrors
    cudaStream_t copy_stream = nullptr;  // host -> device image copies
    cudaStream_t d2h_stream = nullptr;   // device -> host record copies (PCIe is full duplex: never queued behind an upload)
    cudaEvent_t ev_h2d[NSLOT] = {}, ev_done[NSLOT] = {}, ev_d2h[NSLOT] = {};
    cudaEvent_t ev_ext = nullptr;  // siftb_plan_wait_stream
    cudaEvent_t ev_hold[NSLOT] = {};  // siftb_plan_hold_records: readers of outs[slot] on other streams
    bool held[NSLOT] = {};
    const void *src_ptr[NSLOT] = {};  // device pointer of the image submitted to each slot, and its pixel type
    int src_dtype[NSLOT] = {};
    GrowBuf d_warp;                   // output of siftb_plan_warp_last
    int head = 0, n_flight = 0;  // slots [head, head + n_flight) are submitted and not yet collected
    int last = 0;                // slot of the most recently collected image
    int cnt_ints = 0;
    bool profile = false;
    uint64_t launches = 0;
    TbMaps tmap_raws[NSLOT];  // first blur from the host-staging buffers
    bool tmap_raw_ok = false;
    int force_generic = 0;
    int variant = 0;  // 0: orientation_cpu.cl + keypoints_cpu.cl semantics, 1: orientation_gpu.cl + keypoints_gpu2.cl
    std::vector<Event> events_s[NSLOT];  // profiling events of the image in each slot
    int cur = 0;                         // slot being submitted (ProfScope)
    std::vector<const char *> ev_names;
    std::vector<float> ev_ms;

    int *c_nout(int s) const { return d_cnts[s]; }
    int *c_oct(int s, int o) const { return d_cnts[s] + 1 + 4 * o; }
    int *c_stage(int s, int o) const { return d_cnts[s] + 1 + 4 * n_oct + 9 * o; }
    unsigned *c_mm(int s) const { return reinterpret_cast<unsigned *>(d_cnts[s] + 1 + 13 * n_oct); }
};

template <typename T>
static int dalloc(siftb_plan *p, T **ptr, size_t bytes) {
    CK(cudaMalloc((void **)ptr, bytes));
    // once, at plan creation: the padding columns of the pitched planes are read (never used) by the 128-bit loads
    // of the vector kernels; defined contents keep `compute-sanitizer --tool initcheck` quiet
    CK(cudaMemset(*ptr, 0, bytes));
    p->dev_bytes += bytes;
    return 0;
}

static size_t dtype_bytes(int dtype) {
    switch (dtype) {
    case SIFTB_F32: case SIFTB_U32: case SIFTB_I32: return 4;
    case SIFTB_U8: return 1;
    case SIFTB_U16: return 2;
    case SIFTB_U64: case SIFTB_I64: case SIFTB_F64: return 8;
    case SIFTB_RGB8: return 3;
    }
    return 0;
}

extern "C" const char *siftb_last_error(void) { return g_siftb_err.c_str(); }
extern "C" int siftb_version(void) { return SIFTB_VERSION; }
extern "C" int siftb_device_count(int *n) {
    if (!n) return fail(SIFTB_EINVAL, "n is null");
    CK(cudaGetDeviceCount(n));
    return 0;
}
extern "C" int siftb_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(SIFTB_EINVAL, "ptr is null");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return 0;
}
// write-combined page-locked memory: for buffers the CPU only WRITES (input images); the device reads them over
// PCIe without snooping the CPU caches.  CPU reads from it are very slow.
extern "C" int siftb_host_alloc_wc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(SIFTB_EINVAL, "ptr is null");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined));
    return 0;
}
extern "C" int siftb_host_free(void *ptr) {
    CK(cudaFreeHost(ptr));
    return 0;
}

// frees what lane_alloc() allocated (also a partly allocated lane); the lane's stream stays
static void lane_release(siftb_plan *p, siftb_plan::Lane &L) {
    auto drop = [&](auto *&ptr, size_t bytes) {
        if (!ptr) return;
        cudaFree(ptr);
        ptr = nullptr;
        p->dev_bytes -= bytes < p->dev_bytes ? bytes : p->dev_bytes;
    };
    const size_t N = (size_t)p->h * p->w;
    drop(L.d_img, N * sizeof(float));
    for (int o = 0; o < MAX_OCT; o++) {
        const size_t pl = o < p->n_oct ? (size_t)p->opitch[o] * p->oh[o] * sizeof(float) : 0;
        for (auto &q : L.G[o]) drop(q, pl);
        for (auto &q : L.D[o]) drop(q, pl);
        drop(L.cands[o], (size_t)p->cand_cap[o] * sizeof(float4));
    }
    for (int s = 0; s < NSLOT; s++) drop(L.d_pyr[s], sizeof(PyrTable));
    drop(L.d_grad, sizeof(GradTable));
    for (int o = 0; o < SIFTB_KOCT; o++) {
        const size_t pl = o < p->n_oct ? (size_t)p->opitch[o] * p->oh[o] * sizeof(float) : 0;
        for (int i = 0; i < 3; i++) drop(L.gop[o][i], 2 * pl);
    }
    drop(L.kp, (size_t)p->kp_cap * sizeof(float4));
    drop(L.kp_tag, (size_t)p->kp_cap * sizeof(int));
    drop(L.kp_order, (size_t)p->kp_cap * sizeof(int));
    L.ready = false;
}

extern "C" int siftb_plan_destroy(siftb_plan *p) {
    if (!p) return 0;
    DeviceGuard dg_(p->device);
    for (auto &L : p->lanes) if (L.stream) cudaStreamSynchronize(L.stream);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    if (p->d2h_stream) cudaStreamSynchronize(p->d2h_stream);
    for (int s = 0; s < NSLOT; s++) { cudaFree(p->d_raws[s]); cudaFree(p->outs[s]); cudaFree(p->d_cnts[s]); }
    p->d_warp.release();
    for (auto &L : p->lanes) {
        lane_release(p, L);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    cudaFree(p->d_queue);
    for (int s = 0; s < NSLOT; s++) {
        if (p->h_cnts[s]) cudaFreeHost(p->h_cnts[s]);
        if (p->ev_h2d[s]) cudaEventDestroy(p->ev_h2d[s]);
        if (p->ev_done[s]) cudaEventDestroy(p->ev_done[s]);
        if (p->ev_d2h[s]) cudaEventDestroy(p->ev_d2h[s]);
        if (p->ev_hold[s]) cudaEventDestroy(p->ev_hold[s]);
    }
    if (p->ev_ext) cudaEventDestroy(p->ev_ext);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
    for (auto &ev : p->events_s) for (auto &e : ev) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return 0;
}

// planes, lists and device tables of one lane (plan.py:268-295), planes pitched to 128 B
static int lane_alloc(siftb_plan *p, siftb_plan::Lane &L) {
    DeviceGuard dg_(p->device);
    const size_t N = (size_t)p->h * p->w;
    int rc;
    if (p->dtype != SIFTB_F32 && (rc = dalloc(p, &L.d_img, N * sizeof(float)))) return rc;
    for (int o = 0; o < p->n_oct; o++) {
        const size_t pl = (size_t)p->opitch[o] * p->oh[o] * sizeof(float);
        for (int i = 0; i < 6; i++) if ((rc = dalloc(p, &L.G[o][i], pl))) return rc;
        for (int i = 0; i < 5; i++) if ((rc = dalloc(p, &L.D[o][i], pl))) return rc;
    }
    memset(&L.table, 0, sizeof(L.table));
    for (int o = 0; o < p->n_oct; o++) {
        const size_t pl = (size_t)p->opitch[o] * p->oh[o] * sizeof(float);
        for (int i = 0; i < 3; i++) {
            if ((rc = dalloc(p, &L.gop[o][i], 2 * pl))) return rc;
            L.table.go[o][i] = L.gop[o][i];
        }
        L.table.pitch[o] = p->opitch[o];
        L.table.w[o] = p->ow[o];
        L.table.h[o] = p->oh[o];
        L.table.octsize[o] = 1 << o;
    }
    for (int o = 0; o < p->n_oct; o++)
        if ((rc = dalloc(p, &L.cands[o], (size_t)p->cand_cap[o] * sizeof(float4)))) return rc;
    if ((rc = dalloc(p, &L.kp, (size_t)p->kp_cap * sizeof(float4)))) return rc;
    if ((rc = dalloc(p, &L.kp_tag, (size_t)p->kp_cap * sizeof(int)))) return rc;
    if ((rc = dalloc(p, &L.kp_order, (size_t)p->kp_cap * sizeof(int)))) return rc;
    if (tb_get_encode()) {
        for (int o = 0; o < p->n_oct; o++)
            for (int s = 0; s < 5; s++)
                if (tb_supported(p->ntaps[s], s == kScales - 1 && o + 1 < p->n_oct ? TB_DOG_HALF : TB_DOG))
                    L.tmaps_ok[o][s] = tb_encode_pair(&L.tmaps[o][s], L.G[o][s], p->ow[o], p->oh[o], p->opitch[o], p->ntaps[s] >> 1) == 0;
        if (L.d_img && tb_supported(p->ntaps[5], TB_NORM) && p->w % 4 == 0)
            L.tmap_img_ok = tb_encode_pair(&L.tmap_img, L.d_img, p->w, p->h, p->w, p->ntaps[5] >> 1) == 0;
    }
    // device tables of the whole-pyramid launches (k_extrema_all, k_refine_all, k_gradient4_all)
    for (int s = 0; s < NSLOT; s++) {
        PyrTable t;
        memset(&t, 0, sizeof(t));
        t.n_oct = p->n_oct;
        int start = 0;
        for (int o = 0; o < p->n_oct; o++) {
            for (int i = 0; i < 5; i++) t.ds[o].d[i] = L.D[o][i];
            t.ds[o].pitch = p->opitch[o]; t.ds[o].w = p->ow[o]; t.ds[o].h = p->oh[o];
            t.cand[o] = L.cands[o];
            t.cap[o] = p->cand_cap[o];
            t.n_cand[o] = p->c_oct(s, o) + 0;
            t.n_kp_oct[o] = p->c_oct(s, o) + 1;
            t.stage[o] = p->c_stage(s, o);
            t.edthresh[o] = (1 << o) <= 1 ? kEdgeThresh1 : kEdgeThresh;  // plan.py:633-634, image.cl:195
            const bool has = p->ow[o] > 2 * kBorderDist && p->oh[o] > 2 * kBorderDist;
            t.ext_bx[o] = (p->ow[o] + 511) / 512;
            t.ext_start[o] = start;
            if (has) start += t.ext_bx[o] * ((p->oh[o] - 2 * kBorderDist + EXT_ROWS - 1) / EXT_ROWS);
        }
        for (int o = p->n_oct; o <= SIFTB_KOCT; o++) t.ext_start[o] = start;
        p->ext_blocks = start;
        if ((rc = dalloc(p, &L.d_pyr[s], sizeof(PyrTable)))) return rc;
        CK(cudaMemcpy(L.d_pyr[s], &t, sizeof(t), cudaMemcpyHostToDevice));
    }
    {
        GradTable g;
        memset(&g, 0, sizeof(g));
        int start = 0, n = 0;
        for (int o = 0; o < p->n_oct; o++)
            for (int i = 0; i < 3; i++, n++) {
                g.plane[n].g = L.G[o][i + 1]; g.plane[n].go = L.gop[o][i];
                g.plane[n].pitch = p->opitch[o]; g.plane[n].w = p->ow[o]; g.plane[n].h = p->oh[o];
                g.bx[n] = (p->ow[o] + 511) / 512;
                g.start[n] = start;
                start += g.bx[n] * ((p->oh[o] + GRAD4_ROWS - 1) / GRAD4_ROWS);
            }
        g.n_planes = n;
        for (int i = n; i <= GRAD_MAXPLANES; i++) g.start[i] = start;
        p->grad_blocks = start;
        if ((rc = dalloc(p, &L.d_grad, sizeof(GradTable)))) return rc;
        CK(cudaMemcpy(L.d_grad, &g, sizeof(g), cudaMemcpyHostToDevice));
    }
    L.ready = true;
    return 0;
}

static int plan_create_impl(siftb_plan *p) {
    DeviceGuard dg_(p->device);
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (auto &L : p->lanes) CK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
    for (int s = 0; s < NSLOT; s++) {
        CK(cudaEventCreateWithFlags(&p->ev_h2d[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev_done[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev_d2h[s], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev_hold[s], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&p->ev_ext, cudaEventDisableTiming));
    // plan.py:213-224 _calc_scales
    {
        int h = p->h, w = p->w, n = 0;
        const int min_size = 2 * kBorderDist + 2;
        p->ow[n] = w; p->oh[n] = h; n++;
        while ((h < w ? h : w) > min_size && n < MAX_OCT) {
            h /= 2; w /= 2;
            p->ow[n] = w; p->oh[n] = h; n++;
        }
        n--;
        p->n_oct = n;
        if (p->octave_limit > 0 && p->octave_limit < p->n_oct) p->n_oct = p->octave_limit;  // par.OctaveMax
    }
    if (p->n_oct < 1) return fail(SIFTB_EINVAL, "image too small: min(shape) must exceed 12 (plan.py:216)");
    for (int o = 0; o < p->n_oct; o++) p->opitch[o] = align_up(p->ow[o], 32);
    const size_t N = (size_t)p->h * p->w;
    p->kpsize = (int)(N / p->pix_per_kp);  // plan.py:243
    // plan.py:297-306 gaussian kernels
    const double sigmaRatio = pow(2.0, 1.0 / kScales);
    const double curSigma = 0.5;
    if (p->init_sigma > curSigma) {
        double s = sqrt(p->init_sigma * p->init_sigma - curSigma * curSigma);
        p->ntaps[5] = kernel_size(s);
        if (p->ntaps[5] > SIFTB_MAX_TAPS) return fail(SIFTB_EINVAL, "init_sigma too large");
        gaussian_taps(s, p->ntaps[5], p->taps[5].f);
        p->has_init = true;
    } else {
        p->ntaps[5] = 1;
        p->taps[5].f[0] = 1.0f;  // identity "blur": fmaf(x, 1, 0) == x
    }
    double prevSigma = p->init_sigma;
    for (int i = 0; i < kScales + 2; i++) {
        double increase = prevSigma * sqrt(sigmaRatio * sigmaRatio - 1.0);
        p->ntaps[i] = kernel_size(increase);
        if (p->ntaps[i] > SIFTB_MAX_TAPS) return fail(SIFTB_EINVAL, "init_sigma too large");
        gaussian_taps(increase, p->ntaps[i], p->taps[i].f);
        prevSigma *= sigmaRatio;
    }
    // buffers shared by the lanes (plan.py:268-295)
    p->raw_bytes = N * (dtype_bytes(p->dtype) > 4 ? dtype_bytes(p->dtype) : 4);  // fp32 input is always accepted
    int rc;
    for (int s = 0; s < NSLOT; s++) if ((rc = dalloc(p, &p->d_raws[s], p->raw_bytes))) return rc;
    if (p->n_oct > SIFTB_KOCT) return fail(SIFTB_EINVAL, "image too large: more than 16 octaves");
    for (int o = 0; o < p->n_oct; o++) {
        // plan.py:243: kpsize slots per octave; an octave cannot produce more than 3 candidates per pixel, so the
        // small octaves get by with less memory without changing what can overflow
        const long most = 3L * p->ow[o] * p->oh[o];
        p->cand_cap[o] = (int)(most < p->kpsize ? most : p->kpsize);
        if (p->cand_cap[o] < 1) p->cand_cap[o] = 1;
    }
    p->kp_cap = 2 * p->kpsize;
    if ((rc = dalloc(p, &p->d_queue, NSLOT * AUX_INTS * sizeof(int)))) return rc;
    p->out_cap = 2 * p->kpsize;
    for (int s = 0; s < NSLOT; s++) if ((rc = dalloc(p, &p->outs[s], (size_t)p->out_cap * sizeof(KpRecord)))) return rc;
    if (tb_get_encode() && tb_supported(p->ntaps[5], TB_NORM) && p->w % 4 == 0) {
        p->tmap_raw_ok = true;
        for (int s = 0; s < NSLOT; s++)
            p->tmap_raw_ok = p->tmap_raw_ok && tb_encode_pair(&p->tmap_raws[s], (const float *)p->d_raws[s], p->w, p->h,
                                                              p->w, p->ntaps[5] >> 1) == 0;
    }
    p->cnt_ints = 1 + 13 * p->n_oct + 2 + 2;  // ... + min/max + {refined, extra} totals
    for (int s = 0; s < NSLOT; s++) {
        if ((rc = dalloc(p, &p->d_cnts[s], p->cnt_ints * sizeof(int)))) return rc;
        CK(cudaHostAlloc((void **)&p->h_cnts[s], p->cnt_ints * sizeof(int), cudaHostAllocDefault));
        memset(p->h_cnts[s], 0, p->cnt_ints * sizeof(int));
    }
    CK(cudaFuncSetAttribute(k_blur_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)blur_generic_smem(SIFTB_MAX_TAPS - 1)));
    const char *lanes_env = getenv("SIFTB_LANES");
    if (lanes_env && atoi(lanes_env) >= 1 && atoi(lanes_env) <= NLANE) p->max_lanes = atoi(lanes_env);
    return lane_alloc(p, p->lanes[0]);
}

extern "C" int siftb_plan_create(int height, int width, int dtype, int device, int pix_per_kp, double init_sigma,
                                 int octave_max, siftb_plan **out) {
    if (!out) return fail(SIFTB_EINVAL, "out is null");
    *out = nullptr;
    if (height <= 0 || width <= 0) return fail(SIFTB_EINVAL, "bad shape");
    if (dtype_bytes(dtype) == 0) return fail(SIFTB_EINVAL, "invalid input format error (plan.py:488)");
    if (pix_per_kp <= 0) pix_per_kp = 10;
    if (!(init_sigma > 0.)) init_sigma = 1.6;
    siftb_plan *p = new siftb_plan();
    p->device = device; p->h = height; p->w = width; p->dtype = dtype;
    p->pix_per_kp = pix_per_kp; p->init_sigma = init_sigma;
    p->octave_limit = octave_max;
    p->force_generic = env_force_generic();
    int rc = plan_create_impl(p);
    if (rc) {
        std::string keep = g_siftb_err;
        siftb_plan_destroy(p);
        g_siftb_err = keep;
        return rc;
    }
    *out = p;
    return 0;
}

extern "C" int siftb_plan_octaves(const siftb_plan *p) { return p ? p->n_oct : SIFTB_EINVAL; }
extern "C" int siftb_plan_kpsize(const siftb_plan *p) { return p ? p->kpsize : SIFTB_EINVAL; }
extern "C" int siftb_plan_capacity(const siftb_plan *p) { return p ? p->out_cap : SIFTB_EINVAL; }
extern "C" int siftb_plan_octave_shape(const siftb_plan *p, int o, int *w, int *h) {
    if (!p || o < 0 || o >= p->n_oct) return fail(SIFTB_EINVAL, "bad octave");
    if (w) *w = p->ow[o];
    if (h) *h = p->oh[o];
    return 0;
}
extern "C" uint64_t siftb_plan_device_bytes(const siftb_plan *p) { return p ? p->dev_bytes : 0; }
extern "C" void *siftb_plan_stream(const siftb_plan *p) { return p ? (void *)p->stream : nullptr; }
extern "C" uint64_t siftb_plan_launches(const siftb_plan *p) { return p ? p->launches : 0; }
// plan.py:667-725 picks orientation_gpu.cl / keypoints_gpu2.cl for devicetype "GPU" and the *_cpu.cl kernels for "CPU";
// the two families give slightly different numbers.  0 = CPU-variant semantics (default,
// the parity target), 1 = GPU-variant semantics.
extern "C" int siftb_plan_set_variant(siftb_plan *p, int variant) {
    if (!p || variant < 0 || variant > 1) return fail(SIFTB_EINVAL, "variant must be 0 (cpu) or 1 (gpu)");
    std::lock_guard<std::mutex> lk(p->mtx);
    if (p->n_flight) return fail(SIFTB_EINVAL, "images are in flight");
    p->variant = variant;
    return 0;
}
extern "C" int siftb_plan_set_profile(siftb_plan *p, int enable) {
    if (!p) return fail(SIFTB_EINVAL, "plan is null");
    p->profile = enable != 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// launch helpers (shared by the pipeline and the stage hooks)
// Chooses the TMA-staged specialised kernel when one exists for this tap count and the planes meet the
// alignment rules of TMA / float2 stores; otherwise the generic kernel (identical results).
// premap: tensor map already encoded for (in, w, h, in_pitch, C) or null (encoded here).
static int launch_blur(cudaStream_t st, const float *in, int in_pitch, int w, int h, float *outG, int out_pitch,
                       float *outD, float *outHalf, int half_pitch, const Taps &taps, int ntaps,
                       const unsigned *norm_mm, const TbMaps *premap = nullptr, int force_generic = 0) {
    BlurArgs a;
    a.in = in; a.in_pitch = in_pitch; a.outG = outG; a.out_pitch = out_pitch; a.outD = outD;
    a.outHalf = outHalf; a.half_pitch = half_pitch; a.half_w = w / 2; a.half_h = h / 2;
    a.w = w; a.h = h; a.ntaps = ntaps; a.norm_mm = norm_mm;
    int mode = -1;
    if (norm_mm && !outD && !outHalf) mode = TB_NORM;
    else if (!norm_mm && outD && outHalf) mode = TB_DOG_HALF;
    else if (!norm_mm && outD && !outHalf) mode = TB_DOG;
    const bool out_ok = (out_pitch % 4 == 0) && (((uintptr_t)outG & 15) == 0) && (!outD || ((uintptr_t)outD & 15) == 0);
    if (!force_generic && mode >= 0 && tb_supported(ntaps, mode) && tb_source_ok(in, in_pitch) && out_ok &&
        tb_get_encode()) {
        TbMaps local;
        if (!premap) {
            if (tb_encode_pair(&local, in, w, h, in_pitch, ntaps >> 1)) return fail(SIFTB_ECUDA, "cuTensorMapEncodeTiled failed");
            premap = &local;
        }
        CK(tb_launch(st, premap->full, premap->cont, a, taps, mode));
        return 0;
    }
    dim3 grid((w + BLUR_TW - 1) / BLUR_TW, (h + BLUR_TH - 1) / BLUR_TH);
    k_blur_generic<<<grid, BLUR_THREADS, blur_generic_smem(ntaps), st>>>(a, taps);
    CKL();
    return 0;
}

static int launch_minmax_f32(cudaStream_t st, const float *img, long n, unsigned *mm) {
    k_minmax_reset<<<1, 1, 0, st>>>(mm);
    int vec4 = (n % 4 == 0) && (((uintptr_t)img & 15) == 0);
    long work = vec4 ? n / 4 : n;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_minmax_f32<<<blocks, 256, 0, st>>>(img, n, vec4, mm);
    CKL();
    return 0;
}

static int launch_convert(cudaStream_t st, const void *raw, int dtype, long n, float *out, unsigned *mm) {
    k_minmax_reset<<<1, 1, 0, st>>>(mm);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    switch (dtype) {
    case SIFTB_F32: k_convert_minmax<float, false><<<blocks, 256, 0, st>>>((const float *)raw, out, n, mm); break;
    case SIFTB_U8: k_convert_minmax<uint8_t, false><<<blocks, 256, 0, st>>>((const uint8_t *)raw, out, n, mm); break;
    case SIFTB_U16: k_convert_minmax<uint16_t, false><<<blocks, 256, 0, st>>>((const uint16_t *)raw, out, n, mm); break;
    case SIFTB_U32: k_convert_minmax<uint32_t, false><<<blocks, 256, 0, st>>>((const uint32_t *)raw, out, n, mm); break;
    case SIFTB_U64: k_convert_minmax<unsigned long long, false><<<blocks, 256, 0, st>>>((const unsigned long long *)raw, out, n, mm); break;
    case SIFTB_I32: k_convert_minmax<int32_t, false><<<blocks, 256, 0, st>>>((const int32_t *)raw, out, n, mm); break;
    case SIFTB_I64: k_convert_minmax<long long, false><<<blocks, 256, 0, st>>>((const long long *)raw, out, n, mm); break;
    case SIFTB_F64: k_convert_minmax<double, false><<<blocks, 256, 0, st>>>((const double *)raw, out, n, mm); break;
    case SIFTB_RGB8: k_convert_minmax<uint8_t, true><<<blocks, 256, 0, st>>>((const uint8_t *)raw, out, n, mm); break;
    default: return fail(SIFTB_EINVAL, "invalid input format error (plan.py:488)");
    }
    CKL();
    return 0;
}

static DogStack make_dogstack(float *const D[5], int pitch, int w, int h) {
    DogStack s;
    for (int i = 0; i < 5; i++) s.d[i] = D[i];
    s.pitch = pitch; s.w = w; s.h = h;
    return s;
}

// 3x3x3 extrema of one octave: the 4-columns-per-thread kernel on aligned planes, else the scalar one
static int launch_extrema(cudaStream_t st, const DogStack &ds, float gate, float edthresh, float4 *cand, int cap,
                          int *n_cand, int *stage, int scale_lo, int nscales) {
    if (!(ds.w > 2 * kBorderDist && ds.h > 2 * kBorderDist)) return 0;
    bool aligned = ds.pitch % 4 == 0;
    for (int i = 0; i < 5; i++) aligned = aligned && (((uintptr_t)ds.d[i] & 15) == 0);
    const int rows = (ds.h - 2 * kBorderDist + EXT_ROWS - 1) / EXT_ROWS;
    if (aligned) {
        dim3 grid((ds.w + 511) / 512, rows);
        k_extrema<<<grid, 128, 0, st>>>(ds, kBorderDist, gate, edthresh, cand, cap, n_cand, stage, scale_lo, nscales);
    } else {
        dim3 grid((ds.w + 127) / 128, rows);
        k_extrema_scalar<<<grid, 128, 0, st>>>(ds, kBorderDist, gate, edthresh, cand, cap, n_cand, stage, scale_lo, nscales);
    }
    CKL();
    return 0;
}

struct ProfScope {
    siftb_plan *p;
    int idx = -1;
    ProfScope(siftb_plan *p_, const char *name, int o = -1) : p(p_) {
        if (!p->profile) return;
        Event e;
        e.name = name;
        if (o >= 0) e.name += " octave " + std::to_string(o);
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        cudaEventRecord(e.a, p->lanes[p->last_lane].stream);
        p->events_s[p->cur].push_back(e);
        idx = (int)p->events_s[p->cur].size() - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(p->events_s[p->cur][idx].b, p->lanes[p->last_lane].stream);
    }
};

// plan.py:432-543 + :596-756, everything enqueued on the stream of one of the plan's lanes
static int submit_impl(siftb_plan *p, const void *image, int flags) {
    const int on_device = flags & SIFTB_ON_DEVICE;
    const int dtype = (flags & SIFTB_IS_F32) ? SIFTB_F32 : p->dtype;
    DeviceGuard dg_(p->device);
    const long N = (long)p->h * p->w;
    const void *src = image;
    const int slot = (p->head + p->n_flight) % NSLOT;
    p->cur = slot;
    // lane: an image submitted while others are in flight takes the lane the previous one did not (see Lane)
    int li = 0;
    if (p->n_flight > 0 && p->max_lanes > 1 && !p->profile) {
        li = (p->last_lane + 1) % p->max_lanes;
        if (!p->lanes[li].ready && lane_alloc(p, p->lanes[li])) {
            // no memory for a second set of planes: give back what was allocated and stay on one lane
            lane_release(p, p->lanes[li]);
            cudaGetLastError();
            p->max_lanes = 1;
            li = 0;
        }
    }
    siftb_plan::Lane &L = p->lanes[li];
    p->last_lane = li;
    p->slot_lane[slot] = li;
    cudaStream_t st = L.stream;
    for (auto &e : p->events_s[slot]) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    p->events_s[slot].clear();
    if (!on_device) {
        // H->D on the copy stream: overlaps the kernels of the previous image (pinned host memory)
        CK(cudaMemcpyAsync(p->d_raws[slot], image, (size_t)N * dtype_bytes(dtype), cudaMemcpyHostToDevice,
                           p->copy_stream));
        CK(cudaEventRecord(p->ev_h2d[slot], p->copy_stream));
        CK(cudaStreamWaitEvent(st, p->ev_h2d[slot], 0));
        src = p->d_raws[slot];
    }
    p->src_ptr[slot] = src;
    p->src_dtype[slot] = dtype;
    // the records of this slot's previous image must have left the device before they are overwritten
    CK(cudaStreamWaitEvent(st, p->ev_d2h[slot], 0));
    if (p->held[slot]) {  // ... and external readers of that buffer must be done (siftb_plan_hold_records)
        CK(cudaStreamWaitEvent(st, p->ev_hold[slot], 0));
        p->held[slot] = false;
    }
    CK(cudaMemsetAsync(p->d_cnts[slot], 0, p->cnt_ints * sizeof(int), st));
    int *aux = p->d_queue + slot * AUX_INTS;
    CK(cudaMemsetAsync(aux, 0, AUX_INTS * sizeof(int), st));
    int *q_head = aux, *n_kp = aux + 1, *n_extra = aux + 2, *n_order = aux + 3;
    int *oct_valid = aux + 8, *oct_offset = aux + 8 + SIFTB_KOCT, *oct_fill = aux + 8 + 2 * SIFTB_KOCT;
    int *size_hist = aux + 8 + 3 * SIFTB_KOCT, *size_start = size_hist + DESC_CLASSES, *size_fill = size_start + DESC_CLASSES;
    unsigned *mm = p->c_mm(slot);
    const float *img;
    int rc;
    {
        ProfScope ps(p, dtype == SIFTB_F32 ? "max_min" : "convert -> float + max_min");
        if (dtype == SIFTB_F32) {
            if ((rc = launch_minmax_f32(st, (const float *)src, N, mm))) return rc;
            p->launches += 2;
            img = (const float *)src;
        } else {
            if ((rc = launch_convert(st, src, dtype, N, L.d_img, mm))) return rc;
            p->launches += 2;
            img = L.d_img;
        }
    }
    {   // normalize fused into the initial blur (sigma = sqrt(init^2 - 0.5^2)), plan.py:525-539
        ProfScope ps(p, "normalize + init blur");
        const TbMaps *pm = nullptr;
        if (img == (const float *)p->d_raws[slot] && p->tmap_raw_ok) pm = &p->tmap_raws[slot];
        else if (img == L.d_img && L.tmap_img_ok) pm = &L.tmap_img;
        if ((rc = launch_blur(st, img, p->w, p->w, p->h, L.G[0][0], p->opitch[0], nullptr, nullptr, 0, p->taps[5],
                              p->ntaps[5], mm, pm, p->force_generic)))
            return rc;
        p->launches += 1;
    }
    // the whole pyramid first: five blur + DoG launches per octave, chained (the s = 2 launch also writes the next
    // octave's base); then extrema, refinement and gradient planes of ALL octaves with one launch each
    for (int o = 0; o < p->n_oct; o++) {
        const int w = p->ow[o], h = p->oh[o], pitch = p->opitch[o];
        ProfScope ps(p, "blur + DoG", o);
        for (int s = 0; s < kScales + 2; s++) {
            float *half = nullptr;
            int hp = 0;
            if (s == kScales - 1 && o + 1 < p->n_oct) { half = L.G[o + 1][0]; hp = p->opitch[o + 1]; }
            if ((rc = launch_blur(st, L.G[o][s], pitch, w, h, L.G[o][s + 1], pitch, L.D[o][s], half, hp, p->taps[s],
                                  p->ntaps[s], nullptr, L.tmaps_ok[o][s] ? &L.tmaps[o][s] : nullptr,
                                  p->force_generic)))
                return rc;
            p->launches += 1;
        }
    }
    if (p->ext_blocks > 0) {
        ProfScope ps(p, "local_maxmin");
        k_extrema_all<<<p->ext_blocks, 128, 0, st>>>(L.d_pyr[slot], kBorderDist, contrast_gate(kPeakThresh));
        CKL();
        p->launches += 1;
    }
    {
        ProfScope ps(p, "interp_keypoint + compact");
        k_refine_all<<<148 * 8, 128, 0, st>>>(L.d_pyr[slot], kPeakThresh, (float)p->init_sigma, L.kp, L.kp_tag,
                                              p->kp_cap, n_kp);
        CKL();
        p->launches += 1;
    }
    {
        ProfScope ps(p, "compute_gradient_orientation");
        k_gradient4_all<<<p->grad_blocks, 128, 0, st>>>(L.d_grad);
        CKL();
        p->launches += 1;
    }
    // once per image: orientation assignment and descriptors over the keypoints of all octaves
    {
        ProfScope ps(p, "orientation_assignment");
        if (p->variant)
            k_orient<true><<<148 * 5, 256, 0, st>>>(L.table, L.kp, L.kp_tag, n_kp, n_extra, p->kp_cap, kOriSigma,
                                                    p->c_stage(slot, 0), oct_valid, size_hist, aux + 4);
        else
            k_orient<false><<<148 * 5, 256, 0, st>>>(L.table, L.kp, L.kp_tag, n_kp, n_extra, p->kp_cap, kOriSigma,
                                                     p->c_stage(slot, 0), oct_valid, size_hist, aux + 4);
        CKL();
        p->launches += 1;
    }
    {
        ProfScope ps(p, "descriptors");
        k_size_order<<<148, 256, 0, st>>>(L.kp, L.kp_tag, n_kp, n_extra, p->kp_cap, size_hist, size_fill, L.kp_order,
                                          n_order, oct_valid, p->n_oct, oct_offset, p->c_nout(slot),
                                          p->c_oct(slot, 0) + 3);
        if (p->variant)
            k_describe<true><<<148 * 7, DESC_WARPS * 32, 0, st>>>(L.table, L.kp, L.kp_tag, n_order, p->kp_cap,
                                                                  p->outs[slot], p->out_cap, oct_offset, oct_fill,
                                                                  q_head, L.kp_order);
        else
            k_describe<false><<<148 * 7, DESC_WARPS * 32, 0, st>>>(L.table, L.kp, L.kp_tag, n_order, p->kp_cap,
                                                                   p->outs[slot], p->out_cap, oct_offset, oct_fill,
                                                                   q_head, L.kp_order);
        CKL();
        p->launches += 2;
    }
    CK(cudaMemcpyAsync(p->d_cnts[slot] + 1 + 13 * p->n_oct + 2, n_kp, 2 * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(p->h_cnts[slot], p->d_cnts[slot], p->cnt_ints * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(p->ev_done[slot], st));
    CK(cudaStreamWaitEvent(p->stream, p->ev_done[slot], 0));  // the public queue is ordered after every submitted image
    p->n_flight++;
    return 0;
}

static int collect_impl(siftb_plan *p, siftb_kp *out, int cap, int *n_out, int *n_per_octave, float *minmax) {
    if (p->n_flight == 0) return fail(SIFTB_EINVAL, "collect without submit");
    DeviceGuard dg_(p->device);
    const int slot = p->head;
    CK(cudaEventSynchronize(p->ev_done[slot]));
    p->head = (p->head + 1) % NSLOT;
    p->n_flight--;
    p->last = slot;
    const int *h_cnt = p->h_cnts[slot];
    const int n = h_cnt[0];
    int rc = 0;
    int ncopy = n;
    if (ncopy > p->out_cap) { ncopy = p->out_cap; rc = SIFTB_EOVERFLOW; }
    if (out && ncopy > cap) { ncopy = cap; rc = SIFTB_EOVERFLOW; }
    for (int o = 0; o < p->n_oct; o++) {
        const int *c = h_cnt + 1 + 4 * o;
        if (c[0] > p->kpsize || c[1] > p->kpsize) rc = SIFTB_EOVERFLOW;  // per-octave slots, plan.py:243
        if (n_per_octave) n_per_octave[o] = c[3];
    }
    {
        const int *tot = h_cnt + 1 + 13 * p->n_oct + 2;
        if (tot[0] + tot[1] > p->kp_cap) rc = SIFTB_EOVERFLOW;
    }
    if (minmax) {
        const unsigned *mm = reinterpret_cast<const unsigned *>(h_cnt + 1 + 13 * p->n_oct);
        unsigned u0 = mm[0], u1 = mm[1];
        uint32_t a = (u0 & 0x80000000u) ? (u0 & 0x7fffffffu) : ~u0, b = (u1 & 0x80000000u) ? (u1 & 0x7fffffffu) : ~u1;
        memcpy(&minmax[0], &a, 4);
        memcpy(&minmax[1], &b, 4);
    }
    if (out && ncopy > 0) {
        // D->H on its own stream: the compute stream may already be running the next image and the upload stream
        // may hold the images after that (waiting for THEM here serialised upload and download: at 8 GPUs per host,
        // where an upload takes longer than the kernels, that doubled the end-to-end step)
        CK(cudaMemcpyAsync(out, p->outs[slot], (size_t)ncopy * sizeof(siftb_kp), cudaMemcpyDeviceToHost,
                           p->d2h_stream));
        CK(cudaEventRecord(p->ev_d2h[slot], p->d2h_stream));
        CK(cudaStreamSynchronize(p->d2h_stream));
    }
    if (n_out) *n_out = n;
    if (rc == SIFTB_EOVERFLOW) return fail(rc, "keypoint buffer overflow (reference: plan.py:771 warning)");
    return 0;
}

extern "C" int siftb_plan_submit(siftb_plan *p, const void *image, int flags) {
    if (!p || !image) return fail(SIFTB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(p->mtx);
    if (p->n_flight >= NSLOT) return fail(SIFTB_EINVAL, "three submits are already in flight on this plan");
    return submit_impl(p, image, flags);
}
extern "C" int siftb_plan_collect(siftb_plan *p, siftb_kp *out, int cap, int *n_out, int *n_per_octave,
                                  float *minmax) {
    if (!p) return fail(SIFTB_EINVAL, "null plan");
    std::lock_guard<std::mutex> lk(p->mtx);
    return collect_impl(p, out, cap, n_out, n_per_octave, minmax);
}
extern "C" int siftb_plan_keypoints(siftb_plan *p, const void *image, int flags, siftb_kp *out, int cap,
                                    int *n_out, int *n_per_octave, float *minmax) {
    if (!p || !image) return fail(SIFTB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(p->mtx);
    if (p->n_flight != 0) return fail(SIFTB_EINVAL, "a submit is already in flight on this plan");
    int rc = submit_impl(p, image, flags);
    if (rc) return rc;
    return collect_impl(p, out, cap, n_out, n_per_octave, minmax);
}
// host copy of the records of the most recently collected run (after collect(out = NULL))
extern "C" int siftb_plan_fetch_records(siftb_plan *p, siftb_kp *out, int cap, int *n_out) {
    if (!p || !out) return fail(SIFTB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(p->mtx);
    DeviceGuard dg_(p->device);
    const int slot = p->last;
    int n = p->h_cnts[slot][0];
    if (n > p->out_cap) n = p->out_cap;
    if (n > cap) n = cap;
    if (n_out) *n_out = n;
    if (n > 0) {
        CK(cudaMemcpyAsync(out, p->outs[slot], (size_t)n * sizeof(siftb_kp), cudaMemcpyDeviceToHost, p->d2h_stream));
        CK(cudaEventRecord(p->ev_d2h[slot], p->d2h_stream));
        CK(cudaStreamSynchronize(p->d2h_stream));
    }
    return 0;
}
extern "C" int siftb_plan_wait_stream(siftb_plan *p, void *stream) {
    if (!p) return fail(SIFTB_EINVAL, "null plan");
    std::lock_guard<std::mutex> lk(p->mtx);
    DeviceGuard dg_(p->device);
    CK(cudaEventRecord(p->ev_ext, (cudaStream_t)stream));
    CK(cudaStreamWaitEvent(p->stream, p->ev_ext, 0));
    for (auto &L : p->lanes) CK(cudaStreamWaitEvent(L.stream, p->ev_ext, 0));
    return 0;
}
extern "C" int siftb_plan_hold_records(siftb_plan *p, void *stream) {
    if (!p) return fail(SIFTB_EINVAL, "null plan");
    std::lock_guard<std::mutex> lk(p->mtx);
    DeviceGuard dg_(p->device);
    CK(cudaEventRecord(p->ev_hold[p->last], (cudaStream_t)stream));
    p->held[p->last] = true;
    return 0;
}
extern "C" int siftb_plan_device(const siftb_plan *p) { return p ? p->device : SIFTB_EINVAL; }
extern "C" int siftb_plan_result_dev(const siftb_plan *p, const siftb_kp **recs, const int **count) {
    if (!p) return fail(SIFTB_EINVAL, "null plan");
    if (recs) *recs = reinterpret_cast<const siftb_kp *>(p->outs[p->last]);
    if (count) *count = p->d_cnts[p->last];
    return 0;
}
// alignment.py:324-349: warp of the image of the most recently collected run.  That image is still on the device
// (the plan's input staging buffer, or the caller's device array), so LinearAlign uploads a frame once for both
// keypoints() and the warp, like the reference does with buffers["input"] (alignment.py:242-246).
extern "C" int siftb_plan_warp_last(siftb_plan *p, const float matrix[4], const float offset[2], float fill, int mode,
                                    void *out, int out_height, int out_width, int out_on_device) {
    if (!p || !matrix || !offset || !out || out_height <= 0 || out_width <= 0) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(p->mtx);
    DeviceGuard dg_(p->device);
    const int slot = p->last;
    const void *src = p->src_ptr[slot];
    const int dtype = p->src_dtype[slot];
    if (!src) return fail(SIFTB_EINVAL, "no image has been processed by this plan yet");
    if (dtype != SIFTB_F32 && dtype != SIFTB_RGB8)
        return fail(SIFTB_EINVAL, "the warp needs a float32 or RGB8 image (alignment.py:237-240)");
    const size_t bytes = (size_t)out_height * out_width * (dtype == SIFTB_RGB8 ? 3 : 4);
    void *dst = out;
    if (!out_on_device) {
        CK(p->d_warp.reserve(bytes));
        dst = p->d_warp.p;
    }
    const WarpMap m = make_warp_map(matrix, offset, p->h, p->w, out_height, out_width, fill, mode);
    p->cur = slot;
    {
        ProfScope ps(p, "transform");
        if (dtype == SIFTB_RGB8) CK(launch_warp_rgb8(p->stream, (const uint8_t *)src, (uint8_t *)dst, m));
        else CK(launch_warp_f32(p->stream, (const float *)src, (float *)dst, m));
        p->launches += 1;
    }
    if (!out_on_device) {
        ProfScope ps(p, "copy D->H transformed image");
        CK(cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    return 0;
}
extern "C" int siftb_plan_events(siftb_plan *p, const char *const **names, const float **ms, int *n) {
    if (!p) return fail(SIFTB_EINVAL, "null plan");
    std::lock_guard<std::mutex> lk(p->mtx);
    DeviceGuard dg_(p->device);
    // the events of the most recently collected image: all recorded before its ev_done, which collect() waited for
    // (no stream synchronisation here: the next images may already be running)
    auto &events = p->events_s[p->last];
    p->ev_names.clear();
    p->ev_ms.clear();
    for (auto &e : events) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        p->ev_names.push_back(e.name.c_str());
        p->ev_ms.push_back(t);
    }
    if (names) *names = p->ev_names.data();
    if (ms) *ms = p->ev_ms.data();
    if (n) *n = (int)events.size();
    return 0;
}
extern "C" int siftb_plan_stage_counts(siftb_plan *p, int *counts) {
    if (!p || !counts) return fail(SIFTB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(p->mtx);
    memcpy(counts, p->h_cnts[p->last] + 1 + 4 * p->n_oct, sizeof(int) * 9 * p->n_oct);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// stage-level hooks: host in, host out, temporaries on the current device, default stream
extern "C" int siftb_gauss_taps(double sigma, float *taps, int cap, int *n) {
    if (!taps || !n || !(sigma > 0)) return fail(SIFTB_EINVAL, "bad argument");
    int size = kernel_size(sigma);
    if (size > cap) return fail(SIFTB_EINVAL, "taps buffer too small");
    gaussian_taps(sigma, size, taps);
    *n = size;
    return 0;
}

extern "C" int siftb_minmax(const float *image, int height, int width, float *minimum, float *maximum) {
    if (!image || height <= 0 || width <= 0) return fail(SIFTB_EINVAL, "bad argument");
    const long n = (long)height * width;
    DevBuf d, mm;
    DALLOC(d, n * 4); DALLOC(mm, 8);
    CK(cudaMemcpy(d.p, image, n * 4, cudaMemcpyHostToDevice));
    int rc = launch_minmax_f32(0, d.as<float>(), n, mm.as<unsigned>());
    if (rc) return rc;
    unsigned h[2];
    CK(cudaMemcpy(h, mm.p, 8, cudaMemcpyDeviceToHost));
    uint32_t a = (h[0] & 0x80000000u) ? (h[0] & 0x7fffffffu) : ~h[0], b = (h[1] & 0x80000000u) ? (h[1] & 0x7fffffffu) : ~h[1];
    if (minimum) memcpy(minimum, &a, 4);
    if (maximum) memcpy(maximum, &b, 4);
    return 0;
}

extern "C" int siftb_normalize(const float *image, int height, int width, float *out) {
    if (!image || !out || height <= 0 || width <= 0) return fail(SIFTB_EINVAL, "bad argument");
    const long n = (long)height * width;
    DevBuf d, o, mm;
    DALLOC(d, n * 4); DALLOC(o, n * 4); DALLOC(mm, 8);
    CK(cudaMemcpy(d.p, image, n * 4, cudaMemcpyHostToDevice));
    int rc = launch_minmax_f32(0, d.as<float>(), n, mm.as<unsigned>());
    if (rc) return rc;
    k_normalize<<<148 * 8, 256>>>(d.as<float>(), o.as<float>(), n, mm.as<unsigned>());
    CKL();
    CK(cudaMemcpy(out, o.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_to_float(const void *image, int dtype, int height, int width, float *out) {
    if (!image || !out || height <= 0 || width <= 0 || !dtype_bytes(dtype)) return fail(SIFTB_EINVAL, "bad argument");
    const long n = (long)height * width;
    DevBuf d, o, mm;
    DALLOC(d, n * dtype_bytes(dtype)); DALLOC(o, n * 4); DALLOC(mm, 8);
    CK(cudaMemcpy(d.p, image, n * dtype_bytes(dtype), cudaMemcpyHostToDevice));
    int rc = launch_convert(0, d.p, dtype, n, o.as<float>(), mm.as<unsigned>());
    if (rc) return rc;
    CK(cudaMemcpy(out, o.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int ensure_blur_attr() {
    CK(cudaFuncSetAttribute(k_blur_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)blur_generic_smem(SIFTB_MAX_TAPS - 1)));
    return 0;
}

extern "C" int siftb_blur(const float *image, int height, int width, const float *taps, int ntaps, float *out) {
    if (!image || !out || !taps || height <= 0 || width <= 0) return fail(SIFTB_EINVAL, "bad argument");
    if (ntaps < 1 || ntaps > SIFTB_MAX_TAPS || !(ntaps & 1)) return fail(SIFTB_EINVAL, "ntaps must be odd and <= 63");
    if (ntaps / 2 > (height < width ? height : width)) return fail(SIFTB_EINVAL, "kernel wider than the image");
    int rc = ensure_blur_attr();
    if (rc) return rc;
    const int pitch = align_up(width, 32);  // same plane layout as the plan's buffers
    const size_t pb = (size_t)pitch * height * 4, rb = (size_t)width * 4;
    DevBuf d, o, dd;
    DALLOC(d, pb); DALLOC(o, pb); DALLOC(dd, pb);
    CK(cudaMemcpy2D(d.p, (size_t)pitch * 4, image, rb, rb, height, cudaMemcpyHostToDevice));
    Taps t;
    memset(&t, 0, sizeof(t));
    memcpy(t.f, taps, ntaps * sizeof(float));
    rc = launch_blur(0, d.as<float>(), pitch, width, height, o.as<float>(), pitch, dd.as<float>(), nullptr, 0, t, ntaps,
                     nullptr, nullptr, env_force_generic());
    if (rc) return rc;
    CK(cudaMemcpy2D(out, rb, o.p, (size_t)pitch * 4, rb, height, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_pyramid_octave(const float *g0, int height, int width, double init_sigma, float *G5, float *D5,
                                    float *next_base) {
    if (!g0 || height <= 0 || width <= 0) return fail(SIFTB_EINVAL, "bad argument");
    int rc = ensure_blur_attr();
    if (rc) return rc;
    const int pitch = align_up(width, 32);
    const long n = (long)pitch * height;  // pitched planes, as in the plan
    const size_t rb = (size_t)width * 4, pbytes = (size_t)pitch * 4;
    const int hw = width / 2, hh = height / 2;
    DevBuf G, D, half;
    DALLOC(G, 6 * n * 4); DALLOC(D, 5 * n * 4); DALLOC(half, (size_t)hw * hh * 4);
    CK(cudaMemcpy2D(G.p, pbytes, g0, rb, rb, height, cudaMemcpyHostToDevice));
    const double sigmaRatio = pow(2.0, 1.0 / kScales);
    double prevSigma = init_sigma;
    for (int s = 0; s < kScales + 2; s++) {
        double increase = prevSigma * sqrt(sigmaRatio * sigmaRatio - 1.0);
        Taps t;
        memset(&t, 0, sizeof(t));
        int nt = kernel_size(increase);
        if (nt > SIFTB_MAX_TAPS) return fail(SIFTB_EINVAL, "init_sigma too large");
        gaussian_taps(increase, nt, t.f);
        prevSigma *= sigmaRatio;
        rc = launch_blur(0, G.as<float>() + s * n, pitch, width, height, G.as<float>() + (s + 1) * n, pitch,
                         D.as<float>() + s * n, (s == kScales - 1 && hw > 0 && hh > 0) ? half.as<float>() : nullptr, hw,
                         t, nt, nullptr, nullptr, env_force_generic());
        if (rc) return rc;
    }
    if (G5) CK(cudaMemcpy2D(G5, rb, G.as<float>() + n, pbytes, rb, (size_t)5 * height, cudaMemcpyDeviceToHost));
    if (D5) CK(cudaMemcpy2D(D5, rb, D.p, pbytes, rb, (size_t)5 * height, cudaMemcpyDeviceToHost));
    if (next_base && hw > 0 && hh > 0) CK(cudaMemcpy(next_base, half.p, (size_t)hw * hh * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_gradient(const float *image, int height, int width, float *grad, float *ori) {
    if (!image || !grad || !ori || height < 2 || width < 2) return fail(SIFTB_EINVAL, "bad argument");
    const long n = (long)height * width;
    DevBuf d, g, o, go;
    DALLOC(d, n * 4); DALLOC(g, n * 4); DALLOC(o, n * 4); DALLOC(go, n * 8);
    CK(cudaMemcpy(d.p, image, n * 4, cudaMemcpyHostToDevice));
    GradArgs ga;
    for (int i = 0; i < 3; i++) { ga.g[i] = d.as<float>(); ga.go[i] = go.as<float2>(); }
    ga.pitch = width; ga.w = width; ga.h = height;
    if (width % 4 == 0) {  // the form the pipeline uses
        dim3 grid((width + 511) / 512, (height + GRAD4_ROWS - 1) / GRAD4_ROWS, 1);
        k_gradient4<<<grid, 128>>>(ga);
    } else {
        dim3 grid((width + 255) / 256, (height + GRAD_ROWS - 1) / GRAD_ROWS, 1);
        k_gradient<<<grid, 256>>>(ga);
    }
    CKL();
    k_deinterleave<<<148 * 4, 256>>>(go.as<float2>(), n, g.as<float>(), o.as<float>());
    CKL();
    CK(cudaMemcpy(grad, g.p, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ori, o.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_local_maxmin(const float *dogs5, int height, int width, int scale, int octsize, float *kp4,
                                  int cap, int *n) {
    if (!dogs5 || !kp4 || !n || scale < 1 || scale > 3 || cap < 0) return fail(SIFTB_EINVAL, "bad argument");
    const long np = (long)height * width;
    DevBuf D, K, C;
    DALLOC(D, 5 * np * 4); DALLOC(K, (size_t)cap * 16); DALLOC(C, 4);
    CK(cudaMemcpy(D.p, dogs5, 5 * np * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(C.p, 0, 4));
    CK(cudaMemset(K.p, 0, (size_t)cap * 16));
    if (width > 2 * kBorderDist && height > 2 * kBorderDist) {
        float *Dp[5];
        for (int i = 0; i < 5; i++) Dp[i] = D.as<float>() + i * np;
        DogStack ds = make_dogstack(Dp, width, width, height);
        int rc = launch_extrema(0, ds, contrast_gate(kPeakThresh), octsize <= 1 ? kEdgeThresh1 : kEdgeThresh,
                                K.as<float4>(), cap, C.as<int>(), nullptr, scale, 1);
        if (rc) return rc;
    }
    CK(cudaMemcpy(n, C.p, 4, cudaMemcpyDeviceToHost));
    int m = *n < cap ? *n : cap;
    if (m > 0) CK(cudaMemcpy(kp4, K.p, (size_t)m * 16, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_interp(const float *dogs5, int height, int width, const float *kp4_in, int n_in, float init_sigma,
                            float *kp4_out, int *n_out) {
    if (!dogs5 || !kp4_in || !kp4_out || !n_out || n_in < 0) return fail(SIFTB_EINVAL, "bad argument");
    const long np = (long)height * width;
    DevBuf D, Kin, Kout, S, C;
    DALLOC(D, 5 * np * 4); DALLOC(Kin, (size_t)n_in * 16); DALLOC(Kout, (size_t)n_in * 16); DALLOC(S, (size_t)n_in * 4);
    DALLOC(C, 8);
    CK(cudaMemcpy(D.p, dogs5, 5 * np * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Kin.p, kp4_in, (size_t)n_in * 16, cudaMemcpyHostToDevice));
    int cnt[2] = {n_in, 0};
    CK(cudaMemcpy(C.p, cnt, 8, cudaMemcpyHostToDevice));
    float *Dp[5];
    for (int i = 0; i < 5; i++) Dp[i] = D.as<float>() + i * np;
    DogStack ds = make_dogstack(Dp, width, width, height);
    k_refine<<<148 * 4, 128>>>(ds, Kin.as<float4>(), C.as<int>(), n_in, kPeakThresh, init_sigma, Kout.as<float4>(),
                               S.as<int>(), n_in, C.as<int>() + 1, nullptr, 0, nullptr);
    CKL();
    CK(cudaMemcpy(cnt, C.p, 8, cudaMemcpyDeviceToHost));
    *n_out = cnt[1];
    if (cnt[1] > 0) CK(cudaMemcpy(kp4_out, Kout.p, (size_t)cnt[1] * 16, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_orientation_v(const float *kp4_in, int n, const float *grad, const float *ori, int height,
                                   int width, int octsize, float *kp4_out, int cap, int *n_out, int variant) {
    if (!kp4_in || !grad || !ori || !kp4_out || !n_out || n < 0 || cap < n) return fail(SIFTB_EINVAL, "bad argument");
    const long np = (long)height * width;
    DevBuf Gd, Od, GO, K, S, C;
    DALLOC(Gd, np * 4); DALLOC(Od, np * 4); DALLOC(GO, np * 8); DALLOC(K, (size_t)cap * 16); DALLOC(S, (size_t)cap * 4);
    DALLOC(C, 12);
    CK(cudaMemcpy(Gd.p, grad, np * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Od.p, ori, np * 4, cudaMemcpyHostToDevice));
    k_interleave<<<148 * 4, 256>>>(Gd.as<float>(), Od.as<float>(), np, GO.as<float2>());
    CKL();
    CK(cudaMemcpy(K.p, kp4_in, (size_t)n * 16, cudaMemcpyHostToDevice));
    std::vector<int> ones(cap > 0 ? cap : 1, 1);
    CK(cudaMemcpy(S.p, ones.data(), (size_t)cap * 4, cudaMemcpyHostToDevice));
    int cnt[3] = {n, 0, 0};  // keypoints in, extra orientations, work-queue head
    CK(cudaMemcpy(C.p, cnt, 12, cudaMemcpyHostToDevice));
    OctTable tb;  // a one-octave table; every row carries tag (0 << 8) | 1
    memset(&tb, 0, sizeof(tb));
    for (int i = 0; i < 3; i++) tb.go[0][i] = GO.as<float2>();
    tb.pitch[0] = width; tb.w[0] = width; tb.h[0] = height; tb.octsize[0] = octsize;
    if (variant)
        k_orient<true><<<148 * 5, 256>>>(tb, K.as<float4>(), S.as<int>(), C.as<int>(), C.as<int>() + 1, cap, kOriSigma,
                                         nullptr, nullptr, nullptr, C.as<int>() + 2);
    else
        k_orient<false><<<148 * 5, 256>>>(tb, K.as<float4>(), S.as<int>(), C.as<int>(), C.as<int>() + 1, cap, kOriSigma,
                                          nullptr, nullptr, nullptr, C.as<int>() + 2);
    CKL();
    CK(cudaMemcpy(cnt, C.p, 8, cudaMemcpyDeviceToHost));
    int total = n + cnt[1];
    *n_out = total;
    if (total > cap) total = cap;
    if (total > 0) CK(cudaMemcpy(kp4_out, K.p, (size_t)total * 16, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_orientation(const float *kp4_in, int n, const float *grad, const float *ori, int height, int width,
                                 int octsize, float *kp4_out, int cap, int *n_out) {
    return siftb_orientation_v(kp4_in, n, grad, ori, height, width, octsize, kp4_out, cap, n_out, 0);
}

extern "C" int siftb_descriptor_v(const float *kp4, int n, const float *grad, const float *ori, int height, int width,
                                  int octsize, uint8_t *desc, int variant) {
    if (!kp4 || !grad || !ori || !desc || n < 0) return fail(SIFTB_EINVAL, "bad argument");
    const long np = (long)height * width;
    DevBuf Gd, Od, GO, K, Dd;
    DALLOC(Gd, np * 4); DALLOC(Od, np * 4); DALLOC(GO, np * 8); DALLOC(K, (size_t)n * 16); DALLOC(Dd, (size_t)n * 128);
    CK(cudaMemcpy(Gd.p, grad, np * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Od.p, ori, np * 4, cudaMemcpyHostToDevice));
    k_interleave<<<148 * 4, 256>>>(Gd.as<float>(), Od.as<float>(), np, GO.as<float2>());
    CKL();
    CK(cudaMemcpy(K.p, kp4, (size_t)n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemset(Dd.p, 0, (size_t)n * 128));
    if (n > 0) {
        const int blocks = (n + 4 * DESC_WARPS - 1) / (4 * DESC_WARPS);  // 4 keypoints per warp
        if (variant)
            k_describe_rows<true><<<blocks, DESC_WARPS * 32>>>(GO.as<float2>(), width, width, height, K.as<float4>(), n,
                                                              octsize, Dd.as<uint8_t>());
        else
            k_describe_rows<false><<<blocks, DESC_WARPS * 32>>>(GO.as<float2>(), width, width, height, K.as<float4>(), n,
                                                               octsize, Dd.as<uint8_t>());
        CKL();
        CK(cudaMemcpy(desc, Dd.p, (size_t)n * 128, cudaMemcpyDeviceToHost));
    }
    return 0;
}


extern "C" int siftb_descriptor(const float *kp4, int n, const float *grad, const float *ori, int height, int width,
                                int octsize, uint8_t *desc) {
    return siftb_descriptor_v(kp4, n, grad, ori, height, width, octsize, desc, 0);
}
