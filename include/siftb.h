/*
 * siftb.h -- C ABI of libsiftb200.so: the B200 (sm_100a) implementation of the sift_pyocl
 * keypoint path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference has no FFI: its boundary is the Python class API (SiftPlan / MatchPlan /
 * LinearAlign) sitting directly on PyOpenCL enqueue calls.  Each entry point below replaces the
 * PyOpenCL call sequence cited next to it (paths relative to the reference root); the ctypes
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (SIFTB_E*); siftb_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - images are row-major, dense (stride == width), fp32 unless a dtype code says otherwise;
 *   - "host" pointers are ordinary CPU memory (pinned memory makes copies asynchronous),
 *     "dev" pointers are CUDA device memory of the plan's device;
 *   - a plan owns its device buffers, one CUDA stream and a mutex: calls on one plan serialise
 *     (reference: threading.Semaphore per plan, plan.py:156,439), different plans run concurrently;
 *   - keypoint records are the reference's dtype_kp (plan.py:110-115), 144 bytes.
 */
#ifndef SIFTB_H
#define SIFTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SIFTB_API __attribute__((visibility("default")))
#else
#define SIFTB_API
#endif

#define SIFTB_OK 0
#define SIFTB_EINVAL (-1)   /* bad argument (reference: RuntimeError / assert, plan.py:151,443,488) */
#define SIFTB_ECUDA (-2)    /* CUDA runtime error */
#define SIFTB_ENOMEM (-3)   /* allocation failure (reference: MemoryError, plan.py:365) */
#define SIFTB_EOVERFLOW (-4) /* keypoint buffer overflow (reference only warns, plan.py:771) */

/* input pixel types accepted by SiftPlan (plan.py:99-106, preprocess.cl:53-223) */
enum siftb_dtype {
    SIFTB_F32 = 0, SIFTB_U8 = 1, SIFTB_U16 = 2, SIFTB_U32 = 3, SIFTB_U64 = 4,
    SIFTB_I32 = 5, SIFTB_I64 = 6, SIFTB_F64 = 7, SIFTB_RGB8 = 8
};

#define SIFTB_HOST 0
#define SIFTB_ON_DEVICE 1
#define SIFTB_IS_F32 2

/* numpy dtype_kp, plan.py:110-115 */
typedef struct siftb_kp {
    float x, y, scale, angle;
    uint8_t desc[128];
} siftb_kp;

typedef struct siftb_plan siftb_plan;   /* replaces SiftPlan's ctx/queue/buffers, plan.py:117-201 */

SIFTB_API const char *siftb_last_error(void);
SIFTB_API int siftb_version(void);
SIFTB_API int siftb_device_count(int *n);                       /* replaces clinit.py:360 select_device */

/* ---- page-locked host memory for asynchronous copies -------------------------------------- */
SIFTB_API int siftb_host_alloc(void **ptr, uint64_t bytes);
/* write-combined variant for upload-only buffers (the CPU writes, the device reads; CPU reads are very slow) */
SIFTB_API int siftb_host_alloc_wc(void **ptr, uint64_t bytes);
SIFTB_API int siftb_host_free(void *ptr);

/* ---- SiftPlan ------------------------------------------------------------------------------ */
/* plan.py:117-201 (__init__: _calc_scales, _calc_memory, _allocate_buffers, _init_gaussian).
 * octave_max <= 0 means "all octaves" (par.OctaveMax default, param.py:52). */
SIFTB_API int siftb_plan_create(int height, int width, int dtype, int device, int pix_per_kp, double init_sigma,
                      int octave_max, siftb_plan **out);
SIFTB_API int siftb_plan_destroy(siftb_plan *plan);             /* plan.py:203-211 __del__ */

/* read-only facts about a plan: plan.py:213-245 (octave_max, kpsize, scales[o] = (w, h)) */
SIFTB_API int siftb_plan_octaves(const siftb_plan *plan);
SIFTB_API int siftb_plan_kpsize(const siftb_plan *plan);      /* per-octave keypoint slots, plan.py:243 */
SIFTB_API int siftb_plan_capacity(const siftb_plan *plan);    /* records the plan can return for one image (all octaves) */
SIFTB_API int siftb_plan_octave_shape(const siftb_plan *plan, int octave, int *width, int *height);
SIFTB_API uint64_t siftb_plan_device_bytes(const siftb_plan *plan);   /* plan.py:226 _calc_memory; grows once, when the
                                                                        * second lane is first needed */
/* cudaStream_t of the plan's queue: it is ordered after every image submitted so far (work enqueued on it, or an event
 * recorded on it, runs after their kernels).  The kernels themselves run on two private streams ("lanes", each with
 * its own planes: images that are in flight together are computed two at a time); to order a submit after your own
 * producer use siftb_plan_wait_stream, not this stream. */
SIFTB_API void *siftb_plan_stream(const siftb_plan *plan);
SIFTB_API int siftb_plan_set_profile(siftb_plan *plan, int enable);   /* plan.py:185-186 PROFILING_ENABLE */
/* Which of the reference's two kernel families the orientation / descriptor stages reproduce (plan.py:667-725 picks
 * by device type; they give different numbers): 0 = orientation_cpu.cl + keypoints_cpu.cl
 * (default: devicetype "CPU", the parity target), 1 = orientation_gpu.cl + keypoints_gpu2.cl (devicetype "GPU"). */
SIFTB_API int siftb_plan_set_variant(siftb_plan *plan, int variant);
SIFTB_API int siftb_plan_device(const siftb_plan *plan);               /* CUDA ordinal the plan lives on */
/* Orders the plan's queue after everything enqueued so far on `stream` (a cudaStream_t of the plan's device; the
 * handles 0x1 / 0x2 are CUDA's legacy / per-thread default streams): call it before handing a device-resident
 * image to keypoints()/submit() when that image is produced asynchronously on another stream, and after an
 * external stream has been asked to read siftb_plan_result_dev() memory.  Replaces the implicit ordering of the
 * reference's single in-order command queue shared with its pyopencl.array inputs (plan.py:185-188,451). */
SIFTB_API int siftb_plan_wait_stream(siftb_plan *plan, void *stream);
/* Narrower form for readers of siftb_plan_result_dev() memory: the record buffer of the most recently collected
 * image is not rewritten before the work enqueued so far on `stream` has finished.  Only the submit that recycles
 * that buffer (the third after the one that produced these records) waits; images already queued, and the next ones, are not held up. */
SIFTB_API int siftb_plan_hold_records(siftb_plan *plan, void *stream);
/* number of CUDA kernels this plan has launched since it was created (bench.py's gpu_launches) */
SIFTB_API uint64_t siftb_plan_launches(const siftb_plan *plan);

/* plan.py:432-567 keypoints(): the whole path, blocking.
 *   image      : height*width pixels of the plan's dtype (RGB8: height*width*3 bytes)
 *   flags      : SIFTB_HOST (0) = host pointer (copied H->D inside the call);
 *                SIFTB_ON_DEVICE = device pointer (reference: pyopencl.array.Array input, plan.py:451);
 *                SIFTB_IS_F32 = the pixels are float32 although the plan was built for another
 *                dtype (the reference accepts both, plan.py:444,450)
 *   out/cap    : caller-allocated records; at most cap are written
 *   n_out      : number of keypoints found (may exceed cap -> SIFTB_EOVERFLOW, out holds cap)
 *   n_per_octave: optional int[siftb_plan_octaves()] (the "in octave %i found %i kp" log, plan.py:543)
 *   minmax     : optional float[2] = image min, max (self.buffers["min"], used by alignment.py:345) */
SIFTB_API int siftb_plan_keypoints(siftb_plan *plan, const void *image, int flags, siftb_kp *out, int cap,
                         int *n_out, int *n_per_octave, float *minmax);

/* The same path split in two so that callers can overlap the copy of image k+1 with the kernels of
 * image k: submit() enqueues copy + all kernels (on one of the plan's two compute lanes: an image submitted while
 * another is in flight takes the other lane and the two are computed concurrently) and returns immediately;
 * collect() waits for the oldest submitted image and copies its records to the host.  Up to three submits may be
 * in flight per plan (results come back in submission order). */
SIFTB_API int siftb_plan_submit(siftb_plan *plan, const void *image, int flags);
SIFTB_API int siftb_plan_collect(siftb_plan *plan, siftb_kp *out, int cap, int *n_out, int *n_per_octave,
                       float *minmax);   /* out == NULL: wait and return the counts only (records stay on the device) */
/* results of the most recently collected image, left on the device: records, count.  The plan cycles through three
 * record buffers: this one is rewritten by the third submit after the one that produced it (see
 * siftb_plan_hold_records for readers on other streams) */
SIFTB_API int siftb_plan_result_dev(const siftb_plan *plan, const siftb_kp **dev_records, const int **dev_count);

/* host copy of the records of the most recently collected run (for callers that collected with out == NULL) */
SIFTB_API int siftb_plan_fetch_records(siftb_plan *plan, siftb_kp *out, int cap, int *n_out);

/* profile=True event list, plan.py:826-847 log_profile: names[i] ran for ms[i] on the device.
 * Pointers stay valid until the next keypoints()/submit() on the plan. */
SIFTB_API int siftb_plan_events(siftb_plan *plan, const char *const **names, const float **ms, int *n);
/* per (octave, scale) stage counters of the last run: int[octaves][3][3] = extrema, after
 * interpolation, after orientation (what plan.py reads back at :642, :782, :689) */
SIFTB_API int siftb_plan_stage_counts(siftb_plan *plan, int *counts);

/* ---- stage-level entry points (host pointers; used by the per-kernel parity tests, one per
 *      reference kernel, mirroring reference test/test_*.py) --------------------------------- */
/* utils.py:54 kernel_size + plan.py:308-340 _init_gaussian / gaussian.cl:56 */
SIFTB_API int siftb_gauss_taps(double sigma, float *taps, int cap, int *n);
/* reductions.cl:62,142 + preprocess.cl:238 normalizes */
SIFTB_API int siftb_minmax(const float *image, int height, int width, float *minimum, float *maximum);
SIFTB_API int siftb_normalize(const float *image, int height, int width, float *out);
/* preprocess.cl *_to_float / rgb_to_float */
SIFTB_API int siftb_to_float(const void *image, int dtype, int height, int width, float *out);
/* convolution.cl:16,62 via plan.py:571 _gaussian_convolution (horizontal then vertical) */
SIFTB_API int siftb_blur(const float *image, int height, int width, const float *taps, int ntaps, float *out);
/* plan.py:609-625 + :739-745: G[1..5], DoG[0..4] (and G[3][::2, ::2]) of one octave from G[0] */
SIFTB_API int siftb_pyramid_octave(const float *g0, int height, int width, double init_sigma, float *G5, float *D5,
                         float *next_base);
/* image.cl:47 compute_gradient_orientation */
SIFTB_API int siftb_gradient(const float *image, int height, int width, float *grad, float *ori);
/* image.cl:119 local_maxmin for scale in {1,2,3} of a 5-plane DoG stack; rows (val,row,col,scale) */
SIFTB_API int siftb_local_maxmin(const float *dogs5, int height, int width, int scale, int octsize, float *kp4,
                       int cap, int *n);
/* image.cl:235 interp_keypoint + algebra.cl:57 compact: in rows (val,row,col,scale), out the
 * surviving rows (peak,row,col,sigma), compacted */
SIFTB_API int siftb_interp(const float *dogs5, int height, int width, const float *kp4_in, int n_in, float init_sigma,
                 float *kp4_out, int *n_out);
/* orientation_cpu.cl:41: in n rows (peak,row,col,sigma); out n rows (x,y,sigma*oct,angle) followed
 * by the extra-orientation rows; n_out = total */
SIFTB_API int siftb_orientation(const float *kp4_in, int n, const float *grad, const float *ori, int height, int width,
                      int octsize, float *kp4_out, int cap, int *n_out);
/* keypoints_cpu.cl:36: rows (x,y,sigma*oct,angle) -> uint8[n][128] */
SIFTB_API int siftb_descriptor(const float *kp4, int n, const float *grad, const float *ori, int height, int width,
                     int octsize, uint8_t *desc);
/* the same two stages with the kernel family selectable: variant 1 = orientation_gpu.cl:69 / keypoints_gpu2.cl:68 */
SIFTB_API int siftb_orientation_v(const float *kp4_in, int n, const float *grad, const float *ori, int height, int width,
                        int octsize, float *kp4_out, int cap, int *n_out, int variant);
SIFTB_API int siftb_descriptor_v(const float *kp4, int n, const float *grad, const float *ori, int height, int width,
                       int octsize, uint8_t *desc, int variant);

/* ---- MatchPlan (match.py:52-272, matching_{cpu,gpu}.cl:matching) ----------------------------- */
/* The matcher owns persistent device buffers like the reference's buffers["Kp_1"], ["Kp_2"], ["match"], ["cnt"]
 * (match.py:129-160).  Keypoint lists are copied into it from host OR device memory (the reference accepts
 * pyopencl arrays, match.py:216-239) and stay resident until replaced, so a caller that matches many frames
 * against one reference list (LinearAlign, alignment.py:157) sends that list once. */
typedef struct siftb_matcher siftb_matcher;
SIFTB_API int siftb_matcher_create(int device, siftb_matcher **out);          /* match.py:77-127 __init__ */
SIFTB_API int siftb_matcher_destroy(siftb_matcher *m);
SIFTB_API int siftb_matcher_set_profile(siftb_matcher *m, int enable);
SIFTB_API void *siftb_matcher_stream(const siftb_matcher *m);
/* distance: 0 = L1 on the uint8 descriptors, the reference's metric (matching_gpu.cl:79-99; default);
 * 1 = squared L2 (an extra: BASELINE config 4 words the workload as "L2"), ratio test on the squared distances */
SIFTB_API int siftb_matcher_set_metric(siftb_matcher *m, int metric);
/* which = 0 / 1: first / second list of match(); records: n dtype_kp records in host (on_device = 0) or device
 * memory of the matcher's device (on_device = 1, e.g. siftb_plan_result_dev) -- match.py:216-239 */
SIFTB_API int siftb_matcher_set_list(siftb_matcher *m, int which, const siftb_kp *records, int n, int on_device);
/* match.py:241-263: pairs_host (may be NULL): int[cap][2] = (index in list 0, index in list 1); n_found = value of
 * the device counter (may exceed cap; at most cap pairs are stored) */
SIFTB_API int siftb_matcher_run(siftb_matcher *m, float ratio_th, int cap, int *pairs_host, int *n_found);
/* the matched keypoints of the last run, gathered on the device (replaces the host-side fancy indexing of
 * match.py:267-270): out8 = float[m][8] = (x, y, scale, angle) of list 0 then of list 1;
 * out = siftb_kp[m][2], the reference's result recarray.  m = min(n_found, cap) of the last run. */
SIFTB_API int siftb_matcher_pairs(siftb_matcher *m, int *pairs_host);   /* int[m][2], the index pairs again */
SIFTB_API int siftb_matcher_pair_coords(siftb_matcher *m, float *out8);
SIFTB_API int siftb_matcher_pair_records(siftb_matcher *m, siftb_kp *out);
/* profile=True event list (match.py:226-263): names[i] ran for ms[i]; accumulates until reset != 0 */
SIFTB_API int siftb_matcher_events(siftb_matcher *m, const char *const **names, const float **ms, int *n, int reset);

/* stateless form of the above (create, load both lists, run, destroy) */
/* pairs: int[cap][2] = (index in kp1, index in kp2); n = number found (counter, may exceed cap) */
SIFTB_API int siftb_match_l1(const siftb_kp *kp1, int n1, const siftb_kp *kp2, int n2, float ratio_th, int on_device,
                   int device, int *pairs, int cap, int *n);

/* ---- LinearAlign's warp (alignment.py:329-349, transform.cl:22 transform) -------------------- */
/* Warp of the image of the plan's most recent keypoints() run, which is still resident on the device (the
 * reference uploads a frame once into buffers["input"] for both SIFT and the warp, alignment.py:242-246).
 * The image must be float32 or RGB8.  out: out_height*out_width floats (or *3 bytes), host or device memory. */
SIFTB_API int siftb_plan_warp_last(siftb_plan *plan, const float matrix[4], const float offset[2], float fill, int mode,
                                   void *out, int out_height, int out_width, int out_on_device);
/* stateless, host in / host out */
SIFTB_API int siftb_transform(const float *image, int height, int width, float *out, int out_height, int out_width,
                    const float matrix[4], const float offset[2], float fill, int mode, int device);

/* transform.cl:116 transform_RGB: interleaved uint8 [h][w][3] in and out (alignment.py:329-331) */
SIFTB_API int siftb_transform_rgb(const uint8_t *image, int height, int width, uint8_t *out, int out_height,
                                  int out_width, const float matrix[4], const float offset[2], float fill, int mode,
                                  int device);

/* ---- multi-GPU helpers (no reference equivalent: the reference is single-device) ---------------
 * One communicator per process / GPU over NCCL.  Images of a batch are independent, so ranks only exchange
 * results: the ragged keypoint arrays are all-gathered (counts, then records padded to the largest count). */
#define SIFTB_COMM_ID_BYTES 128
typedef struct siftb_comm siftb_comm;
SIFTB_API int siftb_comm_unique_id(char id[SIFTB_COMM_ID_BYTES]);   /* rank 0 creates it, ships it to the others */
SIFTB_API int siftb_comm_init(int rank, int nranks, const char id[SIFTB_COMM_ID_BYTES], int device, siftb_comm **out);
SIFTB_API int siftb_comm_destroy(siftb_comm *comm);
/* dev_records: this rank's n_local records in device memory; counts: int[nranks] (host); out_host (may be NULL):
 * all records grouped by rank, rank 0 first; n_total = sum of counts (> cap_out -> SIFTB_EOVERFLOW) */
SIFTB_API int siftb_allgather_kp(siftb_comm *comm, const siftb_kp *dev_records, int n_local, int *counts,
                                 siftb_kp *out_host, int cap_out, int *n_total);

#ifdef __cplusplus
}
#endif
#endif /* SIFTB_H */
