// siftb_match.cu -- second translation unit of libsiftb200.so: the matcher object behind MatchPlan
// (match.py:52-272), the stateless warp entry points of LinearAlign (alignment.py:329-349) and the NCCL helpers
// a non-Python host uses to shard a batch of images.
#include <dlfcn.h>
#include <nccl.h>  // types only: libnccl.so.2 is resolved at run time (dlopen), the library has no link-time dependency
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "host_common.h"
#include "k_match.cuh"
#include "k_warp.cuh"

// ---------------------------------------------------------------------------------------------
// MatchPlan: persistent device state, like the reference's buffers["Kp_1"], ["Kp_2"], ["match"], ["cnt"]
// (match.py:129-160).  Both keypoint lists are COPIED into the matcher (from host or device memory) and stay
// resident: LinearAlign loads its reference keypoints once (alignment.py:157) and only sends the second list per
// frame; the records of a SiftPlan run go from the plan's device buffer straight into list 2 (device to device).
struct MatchEvent {
    const char *name;
    cudaEvent_t a, b;
    bool used;
};
struct siftb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mtx;
    GrowBuf recs[2], desc[2], pairs, gather, partial;
    int n[2] = {0, 0};
    int pairs_cap = 0, n_match = 0;  // n_match: pairs stored by the last run (<= pairs_cap)
    int metric = 0;                  // 0: L1 (the reference's), 1: squared L2 (extra)
    int *d_cnt = nullptr, *h_cnt = nullptr;
    bool profile = false;
    std::vector<MatchEvent> events;
    std::vector<const char *> ev_names;
    std::vector<float> ev_ms;
};

namespace {
struct EvScope {  // profile=True: one (label, event pair) per enqueue, like match.py:226-263
    siftb_matcher *m;
    int idx = -1;
    EvScope(siftb_matcher *m_, const char *name) : m(m_) {
        if (!m->profile) return;
        MatchEvent e{name, nullptr, nullptr, true};
        cudaEventCreate(&e.a);
        cudaEventCreate(&e.b);
        cudaEventRecord(e.a, m->stream);
        m->events.push_back(e);
        idx = (int)m->events.size() - 1;
    }
    ~EvScope() {
        if (idx >= 0) cudaEventRecord(m->events[idx].b, m->stream);
    }
};
void clear_events(siftb_matcher *m) {
    for (auto &e : m->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    m->events.clear();
}
}  // namespace

extern "C" int siftb_matcher_destroy(siftb_matcher *m) {
    if (!m) return 0;
    {
        DeviceGuard dg_(m->device);
        if (m->stream) cudaStreamSynchronize(m->stream);
        clear_events(m);
        cudaFree(m->d_cnt);
        if (m->h_cnt) cudaFreeHost(m->h_cnt);
        for (int i = 0; i < 2; i++) { m->recs[i].release(); m->desc[i].release(); }
        m->pairs.release();
        m->gather.release();
        m->partial.release();
        if (m->stream) cudaStreamDestroy(m->stream);
    }
    delete m;
    return 0;
}

extern "C" int siftb_matcher_create(int device, siftb_matcher **out) {
    if (!out) return fail(SIFTB_EINVAL, "out is null");
    *out = nullptr;
    siftb_matcher *m = new siftb_matcher();
    m->device = device;
    auto init = [&]() -> int {
        DeviceGuard dg_(device);
        CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        CK(cudaMalloc((void **)&m->d_cnt, sizeof(int)));
        CK(cudaHostAlloc((void **)&m->h_cnt, sizeof(int), cudaHostAllocDefault));
        return 0;
    };
    int rc = init();
    if (rc) {
        std::string keep = g_siftb_err;
        siftb_matcher_destroy(m);
        g_siftb_err = keep;
        return rc;
    }
    *out = m;
    return 0;
}

extern "C" int siftb_matcher_set_profile(siftb_matcher *m, int enable) {
    if (!m) return fail(SIFTB_EINVAL, "null matcher");
    std::lock_guard<std::mutex> lk(m->mtx);
    m->profile = enable != 0;
    return 0;
}
extern "C" int siftb_matcher_set_metric(siftb_matcher *m, int metric) {
    if (!m || metric < 0 || metric > 1) return fail(SIFTB_EINVAL, "metric must be 0 (L1) or 1 (L2)");
    std::lock_guard<std::mutex> lk(m->mtx);
    m->metric = metric;
    return 0;
}
extern "C" void *siftb_matcher_stream(const siftb_matcher *m) { return m ? (void *)m->stream : nullptr; }

// match.py:220-239: (re)size the list buffer, copy the records in, and extract the dense descriptor rows
extern "C" int siftb_matcher_set_list(siftb_matcher *m, int which, const siftb_kp *records, int n, int on_device) {
    if (!m || which < 0 || which > 1 || n < 0 || (n && !records)) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    m->n[which] = n;
    if (n == 0) return 0;
    CK(m->recs[which].reserve((size_t)n * 144));
    CK(m->desc[which].reserve((size_t)n * 128));
    {
        EvScope ev(m, which == 0 ? (on_device ? "copy D->D KP_1" : "copy H->D KP_1")
                                 : (on_device ? "copy D->D KP_2" : "copy H->D KP_2"));
        CK(cudaMemcpyAsync(m->recs[which].p, records, (size_t)n * 144,
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m->stream));
    }
    {
        EvScope ev(m, which == 0 ? "descriptors KP_1" : "descriptors KP_2");
        k_extract_desc<<<(int)(((long)n * 32 + 255) / 256), 256, 0, m->stream>>>(m->recs[which].as<uint8_t>(), n,
                                                                                  m->desc[which].as<uint32_t>());
        CKL();
    }
    if (!on_device) CK(cudaStreamSynchronize(m->stream));  // the caller may reuse its host buffer
    return 0;
}

// match.py:241-263: reset the output, run `matching`, read the counter back, copy the index pairs
extern "C" int siftb_matcher_run(siftb_matcher *m, float ratio_th, int cap, int *pairs_host, int *n_found) {
    if (!m || cap < 0 || !n_found) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    *n_found = 0;
    m->n_match = 0;
    m->pairs_cap = cap;
    if (m->n[0] == 0) return 0;
    CK(m->pairs.reserve((size_t)(cap > 0 ? cap : 1) * sizeof(int2)));
    {
        EvScope ev(m, "memset");
        CK(cudaMemsetAsync(m->d_cnt, 0, sizeof(int), m->stream));
    }
    {
        EvScope ev(m, "matching");
        const int n1 = m->n[0], n2 = m->n[1];
        // two queries per thread once the first list is long enough to fill the GPU that way (k_match.cuh)
        int qpt = n1 >= 148 * 4 * MATCH_THREADS * 2 ? 2 : 1;
        if (const char *e = getenv("SIFTB_MATCH_QPT")) qpt = e[0] == '2' ? 2 : 1;  // A/B switch for measurements
        const int qblocks = (n1 + MATCH_THREADS * qpt - 1) / (MATCH_THREADS * qpt);
        // long second lists are cut into up to 8 segments of >= 16384 rows (k_match.cuh: load balance); the
        // per-segment partial results are folded by k_match_merge
        int nseg = n2 / 16384;
        nseg = nseg < 1 ? 1 : (nseg > 8 ? 8 : nseg);
        const uint32_t *a1 = m->desc[0].as<uint32_t>(), *a2 = m->desc[1].as<uint32_t>();
        const int seg_rows = nseg == 1 ? n2 : ((n2 + nseg - 1) / nseg + MATCH_TILE - 1) / MATCH_TILE * MATCH_TILE;
        if (nseg > 1) CK(m->partial.reserve((size_t)n1 * nseg * sizeof(MatchPartial)));
        int2 *out_pairs = m->pairs.as<int2>();
        MatchPartial *part = m->partial.as<MatchPartial>();
        const dim3 grid(qblocks, nseg);
#define MATCH_LAUNCH(SEG, QPT, L2)                                                                              \
    k_match_l1<SEG, QPT, L2><<<grid, MATCH_THREADS, 0, m->stream>>>(a1, n1, a2, n2, seg_rows, ratio_th, out_pairs, cap, \
                                                                    m->d_cnt, part)
        if (nseg == 1) {
            if (m->metric) { if (qpt == 2) MATCH_LAUNCH(false, 2, true); else MATCH_LAUNCH(false, 1, true); }
            else           { if (qpt == 2) MATCH_LAUNCH(false, 2, false); else MATCH_LAUNCH(false, 1, false); }
            CKL();
        } else {
            if (m->metric) { if (qpt == 2) MATCH_LAUNCH(true, 2, true); else MATCH_LAUNCH(true, 1, true); }
            else           { if (qpt == 2) MATCH_LAUNCH(true, 2, false); else MATCH_LAUNCH(true, 1, false); }
            CKL();
#undef MATCH_LAUNCH
            k_match_merge<<<(n1 + 255) / 256, 256, 0, m->stream>>>(m->partial.as<MatchPartial>(), n1, nseg, ratio_th,
                                                                 m->pairs.as<int2>(), cap, m->d_cnt);
            CKL();
        }
    }
    CK(cudaMemcpyAsync(m->h_cnt, m->d_cnt, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    *n_found = *m->h_cnt;
    m->n_match = *n_found < cap ? *n_found : cap;
    if (pairs_host && m->n_match > 0) {
        EvScope ev(m, "copy D->H match");
        CK(cudaMemcpyAsync(pairs_host, m->pairs.p, (size_t)m->n_match * sizeof(int2), cudaMemcpyDeviceToHost, m->stream));
    }
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int siftb_matcher_pairs(siftb_matcher *m, int *pairs_host) {
    if (!m || !pairs_host) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    if (m->n_match == 0) return 0;
    EvScope ev(m, "copy D->H match");
    CK(cudaMemcpyAsync(pairs_host, m->pairs.p, (size_t)m->n_match * sizeof(int2), cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int siftb_matcher_pair_coords(siftb_matcher *m, float *out8) {
    if (!m || !out8) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    const int n = m->n_match;
    if (n == 0) return 0;
    CK(m->gather.reserve((size_t)n * 32));
    k_pair_coords<<<(2 * n + 255) / 256, 256, 0, m->stream>>>(m->recs[0].as<uint8_t>(), m->recs[1].as<uint8_t>(),
                                                             m->pairs.as<int2>(), n, m->gather.as<float4>());
    CKL();
    CK(cudaMemcpyAsync(out8, m->gather.p, (size_t)n * 32, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

extern "C" int siftb_matcher_pair_records(siftb_matcher *m, siftb_kp *out) {
    if (!m || !out) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    const int n = m->n_match;
    if (n == 0) return 0;
    CK(m->gather.reserve((size_t)n * 288));
    EvScope ev(m, "gather matched records");
    k_pair_records<<<(int)((18L * n + 255) / 256), 256, 0, m->stream>>>(m->recs[0].as<uint8_t>(), m->recs[1].as<uint8_t>(),
                                                                       m->pairs.as<int2>(), n, m->gather.as<uint4>());
    CKL();
    CK(cudaMemcpyAsync(out, m->gather.p, (size_t)n * 288, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return 0;
}

// profile=True event list (match.py:226-263 + MatchPlan.log_profile); the list accumulates until reset
extern "C" int siftb_matcher_events(siftb_matcher *m, const char *const **names, const float **ms, int *n, int reset) {
    if (!m) return fail(SIFTB_EINVAL, "null matcher");
    std::lock_guard<std::mutex> lk(m->mtx);
    DeviceGuard dg_(m->device);
    CK(cudaStreamSynchronize(m->stream));
    m->ev_names.clear();
    m->ev_ms.clear();
    for (auto &e : m->events) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        m->ev_names.push_back(e.name);
        m->ev_ms.push_back(t);
    }
    if (names) *names = m->ev_names.data();
    if (ms) *ms = m->ev_ms.data();
    if (n) *n = (int)m->ev_names.size();
    if (reset) clear_events(m);
    return 0;
}

// stateless form: one call = create, load both lists, run, destroy (kept for hosts that match once)
extern "C" int siftb_match_l1(const siftb_kp *kp1, int n1, const siftb_kp *kp2, int n2, float ratio_th, int on_device,
                              int device, int *pairs, int cap, int *n) {
    if (!n || n1 < 0 || n2 < 0 || cap < 0 || (n1 && !kp1) || (n2 && !kp2)) return fail(SIFTB_EINVAL, "bad argument");
    *n = 0;
    if (n1 == 0) return 0;
    siftb_matcher *m = nullptr;
    int rc = siftb_matcher_create(device, &m);
    if (!rc) rc = siftb_matcher_set_list(m, 0, kp1, n1, on_device);
    if (!rc) rc = siftb_matcher_set_list(m, 1, kp2, n2, on_device);
    if (!rc) rc = siftb_matcher_run(m, ratio_th, cap, pairs, n);
    std::string keep = g_siftb_err;
    siftb_matcher_destroy(m);
    g_siftb_err = keep;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// stateless warps, host in / host out (alignment.py:329-349 for callers without a SiftPlan; LinearAlign itself warps
// the image its SiftPlan already holds on the device: siftb_plan_warp_last)
extern "C" int siftb_transform(const float *image, int height, int width, float *out, int out_height, int out_width,
                               const float matrix[4], const float offset[2], float fill, int mode, int device) {
    if (!image || !out || !matrix || !offset || height <= 0 || width <= 0 || out_height <= 0 || out_width <= 0)
        return fail(SIFTB_EINVAL, "bad argument");
    DeviceGuard dg_(device);
    DevBuf I, O;
    DALLOC(I, (size_t)height * width * 4); DALLOC(O, (size_t)out_height * out_width * 4);
    CK(cudaMemcpy(I.p, image, (size_t)height * width * 4, cudaMemcpyHostToDevice));
    CK(launch_warp_f32(0, I.as<float>(), O.as<float>(), make_warp_map(matrix, offset, height, width, out_height,
                                                                        out_width, fill, mode)));
    CK(cudaMemcpy(out, O.p, (size_t)out_height * out_width * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int siftb_transform_rgb(const uint8_t *image, int height, int width, uint8_t *out, int out_height,
                                   int out_width, const float matrix[4], const float offset[2], float fill, int mode,
                                   int device) {
    if (!image || !out || !matrix || !offset || height <= 0 || width <= 0 || out_height <= 0 || out_width <= 0)
        return fail(SIFTB_EINVAL, "bad argument");
    DeviceGuard dg_(device);
    DevBuf I, O;
    DALLOC(I, (size_t)height * width * 3); DALLOC(O, (size_t)out_height * out_width * 3);
    CK(cudaMemcpy(I.p, image, (size_t)height * width * 3, cudaMemcpyHostToDevice));
    CK(launch_warp_rgb8(0, I.as<uint8_t>(), O.as<uint8_t>(), make_warp_map(matrix, offset, height, width, out_height,
                                                                            out_width, fill, mode)));
    CK(cudaMemcpy(out, O.p, (size_t)out_height * out_width * 3, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// NCCL helpers: one communicator per process / GPU; the ragged keypoint arrays of all ranks are
// all-gathered in two steps, counts first, then the records padded to the largest count.  libnccl.so.2 is looked
// up at run time: inside a torch process that is the copy torch already loaded, in a plain C host the system one.
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString)
            api.error = "libnccl.so.2 lacks a required symbol";
    });
    return &api;
}
}  // namespace
#define CKN(call)                                                                                         \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            return fail(SIFTB_ECUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(r_));        \
    } while (0)

struct siftb_comm {
    int rank = 0, nranks = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mtx;
    GrowBuf send, recv;
    int *d_counts = nullptr, *h_counts = nullptr;  // nranks + 1 ints: [0..nranks) gathered, [nranks] mine
};

extern "C" int siftb_comm_unique_id(char id[SIFTB_COMM_ID_BYTES]) {
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(SIFTB_ECUDA, api->error);
    static_assert(SIFTB_COMM_ID_BYTES == sizeof(ncclUniqueId), "id size");
    ncclUniqueId uid;
    CKN(api->GetUniqueId(&uid));
    memcpy(id, &uid, sizeof(uid));
    return 0;
}

extern "C" int siftb_comm_destroy(siftb_comm *c) {
    if (!c) return 0;
    {
        DeviceGuard dg_(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->comm) nccl_api()->CommDestroy(c->comm);
        cudaFree(c->d_counts);
        if (c->h_counts) cudaFreeHost(c->h_counts);
        c->send.release();
        c->recv.release();
        if (c->stream) cudaStreamDestroy(c->stream);
    }
    delete c;
    return 0;
}

extern "C" int siftb_comm_init(int rank, int nranks, const char id[SIFTB_COMM_ID_BYTES], int device, siftb_comm **out) {
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(SIFTB_EINVAL, "bad argument");
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(SIFTB_ECUDA, api->error);
    siftb_comm *c = new siftb_comm();
    c->rank = rank; c->nranks = nranks; c->device = device;
    auto init = [&]() -> int {
        DeviceGuard dg_(device);
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CK(cudaMalloc((void **)&c->d_counts, (nranks + 1) * sizeof(int)));
        CK(cudaHostAlloc((void **)&c->h_counts, (nranks + 1) * sizeof(int), cudaHostAllocDefault));
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        CKN(api->CommInitRank(&c->comm, nranks, uid, rank));
        return 0;
    };
    int rc = init();
    if (rc) {
        std::string keep = g_siftb_err;
        siftb_comm_destroy(c);
        g_siftb_err = keep;
        return rc;
    }
    *out = c;
    return 0;
}

// All-gather of ragged record arrays.  dev_records: this rank's n_local records in DEVICE memory (e.g.
// siftb_plan_result_dev).  counts[nranks] (host) receives every rank's count; out_host (may be null) receives all
// records grouped by rank, rank 0 first; n_total = sum of counts (may exceed cap_out -> SIFTB_EOVERFLOW).
extern "C" int siftb_allgather_kp(siftb_comm *c, const siftb_kp *dev_records, int n_local, int *counts,
                                  siftb_kp *out_host, int cap_out, int *n_total) {
    if (!c || n_local < 0 || (n_local && !dev_records)) return fail(SIFTB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(c->mtx);
    DeviceGuard dg_(c->device);
    NcclApi *api = nccl_api();
    const int R = c->nranks;
    c->h_counts[R] = n_local;
    CK(cudaMemcpyAsync(c->d_counts + R, c->h_counts + R, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CKN(api->AllGather(c->d_counts + R, c->d_counts, 1, ncclInt32, c->comm, c->stream));
    CK(cudaMemcpyAsync(c->h_counts, c->d_counts, R * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int nmax = 1;
    long total = 0;
    for (int r = 0; r < R; r++) {
        if (counts) counts[r] = c->h_counts[r];
        nmax = c->h_counts[r] > nmax ? c->h_counts[r] : nmax;
        total += c->h_counts[r];
    }
    if (n_total) *n_total = (int)total;
    const size_t slab = (size_t)nmax * 144;
    CK(c->send.reserve(slab));
    CK(c->recv.reserve(slab * R));
    if (n_local) CK(cudaMemcpyAsync(c->send.p, dev_records, (size_t)n_local * 144, cudaMemcpyDeviceToDevice, c->stream));
    CKN(api->AllGather(c->send.p, c->recv.p, slab, ncclUint8, c->comm, c->stream));
    int rc = 0;
    if (out_host) {
        long off = 0;
        for (int r = 0; r < R; r++) {
            long take = c->h_counts[r];
            if (off + take > cap_out) { take = cap_out - off > 0 ? cap_out - off : 0; rc = SIFTB_EOVERFLOW; }
            if (take > 0)
                CK(cudaMemcpyAsync(out_host + off, c->recv.as<uint8_t>() + slab * r, (size_t)take * 144,
                                   cudaMemcpyDeviceToHost, c->stream));
            off += take;
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    if (rc) return fail(rc, "gathered records exceed cap_out");
    return 0;
}
