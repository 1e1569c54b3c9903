// host_common.h -- error reporting, device switching and scratch buffers shared by the translation units of
// libsiftb200.so (siftb_api.cu: SiftPlan path + stage hooks; siftb_match.cu: MatchPlan, warp, NCCL helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/siftb.h"

// message of the last failure on the calling thread (siftb_last_error)
inline thread_local std::string g_siftb_err;
static inline int fail(int code, const std::string &msg) {
    g_siftb_err = msg;
    return code;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? SIFTB_ENOMEM : SIFTB_ECUDA,                  \
                        std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +       \
                            std::to_string(__LINE__) + ")");                                            \
    } while (0)
#define CKL() CK(cudaGetLastError())

// Entry points switch to the plan's device for the duration of the call and restore the caller's current device
// (a host that drives several GPUs from one thread, or torch, keeps its own notion of "current").
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) prev = cur;
        cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// scratch device buffer of a stateless entry point
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T *as() { return (T *)p; }
};
#define DALLOC(buf, bytes) CK((buf).alloc(bytes))

// persistent device buffer that only grows (the reference's plans keep their buffers between calls and enlarge
// them on demand: match.py:220-239)
struct GrowBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~GrowBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes + bytes / 4;  // headroom: lists of slightly different lengths do not reallocate
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T *as() const { return (T *)p; }
};

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
