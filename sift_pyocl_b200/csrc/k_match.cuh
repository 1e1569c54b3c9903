// k_match.cuh -- brute-force L1 matcher with ratio test, and the device-side gathers of the matched pairs.
//
// k_match_l1 replaces matching_gpu.cl:52 / matching_cpu.cl:57 `matching` plus memset.cl
// memset_kp/memset_int (match.py:244-255).  k_pair_coords / k_pair_records replace the host-side fancy
// indexing of match.py:267-270 (result[:, 0] = nkp1[match[:, 0]] ...).
#pragma once
#include "common.cuh"

// 144-byte AoS records -> dense 128-byte descriptor rows (16-B aligned for vector loads)
__global__ void __launch_bounds__(256) k_extract_desc(const uint8_t *__restrict__ recs, int n,
                                                       uint32_t *__restrict__ desc) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one 32-bit word each
    if (i >= (long)n * 32) return;
    long row = i >> 5;
    int wd = (int)(i & 31);
    desc[i] = *reinterpret_cast<const uint32_t *>(recs + row * 144 + 16 + 4 * wd);
}

// acc + sum |a-b| over the 4 packed bytes, as ONE VABSDIFF4.U8.ACC (written in PTX so that the accumulation
// stays a chain: with __vsadu4(a, b) + acc the compiler re-associates the sums into IADD3 trees, which costs
// one extra ALU-pipe instruction per two SADs on the pipe that bounds the kernel)
__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned acc) {
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

#define MATCH_TILE 64      // rows of list 2 per shared-memory stage (8 KB), two stages in flight
#define MATCH_THREADS 64   // small CTAs: 100 000 queries -> 1563 CTAs, ~10.5 per SM (balance), all resident
#define MATCH_NONE 0xffffffffu  // "no row yet": stands for the reference's initial 1e12f (matching_cpu.cl:71)

// Running result of one query: the reference's (dist1, current_min, dist2) with its compare-and-select
// (matching_cpu.cl:92-97): strict '<', so the first row wins a tie and dist2 counts multiplicity.
struct MatchState {
    unsigned d1 = MATCH_NONE, d2 = MATCH_NONE;
    int idx = 0;
    __device__ __forceinline__ void offer(unsigned d, int row) {
        if (d < d1) { d2 = d1; d1 = d; idx = row; }
        else if (d < d2) d2 = d;
    }
};

// the ratio test and the append of matching_cpu.cl:98-108; all 32 lanes of a warp must call
__device__ __forceinline__ void match_emit(const MatchState &best, bool active, int gid0, float ratio_th,
                                           int2 *__restrict__ pairs, int cap, int *__restrict__ counter) {
    const float dist1 = best.d1 == MATCH_NONE ? 1000000000000.0f : (float)best.d1;
    const float dist2 = best.d2 == MATCH_NONE ? 1000000000000.0f : (float)best.d2;
    const bool emit = active && (dist2 != 0.0f) && (dist1 / dist2 < ratio_th);  // matching_cpu.cl:100
    const int slot = warp_append(emit, counter);
    if (emit && slot < cap) pairs[slot] = make_int2(gid0, best.idx);
}

// Partial result of one query over one segment of list 2 (k_match_l1<true>), folded in segment order by
// k_match_merge: the two smallest distances of the union lie among the segments' two smallest, and offering
// (d1, idx) then d2 of every segment in order reproduces the sequential scan's tie rules.
struct MatchPartial {
    unsigned d1, d2;
    int idx, pad;
};

// One thread per query row of list 1 (its 128 bytes live in 32 registers); list 2 is streamed through shared
// memory in double-buffered cp.async stages and every row is broadcast to the whole CTA (LDS.128).
// sum |a-b| over 4 packed bytes = one VABSDIFF4.U8.ACC (64 lanes/clk/SM, tools/sad_probe.cu: the pipe that
// bounds this kernel), so everything else on that pipe is kept to ONE instruction per row: a row can only change
// the running result when its distance is below the current second best (d < dist2; dist1 <= dist2), which after
// the first rows happens ~2 ln(n2) times per query.  Four rows are evaluated together (four independent SAD
// chains hide the pipe latency); their four "d < dist2" predicates are OR-ed and only when one fires are the four
// rows offered, in order, to the reference's compare-and-select.  Distances stay integers (<= 32640), so the
// integer compares decide exactly like the reference's fp32 ones.
// SEGMENTED: blockIdx.y selects a segment of seg_rows rows of list 2 and the kernel writes a MatchPartial per
// (query, segment) instead of emitting pairs.  The pipe-bound CTAs all take equally long, so a grid of only
// n1/64 CTAs (10.5 per SM for 100 000 queries, placed unevenly by the block scheduler: 9 to 12 per SM) leaves SMs
// idle at the end; with several segments per query block the scheduler hands out ~6x more, shorter CTAs as earlier
// ones retire and the SMs finish together.
// QPT: queries per thread.  With two queries per thread every 16-byte chunk of a list-2 row fetched from shared
// memory feeds eight SADs instead of four: the LDS.128 broadcasts (8 per row and warp) otherwise keep the
// load/store data path as busy as the SADs keep the ALU pipe.  QPT = 2 needs ~64 more registers, so it is used for
// long first lists only (enough warps either way).
// L2: squared Euclidean distance instead of the reference's L1 (an extra, BASELINE config 4 words the workload as
// "brute-force L2"; the reference's metric and this package's default is L1): per 4 bytes one VABSDIFF4 (packed
// |a-b|) and one IDP4A (sum of their squares), i.e. two integer-pipe instructions where L1 needs one.  Same scan,
// same tie rules, the ratio test is applied to the squared distances with the same threshold (0.73^2).
template <bool SEGMENTED, int QPT, bool L2>
__global__ void __launch_bounds__(MATCH_THREADS) k_match_l1(const uint32_t *__restrict__ d1, int n1,
                                                             const uint32_t *__restrict__ d2_all, int n2_all,
                                                             int seg_rows, float ratio_th, int2 *__restrict__ pairs,
                                                             int cap, int *__restrict__ counter,
                                                             MatchPartial *__restrict__ partial) {
    __shared__ uint4 tile[2][MATCH_TILE * 8];
    // query u of this thread: consecutive threads own consecutive queries within each of the QPT groups
    const int gid_first = blockIdx.x * (MATCH_THREADS * QPT) + threadIdx.x;
    const int row_first = SEGMENTED ? blockIdx.y * seg_rows : 0;
    const int n2 = SEGMENTED ? max(0, min(seg_rows, n2_all - row_first)) : n2_all;
    const uint32_t *d2 = d2_all + (long)row_first * 32;
    uint32_t q[QPT][32];
#pragma unroll
    for (int u = 0; u < QPT; u++) {
        const int gid = gid_first + u * MATCH_THREADS;
        const uint4 *p = reinterpret_cast<const uint4 *>(d1) + (long)(gid < n1 ? gid : 0) * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = __ldg(p + i);
            q[u][4 * i] = v.x; q[u][4 * i + 1] = v.y; q[u][4 * i + 2] = v.z; q[u][4 * i + 3] = v.w;
        }
    }
    auto stage_load = [&](int st, int base) {  // rows [base, base + MATCH_TILE) of list 2 -> tile[st]
        const int rows = min(MATCH_TILE, n2 - base);
        for (int i = threadIdx.x; i < rows * 8; i += MATCH_THREADS) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[st][i]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                         "l"(reinterpret_cast<const uint4 *>(d2) + (long)base * 8 + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // distances of one row to the thread's QPT queries: QPT chains of 32 VABSDIFF4.U8.ACC
    auto row_dist = [&](const uint4 *row, unsigned d[QPT]) {
#pragma unroll
        for (int u = 0; u < QPT; u++) d[u] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = row[i];
#pragma unroll
            for (int u = 0; u < QPT; u++) {
                if (L2) {
                    unsigned t;
                    t = __vabsdiffu4(q[u][4 * i], v.x);     d[u] = __dp4a(t, t, d[u]);
                    t = __vabsdiffu4(q[u][4 * i + 1], v.y); d[u] = __dp4a(t, t, d[u]);
                    t = __vabsdiffu4(q[u][4 * i + 2], v.z); d[u] = __dp4a(t, t, d[u]);
                    t = __vabsdiffu4(q[u][4 * i + 3], v.w); d[u] = __dp4a(t, t, d[u]);
                } else {
                    d[u] = sad4_acc(q[u][4 * i], v.x, d[u]);
                    d[u] = sad4_acc(q[u][4 * i + 1], v.y, d[u]);
                    d[u] = sad4_acc(q[u][4 * i + 2], v.z, d[u]);
                    d[u] = sad4_acc(q[u][4 * i + 3], v.w, d[u]);
                }
            }
        }
    };
    MatchState best[QPT];
    if (n2 > 0) stage_load(0, 0);
    int st = 0;
    for (int base = 0; base < n2; base += MATCH_TILE, st ^= 1) {
        const int rows = min(MATCH_TILE, n2 - base);
        if (base + MATCH_TILE < n2) {
            stage_load(st ^ 1, base + MATCH_TILE);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const uint4 *t = tile[st];
        if (rows == MATCH_TILE) {
#pragma unroll 2
            for (int r = 0; r < MATCH_TILE; r += 4) {
                const uint4 *p4 = t + r * 8;
                unsigned dA[QPT], dB[QPT], dC[QPT], dD[QPT];
                row_dist(p4, dA);
                row_dist(p4 + 8, dB);
                row_dist(p4 + 16, dC);
                row_dist(p4 + 24, dD);
#pragma unroll
                for (int u = 0; u < QPT; u++) {
                    const unsigned thr = best[u].d2;
                    if ((dA[u] < thr) | (dB[u] < thr) | (dC[u] < thr) | (dD[u] < thr)) {  // rare: ~2 ln(n2) times per query
                        best[u].offer(dA[u], row_first + base + r);
                        best[u].offer(dB[u], row_first + base + r + 1);
                        best[u].offer(dC[u], row_first + base + r + 2);
                        best[u].offer(dD[u], row_first + base + r + 3);
                    }
                }
            }
        } else {
            for (int r = 0; r < rows; r++) {
                unsigned d[QPT];
                row_dist(t + r * 8, d);
#pragma unroll
                for (int u = 0; u < QPT; u++) best[u].offer(d[u], row_first + base + r);
            }
        }
        __syncthreads();  // everyone is done with tile[st] before it is refilled two iterations later
    }
#pragma unroll
    for (int u = 0; u < QPT; u++) {
        const int gid = gid_first + u * MATCH_THREADS;
        const bool active = gid < n1;
        if (SEGMENTED) {
            if (active) {
                MatchPartial o;
                o.d1 = best[u].d1; o.d2 = best[u].d2; o.idx = best[u].idx; o.pad = 0;
                partial[(long)gid * gridDim.y + blockIdx.y] = o;
            }
        } else {
            match_emit(best[u], active, gid, ratio_th, pairs, cap, counter);
        }
    }
}

// fold the per-segment partial results of every query in segment order and apply the ratio test
__global__ void __launch_bounds__(256) k_match_merge(const MatchPartial *__restrict__ partial, int n1, int nseg,
                                                      float ratio_th, int2 *__restrict__ pairs, int cap,
                                                      int *__restrict__ counter) {
    const int gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid0 < n1;
    MatchState best;
    if (active) {
        for (int sgm = 0; sgm < nseg; sgm++) {
            const MatchPartial o = partial[(long)gid0 * nseg + sgm];
            if (o.d1 != MATCH_NONE) best.offer(o.d1, o.idx);
            if (o.d2 != MATCH_NONE) best.offer(o.d2, o.idx);  // d2 >= the running d1: only the value is used
        }
    }
    match_emit(best, active, gid0, ratio_th, pairs, cap, counter);
}

// (x, y, scale, angle) of both keypoints of every matched pair: out[m] = {kp1[i].xysa, kp2[j].xysa}.
// What LinearAlign needs for its least-squares fit (alignment.py:266-302) -- 32 bytes per match instead of the
// 288 bytes of the two full records.
__global__ void __launch_bounds__(256) k_pair_coords(const uint8_t *__restrict__ r1, const uint8_t *__restrict__ r2,
                                                      const int2 *__restrict__ pairs, int m, float4 *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // one float4 (= one keypoint of a pair) each
    if (t >= 2 * m) return;
    const int2 pr = pairs[t >> 1];
    const uint8_t *src = (t & 1) ? r2 + (long)pr.y * 144 : r1 + (long)pr.x * 144;
    out[t] = *reinterpret_cast<const float4 *>(src);  // records are 144 B apart: 16-B aligned
}

// full 144-byte records of every matched pair, laid out like the reference's result recarray (m, 2)
// (match.py:267-270): out[m][0] = kp1[i], out[m][1] = kp2[j]
__global__ void __launch_bounds__(256) k_pair_records(const uint8_t *__restrict__ r1, const uint8_t *__restrict__ r2,
                                                       const int2 *__restrict__ pairs, int m, uint4 *__restrict__ out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk each, 9 per record
    if (t >= 18L * m) return;
    const int rec = (int)(t / 9), chunk = (int)(t - 9L * rec);
    const int2 pr = pairs[rec >> 1];
    const uint8_t *src = (rec & 1) ? r2 + (long)pr.y * 144 : r1 + (long)pr.x * 144;
    out[t] = reinterpret_cast<const uint4 *>(src)[chunk];
}
