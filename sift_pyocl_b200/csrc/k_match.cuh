// k_match.cuh -- brute-force L1 matcher with ratio test, and the affine warps (grey and RGB).
//
// k_match_l1 replaces matching_gpu.cl:52 / matching_cpu.cl:57 `matching` plus memset.cl
// memset_kp/memset_int (match.py:244-255).  k_transform replaces transform.cl:22 `transform`
// (alignment.py:336-346).
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
// One thread per query row of list 1 (its 128 bytes live in 32 registers); list 2 is streamed through shared
// memory in double-buffered cp.async stages and every row is broadcast to the whole CTA (LDS.128, one wavefront).
// sum |a-b| over 4 packed bytes = one VABSDIFF4.U8.ACC (64 lanes/clk/SM, measured with tools/sad_probe.cu: the
// bound of this kernel); two rows are processed together with two accumulators each, so four independent
// chains hide the latency of that pipe.  Distances stay integers (<= 32640); 0xffffffff stands for the
// reference's initial 1e12f (matching_cpu.cl:71), so the integer compares decide exactly like the fp32 ones.
__global__ void __launch_bounds__(MATCH_THREADS) k_match_l1(const uint32_t *__restrict__ d1, int n1,
                                                             const uint32_t *__restrict__ d2, int n2, float ratio_th,
                                                             int2 *__restrict__ pairs, int cap,
                                                             int *__restrict__ counter) {
    __shared__ uint4 tile[2][MATCH_TILE * 8];
    const int gid0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid0 < n1;
    uint32_t q[32];
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(d1) + (long)(active ? gid0 : 0) * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = __ldg(p + i);
            q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
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
    // (distance << 17 | row index) keys: the two smallest keys give dist1 / dist2 with the reference's tie rules
    // (strict '<': the first row wins a tie, the second-best value counts multiplicity) in three min/max
    // operations per row.  17 index bits: list 2 is walked in chunks of 2^17 rows whose results are folded, in
    // order, into the running (dist1, index, dist2) with the reference's compare-and-select (matching_cpu.cl:92-97).
    constexpr int CHUNK = 1 << 17;
    unsigned best1 = 0xffffffffu, best2 = 0xffffffffu;  // running result, plain distances
    int current_min = 0;
    unsigned key1 = 0xffffffffu, key2 = 0xffffffffu;    // current chunk, packed keys
    auto row_dist = [&](const uint4 *row) {
        unsigned da = 0, db = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = row[i];
            da = sad4_acc(q[4 * i], v.x, da);
            db = sad4_acc(q[4 * i + 1], v.y, db);
            da = sad4_acc(q[4 * i + 2], v.z, da);
            db = sad4_acc(q[4 * i + 3], v.w, db);
        }
        return da + db;
    };
    auto update = [&](unsigned d, int idx_in_chunk) {
        const unsigned key = d * (unsigned)CHUNK + (unsigned)idx_in_chunk;
        key2 = min(key2, max(key1, key));
        key1 = min(key1, key);
    };
    auto fold = [&](int chunk_base) {
        if (key1 != 0xffffffffu) {
            const unsigned d = key1 >> 17;
            if (d < best1) { best2 = best1; best1 = d; current_min = chunk_base + (int)(key1 & (CHUNK - 1)); }
            else if (d < best2) { best2 = d; }
        }
        if (key2 != 0xffffffffu) {
            const unsigned d = key2 >> 17;
            if (d < best2) best2 = d;  // d >= best1 here
        }
        key1 = key2 = 0xffffffffu;
    };
    if (n2 > 0) stage_load(0, 0);
    int st = 0, chunk_base = 0;
    for (int base = 0; base < n2; base += MATCH_TILE, st ^= 1) {
        const int rows = min(MATCH_TILE, n2 - base);
        if (base + MATCH_TILE < n2) {
            stage_load(st ^ 1, base + MATCH_TILE);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (base - chunk_base >= CHUNK) { fold(chunk_base); chunk_base = base; }  // CHUNK is a multiple of the tile
        const uint4 *t = tile[st];
        const int loc = base - chunk_base;
        if (rows == MATCH_TILE) {
#pragma unroll 1
            for (int r = 0; r < MATCH_TILE; r += 4) {
                const uint4 *p4 = t + r * 8;
                const unsigned dA = row_dist(p4), dB = row_dist(p4 + 8), dC = row_dist(p4 + 16), dD = row_dist(p4 + 24);
                update(dA, loc + r);
                update(dB, loc + r + 1);
                update(dC, loc + r + 2);
                update(dD, loc + r + 3);
            }
        } else {
            for (int r = 0; r < rows; r++) update(row_dist(t + r * 8), loc + r);
        }
        __syncthreads();  // everyone is done with tile[st] before it is refilled two iterations later
    }
    fold(chunk_base);
    const float dist1 = best1 == 0xffffffffu ? 1000000000000.0f : (float)best1;
    const float dist2 = best2 == 0xffffffffu ? 1000000000000.0f : (float)best2;
    const bool emit = active && (dist2 != 0.0f) && (dist1 / dist2 < ratio_th);  // matching_cpu.cl:100
    const int slot = warp_append(emit, counter);
    if (emit && slot < cap) pairs[slot] = make_int2(gid0, current_min);
}

// transform.cl:34-106; one thread per output pixel
__global__ void __launch_bounds__(256) k_transform(const float *__restrict__ image, float *__restrict__ output,
                                                    float m0, float m1, float m2, float m3, float off0, float off1,
                                                    int image_width, int image_height, int output_width,
                                                    int output_height, float fill, int mode) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= output_width || y >= output_height) return;
    float tx = m2 * (float)y + m3 * (float)x;  // dot(mat.s23, (y, x))
    float ty = m0 * (float)y + m1 * (float)x;
    tx += off1;
    ty += off0;
    const int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
    float interp = fill;
    if (0.0f <= tx && tx < (float)image_width && 0.0f <= ty && ty < (float)image_height) {
        if (mode == 1) {
            const float image_p = image[(long)ty_prev * image_width + tx_prev];
            const bool xo = tx_next >= image_width, yo = ty_next >= image_height;
            const float image_x = xo ? fill : image[(long)ty_prev * image_width + tx_next];
            const float image_y = yo ? fill : image[(long)ty_next * image_width + tx_prev];
            const float image_n = (xo || yo) ? fill : image[(long)ty_next * image_width + tx_next];
            const float wxn = (float)tx_next - tx, wxp = tx - (float)tx_prev;
            const float wyn = (float)ty_next - ty, wyp = ty - (float)ty_prev;
            const float interp1 = wxn * image_p + wxp * image_x;
            const float interp2 = wxn * image_y + wxp * image_n;
            interp = wyn * interp1 + wyp * interp2;
        } else {
            interp = image[(long)((int)ty) * image_width + ((int)tx)];
        }
    }
    if (tx >= (float)image_width + -0.5f) interp = fill;
    if (ty >= (float)image_height + -0.5f) interp = fill;
    output[(long)y * output_width + x] = interp;
}

// transform.cl:116-203 transform_RGB: same mapping per colour channel of an interleaved uint8 image.
// grid (ceil(3*out_w/256), out_h): one thread per output byte.
__global__ void __launch_bounds__(256) k_transform_rgb(const uint8_t *__restrict__ image, uint8_t *__restrict__ output,
                                                        float m0, float m1, float m2, float m3, float off0, float off1,
                                                        int image_width, int image_height, int output_width,
                                                        int output_height, float fill, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (i >= 3 * output_width || y >= output_height) return;
    const int x = i / 3, color = i - 3 * x;
    float tx = m2 * (float)y + m3 * (float)x;
    float ty = m0 * (float)y + m1 * (float)x;
    tx += off1;
    ty += off0;
    const int tx_next = ((int)tx) + 1, tx_prev = (int)tx, ty_next = ((int)ty) + 1, ty_prev = (int)ty;
    float interp = fill;
    if (0.0f <= tx && tx < (float)image_width && 0.0f <= ty && ty < (float)image_height) {
        if (mode == 1) {
            const bool xo = tx_next >= image_width, yo = ty_next >= image_height;
            const float image_p = (float)image[3 * ((long)ty_prev * image_width + tx_prev) + color];
            const float image_x = xo ? fill : (float)image[3 * ((long)ty_prev * image_width + tx_next) + color];
            const float image_y = yo ? fill : (float)image[3 * ((long)ty_next * image_width + tx_prev) + color];
            const float image_n = (xo || yo) ? fill : (float)image[3 * ((long)ty_next * image_width + tx_next) + color];
            const float wxn = (float)tx_next - tx, wxp = tx - (float)tx_prev;
            const float wyn = (float)ty_next - ty, wyp = ty - (float)ty_prev;
            const float interp1 = wxn * image_p + wxp * image_x;
            const float interp2 = wxn * image_y + wxp * image_n;
            interp = wyn * interp1 + wyp * interp2;
        } else {
            interp = (float)image[3 * ((long)((int)ty) * image_width + ((int)tx)) + color];
        }
    }
    if (tx >= (float)image_width + -0.5f) interp = fill;
    if (ty >= (float)image_height + -0.5f) interp = fill;
    output[3 * ((long)y * output_width + x) + color] = (uint8_t)(int)interp;  // implicit float -> uchar store
}
