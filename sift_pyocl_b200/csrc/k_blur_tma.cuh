// k_blur_tma.cuh -- the roofline kernel: fused separable Gaussian blur + DoG (+ decimation /
// + normalisation) with a TMA-staged shared-memory tile and register-tiled FMA chains.
//
// One CTA (256 threads, 2 CTAs/SM) produces a 128 x 64 tile of G[s+1] (64 x 64 on small planes):
//   1. one thread issues a single cp.async.bulk.tensor.2d (TMA) for the (64+2C) x (128+2C+pad) input box
//      at (x0-C-DELTA, y0-C) (x aligned to 16 B); out-of-image parts are zero-filled by the TMA unit and then patched with
//      the reference's mirror rule (convolution.cl:41-50) from the in-tile pixels (border tiles only);
//   2. horizontal pass: each thread owns one tile row and 16 consecutive outputs; the 16+2C inputs
//      are read with conflict-free LDS.128 (row pitch = 4*odd words), every output is one
//      sequential chain sum = fmaf(in, tap, sum), taps coming straight from the constant bank
//      (kernel parameter), results go to a second shared buffer;
//   3. vertical pass: each thread owns 4 adjacent columns x 8 rows (32 independent FMA chains),
//      streaming the 8+2C rows it needs with LDS.128;
//   4. epilogue from registers: G[s+1] (STG.128), DoG[s] = G[s] - G[s+1] with G[s] re-read from L2
//      (so the staged tile is dead after the row pass and the next tile's TMA load overlaps the column
//      pass), and for s == 2 the decimated next-octave base G[3][::2, ::2].
// CTAs are persistent (grid = min(tiles, 2 x 148)) and walk the tiles with a stride of gridDim.x.
// The per-pixel arithmetic (tap order, fused multiply-add, fp32 rounding after each pass) is
// identical to k_blur_generic and to the oracle, so results are bit-identical.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "k_blur.cuh"

#define TB_TH 64
// tile widths: 128 (512 threads) for large planes, 64 (256 threads) when a plane has too few 128-wide tiles
// to fill the GPU.  hbuf pitch (words) = TW + 4: a multiple of 4 with an odd quarter -> conflict-free
// STS.128 across rows and LDS.64 along a row.

// TMA needs the innermost box coordinate 16-byte aligned (measured on B200: a misaligned x raises
// "illegal instruction"), so the box starts DELTA = (-C mod 4) columns left of x0 - C.
__host__ __device__ constexpr int tb_delta(int C) { return (4 - C % 4) % 4; }
// outputs per thread in the row pass and rows per thread in the column pass: 16 on 128-wide tiles, 8 on 64-wide
__host__ __device__ constexpr int tb_r(int TW) { return TW / 8; }
__host__ __device__ constexpr int tb_win4(int C, int TW) { return (tb_r(TW) + 2 * C + tb_delta(C) + 3) / 4; }
__host__ __device__ constexpr int tb_box_w(int C, int TW) {
    // widest column touched by the horizontal pass, rounded to a multiple of 4 whose quarter is odd
    // (conflict-free LDS.128 across consecutive rows)
    int w = (TW - tb_r(TW)) + 4 * tb_win4(C, TW);
    return ((w / 4) & 1) ? w : w + 4;
}
__host__ __device__ constexpr int tb_box_h(int C) { return TB_TH + 2 * C; }
__host__ __device__ constexpr int tb_hp(int TW) { return TW + 4; }
__host__ __device__ constexpr size_t tb_smem_bytes(int C, int TW) {
    return (size_t)(tb_box_h(C) * tb_box_w(C, TW) + tb_box_h(C) * tb_hp(TW)) * sizeof(float) + 16;
}

enum { TB_DOG = 0, TB_DOG_HALF = 1, TB_NORM = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One row-pass task: R consecutive outputs of one tile row.  src points at the (16-byte aligned) tile column of
// the first output minus C + DELTA... i.e. input i of output o is src[DELTA + o + i]; dst receives the R sums.
template <int C, int R, int DELTA>
__device__ __forceinline__ void tb_row_task(const float *__restrict__ src, float *__restrict__ dst, const Taps &taps) {
    constexpr int N = 2 * C + 1;
    constexpr int W4 = (R + 2 * C + DELTA + 3) / 4;
    float in[W4 * 4];
#pragma unroll
    for (int i = 0; i < W4; i++) {
        const float4 v = reinterpret_cast<const float4 *>(src)[i];
        in[4 * i] = v.x; in[4 * i + 1] = v.y; in[4 * i + 2] = v.z; in[4 * i + 3] = v.w;
    }
    float acc[R];
#pragma unroll
    for (int o = 0; o < R; o++) acc[o] = 0.0f;
#pragma unroll
    for (int j = 0; j < N; j++) {
#pragma unroll
        for (int o = 0; o < R; o++) acc[o] = __fmaf_rn(in[DELTA + o + j], taps.f[N - 1 - j], acc[o]);
    }
#pragma unroll
    for (int i = 0; i < R / 4; i++)
        reinterpret_cast<float4 *>(dst)[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
}

// TMA issue helper: one 2-D box of `rows` x BW floats at (x, y) into dst, completion on bar
template <int BW>
__device__ __forceinline__ void tb_issue_load(const CUtensorMap *map, float *dst, uint64_t *bar, int x, int y,
                                              int rows) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"((uint32_t)(BW * rows * sizeof(float))) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// tmap: box (BH rows) for the first tile of a vertical run; tmap_c: box (TB_TH rows) for the continuation tiles.
template <int C, int MODE, int TB_TW>
__global__ void __launch_bounds__(256, 2)
k_blur_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_c, BlurArgs a, Taps taps,
           int nty, int ntiles) {
    constexpr int N = 2 * C + 1;
    constexpr int TB_THREADS = 256;        // column pass: 32 column groups x (TH/RV) row groups
    constexpr int TB_R = tb_r(TB_TW);
    constexpr int TB_HP = tb_hp(TB_TW);
    constexpr int BW = tb_box_w(C, TB_TW), BH = tb_box_h(C);
    constexpr int RH = TB_R;               // outputs per thread in the horizontal pass
    constexpr int NSEG = TB_TW / RH;       // segments per row
    constexpr int DELTA = tb_delta(C);     // tile column of global x is x - (x0 - C - DELTA)
    constexpr int RV = 8;                  // rows per thread in the vertical pass
    static_assert(32 * (TB_TH / RV) == TB_THREADS, "column-pass mapping");
    static_assert(2 * C <= TB_TH, "halo rows are carried between vertically adjacent tiles");
    constexpr int LW = C + DELTA;          // left halo width in tile columns
    extern __shared__ __align__(128) float smem[];
    float *tile = smem;                    // BH x BW staged input rows
    float *hbuf = smem + BH * BW;          // BH x TB_HP row-pass results; row hr <-> image row y0 - C + hr
    uint64_t *bar = reinterpret_cast<uint64_t *>(hbuf + BH * TB_HP);
    const int tid = threadIdx.x;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Programmatic dependent launch: this grid may become resident while the previous kernel of the stream
    // (the blur that produces our input) is still draining; everything above overlaps its tail.  Nothing
    // below may run before that kernel has completed and flushed its writes.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the next blur get resident early too
    __syncthreads();
    // Persistent CTA over a contiguous range of the COLUMN-MAJOR tile list (t -> tile column t / nty, tile row
    // t % nty): consecutive tiles are vertically adjacent, so all but the first tile of a run ("continuation"
    // tiles) keep the last 2C row-pass rows of the tile above, load only TH new input rows and run the row pass
    // on TH instead of TH + 2C rows.  The TMA load of the next tile is issued as soon as the row pass has consumed
    // the staged rows, so it overlaps the column pass + epilogue.
    const int t_begin = (int)(((long)ntiles * blockIdx.x) / gridDim.x);
    const int t_end = (int)(((long)ntiles * (blockIdx.x + 1)) / gridDim.x);
    if (tid == 0 && t_begin < t_end) {
        const int tx0 = (t_begin / nty) * TB_TW, ty0 = (t_begin % nty) * TB_TH;
        tb_issue_load<BW>(&tmap, tile, bar, tx0 - C - DELTA, ty0 - C, BH);
    }
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; t++) {
        const int tyi = t % nty;
        const int x0 = (t / nty) * TB_TW, y0 = tyi * TB_TH;
        const bool cont = (t > t_begin) && (tyi != 0);  // the tile above was the previous one of this CTA
        const int nrows = cont ? TB_TH : BH;             // staged input rows = row-pass rows of this tile
        const int hrow0 = cont ? 2 * C : 0;              // first hbuf row they produce
        if (cont) {  // carry the 2C halo rows: hbuf rows [TH, TH + 2C) of the tile above are rows [0, 2C) here
            for (int i = tid; i < 2 * C * (TB_TW / 4); i += TB_THREADS) {
                const int r = i / (TB_TW / 4), c4 = i - r * (TB_TW / 4);
                reinterpret_cast<float4 *>(hbuf + r * TB_HP)[c4] = reinterpret_cast<const float4 *>(hbuf + (TB_TH + r) * TB_HP)[c4];
            }
        }
        {   // wait for the staged rows
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                    : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        if (MODE == TB_NORM) {  // preprocess.cl:250 on the staged pixels
            // (255 * (x - min)) / (max - min): the division by the image-wide constant is the exact double-reciprocal
            // form of common.cuh (3 instructions instead of the ~12 of an IEEE fp32 division, same result)
            const float mn = ordered_to_float(a.norm_mm[0]);
            const float den = ordered_to_float(a.norm_mm[1]) - mn;
            const double inv_den = div_prepare(den);
            float4 *t4 = reinterpret_cast<float4 *>(tile);
            for (int i = tid; i < nrows * BW / 4; i += TB_THREADS) {
                float4 v = t4[i];
                v.x = div_by(255.0f * (v.x - mn), inv_den); v.y = div_by(255.0f * (v.y - mn), inv_den);
                v.z = div_by(255.0f * (v.z - mn), inv_den); v.w = div_by(255.0f * (v.w - mn), inv_den);
                t4[i] = v;
            }
            __syncthreads();
        }
        // horizontal mirror rule of convolution.cl:41-50 on the staged rows (tiles touching the left / right edge)
        const bool left = x0 == 0, right = x0 + TB_TW + C > a.w;
        if (left) {
            for (int i = tid; i < nrows * LW; i += TB_THREADS) {
                const int ty = i / LW, tx = i - ty * LW;
                const int gx = tx - LW;                 // < 0
                const int mx = -gx - 1, sx = mx + LW;   // source column inside the tile
                if (mx < a.w && sx < BW) tile[ty * BW + tx] = tile[ty * BW + sx];
            }
        }
        if (right) {
            const int txr = a.w - x0 + LW;              // first tile column beyond the image
            const int nr = BW - txr;
            for (int i = tid; i < nrows * nr; i += TB_THREADS) {
                const int ty = i / nr, tx = txr + (i - ty * nr);
                const int gx = x0 - LW + tx;            // >= w
                const int mx = 2 * a.w - 1 - gx, sx = mx - x0 + LW;
                if (mx >= 0 && sx >= 0) tile[ty * BW + tx] = tile[ty * BW + sx];
            }
        }
        if (cont || left || right) __syncthreads();     // halo carry / patches done before the row pass
        // ---- horizontal pass: task q -> (row = q % nrows, segment = q / nrows) --------------------------
        // q / nrows by multiply-shift (exact for q < 2^10: the error of the scaled reciprocal is < q / 2^20)
        const int rdiv = cont ? (1 << 20) / TB_TH + 1 : (1 << 20) / BH + 1;
        for (int q = tid; q < nrows * NSEG; q += TB_THREADS) {
            const int seg = (q * rdiv) >> 20, row = q - seg * nrows;
            tb_row_task<C, RH, DELTA>(tile + row * BW + seg * RH, hbuf + (hrow0 + row) * TB_HP + seg * RH, taps);
        }
        __syncthreads();
        // the staged rows are dead (the DoG centre is re-read from global/L2): prefetch the next tile's rows
        if (tid == 0 && t + 1 < t_end) {
            const int tn = t + 1, nyi = tn % nty;
            const int nx0 = (tn / nty) * TB_TW, ny0 = nyi * TB_TH;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses above -> async write
            if (nyi != 0) tb_issue_load<BW>(&tmap_c, tile, bar, nx0 - C - DELTA, ny0 + C, TB_TH);
            else tb_issue_load<BW>(&tmap, tile, bar, nx0 - C - DELTA, ny0 - C, BH);
        }
        // vertical mirror rule on the row-pass results (the row pass commutes with mirroring rows): hbuf rows whose
        // image row is above / below the image are copies of the mirrored image row's hbuf row
        const bool top = y0 == 0, bottom = y0 + TB_TH + C > a.h;
        if (top || bottom) {
            for (int i = tid; i < BH * (TB_TW / 4); i += TB_THREADS) {
                const int hr = i / (TB_TW / 4), c4 = i - hr * (TB_TW / 4);
                const int gy = y0 - C + hr;
                if (gy < 0 || gy >= a.h) {
                    const int my = (gy < 0) ? -gy - 1 : 2 * a.h - 1 - gy;
                    const int sr = my - y0 + C;
                    if (my >= 0 && my < a.h && sr >= 0 && sr < BH)
                        reinterpret_cast<float4 *>(hbuf + hr * TB_HP)[c4] = reinterpret_cast<const float4 *>(hbuf + sr * TB_HP)[c4];
                }
            }
            __syncthreads();
        }
        // ---- vertical pass: thread -> CV adjacent columns x RV rows (CV*RV independent FMA chains) -------
        constexpr int CV = TB_TW / 32;  // 4 columns (LDS.128 / STG.128) on 128-wide tiles, 2 on 64-wide
        const int cq = tid & 31, r0 = (tid >> 5) * RV;
        const int gx = x0 + CV * cq, gy0 = y0 + r0;
        const bool full = x0 + TB_TW <= a.w && y0 + TB_TH <= a.h;
        // G[s] centre values for the DoG: requested from L2 now, consumed after the column pass, so that
        // their latency hides behind the FMA chains
        float ctr[RV][CV];
        if (MODE != TB_NORM && full) {
            const float *pC = a.in + (size_t)gy0 * a.in_pitch + gx;
#pragma unroll
            for (int o = 0; o < RV; o++) {
                if (CV == 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(pC + o * a.in_pitch));
                    ctr[o][0] = v.x; ctr[o][1] = v.y; ctr[o][CV - 2] = v.z; ctr[o][CV - 1] = v.w;
                } else {
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(pC + o * a.in_pitch));
                    ctr[o][0] = v.x; ctr[o][1] = v.y;
                }
            }
        }
        float acc[RV][CV];
#pragma unroll
        for (int o = 0; o < RV; o++)
#pragma unroll
            for (int c = 0; c < CV; c++) acc[o][c] = 0.0f;
        const float *col = hbuf + r0 * TB_HP + CV * cq;
#pragma unroll
        for (int k = 0; k < RV + 2 * C; k++) {
            float v[CV];
            if (CV == 4) {
                const float4 t4 = *reinterpret_cast<const float4 *>(col + k * TB_HP);
                v[0] = t4.x; v[1] = t4.y; v[CV - 2] = t4.z; v[CV - 1] = t4.w;
            } else {
                const float2 t2 = *reinterpret_cast<const float2 *>(col + k * TB_HP);
                v[0] = t2.x; v[1] = t2.y;
            }
#pragma unroll
            for (int o = 0; o < RV; o++) {
                const int j = k - o;  // tap index of row k for output o: ascending in k, as the reference
                if (j >= 0 && j < N) {
#pragma unroll
                    for (int c = 0; c < CV; c++) acc[o][c] = __fmaf_rn(v[c], taps.f[N - 1 - j], acc[o][c]);
                }
            }
        }
        // ---- epilogue ----------------------------------------------------------------------------------
        if (full) {  // full tile: no per-element predicates, vector stores
            float *pG = a.outG + (size_t)gy0 * a.out_pitch + gx;
            float *pD = (MODE != TB_NORM) ? a.outD + (size_t)gy0 * a.out_pitch + gx : nullptr;
#pragma unroll
            for (int o = 0; o < RV; o++) {
                if (CV == 4) {
                    *reinterpret_cast<float4 *>(pG + o * a.out_pitch) =
                        make_float4(acc[o][0], acc[o][1], acc[o][CV - 2], acc[o][CV - 1]);
                    if (MODE != TB_NORM)
                        *reinterpret_cast<float4 *>(pD + o * a.out_pitch) =
                            make_float4(ctr[o][0] - acc[o][0], ctr[o][1] - acc[o][1], ctr[o][CV - 2] - acc[o][CV - 2],
                                        ctr[o][CV - 1] - acc[o][CV - 1]);
                } else {
                    *reinterpret_cast<float2 *>(pG + o * a.out_pitch) = make_float2(acc[o][0], acc[o][1]);
                    if (MODE != TB_NORM)
                        *reinterpret_cast<float2 *>(pD + o * a.out_pitch) =
                            make_float2(ctr[o][0] - acc[o][0], ctr[o][1] - acc[o][1]);
                }
            }
            if (MODE == TB_DOG_HALF) {
#pragma unroll
                for (int o = 0; o < RV; o += 2) {
                    const int hy = (gy0 + o) >> 1;
                    if (hy < a.half_h) {
#pragma unroll
                        for (int c = 0; c < CV; c += 2)
                            if (((gx + c) >> 1) < a.half_w)
                                a.outHalf[(size_t)hy * a.half_pitch + ((gx + c) >> 1)] = acc[o][c];
                    }
                }
            }
        } else {
#pragma unroll
            for (int o = 0; o < RV; o++) {
                const int gy = gy0 + o;
                if (gy < a.h) {
#pragma unroll
                    for (int c = 0; c < CV; c++) {
                        const int x = gx + c;
                        if (x < a.w) {
                            const size_t p = (size_t)gy * a.out_pitch + x;
                            a.outG[p] = acc[o][c];
                            if (MODE != TB_NORM) a.outD[p] = a.in[(size_t)gy * a.in_pitch + x] - acc[o][c];
                            if (MODE == TB_DOG_HALF) {
                                if (!(o & 1) && !(c & 1) && (gy >> 1) < a.half_h && (x >> 1) < a.half_w)
                                    a.outHalf[(size_t)(gy >> 1) * a.half_pitch + (x >> 1)] = acc[o][c];
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // hbuf is reused (halo carry / row pass) by the next tile
    }
}

// ---- host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*tb_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tb_encode_fn tb_get_encode() {
    static tb_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (tb_encode_fn)p;
    }
    return fn;
}

// source plane usable by the TMA path: 16-B aligned base, row stride multiple of 16 B
static inline bool tb_source_ok(const float *base, int pitch) {
    return (((uintptr_t)base & 15) == 0) && (pitch % 4 == 0);
}

// true when a specialised instantiation exists for this half width / mode
static inline bool tb_supported(int ntaps, int mode) {
    const int C = ntaps >> 1;
    if (!(ntaps & 1)) return false;
    if (mode == TB_DOG) return C == 5 || C == 7 || C == 8 || C == 10 || C == 13;
    if (mode == TB_DOG_HALF) return C == 8;
    if (mode == TB_NORM) return C == 7;
    return false;
}

// tile width used for a plane: 128 unless that leaves fewer tiles than resident CTAs (2 per SM)
static inline int tb_tile_w(int w, int h) {
    const int n128 = ((w + 127) / 128) * ((h + TB_TH - 1) / TB_TH);
    return n128 >= 2 * 148 ? 128 : 64;
}

// rows < 0: the full box (TH + 2C rows, first tile of a vertical run); rows > 0: that many rows (continuation)
static int tb_encode(CUtensorMap *map, const float *base, int w, int h, int pitch, int C, int rows = -1) {
    tb_encode_fn enc = tb_get_encode();
    if (!enc) return -1;
    cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)tb_box_w(C, tb_tile_w(w, h)), (cuuint32_t)(rows > 0 ? rows : tb_box_h(C))};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

template <int C, int MODE, int TW>
static cudaError_t tb_launch_tw(cudaStream_t st, const CUtensorMap &map, const CUtensorMap &map_c, const BlurArgs &a,
                                const Taps &taps) {
    static bool attr_done[64] = {};  // per device: the attribute belongs to the function on the current device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_blur_tma<C, MODE, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tb_smem_bytes(C, TW));
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    const int ntx = (a.w + TW - 1) / TW, nty = (a.h + TB_TH - 1) / TB_TH;
    const int ntiles = ntx * nty;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;  // persistent: 2 CTAs per SM, 148 SMs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = tb_smem_bytes(C, TW);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_blur_tma<C, MODE, TW>, map, map_c, a, taps, nty, ntiles);
}

template <int C, int MODE>
static cudaError_t tb_launch_one(cudaStream_t st, const CUtensorMap &map, const CUtensorMap &map_c, const BlurArgs &a,
                                 const Taps &taps) {
    if (tb_tile_w(a.w, a.h) == 128) return tb_launch_tw<C, MODE, 128>(st, map, map_c, a, taps);
    return tb_launch_tw<C, MODE, 64>(st, map, map_c, a, taps);
}

static cudaError_t tb_launch(cudaStream_t st, const CUtensorMap &map, const CUtensorMap &map_c, const BlurArgs &a,
                             const Taps &taps, int mode) {
    const int C = a.ntaps >> 1;
    if (mode == TB_DOG_HALF) return tb_launch_one<8, TB_DOG_HALF>(st, map, map_c, a, taps);
    if (mode == TB_NORM) return tb_launch_one<7, TB_NORM>(st, map, map_c, a, taps);
    switch (C) {
    case 5: return tb_launch_one<5, TB_DOG>(st, map, map_c, a, taps);
    case 7: return tb_launch_one<7, TB_DOG>(st, map, map_c, a, taps);
    case 8: return tb_launch_one<8, TB_DOG>(st, map, map_c, a, taps);
    case 10: return tb_launch_one<10, TB_DOG>(st, map, map_c, a, taps);
    case 13: return tb_launch_one<13, TB_DOG>(st, map, map_c, a, taps);
    }
    return cudaErrorInvalidValue;
}
