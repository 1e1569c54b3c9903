// Stand-alone probe of cp.async.bulk.tensor.2d with odd box sizes / negative coordinates (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float *out, int bw, int bh, int cx, int cy) {
    extern __shared__ __align__(128) float smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + bw * bh);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)(bw * bh * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(cx), "r"(cy), "r"(s32(bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = smem[i];
}
int main() {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)fp;
    const int W = 512, H = 256;
    std::vector<float> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (float)i;
    float *d, *o; cudaMalloc(&d, W * H * 4); cudaMalloc(&o, 256 * 256 * 4);
    cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    int boxes[][2] = {{128, 64}, {140, 74}, {156, 90}, {148, 78}, {144, 64}, {132, 64}, {128, 74}};
    int coords[][2] = {{0, 0}, {-8, -5}, {248, 59}, {400, 200}, {-4, -13}, {508, 250}};
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    for (auto &b : boxes) for (auto &c : coords) {
        CUtensorMap m; cuuint64_t gd[2] = {W, H}, gs[1] = {W * 4}; cuuint32_t bx[2] = {(cuuint32_t)b[0], (cuuint32_t)b[1]}, es[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        probe<<<1, 256, b[0] * b[1] * 4 + 16>>>(m, o, b[0], b[1], c[0], c[1]);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> res(b[0] * b[1]);
        int bad = 0;
        if (e == cudaSuccess) {
            cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
            for (int y = 0; y < b[1]; y++) for (int x = 0; x < b[0]; x++) {
                int gx = c[0] + x, gy = c[1] + y;
                float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0.f;
                if (res[y * b[0] + x] != want) bad++;
            }
        }
        printf("box %dx%d at (%d,%d): encode=%d run=%s bad=%d\n", b[0], b[1], c[0], c[1], (int)r, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
