// Probe: which 4-D fp32 tiled tensor-map / box shapes does TMA accept on B200?  (debugging reduce3x3s2_tma_kernel)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int bytes, int x0, int y0, int c0, int b) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 90112);
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar, bytes);
        tc::tma_load_4d(smem, &tmap, bar, x0, y0, c0, b);
    }
    tc::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main() {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    const int B = 2, C = 19, H = 64, W = 128;
    std::vector<float> h((size_t)B * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    struct Case { int bx, by, bc; CUtensorMapL2promotion l2; int x0, y0, c0; };
    Case cases[] = {{32, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 0, 0, 0}, {20, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 0, 0, 0},
                    {20, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, 0, 0}, {20, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -4, -1, 16}, {20, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 124, 60, 16},
                    {20, 32, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -4, -1, 0}, {20, 33, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -4, -1, 16},
                    {24, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -8, -1, 16}, {20, 33, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, -1, -1, 16}};
    for (auto& cs : cases) {
        CUtensorMap map;
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)cs.bx, (cuuint32_t)cs.by, (cuuint32_t)cs.bc, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, cs.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int bytes = cs.bx * cs.by * cs.bc * 4;
        printf("box %dx%dx%d l2=%d at (%d,%d,%d): encode=%d bytes=%d ", cs.bx, cs.by, cs.bc, (int)cs.l2, cs.x0, cs.y0, cs.c0, (int)r, bytes);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        probe<<<1, 128, 98304>>>(map, o, bytes, cs.x0, cs.y0, cs.c0, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("KERNEL: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> got(bytes / 4);
        cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
        // check element (ch 1 of the box, row 2, col 3)
        int bad = 0;
        for (int c = 0; c < cs.bc; ++c) for (int y = 0; y < cs.by; ++y) for (int x = 0; x < cs.bx; ++x) {
            int gc = cs.c0 + c, gy = cs.y0 + y, gx = cs.x0 + x;
            float ref = (gc >= 0 && gc < C && gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[(((size_t)1 * C + gc) * H + gy) * W + gx] : 0.f;
            if (got[((size_t)c * cs.by + y) * cs.bx + x] != ref) ++bad;
        }
        printf("ok, mismatches=%d\n", bad);
    }
    return 0;
}
