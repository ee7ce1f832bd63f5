// micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, no-swizzle K-major) as a function of N
#include <cstdio>
#include <cuda_runtime.h>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int distinct_a, int swz, int msel) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;              // up to 16 different 4 KB A tiles (2 K chunks x 2 KB)
    uint8_t* bbuf = smem + 65536;      // N x 32 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&slot, 256);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        uint32_t idesc = tc::umma_idesc_f16(N);
        if (msel == 64) idesc = (idesc & ~(0x1fu << 24)) | ((uint32_t)(64 >> 4) << 24);
        // swz: 0 = interleaved (SBO 128, LBO 2048 / N*16); 2 = SWIZZLE_128B (rows of 128 B, SBO 1024), 4 = 64B (SBO 512), 6 = 32B (SBO 256)
        uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        uint32_t a_lo = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(2048 >> 4) << 16);
        uint32_t b_lo = (tc::smem_addr(bbuf) >> 4) + ((uint32_t)((N * 16) >> 4) << 16);
        if (swz) {
            const uint32_t sbo = swz == 2 ? 1024 : (swz == 4 ? 512 : 256);
            hi = (sbo >> 4) | (1u << 14) | ((uint32_t)swz << 29);
            a_lo = (tc::smem_addr(abuf) >> 4) + (1u << 16);
            b_lo = (tc::smem_addr(bbuf) >> 4) + (1u << 16);
        }
        long long t0 = 0, t1 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = (uint32_t)(((i + j) % distinct_a) * (4096 >> 4));
                    tc::umma_f16(tm, ((uint64_t)hi << 32) | (a_lo + off), ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                }
            }
            tc::umma_commit(&bar);
        }
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        if (tc::elect_one()) { t1 = clock64(); out[0] = t1 - t0; }
        __syncwarp();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}
template <int N> void run(long long* d, int da, int swz = 0, int msel = 128) {
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
    k<N><<<1, 128, 65536 + 8192>>>(d, 4096, da, swz, msel);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d swizzle=%d distinct_A=%2d  cycles/MMA = %.1f  (%s)\n", msel, N, swz, da, (double)h / 4096.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    // layout_type 0 = no swizzle (what the product kernels use), 6 = 32 B swizzle; the operand buffers of this benchmark
    // are laid out for these two only (other swizzle modes / M = 64 need different strides and fault with these buffers)
    for (int swz : {0, 6}) { run<16>(d, 16, swz); run<32>(d, 16, swz); run<64>(d, 16, swz); run<128>(d, 16, swz); run<256>(d, 16, swz); }
    return 0;
}
