// micro-benchmark: is the ~57-cycle cost of a small-N tcgen05.mma the shared-memory operand fetch or the accumulate dependency
// on the TMEM tile?  Rotate consecutive MMAs over `nacc` different accumulator column ranges (independent accumulators),
// with plain MMAs and with A-collector pairs.  M = 128, K = 16, SS operands, no-swizzle K-major, 16 distinct A tiles.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int nacc, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;
    uint8_t* bbuf = smem + 65536;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        const uint32_t idesc = tc::umma_idesc_f16(N);
        const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(2048 >> 4) << 16);
        const uint32_t b_lo = (tc::smem_addr(bbuf) >> 4) + ((uint32_t)((N * 16) >> 4) << 16);
        const uint32_t b1 = b_lo + (uint32_t)((N * 32) >> 4);
        long long t0 = 0, t1 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = (uint32_t)(((i + j) % 16) * (4096 >> 4));
                    const uint64_t a = ((uint64_t)hi << 32) | (a_lo + off);
                    const uint32_t d = tm + (uint32_t)(((i + j) % nacc) * N);
                    if (mode == 0) {
                        tc::umma_f16(d, a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                    } else {
                        tc::umma_f16_keep_a(d, a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                        tc::umma_f16_reuse_a(d, a, ((uint64_t)hi << 32) | b1, idesc);
                    }
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
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}
template <int N> void run(long long* d, int nacc, int mode) {
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16384);
    const int iters = 4096;
    k<N><<<1, 128, 65536 + 16384>>>(d, iters, nacc, mode);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d accumulators=%d %s cycles per %s = %.1f  (%s)\n", N, nacc, mode ? "A-collector pair" : "plain MMA       ", mode ? "pair" : "MMA ",
           (double)h / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int mode : {0, 1})
        for (int nacc : {1, 2, 4, 8}) { run<32>(d, nacc, mode); }
    for (int nacc : {1, 4}) { run<16>(d, nacc, 0); run<64>(d, nacc, 0); run<64>(d, nacc, 1); }
    return 0;
}
