// micro-benchmark: is the ~57-cycle cost of a small-N tcgen05.mma a limit of the ISSUING THREAD or of the tensor pipe?
// W warps (one elected lane each) issue MMAs concurrently, each warp into its own accumulator columns.  If W = 2 finishes 2x the
// MMAs in the time W = 1 needs for 1x, the single issuing thread is the bottleneck.  M = 128, K = 16, N = 32, SS operands.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
template <int N, int MODE>
__global__ void __launch_bounds__(256, 1) k(long long* out, int iters, int nwarps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;
    uint8_t* bbuf = smem + 65536;
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    __shared__ long long tstart[4], tend[4];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (65536 + 16384) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(bar + i, 1); tc::mbar_fence_init(); }
    if (warp == 7) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = slot;
    if (warp < nwarps) {
        const uint32_t idesc = tc::umma_idesc_f16(N);
        const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(2048 >> 4) << 16) + (uint32_t)(warp * 4 * (4096 >> 4));
        const uint32_t b_lo = (tc::smem_addr(bbuf) >> 4) + ((uint32_t)((N * 16) >> 4) << 16);
        const uint32_t b1 = b_lo + (uint32_t)((N * 32) >> 4);
        const uint32_t d = tm + (uint32_t)(warp * 64);
        long long t0 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t a = ((uint64_t)hi << 32) | (a_lo + (uint32_t)(j * (4096 >> 4)));
                    if (MODE == 0) {
                        tc::umma_f16(d, a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                    } else {
                        tc::umma_f16_keep_a(d, a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                        tc::umma_f16_reuse_a(d, a, ((uint64_t)hi << 32) | b1, idesc);
                    }
                }
            }
            tc::umma_commit(bar + warp);
            tstart[warp] = t0;
        }
        __syncwarp();
        tc::mbar_wait(bar + warp, 0);
        if (tc::elect_one()) tend[warp] = clock64();
        __syncwarp();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        long long lo = tstart[0], hi2 = tend[0];
        for (int w = 1; w < nwarps; ++w) { lo = min(lo, tstart[w]); hi2 = max(hi2, tend[w]); }
        out[0] = hi2 - lo;
    }
    if (warp == 7) tc::tmem_dealloc(tm, 512);
}
template <int N, int MODE> void run(long long* d, int nwarps) {
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16384);
    const int iters = 4096;
    k<N, MODE><<<1, 256, 65536 + 16384>>>(d, iters, nwarps);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)iters * nwarps * (MODE ? 2 : 1));
    printf("N=%3d issuing warps=%d %s  total cycles = %lld  -> %.1f cycles per MMA over all warps  (%s)\n", N, nwarps,
           MODE ? "A-collector pairs" : "plain MMAs       ", h, per, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int w : {1, 2, 4}) run<32, 0>(d, w);
    for (int w : {1, 2, 4}) run<32, 1>(d, w);
    for (int w : {1, 2}) run<16, 0>(d, w);
    for (int w : {1, 2}) run<64, 0>(d, w);
    for (int w : {1, 2}) run<160, 0>(d, w);
    return 0;
}
