// micro-benchmark: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, each CTA supplies 128 rows of A and N/2 rows of B)
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
namespace cg = cooperative_groups;

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(long long* out, float* dout, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;              // 4 KB A tile (2 K chunks x 2 KB)
    uint8_t* bbuf = smem + 8192;       // (N/2) x 32 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // A = 1.0 in CTA0, 2.0 in CTA1; B = 1.0 everywhere -> D rows of CTA r = 16 * (r + 1) per MMA
    const uint32_t aval = rank == 0 ? 0x3c003c00u : 0x40004000u;
    for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(abuf)[i] = aval;
    for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(bbuf)[i] = 0x3c003c00u;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_addr(&slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo = (tc::smem_addr(abuf) >> 4) + ((uint32_t)(2048 >> 4) << 16);
        const uint32_t b_lo = (tc::smem_addr(bbuf) >> 4) + ((uint32_t)(((N / 2) * 16) >> 4) << 16);
        long long t0 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tm), "l"(((uint64_t)hi << 32) | a_lo), "l"(((uint64_t)hi << 32) | b_lo), "r"(idesc), "r"(i != 0 ? 1u : 0u)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(tc::smem_addr(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        if (tc::elect_one()) out[0] = clock64() - t0;
        __syncwarp();
    }
    if (!(warp == 1 && rank == 0)) tc::mbar_wait(&bar, 0);     // both CTAs' barriers receive the multicast commit
    tc::tc_fence_after();
    __syncthreads();
    // every warp reads its 32 lanes x first 16 columns
    float v[16];
    tc::tmem_ld16(tm + ((uint32_t)(32 * warp) << 16), v);
    if ((tid & 31) == 0) { dout[(rank * 4 + warp) * 2 + 0] = v[0]; dout[(rank * 4 + warp) * 2 + 1] = v[N > 16 ? 15 : 7]; }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}
template <int N> void run(long long* d, float* dv, int iters) {
    cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaMemset(dv, 0, 64);
    k2<N><<<2, 128, 16384>>>(d, dv, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; float hv[16];
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hv, dv, 64, cudaMemcpyDeviceToHost);
    printf("2-CTA M=256 N=%3d iters=%d cycles/MMA = %.1f  D(cta0)=%g,%g D(cta1)=%g,%g  expect %g / %g  (%s)\n", N, iters, (double)h / iters,
           hv[0], hv[1], hv[8], hv[9], 16.0 * iters, 32.0 * iters, cudaGetErrorString(e));
}
int main() {
    long long* d; float* dv; cudaMalloc(&d, 8); cudaMalloc(&dv, 64);
    run<32>(d, dv, 4); run<32>(d, dv, 4096); run<64>(d, dv, 4096); run<16>(d, dv, 4096);
    return 0;
}
