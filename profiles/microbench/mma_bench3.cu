// micro-benchmark: does the A collector (tcgen05.mma ... .collector::a::fill / ::lastuse) make the SECOND of two consecutive
// MMAs that share their A operand cheaper?  (the 3-term split issues A_hi x W_hi and A_hi x W_lo: same A, different B)
// M = 128, K = 16, N = 32, SS operands, no-swizzle K-major, 16 distinct A tiles; pairs (A_i, B0), (A_i, B1).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../glomeruli_segmentation_b200/csrc/tc_common.cuh"
using namespace espnet;
__device__ __forceinline__ void mma_fill(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_lastuse(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, float* acc_out, int iters, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* abuf = smem;              // 16 different 4 KB A tiles
    uint8_t* bbuf = smem + 65536;      // two B tiles of N x 32 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;          // A = 1.0
    for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(bbuf)[i] = i < (N * 32) / 4 ? 0x3c003c00u : 0x40004000u;   // B0 = 1, B1 = 2
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&slot, 256);
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
            // zero the accumulator once
            tc::umma_f16(tm, ((uint64_t)hi << 32) | a_lo, ((uint64_t)hi << 32) | b_lo, idesc, 0u);
            t0 = clock64();
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = (uint32_t)(((i + j) % 16) * (4096 >> 4));
                    const uint64_t a = ((uint64_t)hi << 32) | (a_lo + off);
                    if (mode == 0) {        // plain: two independent MMAs
                        tc::umma_f16(tm, a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
                        tc::umma_f16(tm, a, ((uint64_t)hi << 32) | b1, idesc, 1u);
                    } else {                // collector: fill, then reuse A
                        mma_fill(tm, a, ((uint64_t)hi << 32) | b_lo, idesc);
                        mma_lastuse(tm, a, ((uint64_t)hi << 32) | b1, idesc);
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
    tc::tc_fence_after();
    if (warp == 0) {
        float v[16];
        tc::tmem_ld16(tm, v);
        if (tid == 0) { acc_out[0] = v[0]; acc_out[1] = v[15]; }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}
template <int N> void run(long long* d, float* a, int mode) {
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
    const int iters = 4096;
    k<N><<<1, 128, 65536 + 8192>>>(d, a, iters, mode);
    cudaDeviceSynchronize();
    long long h; float acc[2];
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(acc, a, 8, cudaMemcpyDeviceToHost);
    // every pair adds 16 * 1 * 1 + 16 * 1 * 2 = 48 to each accumulator element, + 16 from the zeroing MMA
    printf("N=%3d %s  cycles per PAIR = %.1f  acc = %.0f / %.0f (expect %.0f)  (%s)\n", N, mode ? "collector::a fill+lastuse" : "plain                    ",
           (double)h / iters, acc[0], acc[1], 16.0 + 48.0 * iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; float* a; cudaMalloc(&d, 8); cudaMalloc(&a, 8);
    for (int mode : {0, 1}) { run<16>(d, a, mode); run<32>(d, a, mode); run<64>(d, a, mode); }
    return 0;
}
