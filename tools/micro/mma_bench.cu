// Micro-benchmark: issue rate of tcgen05.mma.kind::f16 (bf16, M=128, K=16) as a function of N and of the
// accumulator pattern.  One CTA, operands = zeros in shared memory (K-major SWIZZLE_128B), R instructions
// back to back, one commit, clock64 around issue .. completion.  Build: nvcc -arch=sm_100a -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../cpd_b200/csrc/tc_common.cuh"
using namespace cpd::tc;
namespace cpd { void set_error(const char *, ...) {} void count_launch(int) {} }

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// pattern 0: all into one accumulator; 1: (main, corr, corr) like the 3-product kernels; 2: round-robin over 4 accumulators (N<=128)
// pattern 3: like 1 but operands alternate between 2 smem stages every 12 instructions
__global__ void __launch_bounds__(128, 1) bench(int N, int R, int pattern, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tiles)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        tmem_alloc(smem_u32(&tbase), 512);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t st = smem_u32(tiles);
        long long t0 = clock64();
        for (int r = 0; r < R; ++r) {
            const uint32_t stage = pattern == 3 ? ((r / 12) & 1) * 80 * 1024 : 0;
            const uint64_t a = make_desc(st + stage + ((r % 3) == 1 ? 16384 : 0)) + (uint64_t)(((r / 3) % 4) * 2);
            const uint64_t b = make_desc(st + stage + 32768 + ((r % 3) == 2 ? N * 128 : 0)) + (uint64_t)(((r / 3) % 4) * 2);
            uint32_t d = tbase;
            if (pattern == 1 || pattern == 3) d += (r % 3) ? N : 0;
            if (pattern == 2) d += (r % 4) * N;
            umma_bf16(d, a, b, idesc, r ? 1u : 0u);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int R = 1200;
    for (int pattern = 0; pattern < 4; ++pattern)
        for (int N : {16, 32, 64, 128, 256}) {
            if (pattern == 2 && N > 128) continue;
            if ((pattern == 1 || pattern == 3) && N > 256) continue;
            for (int rep = 0; rep < 2; ++rep) {
                bench<<<1, 128, 200 * 1024>>>(N, R, pattern, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("pattern %d N=%3d: issue %.1f clk/mma, issue..complete %.1f clk/mma (floor %d)\n", pattern, N, (double)h[0] / R,
                   (double)h[1] / R, 128 * N / 256);
        }
    return 0;
}
