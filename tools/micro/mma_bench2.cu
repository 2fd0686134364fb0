// Micro-benchmark 2: how fast can tcgen05.mma be ISSUED?  Tight unrolled loop, descriptors precomputed,
// 1..4 issuing warps (each into its own TMEM accumulator), N = 16..256.
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
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(int N, int R, int nwarps, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tiles)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        tmem_alloc(smem_u32(&tbase), 512);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && warp < nwarps) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t st = smem_u32(tiles);
        const uint64_t a0 = make_desc(st + warp * 16384), b0 = make_desc(st + 65536);
        const uint32_t d = tbase + (nwarps > 1 ? warp * 128 : 0);
        long long t0 = clock64();
        for (int r = 0; r < R; r += 4) {
            umma_acc(d, a0, b0, idesc);
            umma_acc(d, a0 + 2, b0 + 2, idesc);
            umma_acc(d, a0 + 4, b0 + 4, idesc);
            umma_acc(d, a0 + 6, b0 + 6, idesc);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0);
        long long t2 = clock64();
        out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main()
{
    long long *d, h[8];
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int R = 2000;
    for (int nw : {1, 2, 4})
        for (int N : {16, 32, 64, 128, 256}) {
            if (nw > 1 && N > 128) continue;
            for (int rep = 0; rep < 2; ++rep) {
                bench<<<1, 128, 200 * 1024>>>(N, R, nw, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            printf("warps %d N=%3d: issue %.1f clk/mma/warp, complete %.1f clk/mma/warp -> %.1f clk per mma overall (floor %d)\n", nw, N,
                   (double)h[0] / R, (double)h[1] / R, (double)h[1] / R / nw, 128 * N / 256);
        }
    return 0;
}
