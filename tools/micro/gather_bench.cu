// Micro-benchmark: what gather bandwidth can one SM pull out of L2?  Persistent CTAs (1/SM, 256 gather threads),
// rows of `row_bytes` picked pseudo-randomly from an L2-resident (or not) table, 8 lanes x 16 B per 128 B, cp.async
// into a shared-memory ring with D groups in flight, or plain LDG.128 into registers (U loads in flight per thread).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp_async16(uint32_t d, const void *s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(256, 1) gather_cp(const uint8_t *tab, unsigned n_rows, int row_bytes, int iters, int local, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, c = tid & 7, r = tid >> 3;          // 32 rows per pass, 4 passes = 128 rows x 128 B per group x 2 (hi, lo)
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    unsigned seed = blockIdx.x * 7919u + r * 104729u + 12345u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t dst = s0 + (it % (D + 1)) * 32768u + r * 128 + c * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            seed = seed * 1664525u + 1013904223u;
            unsigned row = local ? (blockIdx.x * 4096u + ((seed >> 8) & 4095u)) % n_rows : (seed >> 4) % n_rows;
            const uint8_t *src = tab + (size_t)row * row_bytes + c * 16;
            cp_async16(dst + j * 4096, src);
            cp_async16(dst + 16384 + j * 4096, src + (row_bytes >= 256 ? 128 : row_bytes / 2 >= 16 ? 0 : 0) );
        }
        commit();
        wait<D>();
    }
    wait<0>();
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
}

template <int U>
__global__ void __launch_bounds__(256, 1) gather_ldg(const uint8_t *tab, unsigned n_rows, int row_bytes, int iters, int local, long long *out, float *sink)
{
    const int tid = threadIdx.x, c = tid & 7, r = tid >> 3;
    unsigned seed = blockIdx.x * 7919u + r * 104729u + 12345u;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            seed = seed * 1664525u + 1013904223u;
            unsigned row = local ? (blockIdx.x * 4096u + ((seed >> 8) & 4095u)) % n_rows : (seed >> 4) % n_rows;
            v[j] = __ldg(reinterpret_cast<const float4 *>(tab + (size_t)row * row_bytes + c * 16));
        }
#pragma unroll
        for (int j = 0; j < U; ++j) acc += v[j].x + v[j].w;
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    const unsigned n_rows = 400000;               // x 128 B = 51 MB (L2 resident), x 256 = 102 MB
    uint8_t *tab; long long *out, h[148]; float *sink;
    cudaMalloc(&tab, (size_t)n_rows * 512); cudaMemset(tab, 1, (size_t)n_rows * 512);
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(gather_cp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gather_cp<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gather_cp<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 400;
    auto report = [&](const char *name, double bytes_per_cta) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-34s %6.1f B/clk/SM  (%.2f TB/s at 1.9 GHz x 148)\n", name, bytes_per_cta / mx, bytes_per_cta / mx * 1.9e9 * 148 / 1e12);
    };
    for (int local = 0; local < 2; ++local)
        for (int rb : {128, 256, 512}) {
            char nm[96];
            const double bytes = (double)iters * 256 * 8 * 16;
            for (int rep = 0; rep < 2; ++rep) gather_cp<1><<<148, 256, 2 * 32768>>>(tab, n_rows, rb, iters, local, out);
            snprintf(nm, 96, "cp.async D=1 row %dB local=%d", rb, local); report(nm, bytes);
            for (int rep = 0; rep < 2; ++rep) gather_cp<3><<<148, 256, 4 * 32768>>>(tab, n_rows, rb, iters, local, out);
            snprintf(nm, 96, "cp.async D=3 row %dB local=%d", rb, local); report(nm, bytes);
            for (int rep = 0; rep < 2; ++rep) gather_cp<5><<<148, 256, 6 * 32768>>>(tab, n_rows, rb, iters, local, out);
            snprintf(nm, 96, "cp.async D=5 row %dB local=%d", rb, local); report(nm, bytes);
            for (int rep = 0; rep < 2; ++rep) gather_ldg<8><<<148, 256>>>(tab, n_rows, rb, iters, local, out, sink);
            snprintf(nm, 96, "ldg U=8 row %dB local=%d", rb, local); report(nm, (double)iters * 256 * 8 * 16);
            for (int rep = 0; rep < 2; ++rep) gather_ldg<16><<<148, 256>>>(tab, n_rows, rb, iters / 2, local, out, sink);
            snprintf(nm, 96, "ldg U=16 row %dB local=%d", rb, local); report(nm, (double)(iters / 2) * 256 * 16 * 16);
        }
    return 0;
}
