// Micro-benchmark + semantics check of the TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) on sm_100a:
// can the gather of a sparse convolution (random 64..256-byte rows of an L2-resident matrix into a K-major
// SWIZZLE_128B shared-memory tile) move off the LSU (16-byte cp.async, 8 clk per warp instruction = 64 B/clk/SM)?
//   1. validate: 128 rows (some indices out of range -> must read as zeros) land where the UMMA descriptor expects
//      them (row r at r*128 B, 16-byte chunk c at c ^ (r & 7)) and complete_tx counts the full box also for OOB rows;
//   2. throughput: persistent CTAs, one warp issues 32 x gather4 (= one 128-row x ROWB tile) per stage, ring of stages.
// Every wait is bounded: a wrong assumption reports an error instead of hanging the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_bench gather4_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); exit(1); }
    return (EncodeTiled)fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, int *err)
{
    for (int i = 0; i < (1 << 22); ++i)
        if (mbar_try(bar, parity)) return true;
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, uint32_t bar, int col, int r0, int r1, int r2, int r3)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// ---------------------------------------------------------------- validation
// tile: 128 rows x ROWB bytes; out: copy of the shared-memory tile
template <int ROWB>
__global__ void validate_kernel(const __grid_constant__ CUtensorMap map, const int *idx, uint8_t *out, int *err)
{
    extern __shared__ uint8_t raw[];
    uint8_t *tile = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const int lane = threadIdx.x;
    for (int i = lane; i < 128 * ROWB / 4; i += 32) reinterpret_cast<uint32_t *>(tile)[i] = 0xdeadbeefu;
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_expect(smem_u32(&bar), 128 * ROWB);
    __syncwarp();
    gather4(smem_u32(tile) + lane * 4 * ROWB, &map, smem_u32(&bar), 0, idx[4 * lane], idx[4 * lane + 1], idx[4 * lane + 2], idx[4 * lane + 3]);
    if (!mbar_wait_bounded(smem_u32(&bar), 0, err)) return;
    for (int i = lane; i < 128 * ROWB / 4; i += 32) reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<uint32_t *>(tile)[i];
}

// ---------------------------------------------------------------- throughput
// warps 0..W-1: producers (32 lanes x gather4 per stage; warp w takes the iterations it % W == w -- UTMALDG reads its
// operands from uniform registers, so a warp issues its 32 gather4 one after the other through an ELECT loop and more
// issuing warps = more SMSPs' uniform datapaths); warp W: consumer (frees the stage as soon as it is full)
template <int ROWB, int STAGES, int W>
__global__ void __launch_bounds__(32 * W + 32, 1) tput_kernel(const __grid_constant__ CUtensorMap map, unsigned n_rows, int iters, int local,
                                                     unsigned miss_per_256, long long *cycles, int *err)
{
    extern __shared__ uint8_t raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp < W) {
        unsigned seed = blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u;
        for (int it = warp; it < iters; it += W) {
            const int s = it % STAGES;
            if (it >= STAGES && !mbar_wait_bounded(smem_u32(&empty[s]), ((it / STAGES) - 1) & 1, err)) return;
            int r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                seed = seed * 1664525u + 1013904223u;
                unsigned row = local ? (blockIdx.x * 4096u + ((seed >> 8) & 4095u)) % n_rows : (seed >> 4) % n_rows;
                r[j] = ((seed >> 24) < miss_per_256) ? (int)n_rows : (int)row;            // out of range -> zero fill
            }
            if (lane == 0) mbar_expect(smem_u32(&full[s]), 128 * ROWB);
            __syncwarp();
            gather4(smem_u32(tiles) + s * 128 * ROWB + lane * 4 * ROWB, &map, smem_u32(&full[s]), 0, r[0], r[1], r[2], r[3]);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            if (!mbar_wait_bounded(smem_u32(&full[s]), (it / STAGES) & 1, err)) return;
            if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
            __syncwarp();
        }
        if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

template <int ROWB>
static CUtensorMap make_map(EncodeTiled enc, void *base, unsigned n_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)(ROWB / 2), n_rows};          // bf16 elements per row, rows
    cuuint64_t strides[1] = {(cuuint64_t)ROWB};
    cuuint32_t box[2] = {(cuuint32_t)(ROWB / 2), 1};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (ROWB %d)\n", (int)r, ROWB); exit(1); }
    return m;
}

template <int ROWB>
static void validate(EncodeTiled enc, uint32_t *tab, unsigned n_rows)
{
    CUtensorMap map = make_map<ROWB>(enc, tab, n_rows);
    std::vector<int> idx(128);
    for (int r = 0; r < 128; ++r) idx[r] = (r * 7919 + 13) % n_rows;
    idx[5] = -1; idx[6] = (int)n_rows; idx[7] = 0x7fffffff; idx[64] = (int)n_rows + 5; idx[127] = -7;     // out of range: zeros expected
    int *d_idx, *d_err; uint8_t *d_out;
    CHECK(cudaMalloc(&d_idx, 512)); CHECK(cudaMalloc(&d_err, 4)); CHECK(cudaMalloc(&d_out, 128 * ROWB));
    CHECK(cudaMemcpy(d_idx, idx.data(), 512, cudaMemcpyHostToDevice)); CHECK(cudaMemset(d_err, 0, 4));
    validate_kernel<ROWB><<<1, 32, 128 * ROWB + 1024>>>(map, d_idx, d_out, d_err);
    CHECK(cudaDeviceSynchronize());
    int err; CHECK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> out(128 * ROWB / 4);
    CHECK(cudaMemcpy(out.data(), d_out, 128 * ROWB, cudaMemcpyDeviceToHost));
    if (err) { printf("validate ROWB=%d: TIMEOUT waiting for complete_tx (OOB rows do not count?)\n", ROWB); return; }
    const int chunks = ROWB / 16, swz_mask = chunks >= 8 ? 7 : chunks - 1;      // 128B: c ^ (r & 7); 64B: c ^ ((r >> 1) & 3); 32B: c ^ ((r >> 2) & 1)
    int bad = 0, bad_oob = 0;
    for (int r = 0; r < 128; ++r) {
        const bool oob = idx[r] < 0 || idx[r] >= (int)n_rows;
        const int x = ROWB == 128 ? (r & 7) : ROWB == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
        for (int c = 0; c < chunks; ++c)
            for (int w = 0; w < 4; ++w) {
                const uint32_t got = out[(r * ROWB + ((c ^ x) & swz_mask) * 16) / 4 + w];
                const uint32_t want = oob ? 0u : (uint32_t)idx[r] * (ROWB / 4) + c * 4 + w;
                if (got != want) { if (oob) ++bad_oob; else ++bad; }
            }
    }
    printf("validate ROWB=%d: %s (bad words: %d in-range, %d out-of-range rows' words; row 5 word0 = 0x%08x)\n", ROWB,
           bad == 0 && bad_oob == 0 ? "OK  layout = row r at r*ROWB, chunk c at c ^ f(r); OOB rows zero-filled, tx counts full box" : "MISMATCH",
           bad, bad_oob, out[5 * ROWB / 4]);
    cudaFree(d_idx); cudaFree(d_err); cudaFree(d_out);
}

template <int ROWB, int STAGES, int W>
static void tput(EncodeTiled enc, uint32_t *tab, unsigned n_rows, int local, unsigned miss)
{
    CUtensorMap map = make_map<ROWB>(enc, tab, n_rows);
    long long *cyc, h[148]; int *d_err;
    CHECK(cudaMalloc(&cyc, 148 * 8)); CHECK(cudaMalloc(&d_err, 4)); CHECK(cudaMemset(d_err, 0, 4));
    const int smem = STAGES * 128 * ROWB + 1024, iters = 2000;
    CHECK(cudaFuncSetAttribute(tput_kernel<ROWB, STAGES, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) tput_kernel<ROWB, STAGES, W><<<148, 32 * W + 32, smem>>>(map, n_rows, iters, local, miss, cyc, d_err);
    CHECK(cudaDeviceSynchronize());
    int err; CHECK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)iters * 128 * ROWB;
    printf("gather4 row %3dB stages %d warps %d local=%d missing=%2u%%: %s %6.1f B/clk/SM slot bytes (%5.1f real) = %.2f TB/s slots at 1.9 GHz x 148; %.0f clk per 128-row tile\n",
           ROWB, STAGES, W, local, miss * 100 / 256, err ? "TIMEOUT" : "", bytes / mx, bytes / mx * (256 - miss) / 256.0, bytes / mx * 1.9e9 * 148 / 1e12, mx / iters);
    cudaFree(cyc); cudaFree(d_err);
}

int main()
{
    EncodeTiled enc = get_encode();
    const unsigned n_rows = 600000;                     // x 128 B = 77 MB: L2 resident like a stage-2 feature image
    uint32_t *tab;
    CHECK(cudaMalloc(&tab, (size_t)n_rows * 128));
    {   // word w of the buffer holds w: row r of a ROWB-byte-row view starts at word r * ROWB / 4
        std::vector<uint32_t> h((size_t)n_rows * 32);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i;
        CHECK(cudaMemcpy(tab, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    validate<128>(enc, tab, n_rows);
    validate<64>(enc, tab, n_rows);
    validate<32>(enc, tab, n_rows);
    for (int local = 0; local < 2; ++local)
        for (unsigned miss : {0u, 120u}) {              // 120 / 256 = 47 % missing neighbours
            tput<128, 4, 1>(enc, tab, n_rows, local, miss);
            tput<128, 8, 1>(enc, tab, n_rows, local, miss);
            tput<128, 8, 2>(enc, tab, n_rows, local, miss);
            tput<128, 8, 4>(enc, tab, n_rows, local, miss);
            tput<64, 8, 4>(enc, tab, n_rows, local, miss);
            tput<32, 8, 4>(enc, tab, n_rows, local, miss);
        }
    return 0;
}
