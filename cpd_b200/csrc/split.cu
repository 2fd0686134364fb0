// Split-row images: the operand format of the tcgen05 kernels.
//
// A row matrix X (m, c) fp32, c % 8 == 0, becomes XS (m, 2, c) bf16: row i = [hi(c) | lo(c)] with
// hi = RN_bf16(x), lo = RN_bf16(x - hi) (x = hi + lo up to 2^-18 |x|).  The image has exactly the byte
// size and row pitch (4 c bytes) of X, so a gather moves the same bytes as an fp32 gather -- but the
// rows can then be copied global -> shared with 16-byte cp.async (no register staging, no conversion
// ALU work in the gather loop, deep pipelining), and every row is split ONCE per layer instead of
// once per rulebook pair (a voxel row is gathered by ~12-17 neighbours).
// HBM-bound streaming kernel: 4 B read + 4 B written per element.
#include "tc_common.cuh"

namespace cpd {
namespace {

// COLSUM: also accumulate the per-channel column sums of x (the bias gradient when x = dy) while the rows stream
// through -- requires gridDim.x * blockDim.x % c8_per_row == 0 so that a thread keeps the same 8 channels.
template <bool COLSUM>
__global__ void split_rows_kernel(const float *__restrict__ x, long long n_chunks, int c8_per_row, uint8_t *__restrict__ xs,
                                  float *__restrict__ colsum)
{
    extern __shared__ float acc[];               // COLSUM: [8 * c8_per_row]
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (COLSUM) {
        for (int i = threadIdx.x; i < 8 * c8_per_row; i += blockDim.x) acc[i] = 0.f;
        __syncthreads();
    }
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_chunks; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / c8_per_row;
        const int c8 = (int)(t - row * c8_per_row);
        const float4 *p = reinterpret_cast<const float4 *>(x) + t * 2;
        const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
        if (COLSUM) { s[0] += v0.x; s[1] += v0.y; s[2] += v0.z; s[3] += v0.w; s[4] += v1.x; s[5] += v1.y; s[6] += v1.z; s[7] += v1.w; }
        uint4 h, l;
        tc::split8(v0, v1, h, l);
        uint8_t *dst = xs + (size_t)row * (size_t)(c8_per_row * 32) + (size_t)c8 * 16;
        *reinterpret_cast<uint4 *>(dst) = h;
        *reinterpret_cast<uint4 *>(dst + (size_t)c8_per_row * 16) = l;
    }
    if (COLSUM) {
        const int c8 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % c8_per_row);
        // lanes l, l + c8_per_row, ... of a warp hold the same channels: butterfly them together first
        for (int off = c8_per_row; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
        }
        if ((int)(threadIdx.x & 31) < c8_per_row) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&acc[c8 * 8 + j], s[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * c8_per_row; i += blockDim.x) atomicAdd(colsum + i, acc[i]);
    }
}

}  // namespace

bool split_rows_colsum_supported(int32_t c) { return c >= 8 && c <= 2048 && (c & (c - 1)) == 0; }

int32_t split_rows(const float *x, int64_t m, int32_t c, void *xs, float *colsum, cudaStream_t stream)
{
    CPD_REQUIRE(c >= 8 && c % 8 == 0, CPD_ERR_UNSUPPORTED, "cpd_split_rows: channels must be a multiple of 8");
    CPD_REQUIRE((((uintptr_t)x | (uintptr_t)xs) & 15) == 0, CPD_ERR_MISALIGNED, "cpd_split_rows: pointers must be 16-byte aligned");
    CPD_REQUIRE(!colsum || split_rows_colsum_supported(c), CPD_ERR_UNSUPPORTED, "cpd_split_rows: column sums need a power-of-two channel count <= 2048");
    if (colsum) CPD_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)c, stream));
    if (m == 0) return CPD_OK;
    const long long n_chunks = (long long)m * (c / 8);
    long long blocks = div_up(n_chunks, 256);
    const long long cap = colsum ? 148 * 6 : 148 * 16;   // grid-stride CTAs of 256 threads (256 * blocks % (c / 8) == 0); fewer when
    if (blocks > cap) blocks = cap;                      // every CTA ends with c global atomics
    if (colsum) split_rows_kernel<true><<<(unsigned)blocks, 256, (size_t)c * sizeof(float), stream>>>(x, n_chunks, c / 8, reinterpret_cast<uint8_t *>(xs), colsum);
    else split_rows_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, n_chunks, c / 8, reinterpret_cast<uint8_t *>(xs), nullptr);
    count_launch();
    return launch_status("cpd_split_rows");
}

}  // namespace cpd

extern "C" int32_t cpd_split_rows(const float *x, int64_t m, int32_t c, void *xs, float *colsum, cpd_stream_t stream)
{
    CPD_REQUIRE(x && xs && m >= 0, CPD_ERR_BAD_ARG, "cpd_split_rows: bad argument");
    return cpd::split_rows(x, m, c, xs, colsum, (cudaStream_t)stream);
}
