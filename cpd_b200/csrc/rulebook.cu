// Rulebook construction for sparse 3-D convolution on sm_100a.
//
// Replaces the indice-pair generation inside spconv.SubMConv3d / spconv.SparseConv3d
// (call sites cpd/models/backbones_3d/spconv_backbone.py:17,20-21,108-115,189-190,415,
// 451-452; semantics SURVEY.md Appendix A.3/A.4).  Design choices (B200-first):
//   * output-stationary neighbour tables nbr[row][tap] instead of per-tap pair lists:
//     the convolution kernel then owns its output rows (no atomics, fused epilogue);
//   * the active output set of a strided conv is built with a cell bitmap + popcount
//     scan, which yields rows already sorted by linear key ((b*D+z)*H+y)*W+x --
//     deterministic and batch-major without a sort;
//   * one open-addressing uint32 hash per resolution level serves every lookup.
// All kernels are HBM/L2-latency bound integer work; grids are flat 1-D, 256 threads.
#include "common.cuh"

namespace cpd {
namespace {

struct Geo {
    int d, h, w;  // spatial shape (z, y, x)
    int batch;
};

__device__ __forceinline__ uint32_t lin_key(const Geo &g, int b, int z, int y, int x)
{
    return (uint32_t)((((long long)b * g.d + z) * g.h + y) * g.w + x);
}

__global__ void hash_insert_kernel(const int32_t *__restrict__ coords, long long m, Geo g, HashView h)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + i);
    uint32_t key = lin_key(g, c.x, c.y, c.z, c.w);
    uint32_t s = hash_mix(key) & h.mask;
    for (;;) {
        uint32_t prev = atomicCAS(h.keys + s, HASH_EMPTY, key);
        // vals start at 0x7f7f7f7f (memset by the host wrapper); atomicMin in BOTH branches, so that a duplicate
        // coordinate keeps its first row whatever the interleaving of the claiming and the later thread
        if (prev == HASH_EMPTY || prev == key) { atomicMin(h.vals + s, (int32_t)i); return; }
        s = (s + 1) & h.mask;
    }
}

struct Ker { int kd, kh, kw, sd, sh, sw, pd, ph, pw; };

// The table kernels run one thread per (row, tap).  Their arithmetic is index decoding -- t / K, tap % kw, n % stride -- which
// with RUNTIME divisors costs more than the hash probe it guards (64-bit division by K alone is ~40 instructions).  The
// geometries of the CPD towers are fixed (3x3x3 stride 1 / stride 2, 3x1x1 stride 2x1x1), so the kernels are instantiated
// with those constants (template value 0 = read the runtime field) and a 32-bit flat index whenever m * K fits.
template <int KD, int KH, int KW, int SD, int SH, int SW>
struct KerC {
    const Ker &k;
    __device__ __forceinline__ int kd() const { return KD ? KD : k.kd; }
    __device__ __forceinline__ int kh() const { return KH ? KH : k.kh; }
    __device__ __forceinline__ int kw() const { return KW ? KW : k.kw; }
    __device__ __forceinline__ int sd() const { return SD ? SD : k.sd; }
    __device__ __forceinline__ int sh() const { return SH ? SH : k.sh; }
    __device__ __forceinline__ int sw() const { return SW ? SW : k.sw; }
};

template <typename IdxT, int KD, int KH, int KW>
__global__ void subm_table_kernel(const int32_t *__restrict__ coords, long long m, Geo g, Ker k_, HashView h,
                                  int32_t *__restrict__ nbr)
{
    const KerC<KD, KH, KW, 1, 1, 1> k{k_};
    const int K = k.kd() * k.kh() * k.kw();
    const IdxT t = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)t >= m * K) return;
    const IdxT o = t / (IdxT)K;
    const int tap = (int)(t - o * (IdxT)K);
    const int kx = tap % k.kw(), ky = (tap / k.kw()) % k.kh(), kz = tap / (k.kw() * k.kh());
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + o);
    const int z = c.y + kz - k.kd() / 2, y = c.z + ky - k.kh() / 2, x = c.w + kx - k.kw() / 2;
    int32_t r = -1;
    if (z >= 0 && y >= 0 && x >= 0 && z < g.d && y < g.h && x < g.w)
        r = (kz == k.kd() / 2 && ky == k.kh() / 2 && kx == k.kw() / 2) ? (int32_t)o : hash_find(h, lin_key(g, c.x, z, y, x));
    nbr[t] = r;
}

// strided conv, step 1: every (input, tap) marks its output cell in the bitmap
template <typename IdxT, int KD, int KH, int KW, int SD, int SH, int SW>
__global__ void strided_mark_kernel(const int32_t *__restrict__ coords, long long m, Geo go, Ker k_,
                                    uint32_t *__restrict__ bitmap)
{
    const KerC<KD, KH, KW, SD, SH, SW> k{k_};
    const int K = k.kd() * k.kh() * k.kw();
    const IdxT t = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)t >= m * K) return;
    const IdxT i = t / (IdxT)K;
    const int tap = (int)(t - i * (IdxT)K);
    const int kx = tap % k.kw(), ky = (tap / k.kw()) % k.kh(), kz = tap / (k.kw() * k.kh());
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + i);
    const int nz = c.y + k_.pd - kz, ny = c.z + k_.ph - ky, nx = c.w + k_.pw - kx;
    if (nz < 0 || ny < 0 || nx < 0) return;
    if (nz % k.sd() || ny % k.sh() || nx % k.sw()) return;
    const int z = nz / k.sd(), y = ny / k.sh(), x = nx / k.sw();
    if (z >= go.d || y >= go.h || x >= go.w) return;
    const uint32_t key = lin_key(go, c.x, z, y, x);
    const uint32_t bit = 1u << (key & 31);
    uint32_t *wd = bitmap + (key >> 5);
    if (!(*wd & bit)) atomicOr(wd, bit);   // plain read first: most cells are marked 4-8 times
}

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_WORDS_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_WORDS_PER_THREAD;

__global__ void __launch_bounds__(SCAN_THREADS) bitmap_block_count(const uint32_t *__restrict__ bitmap, long long nwords,
                                                                   int32_t *block_sums)
{
    long long w0 = ((long long)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_WORDS_PER_THREAD;
    int c = 0;
    if (w0 + 3 < nwords) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(bitmap + w0));
        c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    } else {
        for (int j = 0; j < SCAN_WORDS_PER_THREAD; ++j)
            if (w0 + j < nwords) c += __popc(bitmap[w0 + j]);
    }
    int tot;
    block_exclusive_scan(c, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(int32_t *block_sums, int nblocks, int32_t *total)
{
    int running = 0;
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
        int b = b0 + threadIdx.x;
        int v = b < nblocks ? block_sums[b] : 0, tot;
        int ex = block_exclusive_scan(v, &tot);
        if (b < nblocks) block_sums[b] = running + ex;
        running += tot;
    }
    if (threadIdx.x == 0) *total = running;
}

__global__ void __launch_bounds__(SCAN_THREADS) bitmap_compact(const uint32_t *__restrict__ bitmap, long long nwords,
                                                               const int32_t *__restrict__ block_prefix, Geo go,
                                                               long long cap_out, int32_t *__restrict__ out_coords)
{
    long long w0 = ((long long)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_WORDS_PER_THREAD;
    uint32_t v[SCAN_WORDS_PER_THREAD];
    int c = 0;
#pragma unroll
    for (int j = 0; j < SCAN_WORDS_PER_THREAD; ++j) {
        v[j] = (w0 + j < nwords) ? __ldg(bitmap + w0 + j) : 0u;
        c += __popc(v[j]);
    }
    int tot;
    int ex = block_exclusive_scan(c, &tot);
    if (c == 0) return;
    long long row = (long long)block_prefix[blockIdx.x] + ex;
#pragma unroll
    for (int j = 0; j < SCAN_WORDS_PER_THREAD; ++j) {
        uint32_t bits = v[j];
        while (bits) {
            int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (row < cap_out) {
                unsigned long long key = (unsigned long long)(w0 + j) * 32 + b;
                int x = (int)(key % go.w); key /= go.w;
                int y = (int)(key % go.h); key /= go.h;
                int z = (int)(key % go.d); key /= go.d;
                *reinterpret_cast<int4 *>(out_coords + 4 * row) = make_int4((int)key, z, y, x);
            }
            ++row;
        }
    }
}

// nbr_fwd: thread per (output row, tap): input at o*s - p + k
template <typename IdxT, int KD, int KH, int KW, int SD, int SH, int SW>
__global__ void strided_fwd_table_kernel(const int32_t *__restrict__ out_coords, long long m_out, Geo gi, Ker k_,
                                         HashView hin, int32_t *__restrict__ nbr)
{
    const KerC<KD, KH, KW, SD, SH, SW> k{k_};
    const int K = k.kd() * k.kh() * k.kw();
    const IdxT t = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)t >= m_out * K) return;
    const IdxT o = t / (IdxT)K;
    const int tap = (int)(t - o * (IdxT)K);
    const int kx = tap % k.kw(), ky = (tap / k.kw()) % k.kh(), kz = tap / (k.kw() * k.kh());
    const int4 c = __ldg(reinterpret_cast<const int4 *>(out_coords) + o);
    const int z = c.y * k.sd() - k_.pd + kz, y = c.z * k.sh() - k_.ph + ky, x = c.w * k.sw() - k_.pw + kx;
    int32_t r = -1;
    if (z >= 0 && y >= 0 && x >= 0 && z < gi.d && y < gi.h && x < gi.w) r = hash_find(hin, lin_key(gi, c.x, z, y, x));
    nbr[t] = r;
}

// nbr_bwd: thread per (input row, tap): output at (i + p - k)/s when divisible
template <typename IdxT, int KD, int KH, int KW, int SD, int SH, int SW>
__global__ void strided_bwd_table_kernel(const int32_t *__restrict__ in_coords, long long m_in, Geo go, Ker k_,
                                         HashView hout, int32_t *__restrict__ nbr)
{
    const KerC<KD, KH, KW, SD, SH, SW> k{k_};
    const int K = k.kd() * k.kh() * k.kw();
    const IdxT t = (IdxT)blockIdx.x * blockDim.x + threadIdx.x;
    if ((long long)t >= m_in * K) return;
    const IdxT i = t / (IdxT)K;
    const int tap = (int)(t - i * (IdxT)K);
    const int kx = tap % k.kw(), ky = (tap / k.kw()) % k.kh(), kz = tap / (k.kw() * k.kh());
    const int4 c = __ldg(reinterpret_cast<const int4 *>(in_coords) + i);
    const int nz = c.y + k_.pd - kz, ny = c.z + k_.ph - ky, nx = c.w + k_.pw - kx;
    int32_t r = -1;
    if (nz >= 0 && ny >= 0 && nx >= 0 && !(nz % k.sd()) && !(ny % k.sh()) && !(nx % k.sw())) {
        const int z = nz / k.sd(), y = ny / k.sh(), x = nx / k.sw();
        if (z < go.d && y < go.h && x < go.w) r = hash_find(hout, lin_key(go, c.x, z, y, x));
    }
    nbr[t] = r;
}

// Instantiation choice: the three geometries of the CPD towers get constants, everything else the runtime kernel.
enum GeoClass { GEO_RUNTIME = 0, GEO_333_S1, GEO_333_S2, GEO_311_S211 };
inline GeoClass geo_class(const Ker &k)
{
    if (k.kd == 3 && k.kh == 3 && k.kw == 3 && k.sd == 1 && k.sh == 1 && k.sw == 1) return GEO_333_S1;
    if (k.kd == 3 && k.kh == 3 && k.kw == 3 && k.sd == 2 && k.sh == 2 && k.sw == 2) return GEO_333_S2;
    if (k.kd == 3 && k.kh == 1 && k.kw == 1 && k.sd == 2 && k.sh == 1 && k.sw == 1) return GEO_311_S211;
    return GEO_RUNTIME;
}
// KERNEL<IdxT, constants...><<<grid, 256, 0, stream>>>(args...) for the geometry class of k and the index width of n threads
#define CPD_GEO_LAUNCH(KERNEL, k, n, stream, ...)                                                                              \
    do {                                                                                                                      \
        const unsigned grid_ = (unsigned)div_up((long long)(n), 256);                                                         \
        const bool small_ = (long long)(n) + 256 < 0xFFFFFFFFll;                                                              \
        switch (geo_class(k)) {                                                                                               \
            case GEO_333_S1:                                                                                                  \
                if (small_) KERNEL<uint32_t, 3, 3, 3, 1, 1, 1><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                       \
                else KERNEL<long long, 3, 3, 3, 1, 1, 1><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                             \
                break;                                                                                                        \
            case GEO_333_S2:                                                                                                  \
                if (small_) KERNEL<uint32_t, 3, 3, 3, 2, 2, 2><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                       \
                else KERNEL<long long, 3, 3, 3, 2, 2, 2><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                             \
                break;                                                                                                        \
            case GEO_311_S211:                                                                                                \
                if (small_) KERNEL<uint32_t, 3, 1, 1, 2, 1, 1><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                       \
                else KERNEL<long long, 3, 1, 1, 2, 1, 1><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                             \
                break;                                                                                                        \
            default:                                                                                                          \
                KERNEL<long long, 0, 0, 0, 0, 0, 0><<<grid_, 256, 0, stream>>>(__VA_ARGS__);                                  \
        }                                                                                                                     \
    } while (0)

// ---- table post-processing (what torch's index_select / .to(int32) / .t().contiguous() did in 3 + 1 strided passes) ----
constexpr int TT_ROWS = 128;     // rows per CTA = one tile of the tensor-core kernels

// out[r, :] = nbr[perm[r], :], out_rows[r] = perm[r], masks[tile] = taps present anywhere in the tile.
// 128 threads; the K entries of a source row are read by consecutive lanes, the 128 x K block is written back contiguously.
__global__ void __launch_bounds__(TT_ROWS) table_permute_kernel(const int32_t *__restrict__ nbr, long long m, int K,
                                                                  const long long *__restrict__ perm, int32_t *__restrict__ out,
                                                                  int32_t *__restrict__ out_rows, uint32_t *__restrict__ masks)
{
    extern __shared__ int32_t tt_s[];            // [TT_ROWS * K] + [TT_ROWS] source rows
    int32_t *src_row = tt_s + TT_ROWS * K;
    __shared__ uint32_t acc;
    const long long r0 = (long long)blockIdx.x * TT_ROWS;
    const int rows = (int)min((long long)TT_ROWS, m - r0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) acc = 0u;
    if (tid < rows) {
        const long long p = perm[r0 + tid];
        src_row[tid] = (int32_t)p;
        out_rows[r0 + tid] = (int32_t)p;
    }
    __syncthreads();
    uint32_t mine = 0u;
    for (int r = warp; r < rows; r += TT_ROWS / 32) {             // one warp per source row: K <= 32 consecutive words
        if (lane < K) {
            const int32_t v = __ldg(nbr + (long long)src_row[r] * K + lane);
            tt_s[r * K + lane] = v;
            if (v >= 0) mine |= 1u << lane;
        }
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(&acc, mine);
    __syncthreads();
    for (int e = tid; e < rows * K; e += TT_ROWS) out[r0 * K + e] = tt_s[e];
    if (tid == 0 && masks) masks[blockIdx.x] = acc;
}

// nbr (m, K) -> nbr_t (K, m): coalesced on both sides through a shared-memory tile of 128 rows.
__global__ void __launch_bounds__(TT_ROWS) table_transpose_kernel(const int32_t *__restrict__ nbr, long long m, int K,
                                                                    int32_t *__restrict__ nbr_t)
{
    extern __shared__ int32_t tt_s[];            // [TT_ROWS][K + 1]  (odd pitch when K is even: column reads stay conflict-free)
    const int pitch = K | 1;
    const long long r0 = (long long)blockIdx.x * TT_ROWS;
    const int rows = (int)min((long long)TT_ROWS, m - r0);
    const int tid = threadIdx.x;
    for (int e = tid; e < rows * K; e += TT_ROWS) tt_s[(e / K) * pitch + e % K] = __ldg(nbr + r0 * K + e);
    __syncthreads();
    if (tid < rows)
        for (int k = 0; k < K; ++k) nbr_t[(long long)k * m + r0 + tid] = tt_s[tid * pitch + k];
}

int32_t make_geo(const int32_t *shape3, int32_t batch, Geo *g, const char *who)
{
    CPD_REQUIRE(shape3 && batch >= 1, CPD_ERR_BAD_ARG, "%s: bad shape/batch", who);
    CPD_REQUIRE(shape3[0] > 0 && shape3[1] > 0 && shape3[2] > 0, CPD_ERR_BAD_ARG, "%s: empty spatial shape", who);
    long long ncell = (long long)batch * shape3[0] * shape3[1] * shape3[2];
    CPD_REQUIRE(ncell < 0xFFFFFFFFll, CPD_ERR_UNSUPPORTED, "%s: batch*volume exceeds 32-bit cell keys", who);
    g->d = shape3[0]; g->h = shape3[1]; g->w = shape3[2]; g->batch = batch;
    return CPD_OK;
}

int32_t make_ker(const int32_t *ks, const int32_t *st, const int32_t *pd, Ker *k, const char *who)
{
    CPD_REQUIRE(ks, CPD_ERR_BAD_ARG, "%s: null kernel size", who);
    k->kd = ks[0]; k->kh = ks[1]; k->kw = ks[2];
    k->sd = st ? st[0] : 1; k->sh = st ? st[1] : 1; k->sw = st ? st[2] : 1;
    k->pd = pd ? pd[0] : 0; k->ph = pd ? pd[1] : 0; k->pw = pd ? pd[2] : 0;
    CPD_REQUIRE(k->kd > 0 && k->kh > 0 && k->kw > 0 && k->sd > 0 && k->sh > 0 && k->sw > 0 &&
                k->kd * k->kh * k->kw <= 64 && k->pd >= 0 && k->ph >= 0 && k->pw >= 0,
                CPD_ERR_BAD_ARG, "%s: bad kernel/stride/padding", who);
    return CPD_OK;
}

int32_t view_hash(const void *hash, size_t bytes, HashView *h, const char *who)
{
    CPD_REQUIRE(hash && bytes >= 8 * 1024 && (bytes & (bytes - 1)) == 0, CPD_ERR_BAD_ARG, "%s: bad hash buffer", who);
    uint64_t cap = bytes / 8;
    h->keys = (uint32_t *)hash;
    h->vals = (int32_t *)((char *)hash + cap * 4);
    h->mask = (uint32_t)(cap - 1);
    return CPD_OK;
}

}  // namespace
}  // namespace cpd

using namespace cpd;

extern "C" size_t cpd_coord_hash_bytes(int64_t m) { return m < 0 ? 0 : (size_t)hash_capacity(m) * 8; }

extern "C" int32_t cpd_coord_hash_build(const int32_t *coords, int64_t m, const int32_t *shape3, int32_t batch,
                                        void *hash, size_t hash_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Geo g; HashView h;
    int32_t st;
    if ((st = make_geo(shape3, batch, &g, "cpd_coord_hash_build"))) return st;
    if ((st = view_hash(hash, hash_bytes, &h, "cpd_coord_hash_build"))) return st;
    CPD_REQUIRE(m >= 0 && hash_bytes >= cpd_coord_hash_bytes(m), CPD_ERR_WORKSPACE, "cpd_coord_hash_build: hash buffer too small");
    CPD_REQUIRE(((uintptr_t)coords & 15) == 0, CPD_ERR_MISALIGNED, "cpd_coord_hash_build: coords must be 16-byte aligned");
    CPD_CUDA(cudaMemsetAsync(hash, 0xff, hash_bytes / 2, stream));                              // keys: HASH_EMPTY
    CPD_CUDA(cudaMemsetAsync((char *)hash + hash_bytes / 2, 0x7f, hash_bytes / 2, stream));      // vals: > any row index
    if (m > 0) {
        hash_insert_kernel<<<(unsigned)div_up(m, 256), 256, 0, stream>>>(coords, m, g, h);
        count_launch();
    }
    return launch_status("cpd_coord_hash_build");
}

extern "C" int32_t cpd_rulebook_subm(const int32_t *coords, int64_t m, const int32_t *shape3, int32_t batch,
                                     const int32_t *ksize3, const void *hash, size_t hash_bytes, int32_t *nbr,
                                     cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Geo g; Ker k; HashView h;
    int32_t st;
    if ((st = make_geo(shape3, batch, &g, "cpd_rulebook_subm"))) return st;
    if ((st = make_ker(ksize3, nullptr, nullptr, &k, "cpd_rulebook_subm"))) return st;
    if ((st = view_hash(hash, hash_bytes, &h, "cpd_rulebook_subm"))) return st;
    CPD_REQUIRE((k.kd & 1) && (k.kh & 1) && (k.kw & 1), CPD_ERR_UNSUPPORTED, "cpd_rulebook_subm: kernel sizes must be odd");
    CPD_REQUIRE(m >= 0 && nbr && ((uintptr_t)coords & 15) == 0, CPD_ERR_BAD_ARG, "cpd_rulebook_subm: bad m/nbr/coords alignment");
    long long t = m * (long long)(k.kd * k.kh * k.kw);
    if (t > 0) {
        if (k.kd == 3 && k.kh == 3 && k.kw == 3 && t + 256 < 0xFFFFFFFFll)
            subm_table_kernel<uint32_t, 3, 3, 3><<<(unsigned)div_up(t, 256), 256, 0, stream>>>(coords, m, g, k, h, nbr);
        else
            subm_table_kernel<long long, 0, 0, 0><<<(unsigned)div_up(t, 256), 256, 0, stream>>>(coords, m, g, k, h, nbr);
        count_launch();
    }
    return launch_status("cpd_rulebook_subm");
}

static void strided_out_shape(const int32_t *shape3, const Ker &k, int32_t *o)
{
    o[0] = (shape3[0] + 2 * k.pd - k.kd) / k.sd + 1;
    o[1] = (shape3[1] + 2 * k.ph - k.kh) / k.sh + 1;
    o[2] = (shape3[2] + 2 * k.pw - k.kw) / k.sw + 1;
}

extern "C" size_t cpd_rulebook_strided_workspace_bytes(const int32_t *shape3, int32_t batch, const int32_t *ksize3,
                                                       const int32_t *stride3, const int32_t *pad3)
{
    Ker k;
    if (!shape3 || make_ker(ksize3, stride3, pad3, &k, "cpd_rulebook_strided_workspace_bytes")) return 0;
    int32_t o[3];
    strided_out_shape(shape3, k, o);
    if (o[0] <= 0 || o[1] <= 0 || o[2] <= 0) return 0;
    long long ncell = (long long)batch * o[0] * o[1] * o[2];
    long long nwords = div_up(ncell, 32);
    long long nblk = div_up(nwords, SCAN_TILE);
    return align_up((size_t)nwords * 4 + 16, 256) + align_up((size_t)nblk * 4, 256);
}

extern "C" int32_t cpd_rulebook_strided_outputs(const int32_t *coords, int64_t m, const int32_t *shape3, int32_t batch,
                                                const int32_t *ksize3, const int32_t *stride3, const int32_t *pad3,
                                                int32_t *out_shape3, int64_t cap_out, int32_t *out_coords,
                                                int32_t *n_out, void *ws, size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Geo gi, go; Ker k;
    int32_t st;
    if ((st = make_geo(shape3, batch, &gi, "cpd_rulebook_strided_outputs"))) return st;
    if ((st = make_ker(ksize3, stride3, pad3, &k, "cpd_rulebook_strided_outputs"))) return st;
    CPD_REQUIRE(out_shape3 && out_coords && n_out, CPD_ERR_BAD_ARG, "cpd_rulebook_strided_outputs: null output");
    strided_out_shape(shape3, k, out_shape3);
    if ((st = make_geo(out_shape3, batch, &go, "cpd_rulebook_strided_outputs"))) return st;
    CPD_REQUIRE((((uintptr_t)coords | (uintptr_t)out_coords) & 15) == 0, CPD_ERR_MISALIGNED, "cpd_rulebook_strided_outputs: coords must be 16-byte aligned");
    size_t need = cpd_rulebook_strided_workspace_bytes(shape3, batch, ksize3, stride3, pad3);
    CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_rulebook_strided_outputs: workspace %zu < %zu", ws_bytes, need);
    long long ncell = (long long)batch * go.d * go.h * go.w;
    long long nwords = div_up(ncell, 32);
    int nblk = (int)div_up(nwords, SCAN_TILE);
    uint32_t *bitmap = (uint32_t *)ws;
    int32_t *block_sums = (int32_t *)((char *)ws + align_up((size_t)nwords * 4 + 16, 256));
    CPD_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)nwords * 4 + 16, stream));
    long long t = m * (long long)(k.kd * k.kh * k.kw);
    if (t > 0) CPD_GEO_LAUNCH(strided_mark_kernel, k, t, stream, coords, m, go, k, bitmap);
    bitmap_block_count<<<nblk, SCAN_THREADS, 0, stream>>>(bitmap, nwords, block_sums);
    scan_block_sums<<<1, SCAN_THREADS, 0, stream>>>(block_sums, nblk, n_out);
    bitmap_compact<<<nblk, SCAN_THREADS, 0, stream>>>(bitmap, nwords, block_sums, go, cap_out, out_coords);
    count_launch(t > 0 ? 4 : 3);
    return launch_status("cpd_rulebook_strided_outputs");
}

extern "C" int32_t cpd_rulebook_strided_tables(const int32_t *in_coords, int64_t m_in, const int32_t *in_shape3,
                                               const void *in_hash, size_t in_hash_bytes, const int32_t *out_coords,
                                               int64_t m_out, const int32_t *out_shape3, const void *out_hash,
                                               size_t out_hash_bytes, int32_t batch, const int32_t *ksize3,
                                               const int32_t *stride3, const int32_t *pad3, int32_t *nbr_fwd,
                                               int32_t *nbr_bwd, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    Geo gi, go; Ker k; HashView hin, hout;
    int32_t st;
    if ((st = make_geo(in_shape3, batch, &gi, "cpd_rulebook_strided_tables"))) return st;
    if ((st = make_geo(out_shape3, batch, &go, "cpd_rulebook_strided_tables"))) return st;
    if ((st = make_ker(ksize3, stride3, pad3, &k, "cpd_rulebook_strided_tables"))) return st;
    const long long K = k.kd * k.kh * k.kw;
    CPD_REQUIRE((((uintptr_t)in_coords | (uintptr_t)out_coords) & 15) == 0, CPD_ERR_MISALIGNED, "cpd_rulebook_strided_tables: coords must be 16-byte aligned");
    if (nbr_fwd) {
        if ((st = view_hash(in_hash, in_hash_bytes, &hin, "cpd_rulebook_strided_tables"))) return st;
        if (m_out > 0) {
            CPD_GEO_LAUNCH(strided_fwd_table_kernel, k, m_out * K, stream, out_coords, m_out, gi, k, hin, nbr_fwd);
            count_launch();
        }
    }
    if (nbr_bwd) {
        if ((st = view_hash(out_hash, out_hash_bytes, &hout, "cpd_rulebook_strided_tables"))) return st;
        if (m_in > 0) {
            CPD_GEO_LAUNCH(strided_bwd_table_kernel, k, m_in * K, stream, in_coords, m_in, go, k, hout, nbr_bwd);
            count_launch();
        }
    }
    return launch_status("cpd_rulebook_strided_tables");
}

extern "C" int32_t cpd_table_permute(const int32_t *nbr, int64_t m, int32_t K, const int64_t *perm, int32_t *out_nbr,
                                     int32_t *out_rows, uint32_t *tile_masks, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(nbr && perm && out_nbr && out_rows && m >= 0 && K >= 1 && K <= 32, CPD_ERR_BAD_ARG, "cpd_table_permute: bad argument (1 <= K <= 32)");
    if (m == 0) return CPD_OK;
    table_permute_kernel<<<(unsigned)div_up(m, TT_ROWS), TT_ROWS, (size_t)(TT_ROWS * K + TT_ROWS) * 4, stream>>>(
        nbr, m, K, reinterpret_cast<const long long *>(perm), out_nbr, out_rows, tile_masks);
    count_launch();
    return launch_status("cpd_table_permute");
}

extern "C" int32_t cpd_table_transpose(const int32_t *nbr, int64_t m, int32_t K, int32_t *nbr_t, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(nbr && nbr_t && m >= 0 && K >= 1 && K <= 64, CPD_ERR_BAD_ARG, "cpd_table_transpose: bad argument (1 <= K <= 64)");
    if (m == 0) return CPD_OK;
    table_transpose_kernel<<<(unsigned)div_up(m, TT_ROWS), TT_ROWS, (size_t)TT_ROWS * (K | 1) * 4, stream>>>(nbr, m, K, nbr_t);
    count_launch();
    return launch_status("cpd_table_transpose");
}
