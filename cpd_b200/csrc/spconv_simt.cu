// fp32 SIMT gather-GEMM / weight-gradient / dense scatter kernels (sm_100a).
//
// This file is the exact-fp32 CUDA path: used for narrow layers (C_in = 5, C <= 32, the
// 1/2/3-channel CenterHead outputs) where the work is gather-bound rather than FLOP
// bound, for every weight gradient, and as the on-device cross-check of the tcgen05 path
// in spconv_tc.cu.  Semantics: SURVEY.md Appendix A.3-A.5; call sites
// cpd/models/backbones_3d/spconv_backbone.py:17,20-21,108-115.
#include <stdlib.h>
#include "common.cuh"

namespace cpd {

int32_t split_rows(const float *x, int64_t m, int32_t c, void *xs, float *colsum, cudaStream_t stream);
int32_t tile_tap_masks(const int32_t *nbr, int64_t m, int32_t K, uint32_t *masks, cudaStream_t stream);
int32_t tap_block_keys(const int32_t *nbr, int64_t m, int32_t K, int32_t tpb, int32_t *keys, cudaStream_t stream);
int32_t gather_gemm_tc(const void *xs, int32_t cin, const float *w, int32_t K, int32_t cout,
                       const int32_t *nbr, const uint32_t *tile_masks, const int32_t *out_rows, int64_t m_out, const float *bias, const float *scale, const float *shift,
                       const float *residual, int32_t relu, float *stats, float *y, void *ws, size_t ws_bytes,
                       cudaStream_t stream);
size_t gather_gemm_tc_workspace(int32_t cin, int32_t K, int32_t cout);
bool gather_gemm_tc_supported(int32_t cin, int32_t K, int32_t cout);
// tcgen05 weight-gradient kernel: row-stationary with taps along N (wgrad_tc.cu; needs the tap-major table)
bool gather_wgrad_rows_supported(int32_t cin, int32_t K, int32_t cout);
int32_t gather_wgrad_rows_tc(const void *xs, int32_t cin, const void *dys, int64_t m_out, int32_t cout, const int32_t *nbr_t,
                             int32_t K, float *dw, cudaStream_t stream);

namespace {

constexpr int KC = 16;        // input-channel chunk staged per step
constexpr int AS_LD = KC + 4; // padded row stride of the gathered tile (keeps float4 alignment)

struct Epilogue {
    const float *bias, *scale, *shift, *residual;
    float *stats;
    int relu;
};

// One CTA owns TM output rows x TN output channels and walks all K taps (output-stationary:
// the scatter of gather-GEMM-scatter is a plain store).  256 threads, RM x RN micro-tiles.
template <int TM, int TN, int RM, int RN, bool VEC>
__global__ void __launch_bounds__(256) gather_gemm_simt(const float *__restrict__ x, int cin,
                                                        const float *__restrict__ w, int K, int cout,
                                                        const int32_t *__restrict__ nbr, long long m_out,
                                                        Epilogue ep, float *__restrict__ y)
{
    static_assert((TM / RM) * (TN / RN) == 256, "tile/thread mismatch");
    constexpr int TX = TN / RN;
    __shared__ __align__(16) float smem[TM * AS_LD + KC * TN];
    __shared__ int32_t rowidx[TM];
    float *As = smem, *Ws = smem + TM * AS_LD;

    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const long long row0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < K; ++tap) {
        int have = 0;
        for (int r = tid; r < TM; r += 256) {
            long long row = row0 + r;
            int32_t idx = row < m_out ? __ldg(nbr + row * K + tap) : -1;
            rowidx[r] = idx;
            have |= idx >= 0;
        }
        if (!__syncthreads_or(have)) continue;  // no row of this tile has a neighbour at this tap
        for (int c0 = 0; c0 < cin; c0 += KC) {
            if (VEC) {
                for (int t = tid; t < TM * (KC / 4); t += 256) {
                    int r = t / (KC / 4), q = t % (KC / 4);
                    int32_t idx = rowidx[r];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx >= 0 && c0 + q * 4 < cin)
                        v = __ldg(reinterpret_cast<const float4 *>(x + (size_t)idx * cin + c0 + q * 4));
                    *reinterpret_cast<float4 *>(&As[r * AS_LD + q * 4]) = v;
                }
                for (int t = tid; t < TN * (KC / 4); t += 256) {
                    int n = t / (KC / 4), q = t % (KC / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n0 + n < cout && c0 + q * 4 < cin)
                        v = __ldg(reinterpret_cast<const float4 *>(w + ((size_t)(n0 + n) * K + tap) * cin + c0 + q * 4));
                    Ws[(q * 4 + 0) * TN + n] = v.x; Ws[(q * 4 + 1) * TN + n] = v.y;
                    Ws[(q * 4 + 2) * TN + n] = v.z; Ws[(q * 4 + 3) * TN + n] = v.w;
                }
            } else {
                for (int t = tid; t < TM * KC; t += 256) {
                    int r = t / KC, kk = t % KC;
                    int32_t idx = rowidx[r];
                    As[r * AS_LD + kk] = (idx >= 0 && c0 + kk < cin) ? __ldg(x + (size_t)idx * cin + c0 + kk) : 0.f;
                }
                for (int t = tid; t < TN * KC; t += 256) {
                    int n = t / KC, kk = t % KC;
                    Ws[kk * TN + n] = (n0 + n < cout && c0 + kk < cin) ? __ldg(w + ((size_t)(n0 + n) * K + tap) * cin + c0 + kk) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                float4 a[RM];
#pragma unroll
                for (int i = 0; i < RM; ++i) a[i] = *reinterpret_cast<const float4 *>(&As[(ty * RM + i) * AS_LD + k4 * 4]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float b[RN];
#pragma unroll
                    for (int j = 0; j < RN; ++j) b[j] = Ws[(k4 * 4 + kk) * TN + tx * RN + j];
#pragma unroll
                    for (int i = 0; i < RM; ++i) {
                        float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- epilogue: bias -> (stats) -> affine -> residual -> relu -> store ----
    float csum[RN], csq[RN];
#pragma unroll
    for (int j = 0; j < RN; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        long long row = row0 + ty * RM + i;
        if (row >= m_out) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            int col = n0 + tx * RN + j;
            if (col >= cout) continue;
            float v = acc[i][j];
            if (ep.bias) v += __ldg(ep.bias + col);
            csum[j] += v; csq[j] += v * v;
            if (ep.scale) v = fmaf(v, __ldg(ep.scale + col), __ldg(ep.shift + col));
            if (ep.residual) v += __ldg(ep.residual + row * cout + col);
            if (ep.relu) v = fmaxf(v, 0.f);
            acc[i][j] = v;
        }
        float *dst = y + row * cout + n0 + tx * RN;
        if (RN == 4 && (cout & 3) == 0 && n0 + tx * RN + 3 < cout) {
            *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < RN; ++j)
                if (n0 + tx * RN + j < cout) dst[j] = acc[i][j];
        }
    }
    if (ep.stats) {  // per-channel sum / sum of squares: smem tree, then one atomic per channel per CTA
        __syncthreads();
        float *red = smem;  // reuse the staging buffers
        static_assert(2 * (TM / RM) * TN <= TM * AS_LD + KC * TN, "reduction scratch too small");
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            red[(ty * TN + tx * RN + j) * 2 + 0] = csum[j];
            red[(ty * TN + tx * RN + j) * 2 + 1] = csq[j];
        }
        __syncthreads();
        for (int c = tid; c < TN; c += 256) {
            if (n0 + c >= cout) continue;
            float s = 0.f, q = 0.f;
            for (int r = 0; r < TM / RM; ++r) { s += red[(r * TN + c) * 2]; q += red[(r * TN + c) * 2 + 1]; }
            atomicAdd(ep.stats + n0 + c, s);
            atomicAdd(ep.stats + cout + n0 + c, q);
        }
    }
}

template <int TM, int TN, int RM, int RN>
void launch_gg(const float *x, int cin, const float *w, int K, int cout, const int32_t *nbr, long long m_out,
               const Epilogue &ep, float *y, cudaStream_t stream)
{
    dim3 grid((unsigned)div_up(m_out, TM), (unsigned)div_up(cout, TN));
    const bool vec = (cin % 4 == 0) && (((uintptr_t)x | (uintptr_t)w) % 16 == 0);
    if (vec) gather_gemm_simt<TM, TN, RM, RN, true><<<grid, 256, 0, stream>>>(x, cin, w, K, cout, nbr, m_out, ep, y);
    else gather_gemm_simt<TM, TN, RM, RN, false><<<grid, 256, 0, stream>>>(x, cin, w, K, cout, nbr, m_out, ep, y);
    count_launch();
}

// ---- weight gradient: dw[co,k,ci] = sum_o dy[o,co] * x[nbr[o,k],ci] --------------------
// grid (K, S, co-tiles*ci-tiles); each CTA reduces its slice of rows into a 64x64 register
// tile and adds it to dw with fp32 atomics (dw is zeroed first).
constexpr int WG_R = 16;
__global__ void __launch_bounds__(256) gather_wgrad_simt(const float *__restrict__ x, int cin,
                                                         const float *__restrict__ dy, int cout,
                                                         const int32_t *__restrict__ nbr, int K, long long m_out,
                                                         int rows_per_cta, int ci_tiles, int tap_major,
                                                         float *__restrict__ dw)
{
    __shared__ __align__(16) float dyS[WG_R * 64];
    __shared__ __align__(16) float xS[WG_R * 64];
    __shared__ int32_t idxS[WG_R];
    const int tap = blockIdx.x;
    const int co0 = (blockIdx.z / ci_tiles) * 64, ci0 = (blockIdx.z % ci_tiles) * 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    long long r_begin = (long long)blockIdx.y * rows_per_cta;
    long long r_end = r_begin + rows_per_cta < m_out ? r_begin + rows_per_cta : m_out;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long long r0 = r_begin; r0 < r_end; r0 += WG_R) {
        int have = 0;
        if (tid < WG_R) {
            long long row = r0 + tid;
            int32_t idx = row < r_end ? __ldg(nbr + (tap_major ? (long long)tap * m_out + row : row * K + tap)) : -1;
            idxS[tid] = idx;
            have = idx >= 0;
        }
        if (!__syncthreads_or(have)) continue;
        for (int t = tid; t < WG_R * 64; t += 256) {
            int r = t >> 6, c = t & 63;
            int32_t idx = idxS[r];
            long long row = r0 + r;
            dyS[t] = (idx >= 0 && co0 + c < cout) ? __ldg(dy + row * cout + co0 + c) : 0.f;
            xS[t] = (idx >= 0 && ci0 + c < cin) ? __ldg(x + (size_t)idx * cin + ci0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WG_R; ++r) {
            float4 a = *reinterpret_cast<const float4 *>(&dyS[r * 64 + ty * 4]);
            float4 b = *reinterpret_cast<const float4 *>(&xS[r * 64 + tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + tx * 4 + j;
            if (ci < cin && acc[i][j] != 0.f) atomicAdd(dw + ((size_t)co * K + tap) * cin + ci, acc[i][j]);
        }
    }
}

__global__ void __launch_bounds__(256) column_sum_kernel(const float *__restrict__ dy, long long m, int c,
                                                         int rows_per_cta, float *__restrict__ out)
{
    // blockDim = 256 = (256/c') row lanes x c' columns, c' = min(c, 256) rounded to pow2 by the host
    extern __shared__ float red[];
    const int cw = blockDim.x;  // columns handled per pass
    long long r_begin = (long long)blockIdx.x * rows_per_cta;
    long long r_end = r_begin + rows_per_cta < m ? r_begin + rows_per_cta : m;
    for (int c0 = 0; c0 < c; c0 += cw) {
        int col = c0 + threadIdx.x;
        float s = 0.f;
        if (col < c)
            for (long long r = r_begin + threadIdx.y; r < r_end; r += blockDim.y) s += __ldg(dy + r * c + col);
        red[threadIdx.y * cw + threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.y == 0 && col < c) {
            float t = 0.f;
            for (int q = 0; q < (int)blockDim.y; ++q) t += red[q * cw + threadIdx.x];
            atomicAdd(out + col, t);
        }
        __syncthreads();
    }
}

__global__ void weight_transpose_kernel(const float *__restrict__ w, int cout, int K, int cin, int flip,
                                        float *__restrict__ wt)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)cout * K * cin;
    if (t >= total) return;
    // t indexes the destination (cin, K, cout) so that writes are coalesced
    int co = (int)(t % cout);
    int k = (int)((t / cout) % K);
    int ci = (int)(t / ((long long)cout * K));
    int ks = flip ? K - 1 - k : k;
    wt[t] = __ldg(w + ((size_t)co * K + ks) * cin + ci);
}

__global__ void conv2d_table_kernel(int n, int h, int w, int kh, int kw, int stride, int pad, int transposed,
                                    int ho, int wo, int32_t *__restrict__ nbr)
{
    const int K = kh * kw;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * ho * wo * K;
    if (t >= total) return;
    int tap = (int)(t % K);
    long long o = t / K;
    int ox = (int)(o % wo), oy = (int)((o / wo) % ho), b = (int)(o / ((long long)wo * ho));
    int ky = tap / kw, kx = tap % kw;
    int32_t r = -1;
    if (!transposed) {
        int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy >= 0 && ix >= 0 && iy < h && ix < w) r = (b * h + iy) * w + ix;
    } else {
        int ny = oy + pad - ky, nx = ox + pad - kx;
        if (ny >= 0 && nx >= 0 && ny % stride == 0 && nx % stride == 0) {
            int iy = ny / stride, ix = nx / stride;
            if (iy < h && ix < w) r = (b * h + iy) * w + ix;
        }
    }
    nbr[t] = r;
}

template <bool BWD>
__global__ void dense_scatter_kernel(const float *__restrict__ src, const int32_t *__restrict__ coords, long long m,
                                     int c, int d, int h, int w, int channels_last, float *__restrict__ dst)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * c) return;
    long long v = t / c;
    int f = (int)(t - v * c);
    int4 q = __ldg(reinterpret_cast<const int4 *>(coords) + v);
    size_t o;
    if (channels_last) o = (((size_t)q.x * h + q.z) * w + q.w) * ((size_t)c * d) + (size_t)f * d + q.y;
    else o = ((((size_t)q.x * c + f) * d + q.y) * h + q.z) * w + q.w;
    if (BWD) dst[t] = __ldg(src + o);
    else dst[o] = src[t];
}

}  // namespace
}  // namespace cpd

using namespace cpd;

static size_t image_bytes(int64_t m, int32_t c) { return align_up((size_t)m * (size_t)c * 4, 256); }   // split-row image of an (m, c) matrix

extern "C" size_t cpd_gather_gemm_workspace_bytes(int64_t m_in, int64_t, int32_t cin, int32_t K, int32_t cout, int32_t algo,
                                                  int32_t have_x_split)
{
    if (algo == CPD_ALGO_SIMT) return 0;
    if (!gather_gemm_tc_supported(cin, K, cout)) return 0;
    return 256 + gather_gemm_tc_workspace(cin, K, cout) + (have_x_split ? 0 : image_bytes(m_in, cin));
}

extern "C" int32_t cpd_tile_tap_masks(const int32_t *nbr, int64_t m, int32_t K, uint32_t *masks, cpd_stream_t stream)
{
    CPD_REQUIRE(nbr && masks && m >= 0, CPD_ERR_BAD_ARG, "cpd_tile_tap_masks: bad argument");
    return tile_tap_masks(nbr, m, K, masks, (cudaStream_t)stream);
}

extern "C" int32_t cpd_tap_block_keys(const int32_t *nbr, int64_t m, int32_t K, int32_t taps_per_block, int32_t *keys, cpd_stream_t stream)
{
    CPD_REQUIRE(nbr && keys && m >= 0, CPD_ERR_BAD_ARG, "cpd_tap_block_keys: bad argument");
    return tap_block_keys(nbr, m, K, taps_per_block, keys, (cudaStream_t)stream);
}

extern "C" int32_t cpd_gather_gemm(const float *x, const void *x_split, int64_t m_in, int32_t cin, const float *w, int32_t K,
                                   int32_t cout, const int32_t *nbr, const uint32_t *tile_masks, const int32_t *out_rows,
                                   int64_t m_out, const float *bias, const float *scale,
                                   const float *shift, const float *residual, int32_t relu, float *stats, float *y,
                                   int32_t algo, void *ws, size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE((x || x_split) && w && nbr && y, CPD_ERR_BAD_ARG, "cpd_gather_gemm: null argument");
    CPD_REQUIRE(m_in >= 0 && m_out >= 0 && cin >= 1 && cout >= 1 && K >= 1 && K <= 64, CPD_ERR_BAD_ARG, "cpd_gather_gemm: bad sizes");
    CPD_REQUIRE((scale == nullptr) == (shift == nullptr), CPD_ERR_BAD_ARG, "cpd_gather_gemm: scale and shift go together");
    CPD_REQUIRE(m_out < (1ll << 31) && m_in < (1ll << 31), CPD_ERR_UNSUPPORTED, "cpd_gather_gemm: more than 2^31 rows");
    if (m_out == 0) return CPD_OK;
    const size_t need = cpd_gather_gemm_workspace_bytes(m_in, m_out, cin, K, cout, CPD_ALGO_TCGEN05, x_split != nullptr);
    bool tc = false;
    if (algo == CPD_ALGO_TCGEN05) {
        CPD_REQUIRE(gather_gemm_tc_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "cpd_gather_gemm: tcgen05 path needs cin%%8==0, cout in {16,32,64,128,256k}");
        CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_gather_gemm: workspace too small (cpd_gather_gemm_workspace_bytes)");
        tc = true;
    } else if (algo == CPD_ALGO_AUTO) {
        // a shape the tensor-core kernel takes never drops to the SIMT kernel silently: a short workspace is an error
        tc = gather_gemm_tc_supported(cin, K, cout);
        CPD_REQUIRE(!tc || (ws && ws_bytes >= need), CPD_ERR_WORKSPACE, "cpd_gather_gemm: workspace too small (cpd_gather_gemm_workspace_bytes)");
    }
    if (tc) {
        uint8_t *p = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)ws, 256));
        const void *xs = x_split;
        if (!xs) {                                   // build the split-row image of x in the workspace
            int32_t st = split_rows(x, m_in, cin, p, nullptr, stream);
            if (st) return st;
            xs = p;
            p += image_bytes(m_in, cin);
        }
        const size_t left = ws_bytes - (size_t)(p - reinterpret_cast<uint8_t *>(ws));
        return gather_gemm_tc(xs, cin, w, K, cout, nbr, tile_masks, out_rows, m_out, bias, scale, shift, residual, relu, stats, y, p, left, stream);
    }
    CPD_REQUIRE(x, CPD_ERR_BAD_ARG, "cpd_gather_gemm: the SIMT kernel needs the fp32 rows");
    CPD_REQUIRE(!out_rows, CPD_ERR_UNSUPPORTED, "cpd_gather_gemm: out_rows needs the tcgen05 kernel");
    if (stats) CPD_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * (size_t)cout, stream));
    Epilogue ep{bias, scale, shift, residual, stats, relu};
    if (cout > 32) launch_gg<64, 64, 4, 4>(x, cin, w, K, cout, nbr, m_out, ep, y, stream);
    else if (cout > 16) launch_gg<128, 32, 4, 4>(x, cin, w, K, cout, nbr, m_out, ep, y, stream);
    else launch_gg<128, 16, 4, 2>(x, cin, w, K, cout, nbr, m_out, ep, y, stream);
    return launch_status("cpd_gather_gemm");
}

extern "C" size_t cpd_gather_wgrad_workspace_bytes(int64_t m_in, int64_t m_out, int32_t cin, int32_t K, int32_t cout,
                                                   int32_t have_x_split, int32_t have_dy_split)
{
    if (!gather_wgrad_rows_supported(cin, K, cout)) return 0;
    return 256 + (have_x_split ? 0 : image_bytes(m_in, cin)) + (have_dy_split ? 0 : image_bytes(m_out, cout));
}

extern "C" int32_t cpd_gather_wgrad(const float *x, const void *x_split, int64_t m_in, int32_t cin, const float *dy,
                                    const void *dy_split, int64_t m_out, int32_t cout, const int32_t *nbr,
                                    int32_t nbr_tap_major, int32_t K, float *dw, float *dbias, int32_t algo, void *ws,
                                    size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(x && dy && nbr && dw, CPD_ERR_BAD_ARG, "cpd_gather_wgrad: null argument");
    CPD_REQUIRE(m_in >= 0 && m_out >= 0 && cin >= 1 && cout >= 1 && K >= 1 && K <= 64, CPD_ERR_BAD_ARG, "cpd_gather_wgrad: bad sizes");
    CPD_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cout * K * cin, stream));
    if (dbias) CPD_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)cout, stream));
    if (m_out == 0) return CPD_OK;
    const bool tap_major_ok = nbr_tap_major || K == 1;
    const size_t need = cpd_gather_wgrad_workspace_bytes(m_in, m_out, cin, K, cout, x_split != nullptr, dy_split != nullptr);
    const bool shape_ok = gather_wgrad_rows_supported(cin, K, cout) && tap_major_ok;
    bool tc = false;
    if (algo == CPD_ALGO_TCGEN05) {
        CPD_REQUIRE(shape_ok, CPD_ERR_UNSUPPORTED, "cpd_gather_wgrad: tcgen05 path needs cin, cout >= 8 and multiples of 8 and the tap-major table");
        tc = true;
    } else if (algo == CPD_ALGO_AUTO) {
        tc = shape_ok;
    }
    if (tc) {
        // never a silent drop to the SIMT kernel: a short workspace is an error
        CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_gather_wgrad: workspace too small (cpd_gather_wgrad_workspace_bytes)");
        int32_t st;
        uint8_t *p = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)ws, 256));
        const void *xs = x_split, *dys = dy_split;
        if (!xs) {
            st = split_rows(x, m_in, cin, p, nullptr, stream);
            if (st) return st;
            xs = p;
            p += image_bytes(m_in, cin);
        }
        if (!dys) {
            st = split_rows(dy, m_out, cout, p, nullptr, stream);
            if (st) return st;
            dys = p;
        }
        st = gather_wgrad_rows_tc(xs, cin, dys, m_out, cout, nbr, K, dw, stream);
        if (st) return st;
    }
    const int co_tiles = (int)div_up(cout, 64), ci_tiles = (int)div_up(cin, 64);
    int S = (int)div_up(148 * 4, (long long)K * co_tiles * ci_tiles);
    int max_s = (int)div_up(m_out, 256);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    int rows_per_cta = (int)div_up(div_up(m_out, S), WG_R) * WG_R;
    S = (int)div_up(m_out, rows_per_cta);
    if (!tc) {
        dim3 grid(K, S, co_tiles * ci_tiles);
        gather_wgrad_simt<<<grid, 256, 0, stream>>>(x, cin, dy, cout, nbr, K, m_out, rows_per_cta, ci_tiles, nbr_tap_major, dw);
        count_launch();
    }
    if (dbias) {
        int cw = 1;
        while (cw < cout && cw < 256) cw <<= 1;
        dim3 block(cw, 256 / cw);
        int ctas = (int)(m_out / 2048 > 0 ? (m_out / 2048 < 592 ? m_out / 2048 : 592) : 1);
        int rpc = (int)div_up(m_out, ctas);
        column_sum_kernel<<<(unsigned)div_up(m_out, rpc), block, 256 * sizeof(float), stream>>>(dy, m_out, cout, rpc, dbias);
        count_launch();
    }
    return launch_status("cpd_gather_wgrad");
}

extern "C" int32_t cpd_weight_transpose(const float *w, int32_t cout, int32_t K, int32_t cin, int32_t flip_taps,
                                        float *wt, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(w && wt && cout >= 1 && K >= 1 && cin >= 1, CPD_ERR_BAD_ARG, "cpd_weight_transpose: bad argument");
    long long total = (long long)cout * K * cin;
    weight_transpose_kernel<<<(unsigned)div_up(total, 256), 256, 0, stream>>>(w, cout, K, cin, flip_taps, wt);
    count_launch();
    return launch_status("cpd_weight_transpose");
}

extern "C" int32_t cpd_conv2d_table(int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                                    int32_t transposed, int32_t ho, int32_t wo, int32_t *nbr, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(nbr && n >= 1 && h >= 1 && w >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 64 && stride >= 1 && pad >= 0 && ho >= 1 && wo >= 1,
                CPD_ERR_BAD_ARG, "cpd_conv2d_table: bad argument");
    CPD_REQUIRE((long long)n * h * w < (1ll << 31) && (long long)n * ho * wo < (1ll << 31), CPD_ERR_UNSUPPORTED, "cpd_conv2d_table: image batch too large");
    long long total = (long long)n * ho * wo * kh * kw;
    conv2d_table_kernel<<<(unsigned)div_up(total, 256), 256, 0, stream>>>(n, h, w, kh, kw, stride, pad, transposed, ho, wo, nbr);
    count_launch();
    return launch_status("cpd_conv2d_table");
}

static int32_t dense_common(const float *src, const int32_t *coords, int64_t m, int32_t c, int32_t batch,
                            const int32_t *shape3, int32_t channels_last, float *dst, bool bwd, cudaStream_t stream)
{
    CPD_REQUIRE(src && dst && shape3 && c >= 1 && batch >= 1 && m >= 0, CPD_ERR_BAD_ARG, "cpd_sparse_to_dense: bad argument");
    CPD_REQUIRE(m == 0 || (coords && ((uintptr_t)coords & 15) == 0), CPD_ERR_MISALIGNED, "cpd_sparse_to_dense: coords must be 16-byte aligned");
    const int d = shape3[0], h = shape3[1], w = shape3[2];
    if (!bwd) CPD_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)batch * c * d * h * w, stream));
    if (m == 0) return CPD_OK;
    unsigned grid = (unsigned)div_up(m * (long long)c, 256);
    if (bwd) dense_scatter_kernel<true><<<grid, 256, 0, stream>>>(src, coords, m, c, d, h, w, channels_last, dst);
    else dense_scatter_kernel<false><<<grid, 256, 0, stream>>>(src, coords, m, c, d, h, w, channels_last, dst);
    count_launch();
    return launch_status("cpd_sparse_to_dense");
}

extern "C" int32_t cpd_sparse_to_dense(const float *feat, const int32_t *coords, int64_t m, int32_t c, int32_t batch,
                                       const int32_t *shape3, int32_t channels_last, float *out, cpd_stream_t stream)
{
    return dense_common(feat, coords, m, c, batch, shape3, channels_last, out, false, (cudaStream_t)stream);
}

extern "C" int32_t cpd_sparse_to_dense_bwd(const float *dout, const int32_t *coords, int64_t m, int32_t c, int32_t batch,
                                           const int32_t *shape3, int32_t channels_last, float *dfeat, cpd_stream_t stream)
{
    return dense_common(dout, coords, m, c, batch, shape3, channels_last, dfeat, true, (cudaStream_t)stream);
}
