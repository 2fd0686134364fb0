// C ABI of the dense 2-D convolutions of the BEV backbone and the CenterHead (NHWC, stride 1 / ConvTranspose k == s):
// the reference runs them through cuDNN (nn.Conv2d / nn.ConvTranspose2d, cpd/models/backbones_2d/base_bev_backbone.py:31-59,
// cpd/models/dense_heads/center_head.py:11-45,73-80).  Forward and input-gradient run on the TMA-fed tcgen05 kernel in
// spconv_tc.cu (tiled cp.async.bulk.tensor loads of shifted pixel patches, no neighbour table); the weight gradient reuses
// the row-stationary tcgen05 kernel of wgrad_tc.cu over a pixel table generated into the workspace.
#include "common.cuh"

namespace cpd {
bool conv2d_tc_supported(int32_t cin, int32_t kh, int32_t kw, int32_t cout);
size_t conv2d_tc_workspace(int32_t cin, int32_t kh, int32_t kw, int32_t cout);
int32_t conv2d_tc(const void *xs, int32_t n, int32_t h, int32_t w_, int32_t cin, const float *w, int32_t kh, int32_t kw, int32_t pad,
                  int32_t cout, const float *bias, const float *scale, const float *shift, const float *residual, int32_t relu,
                  float *stats, int32_t clear_stats, float *y, int32_t out_h, int32_t out_w, int32_t out_sy, int32_t out_sx, int32_t out_oy,
                  int32_t out_ox, void *ws, size_t ws_bytes, cudaStream_t stream);
bool gather_wgrad_rows_supported(int32_t cin, int32_t K, int32_t cout);
int32_t gather_wgrad_rows_tc(const void *xs, int32_t cin, const void *dys, int64_t m_out, int32_t cout, const int32_t *nbr_t,
                             int32_t K, float *dw, cudaStream_t stream);

namespace {

// (cout, K, cin) -> (cin, K, cout) with the taps reversed: the kernel of the input-gradient convolution
__global__ void flip_transpose_kernel(const float *__restrict__ w, int cout, int K, int cin, float *__restrict__ wt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)cout * K * cin) return;
    const int co = (int)(t % cout), k = (int)((t / cout) % K), ci = (int)(t / ((long long)cout * K));
    wt[t] = __ldg(w + ((size_t)co * K + (K - 1 - k)) * cin + ci);
}

// torch ConvTranspose2d weight (cin, cout, s, s) -> s*s matrices (cout, cin), tap-major
__global__ void convt_weight_kernel(const float *__restrict__ w, int cin, int cout, int ss, float *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ss * cout * cin) return;
    const int ci = (int)(t % cin), co = (int)((t / cin) % cout), k = (int)(t / ((long long)cin * cout));
    out[t] = __ldg(w + ((size_t)ci * cout + co) * ss + k);
}

// tap-major pixel table of a stride-1 convolution: nbr_t[k][o] = input pixel of output pixel o under tap k, or -1
__global__ void conv2d_table_t_kernel(int n, int h, int w, int kh, int kw, int pad, int ho, int wo, int32_t *__restrict__ nbr_t)
{
    const long long m = (long long)n * ho * wo;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * kh * kw) return;
    const int tap = (int)(t / m);
    const long long o = t - (long long)tap * m;
    const int ox = (int)(o % wo), oy = (int)((o / wo) % ho), b = (int)(o / ((long long)wo * ho));
    const int iy = oy - pad + tap / kw, ix = ox - pad + tap % kw;
    nbr_t[t] = (iy >= 0 && ix >= 0 && iy < h && ix < w) ? (int32_t)(((long long)b * h + iy) * w + ix) : -1;
}

}  // namespace
}  // namespace cpd

using namespace cpd;

extern "C" int32_t cpd_conv2d_supported(int32_t cin, int32_t kh, int32_t kw, int32_t cout) { return conv2d_tc_supported(cin, kh, kw, cout) ? 1 : 0; }

extern "C" size_t cpd_conv2d_workspace_bytes(int32_t cin, int32_t kh, int32_t kw, int32_t cout)
{
    if (!conv2d_tc_supported(cin, kh, kw, cout)) return 0;
    return 256 + conv2d_tc_workspace(cin, kh, kw, cout);
}

extern "C" int32_t cpd_conv2d_fwd(const void *x_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t kh,
                                  int32_t kw, int32_t pad, int32_t cout, const float *bias, const float *scale, const float *shift,
                                  const float *residual, int32_t relu, float *stats, float *y, void *ws, size_t ws_bytes,
                                  cpd_stream_t stream)
{
    CPD_REQUIRE(x_split && wgt && y && n >= 1 && h >= 1 && w >= 1, CPD_ERR_BAD_ARG, "cpd_conv2d_fwd: bad argument");
    CPD_REQUIRE((scale == nullptr) == (shift == nullptr), CPD_ERR_BAD_ARG, "cpd_conv2d_fwd: scale and shift go together");
    const int ho = h + 2 * pad - kh + 1, wo = w + 2 * pad - kw + 1;
    return conv2d_tc(x_split, n, h, w, cin, wgt, kh, kw, pad, cout, bias, scale, shift, residual, relu, stats, 1, y, ho, wo, 1, 1, 0, 0, ws,
                     ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t cpd_conv2d_dgrad_workspace_bytes(int32_t cin, int32_t kh, int32_t kw, int32_t cout)
{
    if (!conv2d_tc_supported(cout, kh, kw, cin)) return 0;
    return 512 + align_up((size_t)cin * kh * kw * cout * 4, 256) + conv2d_tc_workspace(cout, kh, kw, cin);
}

// dx (n, h, w, cin) from dy (n, ho, wo, cout): the same kernel on dy with the flipped, transposed weights and padding k - 1 - pad
extern "C" int32_t cpd_conv2d_dgrad(const void *dy_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t kh,
                                    int32_t kw, int32_t pad, int32_t cout, float *dx, void *ws, size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(dy_split && wgt && dx && n >= 1, CPD_ERR_BAD_ARG, "cpd_conv2d_dgrad: bad argument");
    CPD_REQUIRE(kh == kw && pad <= kh - 1, CPD_ERR_UNSUPPORTED, "cpd_conv2d_dgrad: square kernels with pad <= k - 1");
    CPD_REQUIRE(conv2d_tc_supported(cout, kh, kw, cin), CPD_ERR_UNSUPPORTED, "cpd_conv2d_dgrad: needs cout %% 64 == 0 and cin in {16,32,64,128,256k}");
    const size_t need = cpd_conv2d_dgrad_workspace_bytes(cin, kh, kw, cout);
    CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_conv2d_dgrad: workspace too small");
    const int ho = h + 2 * pad - kh + 1, wo = w + 2 * pad - kw + 1, K = kh * kw;
    float *wt = reinterpret_cast<float *>(align_up((size_t)(uintptr_t)ws, 256));
    const long long total = (long long)cout * K * cin;
    flip_transpose_kernel<<<(unsigned)div_up(total, 256), 256, 0, stream>>>(wgt, cout, K, cin, wt);
    count_launch();
    uint8_t *rest = reinterpret_cast<uint8_t *>(wt) + align_up((size_t)total * 4, 256);
    return conv2d_tc(dy_split, n, ho, wo, cout, wt, kh, kw, kh - 1 - pad, cin, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, dx, h, w, 1,
                     1, 0, 0, rest, ws_bytes - (size_t)(rest - reinterpret_cast<uint8_t *>(ws)), stream);
}

extern "C" size_t cpd_conv2d_wgrad_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t kh, int32_t kw, int32_t pad)
{
    const long long ho = h + 2 * pad - kh + 1, wo = w + 2 * pad - kw + 1;
    if (ho < 1 || wo < 1) return 0;
    return 256 + (size_t)n * ho * wo * kh * kw * 4;
}

// dw (cout, kh*kw, cin) = sum over pixels of dy[p, co] * x[p + tap, ci]; dw is overwritten
extern "C" int32_t cpd_conv2d_wgrad(const void *x_split, const void *dy_split, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t kh,
                                    int32_t kw, int32_t pad, int32_t cout, float *dw, void *ws, size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(x_split && dy_split && dw && n >= 1, CPD_ERR_BAD_ARG, "cpd_conv2d_wgrad: bad argument");
    const int K = kh * kw;
    CPD_REQUIRE(gather_wgrad_rows_supported(cin, K, cout), CPD_ERR_UNSUPPORTED, "cpd_conv2d_wgrad: needs cin, cout multiples of 8");
    const size_t need = cpd_conv2d_wgrad_workspace_bytes(n, h, w, kh, kw, pad);
    CPD_REQUIRE(need && ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_conv2d_wgrad: workspace too small");
    const int ho = h + 2 * pad - kh + 1, wo = w + 2 * pad - kw + 1;
    const long long m = (long long)n * ho * wo;
    CPD_REQUIRE(m < (1ll << 31) && (long long)n * h * w < (1ll << 31), CPD_ERR_UNSUPPORTED, "cpd_conv2d_wgrad: image batch too large");
    int32_t *nbr_t = reinterpret_cast<int32_t *>(align_up((size_t)(uintptr_t)ws, 256));
    conv2d_table_t_kernel<<<(unsigned)div_up(m * K, 256), 256, 0, stream>>>(n, h, w, kh, kw, pad, ho, wo, nbr_t);
    count_launch();
    CPD_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cout * K * cin, stream));
    return gather_wgrad_rows_tc(x_split, cin, dy_split, m, cout, nbr_t, K, dw, stream);
}

extern "C" size_t cpd_convt2d_workspace_bytes(int32_t cin, int32_t s, int32_t cout)
{
    if (!conv2d_tc_supported(cin, 1, 1, cout)) return 0;
    return 512 + align_up((size_t)s * s * cout * cin * 4, 256) + (size_t)s * s * align_up(conv2d_tc_workspace(cin, 1, 1, cout), 256);
}

// nn.ConvTranspose2d with kernel == stride == s (base_bev_backbone.py:48-59): every output pixel has exactly one contributing
// input pixel, so it is s*s 1x1 GEMMs whose epilogues place their rows at (y*s + ky, x*s + kx) of the (n, h*s, w*s, cout) map.
// wgt: torch layout (cin, cout, s, s).  stats accumulate over all s*s launches (BatchNorm over the whole output map).
extern "C" int32_t cpd_convt2d_fwd(const void *x_split, int32_t n, int32_t h, int32_t w, int32_t cin, const float *wgt, int32_t s,
                                   int32_t cout, const float *scale, const float *shift, int32_t relu, float *stats, float *y, void *ws,
                                   size_t ws_bytes, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(x_split && wgt && y && n >= 1 && s >= 1 && s <= 4, CPD_ERR_BAD_ARG, "cpd_convt2d_fwd: bad argument");
    CPD_REQUIRE(conv2d_tc_supported(cin, 1, 1, cout), CPD_ERR_UNSUPPORTED, "cpd_convt2d_fwd: needs cin %% 64 == 0 and cout in {16,32,64,128,256k}");
    const size_t need = cpd_convt2d_workspace_bytes(cin, s, cout);
    CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_convt2d_fwd: workspace too small");
    float *wk = reinterpret_cast<float *>(align_up((size_t)(uintptr_t)ws, 256));
    const long long total = (long long)s * s * cout * cin;
    convt_weight_kernel<<<(unsigned)div_up(total, 256), 256, 0, stream>>>(wgt, cin, cout, s * s, wk);
    count_launch();
    uint8_t *rest = reinterpret_cast<uint8_t *>(wk) + align_up((size_t)total * 4, 256);
    const size_t per = align_up(conv2d_tc_workspace(cin, 1, 1, cout), 256);
    for (int k = 0; k < s * s; ++k) {
        const int32_t st = conv2d_tc(x_split, n, h, w, cin, wk + (size_t)k * cout * cin, 1, 1, 0, cout, nullptr, scale, shift, nullptr, relu,
                                     stats, k == 0, y, h * s, w * s, s, s, k / s, k % s, rest + (size_t)k * per, per, stream);
        if (st) return st;
    }
    return CPD_OK;
}
