// Training-mode BatchNorm over row matrices (M, C) for sm_100a, fused with ReLU and the residual add.
//
// Replaces, on the training path, nn.BatchNorm1d / nn.BatchNorm2d + nn.ReLU (+ `out + identity`) as the
// reference chains them after every convolution (cpd/models/backbones_3d/spconv_backbone.py:13-35,
// 100-136,410; cpd/models/backbones_2d/base_bev_backbone.py:31-59; cpd/models/dense_heads/center_head.py:
// 22-27,73-80).  The per-channel sum / sum of squares come out of the convolution epilogue (cpd_gather_gemm
// `stats`), so the forward is ONE streaming pass (read x, write y) and the backward is a reduce pass plus
// one streaming pass; all HBM-bound elementwise work with float4 accesses, grids a multiple of 148 SMs.
//   y  = relu?( (x - mean) * invstd * gamma + beta (+ residual) )
//   dz = dy * (y > 0)?          dbeta = sum dz        dgamma = sum dz * xhat
//   dx = gamma * invstd * (dz - dbeta / M - xhat * dgamma / M)          dresidual = dz
// Both streaming passes can also emit the split-row image (bf16 hi | lo, split.cu) of what they write: the
// next convolution (forward) / the previous one (backward) consumes exactly that tensor, so the image costs
// 4 extra bytes written per element instead of a separate read + write pass.
#include "tc_common.cuh"

namespace cpd {
namespace {

constexpr int BN_THREADS = 256;
constexpr int MAX_C = 1024;

__global__ void __launch_bounds__(BN_THREADS) bn_finalize_kernel(const float *__restrict__ stats, long long m, int c, float eps,
                                                                 float momentum, float *__restrict__ mean_invstd,
                                                                 float *__restrict__ running_mean, float *__restrict__ running_var)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const double inv_m = 1.0 / (double)m;
    const double mu = (double)stats[ch] * inv_m;
    double var = (double)stats[c + ch] * inv_m - mu * mu;
    if (var < 0.0) var = 0.0;
    mean_invstd[ch] = (float)mu;
    mean_invstd[c + ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {   // torch semantics: running_var tracks the UNBIASED batch variance
        const double unbiased = m > 1 ? var * (double)m / (double)(m - 1) : var;
        running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mu);
        running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * unbiased);
    }
}

// one float4 (4 consecutive channels of one row) per thread-iteration; c % 4 == 0
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ mean_invstd,
                                                              const float *__restrict__ gamma, const float *__restrict__ beta,
                                                              const float *__restrict__ residual, int relu, long long m, int c,
                                                              float *__restrict__ y, uint8_t *__restrict__ y_split)
{
    extern __shared__ float sh[];   // scale[c], shift[c]
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
        const float sc = g * mean_invstd[c + ch];
        sh[ch] = sc;
        sh[c + ch] = b - mean_invstd[ch] * sc;
    }
    __syncthreads();
    const long long total4 = m * (long long)(c / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % (c / 4)) * 4;
        float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        v.x = fmaf(v.x, sh[ch], sh[c + ch]); v.y = fmaf(v.y, sh[ch + 1], sh[c + ch + 1]);
        v.z = fmaf(v.z, sh[ch + 2], sh[c + ch + 2]); v.w = fmaf(v.w, sh[ch + 3], sh[c + ch + 3]);
        if (residual) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(residual) + i);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        reinterpret_cast<float4 *>(y)[i] = v;
        if (y_split) {                               // row image = hi(c) | lo(c) bf16, same 4c-byte pitch as the fp32 row
            uint2 h, l;
            tc::split4b(v, h, l);
            uint8_t *p = y_split + (size_t)(i / (c / 4)) * (size_t)(4 * c) + (size_t)ch * 2;
            *reinterpret_cast<uint2 *>(p) = h;
            *reinterpret_cast<uint2 *>(p + 2 * c) = l;
        }
    }
}

// per-channel sum dz and sum dz*xhat.  Block = (c4 lanes) x (rows lanes); each block owns a slice of rows.
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                                   const float *__restrict__ dy, const float *__restrict__ mean_invstd,
                                                                   int relu, long long m, int c, int rows_per_cta,
                                                                   float *__restrict__ sums /* (2, c): dbeta, dgamma */)
{
    extern __shared__ float red[];   // [ty][c4*4][2]
    const int c4n = c / 4;
    const int lanes_c = min(c4n, BN_THREADS), lanes_r = BN_THREADS / lanes_c;
    const int tx = threadIdx.x % lanes_c, ty = threadIdx.x / lanes_c;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + (long long)rows_per_cta, m);
    for (int cb = tx; cb < c4n; cb += lanes_c) {
        const int ch = cb * 4;
        const float mu[4] = {mean_invstd[ch], mean_invstd[ch + 1], mean_invstd[ch + 2], mean_invstd[ch + 3]};
        const float is[4] = {mean_invstd[c + ch], mean_invstd[c + ch + 1], mean_invstd[c + ch + 2], mean_invstd[c + ch + 3]};
        float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        if (ty < lanes_r) {
            // UNR rows per iteration, all loads issued before the first use: the loop is a pure stream, so the bytes in
            // flight per thread are what buys bandwidth (one row at a time ran at ~45 % of the apply pass's rate)
            constexpr int UNR = 4;
            long long r = r0 + ty;
            for (; r + (long long)(UNR - 1) * lanes_r < r1; r += (long long)UNR * lanes_r) {
                float4 g[UNR], xv[UNR], yv[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const long long i = (r + (long long)u * lanes_r) * c4n + cb;
                    g[u] = __ldg(reinterpret_cast<const float4 *>(dy) + i);
                    xv[u] = __ldg(reinterpret_cast<const float4 *>(x) + i);
                    if (relu) yv[u] = __ldg(reinterpret_cast<const float4 *>(y) + i);
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    if (relu) {
                        g[u].x = yv[u].x > 0.f ? g[u].x : 0.f; g[u].y = yv[u].y > 0.f ? g[u].y : 0.f;
                        g[u].z = yv[u].z > 0.f ? g[u].z : 0.f; g[u].w = yv[u].w > 0.f ? g[u].w : 0.f;
                    }
                    s[0] += g[u].x; s[1] += g[u].y; s[2] += g[u].z; s[3] += g[u].w;
                    q[0] += g[u].x * (xv[u].x - mu[0]) * is[0]; q[1] += g[u].y * (xv[u].y - mu[1]) * is[1];
                    q[2] += g[u].z * (xv[u].z - mu[2]) * is[2]; q[3] += g[u].w * (xv[u].w - mu[3]) * is[3];
                }
            }
            for (; r < r1; r += lanes_r) {
                const long long i = r * c4n + cb;
                float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + i);
                const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
                if (relu) {
                    const float4 yv = __ldg(reinterpret_cast<const float4 *>(y) + i);
                    g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f; g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
                }
                s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
                q[0] += g.x * (xv.x - mu[0]) * is[0]; q[1] += g.y * (xv.y - mu[1]) * is[1];
                q[2] += g.z * (xv.z - mu[2]) * is[2]; q[3] += g.w * (xv.w - mu[3]) * is[3];
            }
        }
        // reduce over ty through shared memory
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) { red[((ty * lanes_c + tx) * 4 + k) * 2] = s[k]; red[((ty * lanes_c + tx) * 4 + k) * 2 + 1] = q[k]; }
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = 0.f, b = 0.f;
                for (int t = 0; t < lanes_r; ++t) { a += red[((t * lanes_c + tx) * 4 + k) * 2]; b += red[((t * lanes_c + tx) * 4 + k) * 2 + 1]; }
                atomicAdd(sums + ch + k, a);
                atomicAdd(sums + c + ch + k, b);
            }
        }
    }
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                                  const float *__restrict__ dy, const float *__restrict__ mean_invstd,
                                                                  const float *__restrict__ gamma, const float *__restrict__ sums,
                                                                  int relu, long long m, int c, float *__restrict__ dx,
                                                                  uint8_t *__restrict__ dx_split, float *__restrict__ dres)
{
    extern __shared__ float sh[];   // a[c] = gamma*invstd, b[c] = dbeta/M, d[c] = invstd*dgamma/M, mu[c], is[c]
    const float inv_m = 1.f / (float)m;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float is = mean_invstd[c + ch], g = gamma ? gamma[ch] : 1.f;
        sh[ch] = g * is;
        sh[c + ch] = sums[ch] * inv_m;
        sh[2 * c + ch] = sums[c + ch] * inv_m;
        sh[3 * c + ch] = mean_invstd[ch];
        sh[4 * c + ch] = is;
    }
    __syncthreads();
    const long long total4 = m * (long long)(c / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % (c / 4)) * 4;
        float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + i);
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + i);
        if (relu) {
            const float4 yv = __ldg(reinterpret_cast<const float4 *>(y) + i);
            g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f; g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
        }
        if (dres) reinterpret_cast<float4 *>(dres)[i] = g;
        float4 o;
        o.x = sh[ch] * (g.x - sh[c + ch] - (xv.x - sh[3 * c + ch]) * sh[4 * c + ch] * sh[2 * c + ch]);
        o.y = sh[ch + 1] * (g.y - sh[c + ch + 1] - (xv.y - sh[3 * c + ch + 1]) * sh[4 * c + ch + 1] * sh[2 * c + ch + 1]);
        o.z = sh[ch + 2] * (g.z - sh[c + ch + 2] - (xv.z - sh[3 * c + ch + 2]) * sh[4 * c + ch + 2] * sh[2 * c + ch + 2]);
        o.w = sh[ch + 3] * (g.w - sh[c + ch + 3] - (xv.w - sh[3 * c + ch + 3]) * sh[4 * c + ch + 3] * sh[2 * c + ch + 3]);
        reinterpret_cast<float4 *>(dx)[i] = o;
        if (dx_split) {
            uint2 h, l;
            tc::split4b(o, h, l);
            uint8_t *p = dx_split + (size_t)(i / (c / 4)) * (size_t)(4 * c) + (size_t)ch * 2;
            *reinterpret_cast<uint2 *>(p) = h;
            *reinterpret_cast<uint2 *>(p + 2 * c) = l;
        }
    }
}

unsigned stream_grid(long long total4)
{
    long long blocks = div_up(total4, BN_THREADS * 4);
    const long long cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace
}  // namespace cpd

using namespace cpd;

extern "C" int32_t cpd_bn_train_fwd(const float *x, int64_t m, int32_t c, const float *stats, const float *gamma,
                                    const float *beta, const float *residual, int32_t relu, float eps, float momentum,
                                    float *running_mean, float *running_var, float *mean_invstd, float *y,
                                    void *y_split, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(x && stats && mean_invstd && y && m >= 1 && c >= 4 && c % 4 == 0 && c <= MAX_C, CPD_ERR_BAD_ARG,
                "cpd_bn_train_fwd: bad argument (need c %% 4 == 0, c <= %d, m >= 1)", MAX_C);
    CPD_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CPD_ERR_BAD_ARG, "cpd_bn_train_fwd: running stats go together");
    CPD_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual | (uintptr_t)y_split) & 15) == 0, CPD_ERR_MISALIGNED, "cpd_bn_train_fwd: 16-byte alignment");
    CPD_REQUIRE(!y_split || c % 8 == 0, CPD_ERR_UNSUPPORTED, "cpd_bn_train_fwd: the split-row image needs c %% 8 == 0");
    bn_finalize_kernel<<<(unsigned)div_up(c, BN_THREADS), BN_THREADS, 0, stream>>>(stats, m, c, eps, momentum, mean_invstd, running_mean, running_var);
    bn_apply_kernel<<<stream_grid(m * (c / 4)), BN_THREADS, 2 * c * sizeof(float), stream>>>(x, mean_invstd, gamma, beta, residual, relu, m, c, y,
                                                                                            reinterpret_cast<uint8_t *>(y_split));
    count_launch(2);
    return launch_status("cpd_bn_train_fwd");
}

extern "C" int32_t cpd_bn_train_bwd(const float *x, const float *y, const float *dy, int64_t m, int32_t c,
                                    const float *mean_invstd, const float *gamma, int32_t relu, float *dx,
                                    void *dx_split, float *dresidual, float *dgamma_dbeta /* (2,c): dbeta, dgamma */,
                                    cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(x && dy && mean_invstd && dx && dgamma_dbeta && (y || !relu) && m >= 1 && c >= 4 && c % 4 == 0 && c <= MAX_C,
                CPD_ERR_BAD_ARG, "cpd_bn_train_bwd: bad argument");
    CPD_REQUIRE(!dx_split || (c % 8 == 0 && ((uintptr_t)dx_split & 15) == 0), CPD_ERR_UNSUPPORTED, "cpd_bn_train_bwd: the split-row image needs c %% 8 == 0");
    CPD_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, sizeof(float) * 2 * c, stream));
    int ctas = (int)(m / 256 > 0 ? (m / 256 < 148 * 8 ? m / 256 : 148 * 8) : 1);
    int rpc = (int)div_up(m, ctas);
    const int lanes_c = c / 4 < BN_THREADS ? c / 4 : BN_THREADS;
    bn_bwd_reduce_kernel<<<(unsigned)div_up(m, rpc), BN_THREADS, sizeof(float) * 8 * BN_THREADS, stream>>>(
        x, y, dy, mean_invstd, relu, m, c, rpc, dgamma_dbeta);
    (void)lanes_c;
    bn_bwd_apply_kernel<<<stream_grid(m * (c / 4)), BN_THREADS, 5 * c * sizeof(float), stream>>>(x, y, dy, mean_invstd, gamma, dgamma_dbeta,
                                                                                              relu, m, c, dx, reinterpret_cast<uint8_t *>(dx_split),
                                                                                              dresidual);
    count_launch(2);
    return launch_status("cpd_bn_train_bwd");
}
