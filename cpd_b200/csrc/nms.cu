// Rotated / axis-aligned BEV IoU and NMS for sm_100a, built on warp primitives.
//
// Replaces cpd/ops/iou3d_nms (src/iou3d_nms.h:9-12): boxes_overlap_bev_gpu,
// boxes_iou_bev_gpu, nms_gpu, nms_normal_gpu.  Differences in structure, not in results:
//   * the geometry follows the reference's arithmetic step for step (same operand order,
//     same fp32/fp64 promotions, libdevice cosf/sinf/atan2f, default FMA contraction), so
//     IoU values and suppression bits are meant to be bit-identical
//     (iou3d_nms_kernel.cu:35-234);
//   * mask build: one 64-thread CTA per 64 x 64 tile of the UPPER triangle only (the reference
//     computes all N^2 tiles, iou3d_nms_kernel.cu:275, and never reads the lower ones,
//     iou3d_nms.cpp:128); the column tile is staged through shared memory, each lane keeps
//     its row box and its 64-bit word in registers, and pairs whose circumscribed circles
//     are disjoint skip the polygon clipping (the reference's result for them is exactly
//     0.0f: no crossing, no contained corner, zero area);
//   * greedy suppression runs ON DEVICE (single CTA, 64-box chunks: the diagonal word of
//     each chunk is resolved by one lane with shuffles, the survivors' rows are then OR-ed
//     into the removal vector by the whole CTA), replacing cudaMalloc + blocking
//     cudaMemcpy D2H + host loop + cudaFree of iou3d_nms.cpp:103-132.
#include "common.cuh"

namespace cpd {
namespace {

constexpr float IOU_EPS = 1e-8f;

struct P2 { float x, y; };

__device__ __forceinline__ float cross2(const P2 &a, const P2 &b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float cross3(const P2 &p1, const P2 &p2, const P2 &p0)
{
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__device__ __forceinline__ int bbox_overlap(const P2 &p1, const P2 &p2, const P2 &q1, const P2 &q2)
{
    return min(p1.x, p2.x) <= max(q1.x, q2.x) && min(q1.x, q2.x) <= max(p1.x, p2.x) &&
           min(p1.y, p2.y) <= max(q1.y, q2.y) && min(q1.y, q2.y) <= max(p1.y, p2.y);
}

__device__ __forceinline__ int inside_box(const float *box, const P2 &p)
{
    const float MARGIN = 1e-2;
    float center_x = box[0], center_y = box[1];
    float angle_cos = cos(-box[6]), angle_sin = sin(-box[6]);
    float rot_x = (p.x - center_x) * angle_cos + (p.y - center_y) * (-angle_sin);
    float rot_y = (p.x - center_x) * angle_sin + (p.y - center_y) * angle_cos;
    return (fabs(rot_x) < box[3] / 2 + MARGIN && fabs(rot_y) < box[4] / 2 + MARGIN);
}

__device__ __forceinline__ int seg_cross(const P2 &p1, const P2 &p0, const P2 &q1, const P2 &q0, P2 &ans)
{
    if (bbox_overlap(p0, p1, q0, q1) == 0) return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabs(s5 - s1) > IOU_EPS) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

__device__ __forceinline__ void spin(const P2 &c, const float angle_cos, const float angle_sin, P2 &p)
{
    float new_x = (p.x - c.x) * angle_cos + (p.y - c.y) * (-angle_sin) + c.x;
    float new_y = (p.x - c.x) * angle_sin + (p.y - c.y) * angle_cos + c.y;
    p.x = new_x; p.y = new_y;
}

__device__ __forceinline__ int ang_gt(const P2 &a, const P2 &b, const P2 &c)
{
    return atan2(a.y - c.y, a.x - c.x) > atan2(b.y - c.y, b.x - c.x);
}

__device__ float rot_overlap(const float *box_a, const float *box_b)
{
    float a_angle = box_a[6], b_angle = box_b[6];
    float a_dx_half = box_a[3] / 2, b_dx_half = box_b[3] / 2, a_dy_half = box_a[4] / 2, b_dy_half = box_b[4] / 2;
    float a_x1 = box_a[0] - a_dx_half, a_y1 = box_a[1] - a_dy_half;
    float a_x2 = box_a[0] + a_dx_half, a_y2 = box_a[1] + a_dy_half;
    float b_x1 = box_b[0] - b_dx_half, b_y1 = box_b[1] - b_dy_half;
    float b_x2 = box_b[0] + b_dx_half, b_y2 = box_b[1] + b_dy_half;
    P2 ca{box_a[0], box_a[1]}, cb{box_b[0], box_b[1]};
    P2 qa[5], qb[5];
    qa[0] = P2{a_x1, a_y1}; qa[1] = P2{a_x2, a_y1}; qa[2] = P2{a_x2, a_y2}; qa[3] = P2{a_x1, a_y2};
    qb[0] = P2{b_x1, b_y1}; qb[1] = P2{b_x2, b_y1}; qb[2] = P2{b_x2, b_y2}; qb[3] = P2{b_x1, b_y2};
    float a_angle_cos = cos(a_angle), a_angle_sin = sin(a_angle);
    float b_angle_cos = cos(b_angle), b_angle_sin = sin(b_angle);
    for (int k = 0; k < 4; k++) {
        spin(ca, a_angle_cos, a_angle_sin, qa[k]);
        spin(cb, b_angle_cos, b_angle_sin, qb[k]);
    }
    qa[4] = qa[0];
    qb[4] = qb[0];

    P2 poly[16];
    P2 ctr{0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            if (seg_cross(qa[i + 1], qa[i], qb[j + 1], qb[j], poly[cnt])) {
                ctr.x = ctr.x + poly[cnt].x; ctr.y = ctr.y + poly[cnt].y;
                cnt++;
            }
    for (int k = 0; k < 4; k++) {
        if (inside_box(box_a, qb[k])) { ctr.x = ctr.x + qb[k].x; ctr.y = ctr.y + qb[k].y; poly[cnt] = qb[k]; cnt++; }
        if (inside_box(box_b, qa[k])) { ctr.x = ctr.x + qa[k].x; ctr.y = ctr.y + qa[k].y; poly[cnt] = qa[k]; cnt++; }
    }
    ctr.x /= cnt;
    ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; j++)
        for (int i = 0; i < cnt - j - 1; i++)
            if (ang_gt(poly[i], poly[i + 1], ctr)) { P2 t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t; }
    float area = 0;
    for (int k = 0; k < cnt - 1; k++) {
        P2 u{poly[k].x - poly[0].x, poly[k].y - poly[0].y};
        P2 v{poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
        area += cross2(u, v);
    }
    return fabs(area) / 2.0;
}

__device__ __forceinline__ float rot_iou(const float *box_a, const float *box_b)
{
    float sa = box_a[3] * box_a[4];
    float sb = box_b[3] * box_b[4];
    float s_overlap = rot_overlap(box_a, box_b);
    return s_overlap / fmaxf(sa + sb - s_overlap, IOU_EPS);
}

__device__ __forceinline__ float axis_iou(const float *a, const float *b)
{
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    float interS = width * height;
    float Sa = a[3] * a[4];
    float Sb = b[3] * b[4];
    return interS / fmaxf(Sa + Sb - interS, IOU_EPS);
}

// ---- pairwise matrices: 16x16 pair tile per CTA, both box tiles staged in smem ---------
template <bool IOU>
__global__ void __launch_bounds__(256) pair_matrix_kernel(const float *__restrict__ a, int na,
                                                          const float *__restrict__ b, int nb,
                                                          float *__restrict__ out)
{
    __shared__ float sa[16 * 7], sb[16 * 7];
    const int tid = threadIdx.y * 16 + threadIdx.x;
    if (tid < 16 * 7) {
        int r = blockIdx.y * 16 + tid / 7;
        sa[tid] = r < na ? a[(size_t)r * 7 + tid % 7] : 0.f;
    } else if (tid >= 128 && tid < 128 + 16 * 7) {
        int t = tid - 128, r = blockIdx.x * 16 + t / 7;
        sb[t] = r < nb ? b[(size_t)r * 7 + t % 7] : 0.f;
    }
    __syncthreads();
    const int ai = blockIdx.y * 16 + threadIdx.y, bi = blockIdx.x * 16 + threadIdx.x;
    if (ai >= na || bi >= nb) return;
    float v = IOU ? rot_iou(sa + threadIdx.y * 7, sb + threadIdx.x * 7) : rot_overlap(sa + threadIdx.y * 7, sb + threadIdx.x * 7);
    out[(size_t)ai * nb + bi] = v;
}

// ---- suppression bit-matrix: upper-triangle 64x64 tiles, one lane per row box ----------
template <bool ROTATED>
__global__ void __launch_bounds__(64) nms_mask_kernel(const float *__restrict__ boxes, int n, float thresh,
                                                      unsigned long long *__restrict__ mask)
{
    // linear upper-triangle tile id -> (row tile, col tile), col >= row
    const int cb = (n + 63) >> 6;
    int rt = 0, rem = blockIdx.x;
    while (rem >= cb - rt) { rem -= cb - rt; ++rt; }
    const int ct = rt + rem;
    __shared__ float cbox[64 * 7];
    const int col_size = min(n - ct * 64, 64), row_size = min(n - rt * 64, 64);
    for (int t = threadIdx.x; t < col_size * 7; t += 64) cbox[t] = boxes[(size_t)ct * 64 * 7 + t];
    __syncthreads();
    if ((int)threadIdx.x >= row_size) return;
    const int row = rt * 64 + threadIdx.x;
    float rb[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) rb[j] = boxes[(size_t)row * 7 + j];
    const float r_row = 0.5f * sqrtf(rb[3] * rb[3] + rb[4] * rb[4]) + 0.1f;
    const bool may_skip = ROTATED && thresh >= 0.f;
    unsigned long long bits = 0;
    const int start = (rt == ct) ? threadIdx.x + 1 : 0;
    for (int i = start; i < col_size; ++i) {
        const float *cbx = cbox + i * 7;
        if (may_skip) {
            float ddx = rb[0] - cbx[0], ddy = rb[1] - cbx[1];
            float rr = r_row + 0.5f * sqrtf(cbx[3] * cbx[3] + cbx[4] * cbx[4]);
            if (ddx * ddx + ddy * ddy > rr * rr) continue;  // provably disjoint: IoU is exactly 0
        }
        float v = ROTATED ? rot_iou(rb, cbx) : axis_iou(rb, cbx);
        if (v > thresh) bits |= 1ULL << i;
    }
    mask[(size_t)row * cb + ct] = bits;
}

// ---- on-device greedy scan --------------------------------------------------------------
// mask: n x cb words; keep[] receives kept row ids in order; n_keep the count.
constexpr int SCAN_T = 256;
__global__ void __launch_bounds__(SCAN_T) nms_scan_kernel(const unsigned long long *__restrict__ mask, int n,
                                                          long long *__restrict__ keep, int *__restrict__ n_keep)
{
    const int cb = (n + 63) >> 6;
    extern __shared__ unsigned long long sm[];
    unsigned long long *remv = sm;            // cb words
    unsigned long long *diag = sm + cb;       // 64 words
    __shared__ unsigned long long kept_s;
    __shared__ int total_s;
    for (int j = threadIdx.x; j < cb; j += SCAN_T) remv[j] = 0ull;
    if (threadIdx.x == 0) total_s = 0;
    __syncthreads();
    for (int c = 0; c < cb; ++c) {
        const int base = c * 64, sz = min(64, n - base);
        if ((int)threadIdx.x < 64) diag[threadIdx.x] = (int)threadIdx.x < sz ? mask[(size_t)(base + threadIdx.x) * cb + c] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long rem = remv[c], kept = 0ull;
            for (int b = 0; b < sz; ++b)
                if (!((rem >> b) & 1ull)) { kept |= 1ull << b; rem |= diag[b]; }
            kept_s = kept;
        }
        __syncthreads();
        const unsigned long long kept = kept_s;
        const int tot = total_s;
        if ((int)threadIdx.x < sz && ((kept >> threadIdx.x) & 1ull))
            keep[tot + __popcll(kept & ((1ull << threadIdx.x) - 1ull))] = base + threadIdx.x;
        // fold the survivors' rows into remv for the chunks to the right
        for (int j = c + 1 + (int)threadIdx.x; j < cb; j += SCAN_T) {
            unsigned long long acc = 0ull, k = kept;
            while (k) {
                int b = __ffsll((long long)k) - 1;
                k &= k - 1;
                acc |= mask[(size_t)(base + b) * cb + j];
            }
            remv[j] |= acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) total_s = tot + __popcll(kept);
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_keep = total_s;
}

int32_t launch_mask(const float *boxes, int32_t n, float thresh, bool rotated, unsigned long long *mask, cudaStream_t stream)
{
    const int cb = (n + 63) / 64;
    const int tiles = cb * (cb + 1) / 2;
    CPD_CUDA(cudaMemsetAsync(mask, 0, sizeof(unsigned long long) * (size_t)n * cb, stream));
    if (rotated) nms_mask_kernel<true><<<tiles, 64, 0, stream>>>(boxes, n, thresh, mask);
    else nms_mask_kernel<false><<<tiles, 64, 0, stream>>>(boxes, n, thresh, mask);
    count_launch();
    return launch_status("cpd_nms_mask");
}

int32_t nms_common(const float *boxes, int32_t n, float thresh, bool rotated, int64_t *keep, int32_t *n_keep,
                   void *ws, size_t ws_bytes, cudaStream_t stream)
{
    CPD_REQUIRE(n >= 0 && n_keep, CPD_ERR_BAD_ARG, "cpd_nms: bad argument");
    if (n == 0) { CPD_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), stream)); return CPD_OK; }
    CPD_REQUIRE(boxes && keep, CPD_ERR_BAD_ARG, "cpd_nms: null argument");
    CPD_REQUIRE(n <= 65536, CPD_ERR_UNSUPPORTED, "cpd_nms: more than 65536 boxes");
    size_t need = cpd_nms_workspace_bytes(n);
    CPD_REQUIRE(ws && ws_bytes >= need, CPD_ERR_WORKSPACE, "cpd_nms: workspace %zu < %zu", ws_bytes, need);
    unsigned long long *mask = (unsigned long long *)ws;
    int32_t st = launch_mask(boxes, n, thresh, rotated, mask, stream);
    if (st) return st;
    const int cb = (n + 63) / 64;
    nms_scan_kernel<<<1, SCAN_T, sizeof(unsigned long long) * (cb + 64), stream>>>(mask, n, (long long *)keep, n_keep);
    count_launch();
    return launch_status("cpd_nms");
}

}  // namespace
}  // namespace cpd

using namespace cpd;

extern "C" int32_t cpd_overlap_bev(const float *a, int32_t na, const float *b, int32_t nb, float *out, cpd_stream_t stream)
{
    CPD_REQUIRE(na >= 0 && nb >= 0, CPD_ERR_BAD_ARG, "cpd_overlap_bev: negative size");
    if (na == 0 || nb == 0) return CPD_OK;
    CPD_REQUIRE(a && b && out, CPD_ERR_BAD_ARG, "cpd_overlap_bev: null argument");
    dim3 grid((nb + 15) / 16, (na + 15) / 16), block(16, 16);
    pair_matrix_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(a, na, b, nb, out);
    count_launch();
    return launch_status("cpd_overlap_bev");
}

extern "C" int32_t cpd_iou_bev(const float *a, int32_t na, const float *b, int32_t nb, float *out, cpd_stream_t stream)
{
    CPD_REQUIRE(na >= 0 && nb >= 0, CPD_ERR_BAD_ARG, "cpd_iou_bev: negative size");
    if (na == 0 || nb == 0) return CPD_OK;
    CPD_REQUIRE(a && b && out, CPD_ERR_BAD_ARG, "cpd_iou_bev: null argument");
    dim3 grid((nb + 15) / 16, (na + 15) / 16), block(16, 16);
    pair_matrix_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(a, na, b, nb, out);
    count_launch();
    return launch_status("cpd_iou_bev");
}

extern "C" size_t cpd_nms_workspace_bytes(int32_t n)
{
    if (n <= 0) return 0;
    return align_up(sizeof(unsigned long long) * (size_t)n * ((n + 63) / 64), 256);
}

extern "C" int32_t cpd_nms_rotated(const float *boxes, int32_t n, float thresh, int64_t *keep, int32_t *n_keep,
                                   void *ws, size_t ws_bytes, cpd_stream_t stream)
{
    return nms_common(boxes, n, thresh, true, keep, n_keep, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int32_t cpd_nms_normal(const float *boxes, int32_t n, float thresh, int64_t *keep, int32_t *n_keep,
                                  void *ws, size_t ws_bytes, cpd_stream_t stream)
{
    return nms_common(boxes, n, thresh, false, keep, n_keep, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int32_t cpd_nms_mask(const float *boxes, int32_t n, float thresh, int32_t rotated, uint64_t *mask,
                                cpd_stream_t stream)
{
    CPD_REQUIRE(n >= 0, CPD_ERR_BAD_ARG, "cpd_nms_mask: negative size");
    if (n == 0) return CPD_OK;
    CPD_REQUIRE(boxes && mask, CPD_ERR_BAD_ARG, "cpd_nms_mask: null argument");
    return launch_mask(boxes, n, thresh, rotated != 0, (unsigned long long *)mask, (cudaStream_t)stream);
}
