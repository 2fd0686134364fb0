/*
 * cpd_oracle.c -- CPU restatement of the CPD detection hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cpd_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / the CPU arm.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - rotated IoU / NMS: PINNED.  Restates cpd/ops/iou3d_nms/src/iou3d_cpu.cpp
 *     and iou3d_nms.cpp; tests/test_oracle_cpu.py checks it bit-for-bit against
 *     the reference's own iou3d_cpu.cpp compiled into oracle/_ref (when built)
 *     and against tests/golden/iou_ref_*.npz generated from it.
 *   - voxelizer and sparse convolution: PARITY UNPINNED by the reference.  The
 *     arithmetic lives in the un-vendored wheel spconv-cu111==2.1.22 (+cumm),
 *     absent from /root/reference and from this image; the reference ships no
 *     tests or golden vectors.  The restatement follows the published spconv
 *     semantics anchored on the reference's call sites,
 *     and is cross-checked against torch.nn.functional.conv3d on densified
 *     inputs (tests/test_oracle_cpu.py).
 *
 * Build: `make -C oracle` -> oracle/liboracle.so  (gcc -O2 -fopenmp, no -ffast-math,
 * -ffp-contract=off so results do not depend on the host's FMA support).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* A.1  Point2VoxelCPU3d.point_to_voxel                                      */
/*   call site: cpd/datasets/processor/data_processor.py:35-41,53-58         */
/*   semantics: spconv 2.1.22 Point2VoxelCPU:       */
/*   sequential scan, first-come voxel ids, first max_pts points per voxel,  */
/*   coords stored reversed (z,y,x), dense lookup grid.                      */
/* ------------------------------------------------------------------------ */
ORACLE_API int64_t cpd_oracle_voxelize(const float *pts, int64_t n, int c,
                                       const float *range6, const float *vsize3,
                                       int max_pts, int64_t max_voxels,
                                       float *voxels, int32_t *coords_zyx,
                                       int32_t *num_per_voxel)
{
    int64_t grid[3];
    for (int j = 0; j < 3; ++j)
        grid[j] = (int64_t)roundf((range6[3 + j] - range6[j]) / vsize3[j]);
    const int64_t ncell = grid[0] * grid[1] * grid[2];
    int32_t *lut = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    if (!lut) return -1;
    memset(lut, 0xff, sizeof(int32_t) * (size_t)ncell);   /* -1 == empty */
    int64_t nvox = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pts + i * c;
        int32_t cell[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            float q = floorf((p[j] - range6[j]) / vsize3[j]);
            if (!(q >= 0.0f) || q >= (float)grid[j]) { ok = 0; break; }
            cell[2 - j] = (int32_t)q;                       /* reversed: z,y,x */
        }
        if (!ok) continue;
        int64_t lin = ((int64_t)cell[0] * grid[1] + cell[1]) * grid[0] + cell[2];
        int32_t vid = lut[lin];
        if (vid < 0) {
            if (nvox >= max_voxels) continue;
            vid = (int32_t)nvox++;
            lut[lin] = vid;
            coords_zyx[3 * vid + 0] = cell[0];
            coords_zyx[3 * vid + 1] = cell[1];
            coords_zyx[3 * vid + 2] = cell[2];
            num_per_voxel[vid] = 0;
            memset(voxels + (size_t)vid * max_pts * c, 0, sizeof(float) * max_pts * c);
        }
        int32_t k = num_per_voxel[vid];
        if (k < max_pts) {
            memcpy(voxels + ((size_t)vid * max_pts + k) * c, p, sizeof(float) * c);
            num_per_voxel[vid] = k + 1;
        }
    }
    free(lut);
    return nvox;
}

/* MeanVFE: cpd/models/backbones_3d/vfe/mean_vfe.py:41-44.                    */
ORACLE_API void cpd_oracle_mean_vfe(const float *voxels, const int32_t *num,
                                    int64_t m, int max_pts, int c, float *out)
{
    for (int64_t v = 0; v < m; ++v) {
        float norm = (float)(num[v] < 1 ? 1 : num[v]);
        for (int f = 0; f < c; ++f) {
            float s = 0.f;
            for (int k = 0; k < max_pts; ++k) s += voxels[((size_t)v * max_pts + k) * c + f];
            out[v * c + f] = s / norm;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* coordinate hash used by the rulebook builders (oracle-private)            */
/* ------------------------------------------------------------------------ */
typedef struct { int64_t *keys; int32_t *vals; uint64_t mask; } ohash_t;

static int ohash_init(ohash_t *h, int64_t n)
{
    uint64_t cap = 64;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    h->keys = (int64_t *)malloc(sizeof(int64_t) * cap);
    h->vals = (int32_t *)malloc(sizeof(int32_t) * cap);
    if (!h->keys || !h->vals) return -1;
    for (uint64_t i = 0; i < cap; ++i) h->keys[i] = -1;
    h->mask = cap - 1;
    return 0;
}
static void ohash_free(ohash_t *h) { free(h->keys); free(h->vals); }
static inline uint64_t ohash_mix(int64_t k)
{
    uint64_t x = (uint64_t)k * 0x9E3779B97F4A7C15ull;
    return x ^ (x >> 29);
}
/* returns existing value, or inserts val and returns val */
static inline int32_t ohash_put(ohash_t *h, int64_t key, int32_t val)
{
    uint64_t s = ohash_mix(key) & h->mask;
    for (;;) {
        if (h->keys[s] == key) return h->vals[s];
        if (h->keys[s] == -1) { h->keys[s] = key; h->vals[s] = val; return val; }
        s = (s + 1) & h->mask;
    }
}
static inline int32_t ohash_get(const ohash_t *h, int64_t key)
{
    uint64_t s = ohash_mix(key) & h->mask;
    for (;;) {
        if (h->keys[s] == key) return h->vals[s];
        if (h->keys[s] == -1) return -1;
        s = (s + 1) & h->mask;
    }
}
static inline int64_t lin_key(int32_t b, int32_t z, int32_t y, int32_t x, const int32_t *shape)
{
    return (((int64_t)b * shape[0] + z) * shape[1] + y) * shape[2] + x;
}

/* ------------------------------------------------------------------------ */
/* A.3  SubMConv3d rulebook ("indice pairs", spconv Native layout)           */
/*   call sites: spconv_backbone.py:17,108-115,154,415                       */
/*   pair_in/pair_out: (K, M) int32, -1 padded; pair_cnt: (K,)               */
/*   tap k = (kz*KH + ky)*KW + kx; input site = output site + k - K/2.        */
/* ------------------------------------------------------------------------ */
ORACLE_API int cpd_oracle_rulebook_subm(const int32_t *coords, int64_t m,
                                        const int32_t *shape3, const int32_t *ksize3,
                                        int32_t *pair_in, int32_t *pair_out,
                                        int32_t *pair_cnt)
{
    ohash_t h;
    if (ohash_init(&h, m)) return -1;
    for (int64_t i = 0; i < m; ++i) {
        const int32_t *q = coords + 4 * i;
        ohash_put(&h, lin_key(q[0], q[1], q[2], q[3], shape3), (int32_t)i);
    }
    const int K = ksize3[0] * ksize3[1] * ksize3[2];
    for (int k = 0; k < K; ++k) pair_cnt[k] = 0;
    for (int64_t i = 0; i < (int64_t)K * m; ++i) { pair_in[i] = -1; pair_out[i] = -1; }
    for (int64_t o = 0; o < m; ++o) {
        const int32_t *q = coords + 4 * o;
        int k = 0;
        for (int kz = 0; kz < ksize3[0]; ++kz)
            for (int ky = 0; ky < ksize3[1]; ++ky)
                for (int kx = 0; kx < ksize3[2]; ++kx, ++k) {
                    int32_t z = q[1] + kz - ksize3[0] / 2;
                    int32_t y = q[2] + ky - ksize3[1] / 2;
                    int32_t x = q[3] + kx - ksize3[2] / 2;
                    if (z < 0 || y < 0 || x < 0 || z >= shape3[0] || y >= shape3[1] || x >= shape3[2])
                        continue;
                    int32_t i = ohash_get(&h, lin_key(q[0], z, y, x, shape3));
                    if (i < 0) continue;
                    int32_t c = pair_cnt[k]++;
                    pair_in[(int64_t)k * m + c] = i;
                    pair_out[(int64_t)k * m + c] = (int32_t)o;
                }
    }
    ohash_free(&h);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* A.4  SparseConv3d rulebook.  call sites: spconv_backbone.py:20-21,189-190, */
/*   451-452.  out_shape = floor((D + 2p - k)/s) + 1.  Input i at x feeds     */
/*   output o = (x + p - k)/s through tap k iff divisible and in range.       */
/*   Output rows are canonicalised to ascending linear key.    */
/*   out_coords must hold K*m rows in the worst case; returns m_out.          */
/* ------------------------------------------------------------------------ */
typedef struct { int64_t key; int32_t idx; } keyidx_t;
static int keyidx_cmp(const void *a, const void *b)
{
    int64_t x = ((const keyidx_t *)a)->key, y = ((const keyidx_t *)b)->key;
    return (x > y) - (x < y);
}

ORACLE_API int64_t cpd_oracle_rulebook_strided(const int32_t *coords, int64_t m,
                                               const int32_t *shape3, const int32_t *ksize3,
                                               const int32_t *stride3, const int32_t *pad3,
                                               int32_t *out_shape3, int32_t *out_coords,
                                               int32_t *pair_in, int32_t *pair_out,
                                               int32_t *pair_cnt)
{
    for (int j = 0; j < 3; ++j)
        out_shape3[j] = (shape3[j] + 2 * pad3[j] - ksize3[j]) / stride3[j] + 1;
    const int K = ksize3[0] * ksize3[1] * ksize3[2];
    ohash_t h;
    if (ohash_init(&h, m * (K < 8 ? K : 8))) return -1;
    for (int k = 0; k < K; ++k) pair_cnt[k] = 0;
    for (int64_t i = 0; i < (int64_t)K * m; ++i) { pair_in[i] = -1; pair_out[i] = -1; }
    int64_t m_out = 0, cap = (int64_t)(h.mask + 1) / 2;
    for (int64_t i = 0; i < m; ++i) {
        const int32_t *q = coords + 4 * i;
        int k = 0;
        for (int kz = 0; kz < ksize3[0]; ++kz)
            for (int ky = 0; ky < ksize3[1]; ++ky)
                for (int kx = 0; kx < ksize3[2]; ++kx, ++k) {
                    int32_t nz = q[1] + pad3[0] - kz, ny = q[2] + pad3[1] - ky, nx = q[3] + pad3[2] - kx;
                    if (nz < 0 || ny < 0 || nx < 0) continue;
                    if (nz % stride3[0] || ny % stride3[1] || nx % stride3[2]) continue;
                    int32_t z = nz / stride3[0], y = ny / stride3[1], x = nx / stride3[2];
                    if (z >= out_shape3[0] || y >= out_shape3[1] || x >= out_shape3[2]) continue;
                    if (m_out >= cap) { ohash_free(&h); return -2; }
                    int32_t o = ohash_put(&h, lin_key(q[0], z, y, x, out_shape3), (int32_t)m_out);
                    if (o == (int32_t)m_out) {
                        int32_t *w = out_coords + 4 * m_out;
                        w[0] = q[0]; w[1] = z; w[2] = y; w[3] = x;
                        ++m_out;
                    }
                    int32_t c = pair_cnt[k]++;
                    pair_in[(int64_t)k * m + c] = (int32_t)i;
                    pair_out[(int64_t)k * m + c] = o;
                }
    }
    ohash_free(&h);
    /* canonical order: ascending linear key */
    keyidx_t *ord = (keyidx_t *)malloc(sizeof(keyidx_t) * (size_t)(m_out ? m_out : 1));
    int32_t *remap = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m_out ? m_out : 1));
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(m_out ? m_out : 1));
    for (int64_t o = 0; o < m_out; ++o) {
        const int32_t *w = out_coords + 4 * o;
        ord[o].key = lin_key(w[0], w[1], w[2], w[3], out_shape3);
        ord[o].idx = (int32_t)o;
    }
    qsort(ord, (size_t)m_out, sizeof(keyidx_t), keyidx_cmp);
    for (int64_t r = 0; r < m_out; ++r) {
        remap[ord[r].idx] = (int32_t)r;
        memcpy(tmp + 4 * r, out_coords + 4 * (int64_t)ord[r].idx, 4 * sizeof(int32_t));
    }
    memcpy(out_coords, tmp, sizeof(int32_t) * 4 * (size_t)m_out);
    for (int k = 0; k < K; ++k)
        for (int32_t c = 0; c < pair_cnt[k]; ++c)
            pair_out[(int64_t)k * m + c] = remap[pair_out[(int64_t)k * m + c]];
    free(ord); free(remap); free(tmp);
    return m_out;
}

/* ------------------------------------------------------------------------ */
/* A.3-A.5  "Native" gather-GEMM-scatter forward / backward over a rulebook   */
/*   weight layout (Cout, K, Cin): spconv 2.x, detector3d_template.py:399-408 */
/*   y[o] += W[:,k,:] x[i] for every pair (i,o) of tap k; + bias.             */
/*   Within one tap each output appears at most once => taps are sequential,  */
/*   pairs inside a tap run in parallel without conflicts.                    */
/* ------------------------------------------------------------------------ */
ORACLE_API void cpd_oracle_spconv_fwd(const float *x, int64_t m_in, int cin,
                                      const float *w, const float *bias, int cout, int K,
                                      const int32_t *pair_in, const int32_t *pair_out,
                                      const int32_t *pair_cnt, int64_t pair_stride,
                                      float *y, int64_t m_out)
{
    (void)m_in;
    /* per-tap (Cin, Cout) filter matrices, as the Native algorithm's GEMM operand */
    float *wt = (float *)malloc(sizeof(float) * (size_t)K * cin * cout);
    for (int co = 0; co < cout; ++co)
        for (int k = 0; k < K; ++k)
            for (int ci = 0; ci < cin; ++ci)
                wt[((int64_t)k * cin + ci) * cout + co] = w[((int64_t)co * K + k) * cin + ci];
    for (int64_t o = 0; o < m_out; ++o)
        for (int co = 0; co < cout; ++co) y[o * cout + co] = bias ? bias[co] : 0.f;
    for (int k = 0; k < K; ++k) {
        const int32_t *pi = pair_in + (int64_t)k * pair_stride, *po = pair_out + (int64_t)k * pair_stride;
        const float *wk = wt + (int64_t)k * cin * cout;
#pragma omp parallel for schedule(static)
        for (int32_t p = 0; p < pair_cnt[k]; ++p) {
            const float *xi = x + (int64_t)pi[p] * cin;
            float *yo = y + (int64_t)po[p] * cout;
            for (int ci = 0; ci < cin; ++ci) {
                const float xv = xi[ci];
                const float *wr = wk + (int64_t)ci * cout;
                for (int co = 0; co < cout; ++co) yo[co] += wr[co] * xv;
            }
        }
    }
    free(wt);
}

/* dX[i] += W_k^T dY[o];  dW[:,k,:] += dY[o] x[i]^T;  dbias = sum_o dY[o]   (A.5) */
ORACLE_API void cpd_oracle_spconv_bwd(const float *x, int64_t m_in, int cin,
                                      const float *w, int cout, int K,
                                      const int32_t *pair_in, const int32_t *pair_out,
                                      const int32_t *pair_cnt, int64_t pair_stride,
                                      const float *dy, int64_t m_out,
                                      float *dx, float *dw, float *dbias)
{
    if (dx) memset(dx, 0, sizeof(float) * (size_t)m_in * cin);
    if (dw) memset(dw, 0, sizeof(float) * (size_t)cout * K * cin);
    if (dbias) {
        for (int co = 0; co < cout; ++co) {
            double s = 0.0;
            for (int64_t o = 0; o < m_out; ++o) s += dy[o * cout + co];
            dbias[co] = (float)s;
        }
    }
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    const size_t tile = (size_t)cout * cin;
    float *wk = (float *)malloc(sizeof(float) * tile);                 /* W_k as (cout, cin), contiguous */
    double *acc = dw ? (double *)malloc(sizeof(double) * tile * (size_t)nthreads) : NULL;
    for (int k = 0; k < K; ++k) {
        const int32_t *pi = pair_in + (int64_t)k * pair_stride, *po = pair_out + (int64_t)k * pair_stride;
        const int32_t np = pair_cnt[k];
        if (np == 0) continue;
        for (int co = 0; co < cout; ++co)
            memcpy(wk + (size_t)co * cin, w + ((int64_t)co * K + k) * cin, sizeof(float) * cin);
        if (dx) {
            /* within one tap every input row appears at most once: pairs run in parallel */
#pragma omp parallel for schedule(static)
            for (int32_t p = 0; p < np; ++p) {
                float *dxi = dx + (int64_t)pi[p] * cin;
                const float *dyo = dy + (int64_t)po[p] * cout;
                for (int co = 0; co < cout; ++co) {
                    const float g = dyo[co];
                    const float *row = wk + (size_t)co * cin;
                    for (int ci = 0; ci < cin; ++ci) dxi[ci] += row[ci] * g;
                }
            }
        }
        if (dw) {
            /* per-thread double accumulators over a slice of the pairs, then a tree-free reduce */
#pragma omp parallel
            {
                int t = 0, nt = 1;
#ifdef _OPENMP
                t = omp_get_thread_num(); nt = omp_get_num_threads();
#endif
                double *a = acc + tile * (size_t)t;
                for (size_t q = 0; q < tile; ++q) a[q] = 0.0;
                const int32_t lo = (int32_t)((int64_t)np * t / nt), hi = (int32_t)((int64_t)np * (t + 1) / nt);
                for (int32_t p = lo; p < hi; ++p) {
                    const float *xi = x + (int64_t)pi[p] * cin;
                    const float *dyo = dy + (int64_t)po[p] * cout;
                    for (int co = 0; co < cout; ++co) {
                        const double g = (double)dyo[co];
                        double *row = a + (size_t)co * cin;
                        for (int ci = 0; ci < cin; ++ci) row[ci] += g * (double)xi[ci];
                    }
                }
#pragma omp barrier
#pragma omp for schedule(static)
                for (int co = 0; co < cout; ++co) {
                    float *dwk = dw + ((int64_t)co * K + k) * cin;
                    for (int ci = 0; ci < cin; ++ci) {
                        double sum = 0.0;
                        for (int u = 0; u < nt; ++u) sum += acc[tile * (size_t)u + (size_t)co * cin + ci];
                        dwk[ci] = (float)sum;
                    }
                }
            }
        }
    }
    free(wk);
    free(acc);
}

/* SparseConvTensor.dense(): height_compression.py:136-138, Appendix A.2.     */
/* out (B, C, D, H, W) zero-filled then scattered.                            */
ORACLE_API void cpd_oracle_dense(const float *feat, const int32_t *coords, int64_t m, int c,
                                 int batch, const int32_t *shape3, float *out)
{
    const int64_t vol = (int64_t)shape3[0] * shape3[1] * shape3[2];
    memset(out, 0, sizeof(float) * (size_t)batch * c * vol);
    for (int64_t v = 0; v < m; ++v) {
        const int32_t *q = coords + 4 * v;
        int64_t sp = ((int64_t)q[1] * shape3[1] + q[2]) * shape3[2] + q[3];
        for (int f = 0; f < c; ++f)
            out[((int64_t)q[0] * c + f) * vol + sp] = feat[v * c + f];
    }
}

/* ------------------------------------------------------------------------ */
/* Rotated BEV IoU: restates cpd/ops/iou3d_nms/src/iou3d_cpu.cpp:59-230       */
/* (identical algorithm to iou3d_nms_kernel.cu:35-234).  Every float/double   */
/* promotion of the reference is kept: Point(double,double) ctor narrowing,   */
/* `fabs(area) / 2.0` in double, float libm calls (C++ overloads of cos/sin/  */
/* atan2/fabs on float arguments resolve to the float versions).              */
/* ------------------------------------------------------------------------ */
typedef struct { float x, y; } pt_t;
static const float IOU_EPS = 1e-8f;

static inline float cross2(pt_t a, pt_t b) { return a.x * b.y - a.y * b.x; }
static inline float cross3(pt_t p1, pt_t p2, pt_t p0)
{
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static inline float fmin2(float a, float b) { return a > b ? b : a; }   /* iou3d_cpu.cpp:31-37 */
static inline float fmax2(float a, float b) { return a > b ? a : b; }

static int seg_bbox_overlap(pt_t p1, pt_t p2, pt_t q1, pt_t q2)
{
    return fmin2(p1.x, p2.x) <= fmax2(q1.x, q2.x) && fmin2(q1.x, q2.x) <= fmax2(p1.x, p2.x) &&
           fmin2(p1.y, p2.y) <= fmax2(q1.y, q2.y) && fmin2(q1.y, q2.y) <= fmax2(p1.y, p2.y);
}

static int point_in_box(const float *box, pt_t p)
{
    const float MARGIN = 1e-2f;
    float cx = box[0], cy = box[1];
    float ac = cosf(-box[6]), as = sinf(-box[6]);
    float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
    float ry = (p.x - cx) * as + (p.y - cy) * ac;
    return (fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN);
}

static int seg_intersection(pt_t p1, pt_t p0, pt_t q1, pt_t q0, pt_t *ans)
{
    if (!seg_bbox_overlap(p0, p1, q0, q1)) return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > IOU_EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static inline void rot_about(pt_t c, float ac, float as, pt_t *p)
{
    float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
    float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
    p->x = nx; p->y = ny;
}

static float rot_overlap(const float *A, const float *B)
{
    float a_ang = A[6], b_ang = B[6];
    float ahx = A[3] / 2, bhx = B[3] / 2, ahy = A[4] / 2, bhy = B[4] / 2;
    float ax1 = A[0] - ahx, ay1 = A[1] - ahy, ax2 = A[0] + ahx, ay2 = A[1] + ahy;
    float bx1 = B[0] - bhx, by1 = B[1] - bhy, bx2 = B[0] + bhx, by2 = B[1] + bhy;
    pt_t ca = { A[0], A[1] }, cb = { B[0], B[1] };
    pt_t qa[5] = { { ax1, ay1 }, { ax2, ay1 }, { ax2, ay2 }, { ax1, ay2 } };
    pt_t qb[5] = { { bx1, by1 }, { bx2, by1 }, { bx2, by2 }, { bx1, by2 } };
    float aco = cosf(a_ang), asi = sinf(a_ang), bco = cosf(b_ang), bsi = sinf(b_ang);
    for (int k = 0; k < 4; ++k) { rot_about(ca, aco, asi, &qa[k]); rot_about(cb, bco, bsi, &qb[k]); }
    qa[4] = qa[0]; qb[4] = qb[0];

    pt_t poly[16], ctr = { 0.f, 0.f };
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersection(qa[i + 1], qa[i], qb[j + 1], qb[j], &poly[cnt])) {
                ctr.x = ctr.x + poly[cnt].x; ctr.y = ctr.y + poly[cnt].y; ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (point_in_box(A, qb[k])) { ctr.x = ctr.x + qb[k].x; ctr.y = ctr.y + qb[k].y; poly[cnt++] = qb[k]; }
        if (point_in_box(B, qa[k])) { ctr.x = ctr.x + qa[k].x; ctr.y = ctr.y + qa[k].y; poly[cnt++] = qa[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;                          /* cnt==0 -> NaN, as upstream */
    for (int j = 0; j < cnt - 1; ++j)                    /* bubble sort on atan2 keys  */
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(poly[i].y - ctr.y, poly[i].x - ctr.x) > atan2f(poly[i + 1].y - ctr.y, poly[i + 1].x - ctr.x)) {
                pt_t t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        pt_t u = { poly[k].x - poly[0].x, poly[k].y - poly[0].y };
        pt_t v = { poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y };
        area += cross2(u, v);
    }
    return (float)(fabs((double)area) / 2.0);
}

static float rot_iou(const float *A, const float *B)
{
    float sa = A[3] * A[4], sb = B[3] * B[4];
    float so = rot_overlap(A, B);
    return so / fmaxf(sa + sb - so, IOU_EPS);
}

/* iou3d_nms_kernel.cu:314-325 */
static float axis_iou(const float *a, const float *b)
{
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    float inter = width * height;
    float Sa = a[3] * a[4], Sb = b[3] * b[4];
    return inter / fmaxf(Sa + Sb - inter, IOU_EPS);
}

/* boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252) / boxes_iou_bev_gpu / boxes_overlap_bev_gpu */
ORACLE_API void cpd_oracle_iou_bev(const float *a, int na, const float *b, int nb, float *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(int64_t)i * nb + j] = rot_iou(a + 7 * i, b + 7 * j);
}
ORACLE_API void cpd_oracle_overlap_bev(const float *a, int na, const float *b, int nb, float *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(int64_t)i * nb + j] = rot_overlap(a + 7 * i, b + 7 * j);
}

/* nms_gpu / nms_normal_gpu: mask build (iou3d_nms_kernel.cu:267-311, 328-372: row box vs
 * col box, `start = tid + 1` on the diagonal tile) + sequential greedy scan
 * (iou3d_nms.cpp:115-132).  Boxes must already be sorted by descending score.
 * mask_out (optional): n x ceil(n/64) uint64, lower-triangle tiles included like upstream. */
ORACLE_API int64_t cpd_oracle_nms(const float *boxes, int n, float thresh, int rotated,
                                  int64_t *keep, uint64_t *mask_out)
{
    const int cb = (n + 63) / 64;
    uint64_t *mask = mask_out ? mask_out : (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1) * (cb ? cb : 1));
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < cb; ++c) {
            uint64_t t = 0;
            int jend = n - c * 64 < 64 ? n - c * 64 : 64;
            int start = (i / 64 == c) ? (i % 64) + 1 : 0;
            for (int j = start; j < jend; ++j) {
                const float *bj = boxes + 7 * (c * 64 + j);
                float v = rotated ? rot_iou(boxes + 7 * i, bj) : axis_iou(boxes + 7 * i, bj);
                if (v > thresh) t |= 1ull << j;
            }
            mask[(int64_t)i * cb + c] = t;
        }
    uint64_t *remv = (uint64_t *)calloc((size_t)(cb ? cb : 1), sizeof(uint64_t));
    int64_t nk = 0;
    for (int i = 0; i < n; ++i) {
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ull << ib))) {
            keep[nk++] = i;
            const uint64_t *p = mask + (int64_t)i * cb;
            for (int j = nb; j < cb; ++j) remv[j] |= p[j];
        }
    }
    free(remv);
    if (!mask_out) free(mask);
    return nk;
}

ORACLE_API int cpd_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORACLE_API void cpd_oracle_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
