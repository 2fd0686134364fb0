// Deterministic hash-and-scatter voxelizer for sm_100a.
//
// Replaces spconv.utils.Point2VoxelCPU3d.point_to_voxel (call site
// cpd/datasets/processor/data_processor.py:35-41,53-58), collate_batch's batch column
// (cpd/datasets/dataset.py:262-266) and MeanVFE (cpd/models/backbones_3d/vfe/mean_vfe.py:41-50).
//
// The sequential reference is order dependent: voxel id = rank of the cell
// by its first point, each voxel keeps its first max_pts points in input order.  Here:
//   1. vox_insert   : one thread per point; cell key -> open-addressing table
//                     (one CAS per distinct key per warp via __match_any_sync), then a
//                     lock-free sorted insertion network of atomicMin's keeps the max_pts
//                     SMALLEST point indices of every cell, in order, whatever the
//                     interleaving (state+carry is a conserved multiset).
//   2. vox_count    : a point is a "first" iff it sits in slot 0 of its cell; per-block counts.
//   3. vox_scan     : per-frame exclusive scan of the block counts, max_voxels clamp,
//                     frame-major row bases.
//   4. vox_assign   : rank of each first point = voxel row; writes coords / num.
//   5. vox_fill     : one thread per (row, feature): gathers the <= max_pts points, writes
//                     the zero-padded voxel block and the mean.
// HBM traffic: N*C*4 read + M*(max_pts*C*4 + 16 + 4 + C*4) written, plus the table.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace cpd {
namespace {

constexpr int VOX_BLOCK = 1024;
constexpr int32_t SLOT_EMPTY = 0x7f7f7f7f;  // cudaMemset(0x7f)

struct VoxParams {
    float mn[3], vs[3];
    int grid[3];  // x, y, z
    int batch, c, max_pts;
    long long max_voxels, cap_rows;
    long long off[CPD_MAX_BATCH + 1];
    int blk_off[CPD_MAX_BATCH + 1];
    uint32_t cap;  // table capacity
};

__device__ __forceinline__ void block_to_points(const VoxParams &p, int &f, long long &i, bool &in)
{
    f = 0;
    while (f + 1 < p.batch && (int)blockIdx.x >= p.blk_off[f + 1]) ++f;
    i = p.off[f] + (long long)(blockIdx.x - p.blk_off[f]) * VOX_BLOCK + threadIdx.x;
    in = i < p.off[f + 1];
}

__device__ __forceinline__ uint32_t table_home(uint32_t key, uint32_t cap)
{
    return (uint32_t)(((uint64_t)hash_mix(key) * cap) >> 32);
}

__global__ void __launch_bounds__(VOX_BLOCK) vox_insert(const float *__restrict__ pts, VoxParams p,
                                                         uint32_t *keys, int32_t *slots, int32_t *slot_of_pt)
{
    int f; long long i; bool in;
    block_to_points(p, f, i, in);
    uint32_t key = HASH_EMPTY;
    if (in) {
        const float *q = pts + i * p.c;
        int cell[3];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // fp32 IEEE subtract + divide + floor, exactly as the reference computes the cell
            float t = floorf(__fdiv_rn(__fsub_rn(__ldg(q + j), p.mn[j]), p.vs[j]));
            ok = ok && (t >= 0.0f) && (t < (float)p.grid[j]);
            cell[j] = (int)t;
        }
        if (ok) key = (uint32_t)(((f * p.grid[2] + cell[2]) * (long long)p.grid[1] + cell[1]) * p.grid[0] + cell[0]);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    int32_t slot = -1;
    if (key != HASH_EMPTY) {
        const int leader = __ffs(peers) - 1;
        if ((int)(threadIdx.x & 31) == leader) {
            uint32_t s = table_home(key, p.cap);
            for (;;) {
                uint32_t prev = atomicCAS(keys + s, HASH_EMPTY, key);
                if (prev == HASH_EMPTY || prev == key) break;
                if (++s == p.cap) s = 0;
            }
            slot = (int32_t)s;
        }
        slot = __shfl_sync(peers, slot, leader);
        int32_t carry = (int32_t)i;
        int32_t *cellslots = slots + (size_t)slot * p.max_pts;
        for (int r = 0; r < p.max_pts; ++r) {
            int32_t old = atomicMin(cellslots + r, carry);
            if (old == SLOT_EMPTY) break;
            carry = max(old, carry);
        }
    }
    if (in) slot_of_pt[i] = slot;
}

__device__ __forceinline__ bool is_first(const VoxParams &p, const int32_t *slots, const int32_t *slot_of_pt,
                                         long long i, bool in, int32_t &slot)
{
    slot = in ? slot_of_pt[i] : -1;
    return slot >= 0 && slots[(size_t)slot * p.max_pts] == (int32_t)i;
}

__global__ void __launch_bounds__(VOX_BLOCK) vox_count(VoxParams p, const int32_t *slots,
                                                        const int32_t *slot_of_pt, int32_t *block_sums)
{
    int f; long long i; bool in; int32_t slot;
    block_to_points(p, f, i, in);
    int cnt = __syncthreads_count(is_first(p, slots, slot_of_pt, i, in, slot));
    if (threadIdx.x == 0) block_sums[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(VOX_BLOCK) vox_scan(VoxParams p, const int32_t *block_sums,
                                                       int32_t *block_prefix, int32_t *frame_base, int32_t *counts)
{
    long long acc = 0;
    for (int f = 0; f < p.batch; ++f) {
        int running = 0;
        for (int b0 = p.blk_off[f]; b0 < p.blk_off[f + 1]; b0 += VOX_BLOCK) {
            int b = b0 + threadIdx.x;
            int v = b < p.blk_off[f + 1] ? block_sums[b] : 0, tot;
            int ex = block_exclusive_scan(v, &tot);
            if (b < p.blk_off[f + 1]) block_prefix[b] = running + ex;
            running += tot;
        }
        long long kept = running < p.max_voxels ? running : p.max_voxels;
        if (threadIdx.x == 0) { frame_base[f] = (int32_t)acc; counts[f] = (int32_t)kept; }
        acc += kept;
    }
    if (threadIdx.x == 0) { frame_base[p.batch] = (int32_t)acc; counts[p.batch] = (int32_t)acc; }
}

__global__ void __launch_bounds__(VOX_BLOCK) vox_assign(VoxParams p, const uint32_t *keys, const int32_t *slots,
                                                         const int32_t *slot_of_pt, const int32_t *block_prefix,
                                                         const int32_t *frame_base, int32_t *coords, int32_t *num,
                                                         int32_t *row_slot)
{
    int f; long long i; bool in; int32_t slot;
    block_to_points(p, f, i, in);
    const bool first = is_first(p, slots, slot_of_pt, i, in, slot);
    int tot;
    const int ex = block_exclusive_scan(first ? 1 : 0, &tot);
    if (!first) return;
    const long long vid = (long long)block_prefix[blockIdx.x] + ex;
    if (vid >= p.max_voxels) return;
    const long long row = frame_base[f] + vid;
    if (row >= p.cap_rows) return;
    uint32_t key = keys[slot];
    int x = key % p.grid[0]; key /= p.grid[0];
    int y = key % p.grid[1]; key /= p.grid[1];
    int z = key % p.grid[2];
    int4 cz = make_int4(f, z, y, x);
    *reinterpret_cast<int4 *>(coords + 4 * row) = cz;
    int n = 0;
    for (int r = 0; r < p.max_pts; ++r) n += slots[(size_t)slot * p.max_pts + r] != SLOT_EMPTY;
    num[row] = n;
    row_slot[row] = slot;
}

__global__ void __launch_bounds__(256) vox_fill(const float *__restrict__ pts, VoxParams p, const int32_t *slots,
                                                const int32_t *row_slot, const int32_t *counts,
                                                float *voxels, float *mean)
{
    long long total = counts[p.batch];
    if (total > p.cap_rows) total = p.cap_rows;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total * p.c) return;
    const long long row = t / p.c;
    const int feat = (int)(t - row * p.c);
    const int32_t *cellslots = slots + (size_t)row_slot[row] * p.max_pts;
    float s = 0.f;
    int n = 0;
    for (int r = 0; r < p.max_pts; ++r) {
        int32_t idx = cellslots[r];
        float v = 0.f;
        if (idx != SLOT_EMPTY) { v = __ldg(pts + (size_t)idx * p.c + feat); ++n; }
        if (voxels) voxels[((size_t)row * p.max_pts + r) * p.c + feat] = v;
        s += v;  // sum over all max_pts slots (zeros included), like voxels.sum(dim=1)
    }
    if (mean) mean[t] = s / (float)(n < 1 ? 1 : n);
}

struct VoxWs {
    uint32_t *keys; int32_t *slots, *slot_of_pt, *block_sums, *block_prefix, *frame_base, *row_slot;
    size_t bytes;
};

VoxWs carve(void *ws, int64_t n, int32_t batch, int32_t max_pts, int64_t cap_rows, uint32_t *cap_out)
{
    const uint64_t cap = (uint64_t)n + (uint64_t)n / 2 + 64;
    const int64_t nblk = div_up(n, VOX_BLOCK) + batch;
    char *b = (char *)ws;
    size_t o = 0;
    VoxWs w;
    auto take = [&](size_t bytes) { char *q = b ? b + o : nullptr; o += align_up(bytes, 256); return q; };
    w.keys = (uint32_t *)take(cap * 4);
    w.slots = (int32_t *)take(cap * (size_t)max_pts * 4);
    w.slot_of_pt = (int32_t *)take((size_t)n * 4);
    w.block_sums = (int32_t *)take((size_t)nblk * 4);
    w.block_prefix = (int32_t *)take((size_t)nblk * 4);
    w.frame_base = (int32_t *)take((size_t)(batch + 1) * 4);
    w.row_slot = (int32_t *)take((size_t)cap_rows * 4);
    w.bytes = o;
    if (cap_out) *cap_out = (uint32_t)cap;
    return w;
}

}  // namespace
}  // namespace cpd

using namespace cpd;

extern "C" size_t cpd_voxelize_workspace_bytes(int64_t n_points, int32_t batch, int32_t max_pts, int64_t cap_rows)
{
    if (n_points < 0 || batch < 1 || max_pts < 1 || cap_rows < 0) return 0;
    return carve(nullptr, n_points, batch, max_pts, cap_rows, nullptr).bytes;
}

extern "C" int32_t cpd_voxelize(const float *points, int64_t n, int32_t c, const int64_t *frame_offsets_host,
                                int32_t batch, const float *range6, const float *vsize3, int32_t max_pts,
                                int64_t max_voxels, int64_t cap_rows, float *voxels, int32_t *coords,
                                int32_t *num, float *mean, int32_t *counts, void *ws, size_t ws_bytes,
                                cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(batch >= 1 && batch <= CPD_MAX_BATCH, CPD_ERR_BAD_ARG, "cpd_voxelize: batch %d outside [1,%d]", batch, CPD_MAX_BATCH);
    CPD_REQUIRE(n >= 0 && n < 0x7f000000ll && c >= 3 && max_pts >= 1 && max_pts <= 64, CPD_ERR_BAD_ARG, "cpd_voxelize: bad n/c/max_pts");
    CPD_REQUIRE(frame_offsets_host && range6 && vsize3 && coords && num && counts, CPD_ERR_BAD_ARG, "cpd_voxelize: null argument");
    CPD_REQUIRE(frame_offsets_host[0] == 0 && frame_offsets_host[batch] == n, CPD_ERR_BAD_ARG, "cpd_voxelize: frame offsets must span [0, n]");
    CPD_REQUIRE(((uintptr_t)coords & 15) == 0, CPD_ERR_MISALIGNED, "cpd_voxelize: coords must be 16-byte aligned");
    VoxParams p;
    long long ncell = 1;
    for (int j = 0; j < 3; ++j) {
        p.mn[j] = range6[j];
        p.vs[j] = vsize3[j];
        p.grid[j] = (int)roundf((range6[3 + j] - range6[j]) / vsize3[j]);
        CPD_REQUIRE(p.grid[j] > 0, CPD_ERR_BAD_ARG, "cpd_voxelize: empty grid");
        ncell *= p.grid[j];
    }
    CPD_REQUIRE(ncell * batch < 0xFFFFFFFFll, CPD_ERR_UNSUPPORTED, "cpd_voxelize: batch*grid exceeds 32-bit cell keys");
    p.batch = batch; p.c = c; p.max_pts = max_pts; p.max_voxels = max_voxels; p.cap_rows = cap_rows;
    int nblk = 0;
    for (int f = 0; f < batch; ++f) {
        CPD_REQUIRE(frame_offsets_host[f + 1] >= frame_offsets_host[f], CPD_ERR_BAD_ARG, "cpd_voxelize: offsets not monotone");
        p.off[f] = frame_offsets_host[f];
        p.blk_off[f] = nblk;
        nblk += (int)div_up(frame_offsets_host[f + 1] - frame_offsets_host[f], VOX_BLOCK);
    }
    p.off[batch] = n; p.blk_off[batch] = nblk;
    for (int f = batch + 1; f <= CPD_MAX_BATCH; ++f) { p.off[f] = n; p.blk_off[f] = nblk; }
    VoxWs w = carve(ws, n, batch, max_pts, cap_rows, &p.cap);
    CPD_REQUIRE(ws && ws_bytes >= w.bytes, CPD_ERR_WORKSPACE, "cpd_voxelize: workspace %zu < %zu", ws_bytes, w.bytes);

    CPD_CUDA(cudaMemsetAsync(w.keys, 0xff, (size_t)p.cap * 4, stream));
    CPD_CUDA(cudaMemsetAsync(w.slots, 0x7f, (size_t)p.cap * max_pts * 4, stream));
    if (nblk > 0) {
        vox_insert<<<nblk, VOX_BLOCK, 0, stream>>>(points, p, w.keys, w.slots, w.slot_of_pt);
        vox_count<<<nblk, VOX_BLOCK, 0, stream>>>(p, w.slots, w.slot_of_pt, w.block_sums);
    }
    vox_scan<<<1, VOX_BLOCK, 0, stream>>>(p, w.block_sums, w.block_prefix, w.frame_base, counts);
    if (nblk > 0) {
        vox_assign<<<nblk, VOX_BLOCK, 0, stream>>>(p, w.keys, w.slots, w.slot_of_pt, w.block_prefix, w.frame_base,
                                                   coords, num, w.row_slot);
        long long rows = cap_rows < n ? cap_rows : n;
        long long threads = rows * c;
        if (threads > 0 && (voxels || mean))
            vox_fill<<<(unsigned)div_up(threads, 256), 256, 0, stream>>>(points, p, w.slots, w.row_slot, counts, voxels, mean);
        count_launch(5);
    } else {
        count_launch(1);
    }
    return launch_status("cpd_voxelize");
}


// ---------------------------------------------------------------------------------------------------------------
// Host entry with the same semantics for ONE frame: the reference calls
// Point2VoxelCPU3d.point_to_voxel inside Dataset.__getitem__, i.e. in forked DataLoader worker processes
// (cpd/datasets/processor/data_processor.py:133-144), where CUDA must not be touched.  Sequential first-come
// grouping exactly as upstream; an open-addressing table over the occupied cells replaces upstream's dense
// 92.7 M-cell lookup grid.  All pointers are HOST pointers.  Returns the number of voxels (>= 0) or a cpd_status (< 0).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int64_t cpd_voxelize_cpu(const float *points, int64_t n, int32_t c, const float *range6, const float *vsize3,
                                    int32_t max_pts, int64_t max_voxels, float *voxels, int32_t *coords_zyx, int32_t *num_points)
{
    CPD_REQUIRE(n >= 0 && c >= 3 && max_pts >= 1 && max_voxels >= 0, CPD_ERR_BAD_ARG, "cpd_voxelize_cpu: bad n/c/max_pts");
    CPD_REQUIRE((points || n == 0) && range6 && vsize3 && voxels && coords_zyx && num_points, CPD_ERR_BAD_ARG, "cpd_voxelize_cpu: null argument");
    long long grid[3];
    for (int j = 0; j < 3; ++j) {
        grid[j] = (long long)roundf((range6[3 + j] - range6[j]) / vsize3[j]);
        CPD_REQUIRE(grid[j] > 0, CPD_ERR_BAD_ARG, "cpd_voxelize_cpu: empty grid");
    }
    uint64_t cap = 1024;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    long long *keys = (long long *)malloc(sizeof(long long) * cap);
    int32_t *vals = (int32_t *)malloc(sizeof(int32_t) * cap);
    if (!keys || !vals) { free(keys); free(vals); set_error("cpd_voxelize_cpu: out of host memory"); return CPD_ERR_BAD_ARG; }
    for (uint64_t i = 0; i < cap; ++i) keys[i] = -1;
    int64_t nvox = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = points + i * c;
        int cell[3];
        bool ok = true;
        for (int j = 0; j < 3 && ok; ++j) {
            const float t = floorf((p[j] - range6[j]) / vsize3[j]);       // fp32 subtract, divide, floor -- as upstream and as the kernel
            ok = (t >= 0.0f) && (t < (float)grid[j]);
            cell[j] = (int)t;
        }
        if (!ok) continue;
        const long long key = ((long long)cell[2] * grid[1] + cell[1]) * grid[0] + cell[0];
        uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
        uint64_t s = (h >> 20) & (cap - 1);
        while (keys[s] != -1 && keys[s] != key) s = (s + 1) & (cap - 1);
        int32_t vid;
        if (keys[s] == -1) {
            if (nvox >= max_voxels) continue;                               // at capacity: new cells are dropped
            keys[s] = key;
            vid = vals[s] = (int32_t)nvox++;
            coords_zyx[3 * vid] = cell[2]; coords_zyx[3 * vid + 1] = cell[1]; coords_zyx[3 * vid + 2] = cell[0];
            num_points[vid] = 0;
            memset(voxels + (size_t)vid * max_pts * c, 0, sizeof(float) * (size_t)max_pts * c);
        } else {
            vid = vals[s];
        }
        const int32_t k = num_points[vid];
        if (k < max_pts) {
            memcpy(voxels + ((size_t)vid * max_pts + k) * c, p, sizeof(float) * (size_t)c);
            num_points[vid] = k + 1;
        }
    }
    free(keys); free(vals);
    return nvox;
}
