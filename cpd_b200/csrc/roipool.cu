// RoI grid pooling primitives of VoxelRCNN[Proto]Head for sm_100a: voxel query + point grouping.
//
// Replaces cpd/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-89 (voxel_query_kernel_stack) and
// group_points_gpu.cu:15-125 (group_points[_grad]_kernel_stack) as driven by
// cpd/models/roi_heads/voxel_rcnn_head.py:186-273 and pointnet2_stack/voxel_query_utils.py:12-110.
//
// The reference first materialises generate_voxel2pinds (cpd/utils/spconv_utils.py:4-21): a dense int32
// (B, Z, Y, X) map of voxel -> row, ~25 MB per scale per call at (4, 11, 376, 376), filled with -1 and scattered by
// torch.  Here the query probes the coordinate hash the rulebooks of that level already built (cpd_coord_hash_build):
// no dense map, no memset, no scatter.  A `dense map` variant is kept so that the reference's own Python
// (voxel_query_wrapper of its pybind module) runs on the shim unchanged.
//
// Semantics reproduced exactly (results are bit-identical to the reference kernels, tests/test_gpu_roipool.py):
// the (2 rz + 1)(2 ry + 1)(2 rx + 1) neighbourhood is scanned dz -> dy -> dx ascending; a voxel is taken when its centre
// lies within `radius` of the query point (dist2 > radius2 rejects); the first hit fills all nsample slots, later hits
// overwrite slots 1, 2, ... up to nsample; an empty ball is flagged (the reference writes idx[0] = -1 and zeroes the row
// in Python afterwards: here the row is written as zeros and the flag goes to `empty`).
// Integer / latency-bound work: one thread per query point, 256-thread CTAs.
#include "common.cuh"

namespace cpd {
namespace {

struct QueryGeo {
    int d, h, w, batch;       // grid of this scale (z, y, x)
    int rz, ry, rx, nsample;
    float radius2;
};

template <bool DENSE>
__global__ void __launch_bounds__(256) voxel_query_kernel(const float *__restrict__ new_xyz, const int32_t *__restrict__ new_coords, long long m,
                                                          const float *__restrict__ xyz, HashView hash, const int32_t *__restrict__ dense_map,
                                                          QueryGeo g, int32_t *__restrict__ idx, uint8_t *__restrict__ empty)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const float qx = new_xyz[q * 3], qy = new_xyz[q * 3 + 1], qz = new_xyz[q * 3 + 2];
    const int4 c = __ldg(reinterpret_cast<const int4 *>(new_coords) + q);      // [b, z, y, x]
    int32_t *out = idx + q * g.nsample;
    int cnt = 0;
    if (c.x >= 0 && c.x < g.batch) {
        for (int dz = -g.rz; dz <= g.rz; ++dz) {
            const int z = c.y + dz;
            if (z < 0 || z >= g.d) continue;
            for (int dy = -g.ry; dy <= g.ry; ++dy) {
                const int y = c.z + dy;
                if (y < 0 || y >= g.h) continue;
                for (int dx = -g.rx; dx <= g.rx; ++dx) {
                    const int x = c.w + dx;
                    if (x < 0 || x >= g.w) continue;
                    const long long cell = (((long long)c.x * g.d + z) * g.h + y) * g.w + x;
                    const int32_t nb = DENSE ? __ldg(dense_map + cell) : hash_find(hash, (uint32_t)cell);
                    if (nb < 0) continue;
                    const float px = __ldg(xyz + (size_t)nb * 3), py = __ldg(xyz + (size_t)nb * 3 + 1), pz = __ldg(xyz + (size_t)nb * 3 + 2);
                    const float dist2 = (px - qx) * (px - qx) + (py - qy) * (py - qy) + (pz - qz) * (pz - qz);
                    if (dist2 > g.radius2) continue;
                    if (cnt < g.nsample) {
                        if (cnt == 0)
                            for (int l = 0; l < g.nsample; ++l) out[l] = nb;
                        out[cnt] = nb;
                        ++cnt;
                    }
                }
            }
        }
    }
    if (cnt == 0)
        for (int l = 0; l < g.nsample; ++l) out[l] = 0;
    if (empty) empty[q] = cnt == 0;
}

// out[m, c, s] = features[idx[m, s], c]   (the reference's (M, C, nsample) layout; idx are GLOBAL rows)
__global__ void __launch_bounds__(256) group_points_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, long long m, int c,
                                                           int nsample, float *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * c * nsample) return;
    const int s = (int)(t % nsample), ch = (int)((t / nsample) % c);
    const long long q = t / ((long long)nsample * c);
    out[t] = __ldg(feat + (size_t)__ldg(idx + q * nsample + s) * c + ch);
}

__global__ void __launch_bounds__(256) group_points_grad_kernel(const float *__restrict__ gout, const int32_t *__restrict__ idx, long long m, int c,
                                                                int nsample, float *__restrict__ gfeat)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * c * nsample) return;
    const int s = (int)(t % nsample), ch = (int)((t / nsample) % c);
    const long long q = t / ((long long)nsample * c);
    atomicAdd(gfeat + (size_t)__ldg(idx + q * nsample + s) * c + ch, __ldg(gout + t));
}

}  // namespace
}  // namespace cpd

using namespace cpd;

// new_xyz (m, 3) query points, new_coords (m, 4) int32 [b, z, y, x] their cells at this scale, xyz (n, 3) centres of the
// level's voxels (rows as in its SparseConvTensor).  Exactly one of (hash, dense_map) is given: hash = cpd_coord_hash_build of
// the level's coords; dense_map = the reference's (B, Z, Y, X) int32 voxel -> row tensor.  idx (m, nsample) int32 rows
// (global), empty (m,) uint8 (NULL ok).
extern "C" int32_t cpd_voxel_query(const float *new_xyz, const int32_t *new_coords, int64_t m, const float *xyz, const void *hash,
                                   size_t hash_bytes, const int32_t *dense_map, const int32_t *shape3_host, int32_t batch,
                                   const int32_t *range3_host, float radius, int32_t nsample, int32_t *idx, uint8_t *empty,
                                   cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(new_xyz && new_coords && xyz && shape3_host && range3_host && idx && m >= 0, CPD_ERR_BAD_ARG, "cpd_voxel_query: bad argument");
    CPD_REQUIRE((hash != nullptr) != (dense_map != nullptr), CPD_ERR_BAD_ARG, "cpd_voxel_query: pass the coordinate hash or the dense map");
    CPD_REQUIRE(nsample >= 1 && nsample <= 256 && batch >= 1, CPD_ERR_BAD_ARG, "cpd_voxel_query: bad nsample / batch");
    CPD_REQUIRE(((uintptr_t)new_coords & 15) == 0, CPD_ERR_MISALIGNED, "cpd_voxel_query: new_coords must be 16-byte aligned");
    CPD_REQUIRE((long long)batch * shape3_host[0] * shape3_host[1] * shape3_host[2] < 0xFFFFFFFFll, CPD_ERR_UNSUPPORTED,
                "cpd_voxel_query: batch*volume exceeds 32-bit cell keys");
    if (m == 0) return CPD_OK;
    QueryGeo g{shape3_host[0], shape3_host[1], shape3_host[2], batch, range3_host[0], range3_host[1], range3_host[2], nsample, radius * radius};
    const unsigned grid = (unsigned)div_up(m, 256);
    if (hash) {
        CPD_REQUIRE(hash_bytes >= 8 * 1024 && (hash_bytes & (hash_bytes - 1)) == 0, CPD_ERR_BAD_ARG, "cpd_voxel_query: bad hash buffer");
        const uint64_t cap = hash_bytes / 8;
        HashView h{(uint32_t *)hash, (int32_t *)((char *)hash + cap * 4), (uint32_t)(cap - 1)};
        voxel_query_kernel<false><<<grid, 256, 0, stream>>>(new_xyz, new_coords, m, xyz, h, nullptr, g, idx, empty);
    } else {
        voxel_query_kernel<true><<<grid, 256, 0, stream>>>(new_xyz, new_coords, m, xyz, HashView{}, dense_map, g, idx, empty);
    }
    count_launch();
    return launch_status("cpd_voxel_query");
}

extern "C" int32_t cpd_group_points(const float *features, const int32_t *idx, int64_t m, int32_t c, int32_t nsample, float *out,
                                    cpd_stream_t stream)
{
    CPD_REQUIRE(features && idx && out && m >= 0 && c >= 1 && nsample >= 1, CPD_ERR_BAD_ARG, "cpd_group_points: bad argument");
    const long long t = m * (long long)c * nsample;
    if (t == 0) return CPD_OK;
    group_points_kernel<<<(unsigned)div_up(t, 256), 256, 0, (cudaStream_t)stream>>>(features, idx, m, c, nsample, out);
    count_launch();
    return launch_status("cpd_group_points");
}

// grad_features (n, c) is OVERWRITTEN: zeroed, then accumulated with fp32 atomics like the reference
extern "C" int32_t cpd_group_points_bwd(const float *grad_out, const int32_t *idx, int64_t m, int32_t c, int32_t nsample, int64_t n,
                                        float *grad_features, cpd_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    CPD_REQUIRE(grad_out && idx && grad_features && m >= 0 && n >= 0 && c >= 1 && nsample >= 1, CPD_ERR_BAD_ARG, "cpd_group_points_bwd: bad argument");
    CPD_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)n * c, stream));
    const long long t = m * (long long)c * nsample;
    if (t == 0) return CPD_OK;
    group_points_grad_kernel<<<(unsigned)div_up(t, 256), 256, 0, stream>>>(grad_out, idx, m, c, nsample, grad_features);
    count_launch();
    return launch_status("cpd_group_points_bwd");
}
