// C-ABI shim over the launchers of the reference's OWN CUDA files cpd/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:92-117 and
// group_points_gpu.cu:47-66,104-125, compiled unmodified for sm_100a into oracle/_ref/libpointnet2_ref_gpu.so.  Used by the -m gpu
// tests as the bit-exact oracle for cpd_voxel_query / cpd_group_points[_bwd].
#include <cuda_runtime_api.h>
void voxel_query_kernel_launcher_stack(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                                       const float *new_xyz, const float *xyz, const int *new_coords, const int *point_indices, int *idx);
void group_points_kernel_launcher_stack(int B, int M, int C, int nsample, const float *features, const int *features_batch_cnt, const int *idx,
                                        const int *idx_batch_cnt, float *out);
void group_points_grad_kernel_launcher_stack(int B, int M, int C, int N, int nsample, const float *grad_out, const int *idx, const int *idx_batch_cnt,
                                             const int *features_batch_cnt, float *grad_features);
extern "C" {
int ref_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int zr, int yr, int xr, const float *new_xyz, const float *xyz,
                    const int *new_coords, const int *point_indices, int *idx)
{ voxel_query_kernel_launcher_stack(M, R1, R2, R3, nsample, radius, zr, yr, xr, new_xyz, xyz, new_coords, point_indices, idx); return (int)cudaDeviceSynchronize(); }
int ref_group_points(int B, int M, int C, int nsample, const float *f, const int *fcnt, const int *idx, const int *icnt, float *out)
{ group_points_kernel_launcher_stack(B, M, C, nsample, f, fcnt, idx, icnt, out); return (int)cudaDeviceSynchronize(); }
int ref_group_points_grad(int B, int M, int C, int N, int nsample, const float *g, const int *idx, const int *icnt, const int *fcnt, float *gf)
{ group_points_grad_kernel_launcher_stack(B, M, C, N, nsample, g, idx, icnt, fcnt, gf); return (int)cudaDeviceSynchronize(); }
}
