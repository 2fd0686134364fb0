// C-ABI shim over the launchers of the reference's OWN CUDA file
// (cpd/ops/iou3d_nms/src/iou3d_nms_kernel.cu:376-414), compiled unmodified for sm_100a
// into oracle/_ref/libiou3d_ref_gpu.so.  Prototypes as declared in iou3d_nms.cpp:43-46.
// Used by the -m gpu tests as the bit-exact oracle for IoU matrices and NMS masks.
#include <cuda_runtime_api.h>

void boxesoverlapLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_overlap);
void boxesioubevLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_iou);
void nmsLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);

extern "C" {
int ref_overlap_bev(const float *a, int na, const float *b, int nb, float *out)
{ boxesoverlapLauncher(na, a, nb, b, out); return (int)cudaDeviceSynchronize(); }
int ref_iou_bev(const float *a, int na, const float *b, int nb, float *out)
{ boxesioubevLauncher(na, a, nb, b, out); return (int)cudaDeviceSynchronize(); }
int ref_nms_mask(const float *boxes, int n, float thresh, unsigned long long *mask)
{ nmsLauncher(boxes, mask, n, thresh); return (int)cudaDeviceSynchronize(); }
int ref_nms_normal_mask(const float *boxes, int n, float thresh, unsigned long long *mask)
{ nmsNormalLauncher(boxes, mask, n, thresh); return (int)cudaDeviceSynchronize(); }
}
