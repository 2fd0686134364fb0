// Python binding for the reference's OWN CPU IoU (cpd/ops/iou3d_nms/src/iou3d_cpu.cpp),
// compiled from the sources where they lie under /root/reference into oracle/_ref/.
// The reference co-links this file with its .cu, which breaks on this toolchain
// (SURVEY.md section 0.4); built alone it works.  Test infrastructure only.
#include <torch/extension.h>
#include "iou3d_cpu.h"   // -I /root/reference/cpd/ops/iou3d_nms/src

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "reference rotated BEV IoU (CPU)");
}
