"""Compile the reference's own iou3d_cpu.cpp (alone) into oracle/_ref/iou3d_ref_cpu.so.

Run by `make -C oracle ref` in the build container (where /root/reference exists);
the GPU box only ever loads the prebuilt file.  Test infrastructure only.
"""
import os
import sys

from torch.utils.cpp_extension import load

here = os.path.dirname(os.path.abspath(__file__))
ref_src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/cpd/ops/iou3d_nms/src"
out = os.path.join(here, "_ref")
os.makedirs(out, exist_ok=True)
load(
    name="iou3d_ref_cpu",
    sources=[os.path.join(ref_src, "iou3d_cpu.cpp"),
             os.path.join(here, "ref_bindings", "iou3d_cpu_module.cpp")],
    extra_include_paths=[ref_src, "/usr/local/cuda/include"],
    extra_cflags=["-O2", "-w"],
    build_directory=out,
    verbose=False,
)
print("built", os.path.join(out, "iou3d_ref_cpu.so"))
