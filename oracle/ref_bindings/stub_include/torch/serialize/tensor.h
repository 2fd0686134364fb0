// Stand-in for <torch/serialize/tensor.h> when compiling the reference's pointnet2_stack .cu files for oracle/_ref: their
// headers only DECLARE wrappers taking at::Tensor by value (the kernels and launchers use plain pointers), so an
// incomplete type is enough and the 90-second libtorch header parse (and the link dependency) is avoided.
#pragma once
namespace at { class Tensor; }
