// tcgen05 gather-GEMM (placeholder until the tensor-core kernel lands; reports unsupported
// so that CPD_ALGO_AUTO resolves to the SIMT kernel and CPD_ALGO_TCGEN05 fails loudly).
#include "common.cuh"
namespace cpd {
bool gather_gemm_tc_supported(int32_t, int32_t, int32_t) { return false; }
size_t gather_gemm_tc_workspace(int64_t, int32_t, int32_t, int32_t) { return 0; }
int32_t gather_gemm_tc(const float *, int64_t, int32_t, const float *, int32_t, int32_t, const int32_t *, int64_t,
                       const float *, const float *, const float *, const float *, int32_t, float *, float *, void *,
                       size_t, cudaStream_t)
{
    set_error("tcgen05 gather-GEMM not built");
    return CPD_ERR_UNSUPPORTED;
}
}  // namespace cpd
