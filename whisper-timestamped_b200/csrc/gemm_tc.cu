// tcgen05 tensor-core GEMM (placeholder until the kernel lands: reports an error so that nothing
// silently falls back).
#include "common.cuh"

namespace wts {

int gemm_tc_launch(const WtsGemm& g, cudaStream_t st)
{
    (void)g; (void)st;
    set_error("wts_gemm: tcgen05 backend not built yet; pass backend=1");
    return -3;
}

}  // namespace wts
