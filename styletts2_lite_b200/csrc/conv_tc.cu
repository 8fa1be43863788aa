// placeholder: replaced by the tcgen05/TMA implicit-GEMM kernel
#include "common.cuh"
namespace st2 {
bool conv_tc_supported(const ConvArgs&) { return false; }
int launch_conv_tc(const ConvArgs&, cudaStream_t) {
    set_error("tensor-core conv path not built");
    return ST2_ERR_UNSUPPORTED;
}
}  // namespace st2
