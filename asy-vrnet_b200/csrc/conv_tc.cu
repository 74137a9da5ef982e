// tcgen05 (5th-gen tensor core) bf16 variant of the convolution-as-GEMM engine.  [placeholder until the kernel lands]
#include "conv_common.cuh"
namespace vrcoc {
bool conv_tc_supported(const ConvArgs&) { return false; }
int launch_conv_tc(const ConvArgs&, cudaStream_t) { return fail(VRCOC_EINVAL, "tcgen05 conv engine not built"); }
}  // namespace vrcoc
