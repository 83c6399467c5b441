// Placeholder until the fused tcgen05 readout lands: reports "not covered" for every shape.
#include "common.cuh"
namespace swem {
bool fused_readout_supported(const SwemDims&) { return false; }
size_t fused_readout_workspace(const SwemDims&) { return 0; }
int fused_readout_forward(const SwemReadArgs&, cudaStream_t) { set_error("fused readout not built"); return SWEM_ERR_UNSUPPORTED; }
}  // namespace swem
