// tcgen05 NeRF-W MLP — placeholder until the tensor-core kernel lands.
#include "common.cuh"
namespace dfb {
bool tc_supported(const DfbNerf*, int, int) { return false; }
int pack_tc_weights(DfbNerf*, int, const std::vector<std::vector<float>>&) { return DFB_OK; }
int launch_mlp_tc_rays(const DfbNerf*, int, int, int, const float*, const float*, const float*, int64_t, int, float*,
                       cudaStream_t) {
  set_error("tcgen05 MLP not built");
  return DFB_ERR_UNSUPPORTED;
}
}  // namespace dfb
