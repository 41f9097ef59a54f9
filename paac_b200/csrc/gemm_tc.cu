// tcgen05 (5th-gen tensor core) implicit-GEMM path -- placeholder until the kernels land.
#include "common.cuh"
namespace paacb {
int launch_conv_fwd_tc(const paacb_ctx*, const LayerGeom&, const void*, const float*, const float*, float*, int64_t, int,
                       cudaStream_t) {
  return PAACB_EUNSUPPORTED;
}
}  // namespace paacb
