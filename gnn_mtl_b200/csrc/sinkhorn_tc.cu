// (b) TMA + tcgen05 (3xTF32) cost-tile path of the fused Sinkhorn half-sweep.
// Placeholder until the tensor-core kernel lands: reports "unsupported" so callers
// fail loudly instead of silently taking another path.
#include "common.cuh"

namespace eg {

size_t lse_fused_tc_workspace(int64_t, int64_t, int) { return 256; }

int lse_fused_tc(int, int64_t, int64_t, int, const float*, const float*, float, const float*, const float*,
                 float*, float*, const float*, const float*, const float*, const float*, void*, size_t,
                 cudaStream_t) {
  return EG_ERR_UNSUPPORTED;
}

}  // namespace eg
