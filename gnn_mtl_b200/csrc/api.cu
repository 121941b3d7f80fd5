// Library-level entry points of include/eagraft.h: version, errors, launch counter.
#include "common.cuh"

namespace eg {
std::atomic<int64_t> g_launches{0};
thread_local int t_last_cuda_error = 0;
// tuning / diagnostic knobs owned by the kernels' translation units
extern int g_tune_unroll, g_tune_warps, g_tune_hints, g_tune_spmm_persist, g_tune_spmm_slab, g_tune_spmm_bulk, g_tune_spmm_dynamic;                       // spmm.cu
extern int g_tune_persistent, g_tune_resident, g_tune_onchip, g_tune_scaling, g_tune_tile2d;     // sinkhorn_dense.cu
extern int g_tune_absorb_milli, g_tune_force_fallback;
extern int g_tune_tc_pair;                                                                        // sinkhorn_tc.cu
extern int g_tune_gemm_pair;                                                                      // gemm_nt_raw.cu
extern int g_tune_l1_filter;                                                                      // eval_l1.cu
int sinkhorn_redo_count();
int sinkhorn_absorb_count();
}  // namespace eg

extern "C" {

int eg_version(void) { return 100; }

const char* eg_strerror(int status) {
  switch (status) {
    case EG_OK: return "ok";
    case EG_ERR_INVALID: return "invalid argument";
    case EG_ERR_CUDA: return "CUDA runtime error (see eg_last_cuda_error)";
    case EG_ERR_WORKSPACE: return "workspace too small";
    case EG_ERR_UNSUPPORTED: return "unsupported request";
    case EG_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown status";
  }
}

int eg_last_cuda_error(void) { return eg::t_last_cuda_error; }

int eg_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return EG_ERR_NO_DEVICE;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return EG_ERR_NO_DEVICE;
  return major == 10 ? EG_OK : EG_ERR_NO_DEVICE;
}

int eg_debug_set(int key, int value) {
  switch (key) {
    case 0: eg::g_tune_unroll = value; break;
    case 1: eg::g_tune_warps = value; break;
    case 2: eg::g_tune_hints = value; break;
    case 3: eg::g_tune_persistent = value; break;
    case 4: eg::g_tune_resident = value; break;
    case 5: eg::g_tune_onchip = value; break;
    case 6: eg::g_tune_spmm_persist = value; break;
    case 7: eg::g_tune_scaling = value; break;
    case 8: return eg::sinkhorn_redo_count();
    case 9: return eg::sinkhorn_absorb_count();
    case 10: eg::g_tune_absorb_milli = value; break;
    case 11: eg::g_tune_force_fallback = value; break;
    case 12: eg::g_tune_tile2d = value; break;
    case 13: eg::g_tune_l1_filter = value; break;
    case 14: eg::g_tune_spmm_slab = value; break;
    case 16: eg::g_tune_spmm_bulk = value; break;
    case 17: eg::g_tune_spmm_dynamic = value; break;
    case 18: eg::g_tune_gemm_pair = value; break;
    case 19: eg::g_tune_tc_pair = value; break;
    default: return EG_ERR_INVALID;
  }
  return EG_OK;
}

int64_t eg_launch_count(void) { return eg::g_launches.load(); }
void eg_launch_count_reset(void) { eg::g_launches.store(0); }

}  // extern "C"
