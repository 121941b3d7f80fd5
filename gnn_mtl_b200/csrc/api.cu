// Library-level entry points of include/eagraft.h: version, errors, launch counter.
#include "common.cuh"

namespace eg {
std::atomic<int64_t> g_launches{0};
thread_local int t_last_cuda_error = 0;
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

int64_t eg_launch_count(void) { return eg::g_launches.load(); }
void eg_launch_count_reset(void) { eg::g_launches.store(0); }

}  // extern "C"
