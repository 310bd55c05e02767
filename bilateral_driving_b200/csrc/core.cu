// Error reporting and library identity for libbds_b200.so.
#include <stdarg.h>

#include "bds_common.cuh"

namespace bds {
static thread_local char g_err[1024] = "";
unsigned long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace bds

extern "C" const char* bds_last_error(void) { return bds::g_err; }
extern "C" int bds_abi_version(void) { return BDS_ABI_VERSION; }
extern "C" int bds_device_arch(void) {
  int dev = 0;
  cudaDeviceProp prop;
  BDS_CHECK_CUDA(cudaGetDevice(&dev));
  BDS_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  return prop.major * 10 + prop.minor;
}
extern "C" unsigned long long bds_launch_count(void) { return __atomic_load_n(&bds::g_launches, __ATOMIC_RELAXED); }
