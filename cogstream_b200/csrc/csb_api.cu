// C-ABI plumbing shared by all entry points: error text, ABI version, device gate.
#include <stdarg.h>
#include <string.h>

#include "csb_common.cuh"

namespace csb {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels launched by this library in this process (bench bookkeeping)

void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return CSB_ERR_CUDA;
  }
  return CSB_OK;
}

int device_ok() {
  // cached per device id; the kernels are compiled for sm_100a only and there is no fallback
  static int cached_dev = -1, cached_status = CSB_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s (a CUDA device is required; there is no CPU fallback)", cudaGetErrorString(e));
    return CSB_ERR_CUDA;
  }
  if (dev == cached_dev) return cached_status;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major == 10) {
    cached_status = CSB_OK;
  } else {
    set_error("device %d is sm_%d%d; cogstream_b200 kernels are built for sm_100a only", dev, major, minor);
    cached_status = CSB_ERR_ARCH;
  }
  return cached_status;
}

}  // namespace csb

extern "C" {

int csb_abi_version(void) { return CSB_ABI_VERSION; }
const char* csb_last_error(void) { return csb::g_err; }
int csb_check_device(void) { return csb::device_ok(); }
uint64_t csb_kernel_launches(void) { return (uint64_t)__atomic_load_n(&csb::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
