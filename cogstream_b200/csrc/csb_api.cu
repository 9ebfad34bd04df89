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
  // cached per host thread and device id; the kernels are compiled for sm_100a only (an arch-specific cubin does not load on
  // other 10.x parts such as sm_103) and there is no fallback
  static thread_local int cached_dev = -1, cached_status = CSB_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s (a CUDA device is required; there is no CPU fallback)", cudaGetErrorString(e));
    return CSB_ERR_CUDA;
  }
  if (dev == cached_dev) {
    if (cached_status != CSB_OK) set_error("device %d is not sm_100; cogstream_b200 kernels are built for sm_100a only", dev);
    return cached_status;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  int status = CSB_OK;
  if (!(major == 10 && minor == 0)) {
    set_error("device %d is sm_%d%d; cogstream_b200 kernels are built for sm_100a only", dev, major, minor);
    status = CSB_ERR_ARCH;
  }
  cached_status = status;
  cached_dev = dev;
  return status;
}

}  // namespace csb

extern "C" {

int csb_abi_version(void) { return CSB_ABI_VERSION; }
const char* csb_last_error(void) { return csb::g_err; }
int csb_check_device(void) { return csb::device_ok(); }
uint64_t csb_kernel_launches(void) { return (uint64_t)__atomic_load_n(&csb::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
