// Error channel and ABI housekeeping of libvface_b200.so.
#include "vf_common.cuh"

#include <cstring>

namespace vf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 2;
}

// The kernels are compiled for sm_100a only; anything else cannot run them and there is
// deliberately no fallback.
int check_device() {
  static int ok = -1;
  if (ok == 1) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice (no CUDA device: vface_b200 has no CPU fallback)");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    ok = 0;
    return fail("vface_b200 needs an sm_100a device (B200); found sm_%d%d", major, minor);
  }
  ok = 1;
  return 0;
}

}  // namespace vf

extern "C" int vf_abi_version(void) { return VF_ABI_VERSION; }

extern "C" const char* vf_last_error(void) { return vf::g_err; }
