// Error channel and ABI housekeeping of libvface_b200.so.
#include "vf_common.cuh"

#include <atomic>
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
//
// One process drives ONE GPU (include/vface_b200.h): the library keeps per-process caches that belong to
// a device/context (cudaFuncSetAttribute opt-ins for > 48 KB of dynamic shared memory, the cuBLASLt handle
// and its plans, the SM count).  The first entry point binds the library to the device that is current at
// that moment; a later call with another device current is refused instead of launching with attributes,
// handles or grid sizes that belong to the first one.
static std::atomic<int> g_bound_device{-1};

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice (no CUDA device: vface_b200 has no CPU fallback)");
  const int bound = g_bound_device.load(std::memory_order_acquire);
  if (bound == dev) return 0;
  if (bound >= 0)
    return fail("vface_b200 is bound to CUDA device %d (one process per GPU) but device %d is current; "
                "run one process per GPU (torchrun) or make device %d current before calling", bound, dev, bound);
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return fail("vface_b200 needs an sm_100a device (B200); found sm_%d%d", major, minor);
  int expected = -1;
  if (!g_bound_device.compare_exchange_strong(expected, dev) && expected != dev)
    return fail("vface_b200 is bound to CUDA device %d (one process per GPU) but device %d is current", expected, dev);
  return 0;
}

}  // namespace vf

extern "C" int vf_abi_version(void) { return VF_ABI_VERSION; }

extern "C" const char* vf_last_error(void) { return vf::g_err; }
