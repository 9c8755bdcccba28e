// Error plumbing + version / device probes of the C ABI (include/mojo_b200.h).
#include "common.cuh"

namespace mojo {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace mojo

extern "C" {

int mojo_b200_abi_version(void) { return MOJO_B200_ABI_VERSION; }

const char* mojo_b200_last_error(void) { return mojo::last_error_buffer(); }

int mojo_b200_device_ok(void) {
  int dev = 0;
  MOJO_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  MOJO_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

}  // extern "C"
