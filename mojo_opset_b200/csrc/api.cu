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

// Device error word: kernels cannot raise, and a host check of device data would synchronise (the reference's
// `if block_tables[i, 0] < 0: raise ValueError` reads the table on the host, attention.py:186-187).  The caller
// registers one int of device memory per device; kernels OR a MOJO_B200_ERR_* bit into it when they meet malformed
// data and the caller reads it whenever it is convenient (mojo_opset_b200.check_device_errors()).
static int* g_error_word[64] = {nullptr};

int* error_word() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_error_word[dev];
}

// Arrival counters of the split-KV decode's in-kernel fold (paged_decode.cu): `count` zero-initialised ints of device
// memory per device, registered by the caller like the error word; the kernels leave them zero.
static int* g_decode_tickets[64] = {nullptr};
static int64_t g_decode_ticket_count[64] = {0};

int* decode_tickets(int64_t* count) {
  int dev = 0;
  *count = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  *count = g_decode_ticket_count[dev];
  return g_decode_tickets[dev];
}

bool pdl_enabled() {
  const char* v = getenv("MOJO_B200_PDL");
  return !(v && v[0] == '0');
}

}  // namespace mojo

extern "C" {

int mojo_b200_set_error_word(int* device_word) {
  int dev = 0;
  MOJO_CUDA_OK(cudaGetDevice(&dev));
  MOJO_REQUIRE(dev >= 0 && dev < 64, MOJO_B200_EUNSUPPORTED, "set_error_word: device ordinal %d", dev);
  mojo::g_error_word[dev] = device_word;
  return 0;
}

int mojo_b200_set_decode_tickets(int* device_words, int64_t count) {
  int dev = 0;
  MOJO_CUDA_OK(cudaGetDevice(&dev));
  MOJO_REQUIRE(dev >= 0 && dev < 64, MOJO_B200_EUNSUPPORTED, "set_decode_tickets: device ordinal %d", dev);
  MOJO_REQUIRE(count >= 0 && (device_words || count == 0), MOJO_B200_EINVAL, "set_decode_tickets: bad arguments");
  mojo::g_decode_tickets[dev] = count > 0 ? device_words : nullptr;
  mojo::g_decode_ticket_count[dev] = count > 0 ? count : 0;
  return 0;
}

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
