// MojoPagedPrefillGQA / MojoSdpa entry points (placeholder until the tensor-core kernels land).
#include "common.cuh"

extern "C" int mojo_b200_paged_prefill_gqa(
    const void*, const void*, const void*, const int32_t*, const int32_t*, const int32_t*, void*, int64_t, int, int,
    int, int, int64_t, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
    int64_t, int64_t, int64_t, int64_t, float, int, int, int, void*) {
  return mojo::fail(MOJO_B200_EUNSUPPORTED, "paged_prefill_gqa: kernel not built yet");
}

extern "C" int mojo_b200_sdpa(const void*, const void*, const void*, void*, int, int, int, int64_t, int64_t, int,
                              int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                              int64_t, int64_t, int64_t, float, int, void*) {
  return mojo::fail(MOJO_B200_EUNSUPPORTED, "sdpa: kernel not built yet");
}
