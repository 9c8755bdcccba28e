// Host-side cuTensorMapEncodeTiled wrapper with a small descriptor cache.  The driver symbol is resolved
// through the runtime (cudaGetDriverEntryPoint) so the library does not link libcuda.
#include <cstring>
#include <mutex>
#include <vector>

#include "tma.cuh"

namespace mojo {

bool TensorMapKey::operator==(const TensorMapKey& o) const {
  if (base != o.base || rank != o.rank || dtype != o.dtype || swizzle != o.swizzle) return false;
  for (int i = 0; i < rank; ++i)
    if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
  for (int i = 0; i + 1 < rank; ++i)
    if (strides[i] != o.strides[i]) return false;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int get_tensor_map(const TensorMapKey& key, CUtensorMap* out) {
  struct Entry {
    TensorMapKey key;
    CUtensorMap map;
  };
  static std::mutex mu;
  static std::vector<Entry> cache;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
      if (e.key == key) {
        *out = e.map;
        return 0;
      }
  }
  EncodeTiledFn encode = resolve_encode();
  if (!encode) return fail(MOJO_B200_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");

  CUtensorMapDataType dt;
  switch (key.dtype) {
    case MOJO_B200_BF16: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case MOJO_B200_F16: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; break;
    case MOJO_B200_F32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    default: return fail(MOJO_B200_EINVAL, "tensor map: bad dtype %d", key.dtype);
  }
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], elem[5];
  for (int i = 0; i < key.rank; ++i) {
    dims[i] = key.dims[i];
    box[i] = key.box[i];
    elem[i] = 1;
  }
  for (int i = 0; i + 1 < key.rank; ++i) strides[i] = key.strides[i];
  CUtensorMap map;
  CUresult rc = encode(&map, dt, (cuuint32_t)key.rank, const_cast<void*>(key.base), dims, strides, box, elem,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)key.swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    return fail(MOJO_B200_EUNSUPPORTED,
                "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu,%llu] strides "
                "[%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u]",
                (int)rc, key.rank, (unsigned long long)key.dims[0], (unsigned long long)key.dims[1],
                (unsigned long long)(key.rank > 2 ? key.dims[2] : 0), (unsigned long long)(key.rank > 3 ? key.dims[3] : 0),
                (unsigned long long)(key.rank > 4 ? key.dims[4] : 0), (unsigned long long)key.strides[0],
                (unsigned long long)(key.rank > 2 ? key.strides[1] : 0), (unsigned long long)(key.rank > 3 ? key.strides[2] : 0),
                (unsigned long long)(key.rank > 4 ? key.strides[3] : 0), key.box[0], key.box[1], key.rank > 2 ? key.box[2] : 0,
                key.rank > 3 ? key.box[3] : 0, key.rank > 4 ? key.box[4] : 0);
  }
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 256) cache.erase(cache.begin());
    cache.push_back(Entry{key, map});
  }
  *out = map;
  return 0;
}

int build_cache_map(const void* base, int dtype, int head_dim, int64_t block_size, int num_kv_heads,
                    int64_t num_blocks, int64_t s_b, int64_t s_h, int64_t s_t, bool split_halves, int box_rows,
                    CUtensorMap* out) {
  const int nh = head_dim / 64;
  TensorMapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = 5;
  key.dtype = dtype;
  key.swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  key.dims[0] = 64;
  key.box[0] = 64;
  if (split_halves) {  // [64 | token | half | head | block]: a box lands as [half][token][128 B]
    key.dims[1] = (uint64_t)block_size; key.strides[0] = (uint64_t)s_t * 2; key.box[1] = (uint32_t)box_rows;
    key.dims[2] = (uint64_t)nh;         key.strides[1] = 128;               key.box[2] = (uint32_t)nh;
  } else {             // [64 | half | token | head | block]: a box lands as [token][half][128 B]
    key.dims[1] = (uint64_t)nh;         key.strides[0] = 128;               key.box[1] = (uint32_t)nh;
    key.dims[2] = (uint64_t)block_size; key.strides[1] = (uint64_t)s_t * 2; key.box[2] = (uint32_t)box_rows;
  }
  key.dims[3] = (uint64_t)num_kv_heads; key.strides[2] = (uint64_t)s_h * 2; key.box[3] = 1;
  key.dims[4] = (uint64_t)num_blocks;   key.strides[3] = (uint64_t)s_b * 2; key.box[4] = 1;
  return get_tensor_map(key, out);
}

int build_tile_map(const void* base, int dtype, int64_t rows_per_block, int num_heads, int64_t num_blocks, int64_t s_b,
                   int64_t s_h, int64_t s_t, int box_rows, CUtensorMap* out, int halves) {
  TensorMapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = 5;
  key.dtype = dtype;
  key.swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B;
  key.dims[0] = 64;                        key.box[0] = 64;
  key.dims[1] = (uint64_t)rows_per_block;  key.strides[0] = (uint64_t)s_t * 2;  key.box[1] = (uint32_t)box_rows;
  key.dims[2] = (uint64_t)halves;          key.strides[1] = 128;                key.box[2] = 1;
  key.dims[3] = (uint64_t)num_heads;       key.strides[2] = (uint64_t)s_h * 2;  key.box[3] = 1;
  key.dims[4] = (uint64_t)num_blocks;      key.strides[3] = (uint64_t)s_b * 2;  key.box[4] = 1;
  return get_tensor_map(key, out);
}

}  // namespace mojo
