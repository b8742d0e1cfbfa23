// Host-side cache of 2-D bf16 CUtensorMaps (row-major (rows, cols) matrices with an element row stride): a map is a
// pure function of (pointer, shape, box, swizzle), encoding costs ~1 us, so maps are memoised per process.
#pragma once
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace pn_tmap {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct Key {
  const void* ptr;
  long long rows;
  int cols, ld, box_cols, box_rows, swizzle;
  bool operator==(const Key& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols &&
           box_rows == o.box_rows && swizzle == o.swizzle;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    return std::hash<const void*>()(k.ptr) ^ (size_t)k.rows * 1000003u ^ (size_t)k.cols * 10007u ^ (size_t)k.ld * 131u ^
           (size_t)k.box_rows * 31u ^ (size_t)k.box_cols * 7u ^ (size_t)k.swizzle;
  }
};

// bf16 (rows, cols) matrix, row stride ld elements; box = box_cols x box_rows; out-of-bounds elements read as zero.
inline int get(const void* base, long long rows, int cols, int ld, int box_cols, int box_rows, CUtensorMapSwizzle swizzle,
               CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key{base, rows, cols, ld, box_cols, box_rows, (int)swizzle};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return PN_OK;
    }
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return PN_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return PN_ERR_CUDA;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return PN_OK;
}

}  // namespace pn_tmap
