// Host-side helpers shared by the tensor-core kernels: cuTensorMapEncodeTiled through the runtime's driver entry point, with a cache.
// Encoding a tensor map costs a few microseconds on the host; one convolution needs 3-6 of them and a forward pass launches ~250
// convolutions.  torch's caching allocator hands out the same addresses step after step and weights never move, so the maps are cached by
// (pointer, geometry): after the first step every launch finds its maps ready.
#pragma once
#include <cuda.h>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include "ga_common.cuh"

namespace ga {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
  }
  return fn;
}

struct TmapKey {
  uint64_t base;
  uint64_t dims[4];
  uint64_t strides[3];
  uint32_t box[4];
  uint32_t estr[4];
  uint32_t rank, dtype, swizzle, l2;
};

// -> 0 on success.  Zero fill out of bounds, no interleave.
static inline int encode_tiled_cached(CUtensorMap* tm, CUtensorMapDataType dtype, uint32_t rank, const void* base, const cuuint64_t* dims,
                                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swizzle,
                                      CUtensorMapL2promotion l2) {
  static std::unordered_map<std::string, CUtensorMap> cache;
  static std::mutex mu;
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.base = (uint64_t)(uintptr_t)base; k.rank = rank; k.dtype = (uint32_t)dtype; k.swizzle = (uint32_t)swizzle; k.l2 = (uint32_t)l2;
  for (uint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; }
  for (uint32_t i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
  std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *tm = it->second; return 0; }
  }
  PFN_tmapEncodeTiled enc = get_encode_fn();
  GA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  CUresult r = enc(tm, dtype, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %u, dims %llu %llu %llu %llu, box %u %u %u %u) failed: %d", rank,
           (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
           (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 16384) cache.clear();       // bounded: a long-running process with ever-changing shapes starts over
    cache.emplace(std::move(key), *tm);
  }
  return 0;
}

static inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ga
