// Shared tensor-map encoder with a small cache.  cuTensorMapEncodeTiled costs ~2 us per map and the fused path needs 13 of
// them per call; weights, pooled workspaces and (in serving loops) input buffers keep their addresses between calls, so the
// encoded descriptors are memoised on their full argument list.  Bounded (round-robin replacement), mutex-protected.
#pragma once

#include <cuda.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace merv {

struct TmapKey {
  const void* base;
  unsigned long long dims[5];
  unsigned long long strides[4];
  unsigned box[5];
  int dtype, rank, swizzle;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};

// Encodes (or fetches) a tiled tensor map; `dtype` is a CUtensorMapDataType, `swizzle` a CUtensorMapSwizzle.
int encode_tmap_cached(CUtensorMap* out, int dtype, int rank, const void* base, const unsigned long long* dims,
                       const unsigned long long* strides_bytes, const unsigned* box, int swizzle);

}  // namespace merv
