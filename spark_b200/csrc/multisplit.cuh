// multisplit.cuh -- the stable single-pass multisplit shared by the hash partitioner (partition.cu) and the
// LSD radix sort passes (sort.cu).
#pragma once
#include "common.cuh"

namespace sb {

struct PartGeometry {
  int nblocks;
  int64_t chunk;   // rows per histogram column (a multisplit tile for nbuckets <= 256)
};
// big_tiles: 8192-row tiles (wide rows into many buckets) instead of 4096-row tiles; only meaningful for nbuckets <= 256
PartGeometry part_geometry(int64_t n, int32_t nbuckets, bool big_tiles = false);

struct SplitCol {
  int width;
  const void *src;
  void *dst;
  const uint8_t *src_valid;   // optional source validity bitmap
  uint32_t *dst_valid;        // optional destination bitmap, pre-set to all ones
};

// bucket[i] in [0, nbuckets) for every row and hist[bucket][block] (counts) must have been produced with
// the same geometry.  Scans hist in place, moves every column stably and optionally writes
// perm_out[dest] = source row.  offsets_dev (nbuckets+1 int64, optional) receives the bucket boundaries.
void multisplit_scatter(const int32_t *bucket_dev, uint32_t *hist_dev, int32_t nbuckets, const PartGeometry &g,
                        const SplitCol *cols, int ncols, int64_t n, int64_t *perm_out, int64_t *offsets_dev,
                        cudaStream_t st);

constexpr int MULTISPLIT_MAX_BUCKETS = 16384;   // two-level split: 64 x 256

}  // namespace sb
