// primitives.cuh -- device-wide building blocks shared by the operators: exclusive scan,
// row gather (take) with validity, mask compaction.
#pragma once
#include "common.cuh"

namespace sb {

// out[i] = sum_{j<i} in[j]; total (if non-null, device pointer) = sum of all.  in may alias out.
void exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t st);
void exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, int32_t *total_dev, cudaStream_t st);
// widening scan: int32 counts -> int64 offsets
void exclusive_scan_i32_to_i64(const int32_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t st);

// Gather rows of every column of `in` at positions idx[0..nout) (idx < 0 -> NULL row, used for
// outer-join padding).  Result columns carry validity when the source has it or idx may be negative.
// skip_col >= 0: that column of the result is left empty for the caller to fill (SortExec rebuilds its key column from the sorted keys)
sb_table *gather_table(const sb_table *in, const int64_t *idx_dev, int64_t nout, bool idx_may_be_negative,
                       cudaStream_t st, int skip_col = -1);
Column gather_column(const Column &c, const int64_t *idx_dev, int64_t nout, bool idx_may_be_negative, cudaStream_t st);

// indices of rows whose mask byte is non-zero, in row order; returns count (synchronizes the stream)
int64_t compact_mask(const uint8_t *mask_dev, int64_t n, int64_t *out_idx_dev, cudaStream_t st);

// same without the host round trip: the count lands in *total_dev (device); the caller provides the per-tile scratch
// (compact_tiles(n) int32 counts and as many int64 offsets)
inline int64_t compact_tiles(int64_t n) { return (n + 4095) / 4096; }
void compact_mask_async(const uint8_t *mask_dev, int64_t n, int64_t *out_idx_dev, int32_t *tile_counts_scratch,
                        int64_t *tile_offsets_scratch, int64_t *total_dev, cudaStream_t st);

// validity bitmap <-> one byte per row (bitmaps cannot be sliced or concatenated at arbitrary row offsets)
void bitmap_to_bytes(const uint8_t *bitmap_dev, int64_t n, uint8_t *out_dev, cudaStream_t st);   // bitmap == nullptr -> all ones
void bytes_to_bitmap(const uint8_t *bytes_dev, int64_t n, uint32_t *bitmap_dev, cudaStream_t st);

// fills idx[i] = begin + i
void iota_i64(int64_t *out, int64_t n, int64_t begin, cudaStream_t st);

}  // namespace sb
