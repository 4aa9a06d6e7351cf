// radix.cuh -- stable LSD radix sort of (u64 key, u32 value) pairs (csrc/radix.cu)
#pragma once
#include "common.cuh"

namespace sb {

// Sorts the pairs ascending by key, stably; the sorted VALUES end up in `vals` (the keys buffer is scratch afterwards).
// Bytes on which all keys agree are skipped like RadixSort.java:213-236 does.  Returns the number of scatter passes.
// keys_alt (optional, n entries): the second key buffer of the ping-pong, so the caller can keep the sorted keys; *sorted_keys then
// says which of keys / keys_alt holds them.
int radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, cudaStream_t st, uint64_t *keys_alt = nullptr, uint64_t **sorted_keys = nullptr);

}  // namespace sb
