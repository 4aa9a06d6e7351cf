// sort.cuh -- the stable sort permutation of csrc/sort.cu for other operators (WindowExec sorts by partition keys ++ order keys)
#pragma once
#include <memory>
#include "common.cuh"

namespace sb {

// perm[0..n): row ids in sorted order (stable; SortOrder semantics incl. NULL placement; string keys through dictionary codes)
// sk (optional): for ONE integer-typed, NULL-free sort column the sorted 64-bit keys are kept (sk->sorted != nullptr) so the caller can
// rebuild that column of the result from them instead of gathering it.
struct SortedKeys {
  std::unique_ptr<Scratch> a, b;
  uint64_t *sorted = nullptr;
  int col = -1;
  bool desc = false;
};
void sort_permutation_impl(const sb_table *in, const sb_sort_order *orders, int32_t norders, uint32_t *perm, cudaStream_t st, SortedKeys *sk = nullptr);

}  // namespace sb
