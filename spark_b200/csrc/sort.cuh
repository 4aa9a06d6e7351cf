// sort.cuh -- the stable sort permutation of csrc/sort.cu for other operators (WindowExec sorts by partition keys ++ order keys)
#pragma once
#include "common.cuh"

namespace sb {

// perm[0..n): row ids in sorted order (stable; SortOrder semantics incl. NULL placement; string keys through dictionary codes)
void sort_permutation_impl(const sb_table *in, const sb_sort_order *orders, int32_t norders, uint32_t *perm, cudaStream_t st);

}  // namespace sb
