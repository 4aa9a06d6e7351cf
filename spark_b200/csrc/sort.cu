// sort.cu -- SortExec / TakeOrderedAndProjectExec on the GPU: stable LSD radix sort, no spill.
//
// Reference path replaced (citations relative to the reference tree):
//   SQLX/SortExec.scala:39 (createSorter :75-107: radix only for ONE sort column of a prefix-sortable type),
//   SortPrefix (sql/catalyst/.../expressions/SortOrder.scala:128-242), PrefixComparators.java:28-182
//   (double prefix: sign-flip, -0.0 -> 0.0, NaN canonical and largest), UnsafeInMemorySorter.java:241-262
//   (NULL-prefix records are parked at the front by swapping the first non-null record to the end) and
//   :348-390 (NULL block emitted first/last per nullsFirst, independent of ASC/DESC),
//   RadixSort.sortKeyPrefixArray (RadixSort.java:178-259: LSD, 8-bit digits, bytes equal in all records are
//   skipped, signed top byte, DESC = reversed bucket walk keeping in-bucket insertion order),
//   TimSort + RowOrdering for everything else (stable), TakeOrderedAndProjectExec SQLX/limit.scala:347-386.
//
// GPU design: every sort column is turned into an order-preserving unsigned 64-bit key (signed -> flip the sign
// bit, double -> PrefixComparators transform, DESC -> bitwise complement), so one ascending stable LSD pass
// implementation covers all four (ASC|DESC) x (NULLS FIRST|LAST) orders with the reference's tie order.
// The key sort itself is csrc/radix.cu (one histogram read, then one "onesweep" kernel per varying byte; constant bytes
// are skipped like the reference does).  Multi-column orders run column by column from the
// last to the first (stable passes compose into the lexicographic order TimSort + RowOrdering produces).
// Payload columns are gathered once at the end.
#include <algorithm>
#include "common.cuh"
#include "multisplit.cuh"
#include "primitives.cuh"
#include "radix.cuh"
#include "strings.cuh"
#include "sort.cuh"

namespace sb {

constexpr int SORT_THREADS = 256;

// order-preserving key of row `row` of a column for an ascending unsigned sort
__device__ __forceinline__ uint64_t sort_key(const void *data, int32_t type, int64_t row, bool desc) {
  uint64_t k;
  switch (type) {
    case SB_FLOAT32: case SB_FLOAT64: {   // DoublePrefixComparator.computePrefix
      double v = type == SB_FLOAT32 ? (double)((const float *)data)[row] : ((const double *)data)[row];
      if (v == 0.0) v = 0.0;
      int64_t b = double_bits_canonical(v);
      k = (uint64_t)b ^ ((uint64_t)(b >> 63) | 0x8000000000000000ull);
      break;
    }
    case SB_BOOL: k = ((const uint8_t *)data)[row] ? 1 : 0; k ^= 0x8000000000000000ull; break;
    default: k = (uint64_t)load_i64(data, type, row) ^ 0x8000000000000000ull; break;   // signed -> unsigned order
  }
  return desc ? ~k : k;
}

// keys[i] = key of row rows[i] (rows == nullptr -> identity); isnull[i] optional
__global__ void __launch_bounds__(SORT_THREADS) make_keys_kernel(const void *data, const uint8_t *valid, int32_t type, int desc,
                                                                 const uint32_t *__restrict__ rows, int64_t n,
                                                                 uint64_t *__restrict__ keys, uint8_t *__restrict__ isnull) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t r = rows ? rows[i] : i;
  bool v = bit_valid(valid, r);
  keys[i] = v ? sort_key(data, type, r, desc != 0) : 0;
  if (isnull) isnull[i] = !v;
}

// 1-bit stable split used for NULL placement in multi-column orders
__global__ void __launch_bounds__(SORT_THREADS) flag_hist_kernel(const uint8_t *__restrict__ flag, int invert, int64_t n,
                                                                 int64_t chunk, int32_t *__restrict__ bucket,
                                                                 uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[2];
  if (threadIdx.x < 2) sh[threadIdx.x] = 0;
  __syncthreads();
  int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t i = begin + threadIdx.x; i < end; i += SORT_THREADS) {
    int d = (flag[i] != 0) ^ invert;
    bucket[i] = d;
    atomicAdd(&sh[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 2) hist[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}

// stable split of vals by a 0/1 flag (flag[i] belongs to vals[i]): zeros first unless invert
static void stable_split_by_flag(const uint8_t *flag, int invert, uint32_t *vals, int64_t n, cudaStream_t st) {
  if (n <= 1) return;
  PartGeometry g = part_geometry(n, 2);
  Scratch vals2(n * 4 + 16, st), bucket(n * 4 + 16, st), hist((int64_t)2 * g.nblocks * 4 + 16, st);
  flag_hist_kernel<<<g.nblocks, SORT_THREADS, 0, st>>>(flag, invert, n, g.chunk, bucket.as<int32_t>(), hist.as<uint32_t>());
  SB_LAUNCH_CHECK();
  SplitCol col = {4, vals, vals2.ptr, nullptr, nullptr};
  multisplit_scatter(bucket.as<int32_t>(), hist.as<uint32_t>(), 2, g, &col, 1, n, nullptr, nullptr, st);
  SB_CUDA(cudaMemcpyAsync(vals, vals2.ptr, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
}

// ---- UnsafeInMemorySorter.insertRecord replay (radix path with NULLs) ----------------------------------
// Rows from the first non-null row f on form a "tape": a non-null row appends itself, a NULL row re-appends
// the record at the head of the queue (tape[k] for the k-th such NULL).  After all inserts the non-null
// records sit in tape[r .. r+m) (r = NULLs after f, m = non-null count).  ptr[j] starts as j (push) or k
// (copy of tape[k]) and is resolved by pointer jumping.
__global__ void tape_init_kernel(const uint8_t *__restrict__ isnull, const int64_t *__restrict__ nulls_before, int64_t f,
                                 int64_t len, uint32_t *__restrict__ ptr) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  int64_t row = f + j;
  ptr[j] = isnull[row] ? (uint32_t)(nulls_before[row] - f) : (uint32_t)j;
}
__global__ void tape_jump_kernel(uint32_t *ptr, int64_t len, int *changed) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  uint32_t p = ptr[j];
  uint32_t q = ptr[p];
  if (q != p) {
    ptr[j] = q;
    *changed = 1;
  }
}
__global__ void tape_finish_kernel(const uint32_t *__restrict__ ptr, int64_t f, int64_t r, int64_t m, uint32_t *__restrict__ order) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < m) order[q] = (uint32_t)(f + ptr[r + q]);
}
__global__ void u8_to_i32_kernel(const uint8_t *__restrict__ in, int64_t n, int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 1 : 0;
}
__global__ void iota_u32_kernel(uint32_t *out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}
__global__ void null_rows_kernel(const uint8_t *__restrict__ isnull, const int64_t *__restrict__ nulls_before, int64_t n,
                                 uint32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && isnull[i]) out[nulls_before[i]] = (uint32_t)i;
}
__global__ void u32_to_i64_kernel(const uint32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

// The order-preserving key of an integer-typed column is a bijection of the value (sign flip, complement for DESC), so the sorted
// column IS the sorted keys mapped back -- a sequential pass instead of a random gather through the permutation.
static bool key_is_bijective(int32_t type) {
  return type == SB_INT8 || type == SB_INT16 || type == SB_INT32 || type == SB_INT64 || type == SB_DATE32 || type == SB_TIMESTAMP || type == SB_DECIMAL64;
}
__global__ void __launch_bounds__(256) unkey_kernel(const uint64_t *__restrict__ keys, int64_t n, int desc, int width, void *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t k = keys[i];
  if (desc) k = ~k;
  k ^= 0x8000000000000000ull;
  switch (width) {
    case 1: ((uint8_t *)out)[i] = (uint8_t)k; break;
    case 2: ((uint16_t *)out)[i] = (uint16_t)k; break;
    case 4: ((uint32_t *)out)[i] = (uint32_t)k; break;
    default: ((uint64_t *)out)[i] = k; break;
  }
}

// RangePartitioner.getPartition (core/.../Partitioner.scala:241-260): partition = #bounds the key is strictly greater than.
// Keys and bounds are compared as (null rank, order-preserving key): NULLs sort before everything when nulls_first, after
// everything otherwise, exactly as the sort itself orders them.  Bounds are few (numPartitions - 1): binary search in shared memory.
__global__ void __launch_bounds__(SORT_THREADS) range_pid_hist_kernel(const void *data, const uint8_t *valid, int32_t type, int desc,
                                                                      int nulls_first, const void *bdata, const uint8_t *bvalid, int32_t nbounds,
                                                                      int64_t n, int64_t chunk, int32_t *__restrict__ bucket,
                                                                      uint32_t *__restrict__ hist) {
  extern __shared__ uint64_t rp_smem[];
  uint64_t *bkey = rp_smem;                               // [nbounds]
  uint8_t *bnull = (uint8_t *)(bkey + nbounds);           // [nbounds]
  uint32_t *sh = (uint32_t *)(((uintptr_t)(bnull + nbounds) + 3) & ~(uintptr_t)3);   // [nbounds + 1]
  for (int i = threadIdx.x; i < nbounds; i += SORT_THREADS) {
    bool v = bit_valid(bvalid, i);
    bkey[i] = v ? sort_key(bdata, type, i, desc != 0) : 0;
    bnull[i] = !v;
  }
  for (int i = threadIdx.x; i <= nbounds; i += SORT_THREADS) sh[i] = 0;
  __syncthreads();
  int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t i = begin + threadIdx.x; i < end; i += SORT_THREADS) {
    const bool v = bit_valid(valid, i);
    const uint64_t k = v ? sort_key(data, type, i, desc != 0) : 0;
    // rank of a value: NULL is 0 (first) or 2 (last), non-null is 1
    const int kr = v ? 1 : (nulls_first ? 0 : 2);
    int lo = 0, hi = nbounds;                            // count of bounds strictly less than the key
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      const int br = bnull[mid] ? (nulls_first ? 0 : 2) : 1;
      const bool less = br < kr || (br == kr && kr == 1 && bkey[mid] < k);
      if (less) lo = mid + 1; else hi = mid;
    }
    bucket[i] = lo;
    atomicAdd(&sh[lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= nbounds; i += SORT_THREADS) hist[(int64_t)i * gridDim.x + blockIdx.x] = sh[i];
}

// ---- radix select for TakeOrderedAndProject: which keys can be among the k smallest? -------------------------------------------
// 256-bin histogram of the byte below the `bits` already fixed high bits, over the rows whose key starts with `prefix`
__global__ void __launch_bounds__(SORT_THREADS) select_hist_kernel(const void *data, int32_t type, int desc, int64_t n, uint64_t prefix, int bits,
                                                                   uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 56 - bits;
  for (int64_t i = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * SORT_THREADS) {
    const uint64_t k = sort_key(data, type, i, desc != 0);
    if (bits == 0 || (k >> (64 - bits)) == (prefix >> (64 - bits))) atomicAdd(&sh[(k >> shift) & 0xff], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void __launch_bounds__(SORT_THREADS) select_mask_kernel(const void *data, int32_t type, int desc, int64_t n, uint64_t hi, uint8_t *__restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x;
  if (i < n) mask[i] = sort_key(data, type, i, desc != 0) <= hi;
}

static bool radix_eligible(int32_t type) { return type != SB_STRING; }

// writes the sorted permutation (uint32 row ids) into perm (n entries)
void sort_permutation_impl(const sb_table *in, const sb_sort_order *orders, int32_t norders, uint32_t *perm, cudaStream_t st, SortedKeys *sk) {
  const int64_t n = in->nrows;
  SB_REQUIRE(n < (1ll << 32), "tables of 2^32 rows or more must be sorted in chunks");
  SB_REQUIRE(norders >= 1 && orders, "sort needs at least one order");
  std::vector<int> scols;
  for (int k = 0; k < norders; k++) {
    SB_REQUIRE(orders[k].col >= 0 && orders[k].col < (int)in->cols.size(), "sort column %d out of range", orders[k].col);
    if (in->cols[orders[k].col].type == SB_DECIMAL128) fail(SB_ERR_UNSUPPORTED, "decimal(p > 18) sort keys are not supported");
    if (!radix_eligible(in->cols[orders[k].col].type)) scols.push_back(orders[k].col);
  }
  if (n == 0) return;
  if (!scols.empty()) {
    // string sort keys: their order-preserving dictionary codes sort exactly like UTF8String.compareTo sorts the strings
    // (csrc/strings.cu), NULLs stay NULLs -- so the radix path below runs on the codes
    EncodedView ev;
    encode_string_columns(in, scols, nullptr, st, ev);
    sort_permutation_impl(ev.view, orders, norders, perm, st);
    return;
  }
  if (norders == 1) {
    // ---- the reference's radix path --------------------------------------------------------------
    const Column &c = in->cols[orders[0].col];
    const bool desc = !orders[0].ascending, nulls_first = orders[0].nulls_first != 0;
    if (!c.validity && sk && key_is_bijective(c.type)) {   // the caller wants the sorted keys back (sb_sort: the key column of the result)
      sk->a.reset(new Scratch(n * 8 + 16, st));
      sk->b.reset(new Scratch(n * 8 + 16, st));
      iota_u32_kernel<<<nblk(n), 256, 0, st>>>(perm, n);
      SB_LAUNCH_CHECK();
      make_keys_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c.d(), nullptr, c.type, desc, nullptr, n, sk->a->as<uint64_t>(), nullptr);
      SB_LAUNCH_CHECK();
      radix_sort_pairs(sk->a->as<uint64_t>(), perm, n, st, sk->b->as<uint64_t>(), &sk->sorted);
      sk->col = orders[0].col;
      sk->desc = desc;
      return;
    }
    Scratch keys(n * 8 + 16, st);
    if (!c.validity) {
      iota_u32_kernel<<<nblk(n), 256, 0, st>>>(perm, n);
      SB_LAUNCH_CHECK();
      make_keys_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c.d(), nullptr, c.type, desc, nullptr, n, keys.as<uint64_t>(), nullptr);
      SB_LAUNCH_CHECK();
      radix_sort_pairs(keys.as<uint64_t>(), perm, n, st);
      return;
    }
    Scratch isnull(n + 16, st), flags(n * 4 + 16, st), nulls_before(n * 8 + 16, st), total(8, st);
    make_keys_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c.d(), c.v(), c.type, desc, nullptr, n, keys.as<uint64_t>(), isnull.as<uint8_t>());
    SB_LAUNCH_CHECK();
    u8_to_i32_kernel<<<nblk(n), 256, 0, st>>>(isnull.as<uint8_t>(), n, flags.as<int32_t>());
    SB_LAUNCH_CHECK();
    exclusive_scan_i32_to_i64(flags.as<int32_t>(), nulls_before.as<int64_t>(), n, total.as<int64_t>(), st);
    int64_t nnull = 0;
    SB_CUDA(cudaMemcpyAsync(&nnull, total.ptr, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const int64_t m = n - nnull;
    uint32_t *null_part = nulls_first ? perm : perm + m;
    uint32_t *sorted_part = nulls_first ? perm + nnull : perm;
    if (nnull > 0) {   // NULL block keeps arrival order
      null_rows_kernel<<<nblk(n), 256, 0, st>>>(isnull.as<uint8_t>(), nulls_before.as<int64_t>(), n, null_part);
      SB_LAUNCH_CHECK();
    }
    if (m == 0) return;
    if (nnull == 0) {
      iota_u32_kernel<<<nblk(n), 256, 0, st>>>(sorted_part, n);
      SB_LAUNCH_CHECK();
      radix_sort_pairs(keys.as<uint64_t>(), sorted_part, n, st);
      return;
    }
    // first non-null row f: found on the host from the null prefix counts (nulls_before[i] == i up to f)
    // -> f = number of leading NULL rows; binary search over the monotone predicate nulls_before[i] == i
    int64_t f;
    {
      int64_t lo = 0, hi = n;   // invariant: rows < lo are leading NULLs
      while (lo < hi) {
        int64_t mid = (lo + hi) / 2, nb = 0;
        uint8_t isn = 0;
        SB_CUDA(cudaMemcpyAsync(&nb, nulls_before.as<int64_t>() + mid, 8, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaMemcpyAsync(&isn, isnull.as<uint8_t>() + mid, 1, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        if (nb == mid && isn) lo = mid + 1; else hi = mid;
      }
      f = lo;
    }
    const int64_t len = n - f, r = nnull - f;
    Scratch ptr(len * 4 + 16, st), changed(4, st);
    tape_init_kernel<<<nblk(len), 256, 0, st>>>(isnull.as<uint8_t>(), nulls_before.as<int64_t>(), f, len, ptr.as<uint32_t>());
    SB_LAUNCH_CHECK();
    for (;;) {
      SB_CUDA(cudaMemsetAsync(changed.ptr, 0, 4, st));
      tape_jump_kernel<<<nblk(len), 256, 0, st>>>(ptr.as<uint32_t>(), len, changed.as<int>());
      SB_LAUNCH_CHECK();
      int ch = 0;
      SB_CUDA(cudaMemcpyAsync(&ch, changed.ptr, 4, cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      if (!ch) break;
    }
    tape_finish_kernel<<<nblk(m), 256, 0, st>>>(ptr.as<uint32_t>(), f, r, m, sorted_part);
    SB_LAUNCH_CHECK();
    Scratch keys_nn(m * 8 + 16, st);
    make_keys_kernel<<<nblk(m), SORT_THREADS, 0, st>>>(c.d(), nullptr, c.type, desc, sorted_part, m, keys_nn.as<uint64_t>(), nullptr);
    SB_LAUNCH_CHECK();
    radix_sort_pairs(keys_nn.as<uint64_t>(), sorted_part, m, st);
    return;
  }
  // ---- multi-column: stable passes from the last sort column to the first ---------------------------------
  iota_u32_kernel<<<nblk(n), 256, 0, st>>>(perm, n);
  SB_LAUNCH_CHECK();
  Scratch keys(n * 8 + 16, st), isnull(n + 16, st);
  for (int k = norders - 1; k >= 0; k--) {
    const Column &c = in->cols[orders[k].col];
    make_keys_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c.d(), c.v(), c.type, !orders[k].ascending, perm, n, keys.as<uint64_t>(),
                                                       c.validity ? isnull.as<uint8_t>() : nullptr);
    SB_LAUNCH_CHECK();
    radix_sort_pairs(keys.as<uint64_t>(), perm, n, st);
    if (c.validity) {   // NULLs of this column first or last, keeping the order established so far
      make_keys_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c.d(), c.v(), c.type, 0, perm, n, keys.as<uint64_t>(), isnull.as<uint8_t>());
      SB_LAUNCH_CHECK();
      stable_split_by_flag(isnull.as<uint8_t>(), orders[k].nulls_first ? 1 : 0, perm, n, st);
    }
  }
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_sort_permutation(const sb_table *in, const sb_sort_order *orders, int32_t norders, sb_stream *s, int64_t *out_perm_device) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out_perm_device, "null argument");
  cudaStream_t st = stream_of(s);
  int64_t n = in->nrows;
  Scratch perm(n * 4 + 16, st);
  sort_permutation_impl(in, orders, norders, perm.as<uint32_t>(), st);
  if (n > 0) {
    u32_to_i64_kernel<<<nblk(n), 256, 0, st>>>(perm.as<uint32_t>(), n, out_perm_device);
    SB_LAUNCH_CHECK();
  }
  SB_CUDA(cudaStreamSynchronize(st));
  SB_API_END
}

int sb_sort(const sb_table *in, const sb_sort_order *orders, int32_t norders, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out, "null argument");
  cudaStream_t st = stream_of(s);
  int64_t n = in->nrows;
  Scratch perm(n * 4 + 16, st), perm64(n * 8 + 16, st);
  SortedKeys sk;
  sort_permutation_impl(in, orders, norders, perm.as<uint32_t>(), st, &sk);
  if (n > 0 && !(sk.sorted && in->cols.size() == 1)) {
    u32_to_i64_kernel<<<nblk(n), 256, 0, st>>>(perm.as<uint32_t>(), n, perm64.as<int64_t>());
    SB_LAUNCH_CHECK();
  }
  sb_table *t = gather_table(in, perm64.as<int64_t>(), n, false, st, sk.sorted ? sk.col : -1);   // stream-ordered: no host synchronisation needed
  if (sk.sorted) {
    try {
      KernelTimer kt("gather", st);
      const Column &c = in->cols[sk.col];
      Column r = column_alloc(c.type, c.scale, n, false, st);
      r.null_count = 0;
      unkey_kernel<<<nblk(n), 256, 0, st>>>(sk.sorted, n, sk.desc ? 1 : 0, type_width(c.type), r.data->ptr);
      SB_LAUNCH_CHECK();
      t->cols[sk.col] = r;
    } catch (...) {
      table_free(t);
      throw;
    }
  }
  *out = t;
  SB_API_END
}

int sb_range_partition(const sb_table *in, const sb_sort_order *order, const sb_table *bounds, sb_stream *s, sb_table **out,
                       int64_t *out_offsets_host) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && order && bounds && out && out_offsets_host, "null argument");
  SB_REQUIRE(order->col >= 0 && order->col < (int)in->cols.size(), "sort column %d out of range", order->col);
  SB_REQUIRE(bounds->cols.size() == 1, "bounds must be a one-column table");
  cudaStream_t st = stream_of(s);
  const Column &c = in->cols[order->col];
  const Column &b = bounds->cols[0];
  SB_REQUIRE(c.type == b.type, "bounds column type %d differs from the sort column type %d", b.type, c.type);
  if (!radix_eligible(c.type)) fail(SB_ERR_UNSUPPORTED, "range partitioning on string columns is not implemented");
  for (auto &col : in->cols)
    if (col.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "range partitioning of tables with string columns is not implemented");
  const int64_t n = in->nrows;
  const int32_t nbounds = (int32_t)bounds->nrows, nparts = nbounds + 1;
  SB_REQUIRE(nparts <= MULTISPLIT_MAX_BUCKETS && nbounds <= 4096, "too many range bounds (%d)", nbounds);
  PartGeometry g = part_geometry(n, nparts);
  Scratch bucket(n * 4 + 16, st), hist((int64_t)nparts * g.nblocks * 4 + 16, st), offs_dev((int64_t)(nparts + 1) * 8, st);
  size_t smem = (size_t)nbounds * 9 + 8 + (size_t)(nbounds + 1) * 4 + 16;
  range_pid_hist_kernel<<<g.nblocks, SORT_THREADS, smem, st>>>(c.d(), c.v(), c.type, !order->ascending, order->nulls_first, b.d(), b.v(),
                                                               nbounds, n, g.chunk, bucket.as<int32_t>(), hist.as<uint32_t>());
  SB_LAUNCH_CHECK();
  sb_table *t = table_new(n);
  try {
    std::vector<SplitCol> sc;
    for (auto &col : in->cols) {
      Column r = column_alloc(col.type, col.scale, n, col.validity != nullptr, st);
      if (r.validity) SB_CUDA(cudaMemsetAsync(r.validity->ptr, 0xff, (size_t)bitmap_alloc_bytes(n), st));
      r.null_count = col.null_count;
      t->cols.push_back(r);
      sc.push_back({type_width(col.type), col.d(), r.data->ptr, col.v(), r.validity ? (uint32_t *)r.validity->ptr : nullptr});
    }
    multisplit_scatter(bucket.as<int32_t>(), hist.as<uint32_t>(), nparts, g, sc.data(), (int)sc.size(), n, nullptr, offs_dev.as<int64_t>(), st);
    SB_CUDA(cudaMemcpyAsync(out_offsets_host, offs_dev.ptr, (size_t)(nparts + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// ---- RangePartitioner, sampling half (core/src/main/scala/org/apache/spark/Partitioner.scala:175-320, 334-357) -----------------------
// sketch(): every input partition contributes up to sampleSizePerPartition keys drawn uniformly without replacement plus its row
// count; the weight of a candidate is n / sample.length (:229).  The reference draws with a reservoir seeded from the RDD id
// (XORShiftRandom: parity unpinned, like the round-robin start); here row j of the sample is a jittered-stratified draw
// floor((j + u_j) * n / k), which gives every row the same inclusion probability k / n.
__global__ void range_sample_idx_kernel(int64_t n, int64_t k, uint64_t seed, int64_t *__restrict__ idx) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  uint64_t x = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(j + 1);
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
  int64_t r = (int64_t)(((double)j + u) * (double)n / (double)k);
  idx[j] = r < 0 ? 0 : (r >= n ? n - 1 : r);
}
__global__ void fill_f32_kernel(float *out, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

int sb_range_sample(const sb_table *in, const sb_sort_order *order, int64_t sample_size, uint64_t seed, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && order && out && sample_size > 0, "bad argument");
  SB_REQUIRE(order->col >= 0 && order->col < (int)in->cols.size(), "sort column %d out of range", order->col);
  cudaStream_t st = stream_of(s);
  const int64_t n = in->nrows, k = n < sample_size ? n : sample_size;
  const Column &c = in->cols[order->col];
  if (!radix_eligible(c.type)) fail(SB_ERR_UNSUPPORTED, "range partitioning on string columns is not implemented");
  Scratch idx(k * 8 + 16, st);
  if (k > 0) {
    if (k == n) iota_i64(idx.as<int64_t>(), k, 0, st);
    else {
      range_sample_idx_kernel<<<nblk(k), 256, 0, st>>>(n, k, seed, idx.as<int64_t>());
      SB_LAUNCH_CHECK();
    }
  }
  sb_table *t = table_new(k);
  try {
    t->cols.push_back(gather_column(c, idx.as<int64_t>(), k, false, st));
    Column w = column_alloc(SB_FLOAT32, 0, k, false, st);
    t->cols.push_back(w);
    if (k > 0) {
      fill_f32_kernel<<<nblk(k), 256, 0, st>>>((float *)w.data->ptr, k, (float)((double)n / (double)k));   // (n.toDouble / sample.length).toFloat
      SB_LAUNCH_CHECK();
    }
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// determineBounds (:357-388): candidates sorted by key; walk the cumulative weight in steps of sumWeights / partitions and take
// the candidate that crosses each step as a bound, skipping keys equal to the previous bound.  `sample` = (key, weight float32)
// rows of every input partition (after the all-gather in a multi-GPU job); returns a one-column table of <= partitions - 1 bounds.
int sb_range_determine_bounds(const sb_table *sample, const sb_sort_order *order, int32_t num_partitions, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(sample && order && out && sample->cols.size() == 2 && sample->cols[1].type == SB_FLOAT32, "sample must be (key, float32 weight)");
  SB_REQUIRE(num_partitions >= 0, "bad partition count");
  cudaStream_t st = stream_of(s);
  const Column &c = sample->cols[0];
  const int64_t m = sample->nrows;
  const int partitions = (int)std::min<int64_t>(num_partitions, m);    // math.min(partitions, candidates.size)
  std::vector<int64_t> picked;
  if (m > 0 && partitions > 1) {
    Scratch keys(m * 8 + 16, st), isnull(m + 16, st);
    make_keys_kernel<<<nblk(m), SORT_THREADS, 0, st>>>(c.d(), c.v(), c.type, !order->ascending, nullptr, m, keys.as<uint64_t>(), isnull.as<uint8_t>());
    SB_LAUNCH_CHECK();
    std::vector<uint64_t> hk((size_t)m);
    std::vector<uint8_t> hn((size_t)m);
    std::vector<float> hw((size_t)m);
    SB_CUDA(cudaMemcpyAsync(hk.data(), keys.ptr, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(hn.data(), isnull.ptr, (size_t)m, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(hw.data(), sample->cols[1].d(), (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const int null_rank = order->nulls_first ? 0 : 2;
    auto rank_of = [&](int64_t i) { return hn[i] ? null_rank : 1; };
    auto less = [&](int64_t a, int64_t b) {
      const int ra = rank_of(a), rb = rank_of(b);
      if (ra != rb) return ra < rb;
      return ra == 1 && hk[a] < hk[b];
    };
    std::vector<int64_t> ord((size_t)m);
    for (int64_t i = 0; i < m; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), less);                    // candidates.sortBy(_._1)
    double sum = 0;
    for (int64_t i = 0; i < m; i++) sum += (double)hw[i];
    const double step = sum / partitions;
    double cum = 0, target = step;
    int64_t prev = -1;
    for (int64_t i = 0, j = 0; i < m && j < partitions - 1; i++) {
      const int64_t cand = ord[i];
      cum += (double)hw[cand];
      if (cum >= target) {
        if (prev < 0 || less(prev, cand)) {                            // skip duplicate values
          picked.push_back(cand);
          target += step;
          j++;
          prev = cand;
        }
      }
    }
  }
  const int64_t nb = (int64_t)picked.size();
  Scratch idx(nb * 8 + 16, st);
  if (nb > 0) {
    SB_CUDA(cudaMemcpyAsync(idx.ptr, picked.data(), (size_t)nb * 8, cudaMemcpyHostToDevice, st));
  }
  sb_table *t = table_new(nb);
  try {
    t->cols.push_back(gather_column(c, idx.as<int64_t>(), nb, false, st));
    SB_CUDA(cudaStreamSynchronize(st));   // `picked` backs the copy above
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// TakeOrderedAndProjectExec: the reference keeps a bounded priority queue per partition and merges
// (limit.scala:347-386); the result is the first k rows of the total order, which is what this returns.
int sb_top_n(const sb_table *in, const sb_sort_order *orders, int32_t norders, int64_t k, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out && k >= 0, "bad argument");
  SB_REQUIRE(norders >= 1 && orders, "top-n needs at least one order");
  cudaStream_t st = stream_of(s);
  int64_t n = in->nrows;
  int64_t take = k < n ? k : n;
  SB_REQUIRE(orders[0].col >= 0 && orders[0].col < (int)in->cols.size(), "sort column %d out of range", orders[0].col);
  const Column &c0 = in->cols[orders[0].col];
  const sb_table *src = in;
  sb_table *cand = nullptr;
  // Bounded selection instead of a full sort (the reference keeps a k-element heap per partition, Utils.takeOrdered): a radix
  // select on the FIRST sort column finds a key bound such that the rows at or below it number >= k but few; only those rows are
  // sorted (by all the sort columns, stably), so the answer is the first k rows of the full stable sort.
  if (take > 0 && n > 65536 && take * 8 <= n && !c0.validity && radix_eligible(c0.type)) {
    const int desc = !orders[0].ascending;
    Scratch hist(256 * 4, st);
    uint32_t h[256];
    uint64_t prefix = 0;
    int bits = 0;
    int64_t below = 0, k_rem = take, in_bucket = n;
    const int64_t budget = std::max<int64_t>(4 * take, 65536);
    const int grid = grid_for(n, SORT_THREADS * 8, rt().num_sms * 8);
    while (bits < 64 && below + in_bucket > budget) {
      SB_CUDA(cudaMemsetAsync(hist.ptr, 0, 256 * 4, st));
      select_hist_kernel<<<grid, SORT_THREADS, 0, st>>>(c0.d(), c0.type, desc, n, prefix, bits, hist.as<uint32_t>());
      SB_LAUNCH_CHECK();
      SB_CUDA(cudaMemcpyAsync(h, hist.ptr, 256 * 4, cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      int64_t cum = 0;
      int b = 0;
      for (; b < 255; b++) {
        if (cum + h[b] >= k_rem) break;
        cum += h[b];
      }
      below += cum;
      k_rem -= cum;
      in_bucket = h[b];
      prefix |= (uint64_t)b << (56 - bits);
      bits += 8;
    }
    if (bits > 0 && below + in_bucket < n) {
      const uint64_t hi = bits >= 64 ? prefix : (prefix | ((1ull << (64 - bits)) - 1));
      const int64_t ncand = below + in_bucket;
      Scratch mask(n + 16, st), f32(compact_tiles(n) * 4 + 16, st), pos(compact_tiles(n) * 8 + 16, st), total(8, st);
      select_mask_kernel<<<nblk(n), SORT_THREADS, 0, st>>>(c0.d(), c0.type, desc, n, hi, mask.as<uint8_t>());
      SB_LAUNCH_CHECK();
      // the candidate count is known from the histograms: no read-back
      Scratch idx_full(n * 8 + 16, st);
      compact_mask_async(mask.as<uint8_t>(), n, idx_full.as<int64_t>(), f32.as<int32_t>(), pos.as<int64_t>(), total.as<int64_t>(), st);
      cand = gather_table(in, idx_full.as<int64_t>(), ncand, false, st);   // candidates in row order: the stable sort below keeps tie order
      src = cand;
      n = ncand;
    }
  }
  try {
    Scratch perm(n * 4 + 16, st), perm64(take * 8 + 16, st);
    sort_permutation_impl(src, orders, norders, perm.as<uint32_t>(), st);
    if (take > 0) {
      u32_to_i64_kernel<<<nblk(take), 256, 0, st>>>(perm.as<uint32_t>(), take, perm64.as<int64_t>());
      SB_LAUNCH_CHECK();
    }
    *out = gather_table(src, perm64.as<int64_t>(), take, false, st);
  } catch (...) {
    if (cand) table_free(cand);
    throw;
  }
  if (cand) table_free(cand);
  SB_API_END
}

}  // extern "C"
