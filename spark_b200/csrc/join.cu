// join.cu -- equi-joins on the GPU: hash build / probe replacing BroadcastHashJoinExec, ShuffledHashJoinExec
// and SortMergeJoinExec.
//
// Reference path replaced (citations relative to the reference tree, SQLX = sql/core/src/main/scala/org/
// apache/spark/sql/execution):
//   HashedRelation.apply SQLX/joins/HashedRelation.scala:136-168 (LongToUnsafeRowMap :536 for a single long
//   key, UnsafeHashedRelation :209 over BytesToBytesMap otherwise; duplicate keys chain),
//   key packing HashJoin.rewriteKeyExpr SQLX/joins/HashJoin.scala:891-912 (integral keys totalling <= 8 bytes
//   are shifted into one long -- the same packing is used here),
//   probe loops HashJoin.innerJoin :184, outerJoin :215, semiJoin :275, antiJoin :334;
//   a row with any NULL key never matches (HashJoin.scala:160-172);
//   SortMergeJoinScanner SQLX/joins/SortMergeJoinExec.scala:1213-1360 yields the same multiset.
//
// GPU design (random-access bound; HBM for the probe stream, L2/HBM for the table):
//   build : open addressing, capacity = 2^k >= 2 x build rows, linear probing; a build row claims the first
//           free slot of its probe sequence with one 32-bit atomicCAS on the slot's row id (0xFFFFFFFF = free)
//           and then stores its packed 64-bit key -- duplicates simply occupy several slots; no sentinel key.
//           A slot is 16 bytes {key, row id}, so a probe step is ONE 128-bit load (one 32-byte sector).
//   probe : pass 1 counts matches per streamed row (and remembers the first match), exclusive scan,
//           pass 2 writes (probe row, build row) pairs in streamed-row order; rows with <= 1 match do not walk
//           the table twice.  Output columns are gathered once from both sides.
#include "common.cuh"
#include "primitives.cuh"

struct sb_hash_table {
  sb_table *build = nullptr;        // retained build-side batch (payload gathered at probe time)
  struct Slot { uint64_t key; uint32_t row; uint32_t pad; };   // 16 bytes: one sector-aligned load per probe step
  Slot *slots = nullptr;            // [cap]; row == 0xFFFFFFFF means free
  int64_t cap = 0;
  int32_t nkeys = 0;
  int32_t key_type[4];
  int32_t key_bits[4];
  int32_t key_shift[4];
  cudaStream_t st = nullptr;
};

namespace sb {

constexpr int JOIN_THREADS = 256;
constexpr int JOIN_MAX_KEYS = 4;
constexpr uint32_t FREE_SLOT = 0xFFFFFFFFu;

struct JoinKeys {
  int n;
  const void *data[JOIN_MAX_KEYS];
  const uint8_t *valid[JOIN_MAX_KEYS];
  int32_t type[JOIN_MAX_KEYS];
  int32_t bits[JOIN_MAX_KEYS];
  int32_t shift[JOIN_MAX_KEYS];
};

__device__ __forceinline__ uint64_t join_mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// packed key of a row; false when any key column is NULL (such a row never matches)
__device__ __forceinline__ bool join_key(const JoinKeys &k, int64_t row, uint64_t &out) {
  uint64_t w = 0;
#pragma unroll
  for (int i = 0; i < JOIN_MAX_KEYS; i++) {
    if (i >= k.n) break;
    if (!bit_valid(k.valid[i], row)) return false;
    uint64_t v;
    if (k.type[i] == SB_FLOAT64) {          // join keys are normalised like grouping keys (-0.0 == 0.0, one NaN)
      double d = ((const double *)k.data[i])[row];
      v = d == 0.0 ? 0ull : (uint64_t)double_bits_canonical(d);
    } else if (k.type[i] == SB_FLOAT32) {
      float f = ((const float *)k.data[i])[row];
      v = f == 0.0f ? 0u : (uint32_t)float_bits_canonical(f);
    } else {
      v = (uint64_t)load_i64(k.data[i], k.type[i], row);
      if (k.bits[i] < 64) v &= (1ull << k.bits[i]) - 1;
    }
    w |= v << k.shift[i];
  }
  out = w;
  return true;
}

typedef sb_hash_table::Slot JoinSlot;
__device__ __forceinline__ void load_slot(const JoinSlot *p, uint64_t &key, uint32_t &row) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
  key = v.x;
  row = (uint32_t)v.y;
}

__global__ void __launch_bounds__(JOIN_THREADS) join_build_kernel(JoinKeys k, int64_t n, JoinSlot *__restrict__ slots, int64_t cap) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  uint64_t key;
  if (!join_key(k, row, key)) return;
  uint64_t mask = (uint64_t)cap - 1;
  uint64_t h = join_mix(key) & mask;
  for (;;) {
    if (slots[h].row == FREE_SLOT && atomicCAS(&slots[h].row, FREE_SLOT, (uint32_t)row) == FREE_SLOT) {
      slots[h].key = key;
      return;
    }
    h = (h + 1) & mask;
  }
}

// pass 1: matches per streamed row (join-type adjusted) + first matching build row
__global__ void __launch_bounds__(JOIN_THREADS) join_count_kernel(JoinKeys k, int64_t n, const JoinSlot *__restrict__ slots, int64_t cap,
                                                                  int join_type, int32_t *__restrict__ counts, uint32_t *__restrict__ first,
                                                                  int32_t *__restrict__ block_counts) {
  __shared__ int32_t wsum[JOIN_THREADS / 32];
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = row < n;
  uint64_t key;
  int32_t matches = 0;
  uint32_t f = FREE_SLOT;
  if (in_range && join_key(k, row, key)) {
    uint64_t mask = (uint64_t)cap - 1;
    uint64_t h = join_mix(key) & mask;
    for (;;) {
      uint64_t sk;
      uint32_t r;
      load_slot(&slots[h], sk, r);
      if (r == FREE_SLOT) break;
      if (sk == key) {
        if (matches == 0) f = r;
        matches++;
      }
      h = (h + 1) & mask;
    }
  }
  int32_t c;
  switch (join_type) {
    case SB_JOIN_INNER: c = matches; break;
    case SB_JOIN_LEFT_OUTER: c = matches > 0 ? matches : 1; break;
    case SB_JOIN_LEFT_SEMI: c = matches > 0 ? 1 : 0; break;
    default: c = matches > 0 ? 0 : 1; break;   // anti
  }
  if (!in_range) c = 0;
  if (in_range) {
    first[row] = f;
    counts[row] = c;
  }
  // output rows of this block: the scan that turns counts into offsets runs over blocks, not rows (join_fill_kernel redoes
  // the in-block prefix in shared memory), which saves an 8-byte offset per streamed row and two passes over them
  int32_t t = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t b = 0;
#pragma unroll
    for (int w = 0; w < JOIN_THREADS / 32; w++) b += wsum[w];
    block_counts[blockIdx.x] = b;
  }
}

// pass 2: (probe row, build row) pairs at the scanned offsets
__global__ void __launch_bounds__(JOIN_THREADS) join_fill_kernel(JoinKeys k, int64_t n, const JoinSlot *__restrict__ slots, int64_t cap,
                                                                 int join_type, const int32_t *__restrict__ counts,
                                                                 const int64_t *__restrict__ block_offsets, const uint32_t *__restrict__ first,
                                                                 int64_t *__restrict__ out_probe, int64_t *__restrict__ out_build) {
  __shared__ int32_t wsum[JOIN_THREADS / 32];
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t c = row < n ? counts[row] : 0;
  int32_t x = c;   // inclusive prefix inside the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  int32_t woff = 0;
  for (int w = 0; w < warp; w++) woff += wsum[w];
  if (c == 0) return;
  int64_t o = block_offsets[blockIdx.x] + woff + (x - c);
  uint32_t f = first[row];
  if (c == 1 || join_type == SB_JOIN_LEFT_SEMI || join_type == SB_JOIN_LEFT_ANTI) {
    out_probe[o] = row;
    if (out_build) out_build[o] = (f == FREE_SLOT || join_type >= SB_JOIN_LEFT_SEMI) ? -1 : (int64_t)f;
    return;
  }
  uint64_t key;
  join_key(k, row, key);
  uint64_t mask = (uint64_t)cap - 1;
  uint64_t h = join_mix(key) & mask;
  for (;;) {
    uint64_t sk;
    uint32_t r;
    load_slot(&slots[h], sk, r);
    if (r == FREE_SLOT) break;
    if (sk == key) {
      out_probe[o] = row;
      out_build[o] = r;
      o++;
    }
    h = (h + 1) & mask;
  }
}

static JoinKeys make_join_keys(const sb_table *t, const int32_t *key_cols, int32_t nkeys, const sb_hash_table *ht) {
  SB_REQUIRE(nkeys >= 1 && nkeys <= JOIN_MAX_KEYS, "joins support 1..%d key columns (got %d)", JOIN_MAX_KEYS, nkeys);
  JoinKeys k;
  k.n = nkeys;
  int pos = 0;
  for (int i = 0; i < nkeys; i++) {
    SB_REQUIRE(key_cols[i] >= 0 && key_cols[i] < (int)t->cols.size(), "join key column %d out of range", key_cols[i]);
    const Column &c = t->cols[key_cols[i]];
    if (c.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "string join keys are not supported (dictionary-encode them)");
    k.data[i] = c.d();
    k.valid[i] = c.v();
    k.type[i] = c.type;
    int bits = type_width(c.type) * 8;
    if (ht) {   // the probe side must pack exactly like the build side
      SB_REQUIRE(ht->key_bits[i] == bits, "join key %d: probe width %d bits differs from build width %d bits (cast first)", i, bits,
                 ht->key_bits[i]);
    }
    k.bits[i] = bits;
    k.shift[i] = pos;
    pos += bits;
  }
  if (pos > 64) fail(SB_ERR_UNSUPPORTED, "join keys need %d bits; at most 64 bits of fixed-width keys are packed (HashJoin.rewriteKeyExpr)", pos);
  return k;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_join_build(const sb_table *build, const int32_t *key_cols, int32_t nkeys, sb_stream *s, sb_hash_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(build && key_cols && out, "null argument");
  cudaStream_t st = stream_of(s);
  int64_t n = build->nrows;
  SB_REQUIRE(n < 0xFFFFFFFFll, "build side has too many rows for one relation");
  JoinKeys k = make_join_keys(build, key_cols, nkeys, nullptr);
  sb_hash_table *ht = new sb_hash_table();
  ht->st = st;
  ht->nkeys = nkeys;
  for (int i = 0; i < nkeys; i++) {
    ht->key_type[i] = k.type[i];
    ht->key_bits[i] = k.bits[i];
    ht->key_shift[i] = k.shift[i];
  }
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  ht->cap = cap;
  try {
    SB_CUDA(cudaMallocAsync((void **)&ht->slots, (size_t)cap * sizeof(JoinSlot), st));
    SB_CUDA(cudaMemsetAsync(ht->slots, 0xff, (size_t)cap * sizeof(JoinSlot), st));
    if (n > 0) {
      KernelTimer kt("join_build", st);
      join_build_kernel<<<(unsigned)((n + JOIN_THREADS - 1) / JOIN_THREADS), JOIN_THREADS, 0, st>>>(k, n, ht->slots, cap);
      SB_LAUNCH_CHECK();
    }
    ht->build = const_cast<sb_table *>(build);
    ht->build->refs.fetch_add(1);
  } catch (...) {
    if (ht->slots) cudaFreeAsync(ht->slots, st);
    delete ht;
    throw;
  }
  *out = ht;
  SB_API_END
}

int sb_hash_table_release(sb_hash_table *ht) {
  SB_API_BEGIN
  if (ht) {
    if (ht->slots) cudaFreeAsync(ht->slots, ht->st);
    if (ht->build && ht->build->refs.fetch_sub(1) == 1) table_free(ht->build);
    delete ht;
  }
  SB_API_END
}

int sb_join_probe(const sb_hash_table *ht, const sb_table *probe, const int32_t *key_cols, int32_t nkeys, int32_t join_type,
                  sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(ht && probe && key_cols && out, "null argument");
  SB_REQUIRE(nkeys == ht->nkeys, "probe has %d key columns, the relation was built on %d", nkeys, ht->nkeys);
  SB_REQUIRE(join_type >= SB_JOIN_INNER && join_type <= SB_JOIN_LEFT_ANTI, "unknown join type %d", join_type);
  cudaStream_t st = stream_of(s);
  const int64_t n = probe->nrows;
  JoinKeys k = make_join_keys(probe, key_cols, nkeys, ht);
  const bool pairs = join_type == SB_JOIN_INNER || join_type == SB_JOIN_LEFT_OUTER;
  unsigned nb = (unsigned)((n + JOIN_THREADS - 1) / JOIN_THREADS);
  Scratch counts(n * 4 + 16, st), first(n * 4 + 16, st), block_counts((int64_t)nb * 4 + 16, st), offsets((int64_t)nb * 8 + 16, st), total(8, st);
  if (n > 0) {
    KernelTimer kt("join_probe", st);
    join_count_kernel<<<nb, JOIN_THREADS, 0, st>>>(k, n, ht->slots, ht->cap, join_type, counts.as<int32_t>(), first.as<uint32_t>(),
                                                   block_counts.as<int32_t>());
    SB_LAUNCH_CHECK();
  }
  exclusive_scan_i32_to_i64(block_counts.as<int32_t>(), offsets.as<int64_t>(), nb, total.as<int64_t>(), st);
  int64_t nout = 0;
  SB_CUDA(cudaMemcpyAsync(&nout, total.ptr, 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  Scratch out_probe(nout * 8 + 16, st), out_build(pairs ? nout * 8 + 16 : 0, st);
  if (n > 0 && nout > 0) {
    KernelTimer kt("join_fill", st);
    join_fill_kernel<<<nb, JOIN_THREADS, 0, st>>>(k, n, ht->slots, ht->cap, join_type, counts.as<int32_t>(),
                                                  offsets.as<int64_t>(), first.as<uint32_t>(), out_probe.as<int64_t>(),
                                                  pairs ? out_build.as<int64_t>() : nullptr);
    SB_LAUNCH_CHECK();
  }
  sb_table *left = gather_table(probe, out_probe.as<int64_t>(), nout, false, st);
  if (!pairs) {
    *out = left;
  } else {
    sb_table *right = nullptr;
    try {
      right = gather_table(ht->build, out_build.as<int64_t>(), nout, join_type == SB_JOIN_LEFT_OUTER, st);
    } catch (...) {
      table_free(left);
      throw;
    }
    for (auto &c : right->cols) left->cols.push_back(c);   // output = streamed columns ++ build columns (HashJoin.scala:55-70)
    right->cols.clear();
    table_free(right);
    *out = left;
  }
  SB_CUDA(cudaStreamSynchronize(st));
  SB_API_END
}

}  // extern "C"
