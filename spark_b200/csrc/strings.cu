// strings.cu -- string keys as ORDER-PRESERVING dictionary codes.
//
// The fixed-width operators (hash aggregate, hash join, radix sort) pack their keys into 64-bit words; a UTF8String key
// (common/unsafe/src/main/java/org/apache/spark/unsafe/types/UTF8String.java) does not fit.  What Spark does with such keys is
// compare bytes (grouping / join equality = binary equality, ordering = UTF8String.compareTo -> ByteArray.compareBinary:
// unsigned bytes, a proper prefix sorts first).  Both relations survive the map string -> rank of the string among the
// column's distinct values, so a string key column becomes an int32 column of ranks ("codes") plus one dictionary column:
//     equal strings <=> equal codes,   a < b <=> code(a) < code(b),   NULL stays NULL.
// sb_dictionary_encode builds that pair on the device; the operators run on the codes unchanged; sb_dictionary_decode turns
// the key columns of the result back into strings.  sb_dictionary_lookup encodes the OTHER side of a join against the build
// side's dictionary (a string that is not in it gets -1, which equals no build code).
//
//   distinct values : open-addressing table of (32-bit tag, representative row) words; a probe that meets its own tag compares
//                     the actual bytes, so two different strings never share a code whatever their hashes do
//   ordering        : the (few) distinct strings are sorted by LSD radix passes over 4-byte big-endian chunks, most
//                     significant chunk first, each round refining the groups the previous rounds could not separate
//                     (key = group rank : chunk), then by length (a proper prefix sorts first)
#include <memory>
#include "common.cuh"
#include "primitives.cuh"
#include "radix.cuh"
#include "strings.cuh"

namespace sb {

constexpr int STR_THREADS = 256;
constexpr uint64_t DICT_EMPTY = ~0ull;

struct StrView {
  const uint8_t *bytes;
  const int32_t *offs;
  const uint8_t *valid;
};
static StrView view_of(const Column &c) { return StrView{(const uint8_t *)c.d(), c.o(), c.v()}; }

__device__ __forceinline__ uint64_t str_hash(const uint8_t *__restrict__ p, int len) {   // FNV-1a + a finaliser; slot choice only
  uint64_t h = 14695981039346656037ull;
  for (int i = 0; i < len; i++) h = (h ^ p[i]) * 1099511628211ull;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}
__device__ __forceinline__ bool str_equal(const uint8_t *__restrict__ a, int la, const uint8_t *__restrict__ b, int lb) {
  if (la != lb) return false;
  for (int i = 0; i < la; i++)
    if (a[i] != b[i]) return false;
  return true;
}

// find-or-insert every non-NULL row's string; slot word = tag << 32 | representative row (of `in` itself).
// pos_out[row] = slot of the row's string (-1 for NULL rows).  ctl[0] = table too small, ctl[1] = distinct strings so far.
__global__ void __launch_bounds__(STR_THREADS) dict_insert_kernel(StrView in, int64_t n, uint64_t *__restrict__ slots, uint64_t mask,
                                                                  int32_t *__restrict__ pos_out, int32_t *__restrict__ ctl, int32_t max_fill) {
  const int64_t row = (int64_t)blockIdx.x * STR_THREADS + threadIdx.x;
  if (row >= n) return;
  if (!bit_valid(in.valid, row)) {
    pos_out[row] = -1;
    return;
  }
  const int32_t o = in.offs[row], len = in.offs[row + 1] - o;
  const uint8_t *p = in.bytes + o;
  const uint64_t h = str_hash(p, len);
  const uint64_t tag = h >> 32, mine = (tag << 32) | (uint64_t)(uint32_t)row;
  uint64_t pos = h & mask;
  for (uint64_t step = 0; step <= mask; step++) {
    uint64_t cur = *(volatile uint64_t *)&slots[pos];
    if (cur == DICT_EMPTY) {
      const uint64_t old = atomicCAS((unsigned long long *)&slots[pos], (unsigned long long)DICT_EMPTY, (unsigned long long)mine);
      if (old == DICT_EMPTY) {
        if (atomicAdd(&ctl[1], 1) >= max_fill) ctl[0] = 1;
        pos_out[row] = (int32_t)pos;
        return;
      }
      cur = old;
    }
    if ((cur >> 32) == tag) {
      const int64_t rep = (int64_t)(cur & 0xFFFFFFFFull);
      const int32_t ro = in.offs[rep];
      if (str_equal(p, len, in.bytes + ro, in.offs[rep + 1] - ro)) {
        pos_out[row] = (int32_t)pos;
        return;
      }
    }
    pos = (pos + 1) & mask;
    if ((step & 15) == 15 && *(volatile int32_t *)&ctl[0]) return;   // somebody found the table too small: the pass is void
  }
  ctl[0] = 1;
}

// dense ids for the occupied slots (any order: the sort below fixes the final codes)
__global__ void dict_compact_kernel(const uint64_t *__restrict__ slots, int64_t cap, int64_t *__restrict__ reps, int32_t *__restrict__ slot_id,
                                    int32_t *__restrict__ counter) {
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= cap) return;
  const uint64_t w = slots[pos];
  if (w == DICT_EMPTY) return;
  const int32_t id = atomicAdd(counter, 1);
  reps[id] = (int64_t)(w & 0xFFFFFFFFull);
  slot_id[pos] = id;
}

__global__ void dict_codes_kernel(const int32_t *__restrict__ pos, int64_t n, const int32_t *__restrict__ slot_id, const uint32_t *__restrict__ rank,
                                  int32_t *__restrict__ codes) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int32_t p = pos[row];
  codes[row] = p < 0 ? 0 : (int32_t)rank[slot_id[p]];
}

// table over an existing dictionary (distinct strings): slot word = tag << 32 | dictionary index
__global__ void dict_build_kernel(StrView dict, int64_t d, uint64_t *__restrict__ slots, uint64_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  const int32_t o = dict.offs[i], len = dict.offs[i + 1] - o;
  const uint64_t h = str_hash(dict.bytes + o, len);
  const uint64_t mine = ((h >> 32) << 32) | (uint64_t)(uint32_t)i;
  uint64_t pos = h & mask;
  for (;;) {
    if (slots[pos] == DICT_EMPTY && atomicCAS((unsigned long long *)&slots[pos], (unsigned long long)DICT_EMPTY, (unsigned long long)mine) == DICT_EMPTY) return;
    pos = (pos + 1) & mask;
  }
}
__global__ void __launch_bounds__(STR_THREADS) dict_find_kernel(StrView in, int64_t n, StrView dict, const uint64_t *__restrict__ slots, uint64_t mask,
                                                                int32_t *__restrict__ codes) {
  const int64_t row = (int64_t)blockIdx.x * STR_THREADS + threadIdx.x;
  if (row >= n) return;
  int32_t code = -1;
  if (bit_valid(in.valid, row)) {
    const int32_t o = in.offs[row], len = in.offs[row + 1] - o;
    const uint8_t *p = in.bytes + o;
    const uint64_t h = str_hash(p, len);
    const uint64_t tag = h >> 32;
    uint64_t pos = h & mask;
    for (;;) {
      const uint64_t cur = __ldg(&slots[pos]);
      if (cur == DICT_EMPTY) break;
      if ((cur >> 32) == tag) {
        const int64_t i = (int64_t)(cur & 0xFFFFFFFFull);
        const int32_t ro = dict.offs[i];
        if (str_equal(p, len, dict.bytes + ro, dict.offs[i + 1] - ro)) {
          code = (int32_t)i;
          break;
        }
      }
      pos = (pos + 1) & mask;
    }
  } else code = 0;
  codes[row] = code;
}

// ---- sorting the distinct strings ------------------------------------------------------------------------------------------
__global__ void str_maxlen_kernel(const int32_t *__restrict__ offs, int64_t n, int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t len = i < n ? offs[i + 1] - offs[i] : 0;
  len = __reduce_max_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}
// key of the element at sorted position i.  Round 0 (nobody is separated yet): the first 8 bytes big-endian (zero padded).
// Later rounds: the group rank so far, then the next 4 bytes (bytes [8 + 4 (round - 1), + 4)); the final round: rank, then length.
__device__ __forceinline__ uint64_t refine_key(const StrView &s, const uint32_t *__restrict__ seg_of, uint32_t e, int round, int by_length) {
  const int32_t o = s.offs[e], len = s.offs[e + 1] - o;
  if (round == 0 && !by_length) {
    uint64_t chunk = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) chunk = (chunk << 8) | (b < len ? (uint64_t)s.bytes[o + b] : 0ull);
    return chunk;
  }
  uint32_t chunk = 0;
  if (by_length) chunk = (uint32_t)len;
  else {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int at = 8 + (round - 1) * 4 + b;
      chunk = (chunk << 8) | (at < len ? (uint32_t)s.bytes[o + at] : 0u);
    }
  }
  return ((uint64_t)seg_of[e] << 32) | chunk;
}
__global__ void refine_keys_kernel(StrView s, const uint32_t *__restrict__ seg_of, const uint32_t *__restrict__ perm, int64_t n, int round, int by_length,
                                   uint64_t *__restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = refine_key(s, seg_of, perm[i], round, by_length);
}
__global__ void refine_heads_kernel(StrView s, const uint32_t *__restrict__ seg_of, const uint32_t *__restrict__ perm, int64_t n, int round, int by_length,
                                    int32_t *__restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = i == 0 || refine_key(s, seg_of, perm[i], round, by_length) != refine_key(s, seg_of, perm[i - 1], round, by_length);
}
__global__ void refine_assign_kernel(const int32_t *__restrict__ before, const int32_t *__restrict__ head, const uint32_t *__restrict__ perm, int64_t n,
                                     uint32_t *__restrict__ seg_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg_of[perm[i]] = (uint32_t)(before[i] + head[i] - 1);   // dense rank of the group position i belongs to
}
__global__ void iota32_kernel(uint32_t *out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}
__global__ void inverse_rank_kernel(const uint32_t *__restrict__ rank, int64_t d, int64_t *__restrict__ order) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < d) order[rank[e]] = e;
}
// codes -> gather indices; NULL codes and codes outside the dictionary (-1 of sb_dictionary_lookup) decode to NULL
__global__ void widen_i32_kernel(const int32_t *__restrict__ in, const uint8_t *__restrict__ valid, int64_t n, int64_t d, int64_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bit_valid(valid, i) && in[i] >= 0 && in[i] < d ? (int64_t)in[i] : -1;
}

static inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

// rank_of[e] = position of distinct string e in UTF8String order (the strings must be pairwise different)
static void rank_distinct_strings(const Column &c, uint32_t *rank_of, cudaStream_t st) {
  const int64_t d = c.length;
  if (d == 0) return;
  SB_CUDA(cudaMemsetAsync(rank_of, 0, (size_t)d * 4, st));
  if (d == 1) return;
  Scratch maxlen(4, st);
  SB_CUDA(cudaMemsetAsync(maxlen.ptr, 0, 4, st));
  str_maxlen_kernel<<<blocks_for(d), 256, 0, st>>>(c.o(), d, maxlen.as<int32_t>());
  SB_LAUNCH_CHECK();
  int32_t ml = 0;
  SB_CUDA(cudaMemcpyAsync(&ml, maxlen.ptr, 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  const int rounds = ml <= 8 ? 1 : 1 + (ml - 8 + 3) / 4;   // 8 bytes in the first round, 4 in every later one
  const StrView s = view_of(c);
  Scratch perm(d * 4 + 16, st), keys(d * 8 + 16, st), head(d * 4 + 16, st), before(d * 4 + 16, st), total(4, st);
  iota32_kernel<<<blocks_for(d), 256, 0, st>>>(perm.as<uint32_t>(), d);
  SB_LAUNCH_CHECK();
  for (int r = 0; r <= rounds; r++) {
    const int by_length = r == rounds;
    refine_keys_kernel<<<blocks_for(d), 256, 0, st>>>(s, rank_of, perm.as<uint32_t>(), d, r, by_length, keys.as<uint64_t>());
    SB_LAUNCH_CHECK();
    radix_sort_pairs(keys.as<uint64_t>(), perm.as<uint32_t>(), d, st);
    refine_heads_kernel<<<blocks_for(d), 256, 0, st>>>(s, rank_of, perm.as<uint32_t>(), d, r, by_length, head.as<int32_t>());
    SB_LAUNCH_CHECK();
    exclusive_scan_i32(head.as<int32_t>(), before.as<int32_t>(), d, total.as<int32_t>(), st);
    refine_assign_kernel<<<blocks_for(d), 256, 0, st>>>(before.as<int32_t>(), head.as<int32_t>(), perm.as<uint32_t>(), d, rank_of);
    SB_LAUNCH_CHECK();
    int32_t groups = 0;
    SB_CUDA(cudaMemcpyAsync(&groups, total.ptr, 4, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (groups == d) return;   // every string has its own rank
  }
  fail(SB_ERR_INVALID, "rank_distinct_strings: the dictionary holds equal strings");
}

static const Column &string_column(const sb_table *t, int32_t col) {
  SB_REQUIRE(t, "null table");
  SB_REQUIRE(col >= 0 && col < (int)t->cols.size(), "column %d out of range", col);
  SB_REQUIRE(t->cols[col].type == SB_STRING, "column %d is not a string column", col);
  return t->cols[col];
}

// codes share the input's validity bitmap
static void share_validity(const Column &src, Column &codes) {
  if (src.validity) {
    buffer_retain(src.validity);
    codes.validity = src.validity;
    codes.null_count = src.null_count;
  }
}

void dictionary_encode(const Column &c, cudaStream_t st, Column &codes_out, Column &dict_out) {
  const int64_t n = c.length;
  SB_REQUIRE(c.type == SB_STRING, "dictionary_encode takes a string column");
  SB_REQUIRE(n < (1ll << 31), "dictionary encoding takes fewer than 2^31 rows per call");
  KernelTimer kt("dictionary_encode", st);
  const StrView in = view_of(c);
  Scratch pos(n * 4 + 16, st), ctl(8, st);
  int64_t cap = 1 << 16, cap_max = 1 << 16;
  while (cap_max < 2 * n) cap_max <<= 1;
  std::unique_ptr<Scratch> slots;
  int32_t hctl[2] = {0, 0};
  // Size the table from a sample instead of discovering an overflow pass by pass (each failed pass hashes every row): the distinct
  // count of the first 2^20 rows says whether the column is low-cardinality (table = 16 x what the sample saw) or not (table for
  // "every row its own value")
  constexpr int64_t kSample = 1 << 20;
  if (n > 4 * kSample) {
    Scratch sslots(2 * kSample * 8, st);
    SB_CUDA(cudaMemsetAsync(sslots.ptr, 0xff, (size_t)2 * kSample * 8, st));
    SB_CUDA(cudaMemsetAsync(ctl.ptr, 0, 8, st));
    dict_insert_kernel<<<blocks_for(kSample, STR_THREADS), STR_THREADS, 0, st>>>(in, kSample, sslots.as<uint64_t>(), (uint64_t)2 * kSample - 1, pos.as<int32_t>(),
                                                                                ctl.as<int32_t>(), (int32_t)(2 * kSample));
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemcpyAsync(hctl, ctl.ptr, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const int64_t seen = hctl[1];
    if (seen > kSample / 4) cap = cap_max;
    else
      while (cap < 16 * seen && cap < cap_max) cap <<= 1;
  }
  for (;;) {
    slots.reset(new Scratch(cap * 8, st));
    SB_CUDA(cudaMemsetAsync(slots->ptr, 0xff, (size_t)cap * 8, st));
    SB_CUDA(cudaMemsetAsync(ctl.ptr, 0, 8, st));
    if (n > 0) {
      dict_insert_kernel<<<blocks_for(n, STR_THREADS), STR_THREADS, 0, st>>>(in, n, slots->as<uint64_t>(), (uint64_t)cap - 1, pos.as<int32_t>(),
                                                                            ctl.as<int32_t>(), (int32_t)(cap / 2));
      SB_LAUNCH_CHECK();
    }
    SB_CUDA(cudaMemcpyAsync(hctl, ctl.ptr, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (!hctl[0]) break;
    SB_REQUIRE(cap < cap_max, "dictionary table overflow at its maximum size");
    cap = cap * 16 > cap_max ? cap_max : cap * 16;
  }
  const int64_t d = hctl[1];
  Scratch reps(d * 8 + 16, st), slot_id(cap * 4 + 16, st), counter(4, st), rank(d * 4 + 16, st), order(d * 8 + 16, st);
  SB_CUDA(cudaMemsetAsync(counter.ptr, 0, 4, st));
  dict_compact_kernel<<<blocks_for(cap), 256, 0, st>>>(slots->as<uint64_t>(), cap, reps.as<int64_t>(), slot_id.as<int32_t>(), counter.as<int32_t>());
  SB_LAUNCH_CHECK();
  Column distinct = gather_column(c, reps.as<int64_t>(), d, false, st);   // the representatives are non-NULL rows
  if (distinct.validity) {
    buffer_release(distinct.validity);
    distinct.validity = nullptr;
  }
  distinct.null_count = 0;
  Column dict, cc;
  bool have_dict = false;
  try {
    rank_distinct_strings(distinct, rank.as<uint32_t>(), st);
    if (d > 0) {   // dictionary[rank[e]] = distinct[e]: gather through the inverse permutation
      inverse_rank_kernel<<<blocks_for(d), 256, 0, st>>>(rank.as<uint32_t>(), d, order.as<int64_t>());
      SB_LAUNCH_CHECK();
    }
    dict = gather_column(distinct, order.as<int64_t>(), d, false, st);
    have_dict = true;
    cc = column_alloc(SB_INT32, 0, n, false, st);
    if (n > 0) {
      dict_codes_kernel<<<blocks_for(n), 256, 0, st>>>(pos.as<int32_t>(), n, slot_id.as<int32_t>(), rank.as<uint32_t>(), (int32_t *)cc.data->ptr);
      SB_LAUNCH_CHECK();
    }
  } catch (...) {
    column_release(distinct);
    if (have_dict) column_release(dict);
    throw;
  }
  column_release(distinct);
  share_validity(c, cc);
  codes_out = cc;
  dict_out = dict;
}

Column dictionary_lookup(const Column &c, const Column &dc, cudaStream_t st) {
  SB_REQUIRE(c.type == SB_STRING && dc.type == SB_STRING, "dictionary_lookup takes string columns");
  SB_REQUIRE(!dc.validity || dc.null_count == 0, "a dictionary holds no NULLs");
  const int64_t n = c.length, d = dc.length;
  KernelTimer kt("dictionary_encode", st);
  int64_t cap = 1024;
  while (cap < 2 * d) cap <<= 1;
  Scratch slots(cap * 8, st);
  SB_CUDA(cudaMemsetAsync(slots.ptr, 0xff, (size_t)cap * 8, st));
  const StrView dv = view_of(dc);
  if (d > 0) {
    dict_build_kernel<<<blocks_for(d), 256, 0, st>>>(dv, d, slots.as<uint64_t>(), (uint64_t)cap - 1);
    SB_LAUNCH_CHECK();
  }
  Column cc = column_alloc(SB_INT32, 0, n, false, st);
  if (n > 0) {
    dict_find_kernel<<<blocks_for(n, STR_THREADS), STR_THREADS, 0, st>>>(view_of(c), n, dv, slots.as<uint64_t>(), (uint64_t)cap - 1, (int32_t *)cc.data->ptr);
    SB_LAUNCH_CHECK();
  }
  share_validity(c, cc);
  return cc;
}

Column dictionary_decode(const Column &cc, const Column &dc, cudaStream_t st) {
  SB_REQUIRE(cc.type == SB_INT32, "dictionary codes are an int32 column");
  SB_REQUIRE(dc.type == SB_STRING, "a dictionary is a string column");
  const int64_t n = cc.length;
  Scratch idx(n * 8 + 16, st);
  if (n > 0) {
    widen_i32_kernel<<<blocks_for(n), 256, 0, st>>>((const int32_t *)cc.d(), cc.v(), n, dc.length, idx.as<int64_t>());
    SB_LAUNCH_CHECK();
  }
  return gather_column(dc, idx.as<int64_t>(), n, true, st);
}

EncodedView::~EncodedView() {
  if (view && view->refs.fetch_sub(1) == 1) table_free(view);   // a relation with wide keys keeps the view alive
  for (auto &d : dictionaries) column_release(d);
}

void encode_string_columns(const sb_table *t, const std::vector<int> &cols, const std::vector<const Column *> *given, cudaStream_t st, EncodedView &out) {
  out.view = table_new(t->nrows);
  for (auto &c : t->cols) out.view->cols.push_back(column_share(c));
  for (size_t i = 0; i < cols.size(); i++) {
    const int col = cols[i];
    if (out.dictionary_of(col)) continue;
    SB_REQUIRE(col >= 0 && col < (int)t->cols.size() && t->cols[col].type == SB_STRING, "column %d is not a string column", col);
    Column codes, dict;
    if (given) {
      codes = dictionary_lookup(t->cols[col], *(*given)[i], st);
      dict = column_share(*(*given)[i]);
    } else {
      dictionary_encode(t->cols[col], st, codes, dict);
    }
    column_release(out.view->cols[col]);
    out.view->cols[col] = codes;
    out.cols.push_back(col);
    out.dictionaries.push_back(dict);
  }
}

}  // namespace sb

using namespace sb;

static sb_table *one_column_table(Column c) {
  sb_table *t = table_new(c.length);
  t->cols.push_back(c);
  return t;
}

extern "C" {

int sb_dictionary_encode(const sb_table *t, int32_t col, sb_stream *s, sb_table **out_codes, sb_table **out_dictionary) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(out_codes && out_dictionary, "null argument");
  Column codes, dict;
  dictionary_encode(string_column(t, col), stream_of(s), codes, dict);
  *out_codes = one_column_table(codes);
  *out_dictionary = one_column_table(dict);
  SB_API_END
}

int sb_dictionary_lookup(const sb_table *t, int32_t col, const sb_table *dictionary, sb_stream *s, sb_table **out_codes) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(out_codes, "null argument");
  *out_codes = one_column_table(dictionary_lookup(string_column(t, col), string_column(dictionary, 0), stream_of(s)));
  SB_API_END
}

int sb_dictionary_decode(const sb_table *codes, int32_t col, const sb_table *dictionary, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(codes && out, "null argument");
  SB_REQUIRE(col >= 0 && col < (int)codes->cols.size(), "column %d out of range", col);
  *out = one_column_table(dictionary_decode(codes->cols[col], string_column(dictionary, 0), stream_of(s)));
  SB_API_END
}

}  // extern "C"
