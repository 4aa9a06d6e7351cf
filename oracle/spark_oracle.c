/*
 * spark_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the algorithms on apache/spark's shuffle / sort /
 * hash-aggregate / hash-join hot path, written from the reference's behaviour
 * (file:line citations are relative to /root/reference).  It is the checker
 * for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (spark_b200/csrc) never links, includes or calls anything in here.
 *
 * Parity status: the reference itself is pure JVM and cannot run in this
 * image (no JDK), so this restatement is pinned against the reference's own
 * known-answer vectors (tests/test_oracle_golden.py):
 *   - Murmur3_x86_32Suite.java:38-53, hash.scala:844-845, builtin.py:15372-15390
 *   - RadixSortSuite-style differential checks against a stable comparison sort
 *   - PrefixComparatorsSuite double-ordering cases
 *   - InnerJoinSuite / OuterJoinSuite / ExistenceJoinSuite literal fixtures
 * NOT pinned (no golden in the reference): TPC-H answers, round-robin start
 * offset (XORShiftRandom depends on scala/java stdlib code outside the tree).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- column descriptor (Arrow layout; same numeric type ids as include/spark_b200.h) */
enum {
  SO_BOOL = 1, SO_INT8 = 2, SO_INT16 = 3, SO_INT32 = 4, SO_INT64 = 5,
  SO_FLOAT32 = 6, SO_FLOAT64 = 7, SO_DATE32 = 8, SO_TIMESTAMP = 9,
  SO_DECIMAL64 = 10, SO_STRING = 11
};

typedef struct so_column {
  int32_t type;
  int32_t scale;            /* decimal scale (unused otherwise) */
  int64_t length;
  int64_t null_count;
  const void *data;         /* values (string: byte arena) */
  const uint8_t *validity;  /* Arrow validity bitmap, LSB first; NULL = all valid */
  const int32_t *offsets;   /* string only: length+1 int32 offsets */
} so_column;

static inline int so_valid(const so_column *c, int64_t i) {
  return c->validity == NULL || ((c->validity[i >> 3] >> (i & 7)) & 1);
}

int so_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void so_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* =====================================================================
 * Murmur3_x86_32  (common/unsafe/.../hash/Murmur3_x86_32.java:47-150)
 * ===================================================================== */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t mixK1(uint32_t k1) {           /* :124-129 */
  k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u; return k1;
}
static inline uint32_t mixH1(uint32_t h1, uint32_t k1) { /* :131-136 */
  h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5u + 0xe6546b64u; return h1;
}
static inline uint32_t fmix(uint32_t h1, uint32_t len) { /* :139-147 */
  h1 ^= len; h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
  return h1;
}
int32_t so_murmur3_int(int32_t v, int32_t seed) {        /* hashInt :48-53 */
  return (int32_t)fmix(mixH1((uint32_t)seed, mixK1((uint32_t)v)), 4);
}
int32_t so_murmur3_long(int64_t v, int32_t seed) {       /* hashLong :111-122 */
  uint32_t low = (uint32_t)v, high = (uint32_t)((uint64_t)v >> 32);
  uint32_t h1 = mixH1((uint32_t)seed, mixK1(low));
  h1 = mixH1(h1, mixK1(high));
  return (int32_t)fmix(h1, 8);
}
static uint32_t hashBytesByInt(const uint8_t *p, int len, uint32_t seed) { /* :96-106 */
  uint32_t h1 = seed;
  for (int i = 0; i < len; i += 4) {
    uint32_t w; memcpy(&w, p + i, 4);                    /* little-endian host */
    h1 = mixH1(h1, mixK1(w));
  }
  return h1;
}
/* hashUnsafeBytes :66-78 -- the LEGACY tail: each remaining byte is sign-extended and mixed as
 * its own block.  This is the variant Catalyst uses for strings/binary (hash.scala:733-741). */
int32_t so_murmur3_bytes(const uint8_t *p, int32_t len, int32_t seed) {
  int aligned = len - len % 4;
  uint32_t h1 = hashBytesByInt(p, aligned, (uint32_t)seed);
  for (int i = aligned; i < len; i++) {
    int32_t half = (int8_t)p[i];
    h1 = mixH1(h1, mixK1((uint32_t)half));
  }
  return (int32_t)fmix(h1, (uint32_t)len);
}
/* hashUnsafeWords :59-64 (UnsafeRow.hashCode, len % 8 == 0) */
int32_t so_murmur3_words(const uint8_t *p, int32_t len, int32_t seed) {
  return (int32_t)fmix(hashBytesByInt(p, len, (uint32_t)seed), (uint32_t)len);
}

static inline int64_t double_to_long_bits(double d) {   /* Double.doubleToLongBits: NaN canonical */
  if (d != d) return 0x7ff8000000000000LL;
  int64_t b; memcpy(&b, &d, 8); return b;
}
static inline int32_t float_to_int_bits(float f) {
  if (f != f) return 0x7fc00000;
  int32_t b; memcpy(&b, &f, 4); return b;
}

/* One column folded into a running per-row hash: InterpretedHashFunction.hash
 * (sql/catalyst/.../expressions/hash.scala:707-760); NULL leaves the hash unchanged;
 * chaining = the previous column's hash is the next seed (hash.scala:400-409). */
void so_hash_column(const so_column *c, int64_t n, int32_t *inout_hash) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(c, i)) continue;
    int32_t seed = inout_hash[i], h;
    switch (c->type) {
      case SO_BOOL:  h = so_murmur3_int(((const uint8_t *)c->data)[i] ? 1 : 0, seed); break;
      case SO_INT8:  h = so_murmur3_int(((const int8_t *)c->data)[i], seed); break;
      case SO_INT16: h = so_murmur3_int(((const int16_t *)c->data)[i], seed); break;
      case SO_INT32: case SO_DATE32:
                     h = so_murmur3_int(((const int32_t *)c->data)[i], seed); break;
      case SO_INT64: case SO_TIMESTAMP: case SO_DECIMAL64:   /* decimal p<=18: unscaled long :724-727 */
                     h = so_murmur3_long(((const int64_t *)c->data)[i], seed); break;
      case SO_FLOAT32: {
        float f = ((const float *)c->data)[i];
        h = so_murmur3_int(f == 0.0f ? 0 : float_to_int_bits(f), seed); break;  /* -0.0f -> 0 :718 */
      }
      case SO_FLOAT64: {
        double d = ((const double *)c->data)[i];
        h = so_murmur3_long(d == 0.0 ? 0 : double_to_long_bits(d), seed); break; /* -0.0 -> 0 :720 */
      }
      case SO_STRING: {
        int32_t o = c->offsets[i], l = c->offsets[i + 1] - o;
        h = so_murmur3_bytes((const uint8_t *)c->data + o, l, seed); break;
      }
      default: h = seed;
    }
    inout_hash[i] = h;
  }
}

/* Row hash over key columns, initial seed 42 (hash.scala:887 Murmur3Hash default seed) */
void so_hash_rows(const so_column *cols, int32_t ncols, int64_t n, int32_t seed, int32_t *out_hash) {
  for (int64_t i = 0; i < n; i++) out_hash[i] = seed;
  for (int c = 0; c < ncols; c++) so_hash_column(&cols[c], n, out_hash);
}

/* pmod: sql/api/.../catalyst/util/MathUtils.scala:96-99 */
static inline int32_t pmod_i32(int32_t a, int32_t n) {
  int32_t r = a % n;
  return r < 0 ? (r + n) % n : r;
}
/* HashPartitioning.partitionIdExpression = Pmod(Murmur3Hash(exprs), n) (partitioning.scala:339-341) */
void so_partition_ids(const so_column *cols, int32_t ncols, int64_t n, int32_t nparts, int32_t *out_pid) {
  so_hash_rows(cols, ncols, n, 42, out_pid);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) out_pid[i] = pmod_i32(out_pid[i], nparts);
}

/* Group rows by partition id, keeping arrival order inside each partition -- what the shuffle
 * writers guarantee per map task (BypassMergeSortShuffleWriter appends in arrival order;
 * ShuffleInMemorySorter radix-sorts PackedRecordPointers on the partition bytes only, which is a
 * stable LSD sort: ShuffleInMemorySorter.java:202-225, PackedRecordPointer.java:21-52). */
void so_partition_scatter(const int32_t *pid, int64_t n, int32_t nparts, int64_t *out_perm,
                          int64_t *out_offsets /* nparts+1 */) {
  memset(out_offsets, 0, sizeof(int64_t) * (size_t)(nparts + 1));
  for (int64_t i = 0; i < n; i++) out_offsets[pid[i] + 1]++;
  for (int32_t p = 0; p < nparts; p++) out_offsets[p + 1] += out_offsets[p];
  int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)nparts);
  memcpy(cur, out_offsets, sizeof(int64_t) * (size_t)nparts);
  for (int64_t i = 0; i < n; i++) out_perm[cur[pid[i]]++] = i;
  free(cur);
}

/* Round-robin partition ids (ShuffleExchangeExec.scala:428-442): the position counter is
 * incremented BEFORE use, so row i of a map task goes to (start + 1 + i) mod n.  `start` is
 * XORShiftRandom(mapPartitionId).nextInt(n) in the reference -- PARITY UNPINNED (depends on
 * scala.util.hashing.MurmurHash3 + java.util.Random, both outside the tree); the caller passes it. */
void so_round_robin_ids(int64_t n, int32_t nparts, int32_t start, int32_t *out_pid) {
  for (int64_t i = 0; i < n; i++) out_pid[i] = (int32_t)(((int64_t)start + 1 + i) % nparts);
}

/* =====================================================================
 * Sort prefix + radix sort
 * ===================================================================== */
/* DoublePrefixComparator.computePrefix (PrefixComparators.java:66-80) */
static inline int64_t double_prefix(double v) {
  if (v == 0.0) v = 0.0;                       /* -0.0 -> 0.0 */
  int64_t bits = double_to_long_bits(v);
  int64_t mask = -(int64_t)((uint64_t)bits >> 63) | (int64_t)0x8000000000000000ULL;
  return bits ^ mask;
}
int64_t so_double_prefix(double v) { return double_prefix(v); }

static inline int64_t bswap_prefix(const uint8_t *p, int len) { /* UTF8String.getPrefix: first 8 bytes, big-endian */
  uint64_t r = 0;
  for (int i = 0; i < 8; i++) r = (r << 8) | (i < len ? p[i] : 0);
  return (int64_t)r;
}

/* SortPrefix.eval (SortOrder.scala:128-199): 64-bit prefix per row; isnull[i]=1 for NULL rows and
 * prefix = nullValue (:130-149) for them. */
void so_sort_prefix(const so_column *c, int64_t n, int32_t ascending, int32_t nulls_first,
                    int64_t *out_prefix, uint8_t *out_isnull) {
  int null_smallest = (ascending && nulls_first) || (!ascending && !nulls_first);
  int64_t nullv;
  switch (c->type) {
    case SO_FLOAT32: case SO_FLOAT64: case SO_STRING: nullv = null_smallest ? 0 : -1; break;
    default: nullv = null_smallest ? INT64_MIN : INT64_MAX;
  }
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(c, i)) { out_prefix[i] = nullv; out_isnull[i] = 1; continue; }
    out_isnull[i] = 0;
    switch (c->type) {
      case SO_BOOL:  out_prefix[i] = ((const uint8_t *)c->data)[i] ? 1 : 0; break;
      case SO_INT8:  out_prefix[i] = ((const int8_t *)c->data)[i]; break;
      case SO_INT16: out_prefix[i] = ((const int16_t *)c->data)[i]; break;
      case SO_INT32: case SO_DATE32: out_prefix[i] = ((const int32_t *)c->data)[i]; break;
      case SO_INT64: case SO_TIMESTAMP: case SO_DECIMAL64: out_prefix[i] = ((const int64_t *)c->data)[i]; break;
      case SO_FLOAT32: out_prefix[i] = double_prefix((double)((const float *)c->data)[i]); break;
      case SO_FLOAT64: out_prefix[i] = double_prefix(((const double *)c->data)[i]); break;
      case SO_STRING: {
        int32_t o = c->offsets[i];
        out_prefix[i] = bswap_prefix((const uint8_t *)c->data + o, c->offsets[i + 1] - o); break;
      }
      default: out_prefix[i] = 0;
    }
  }
}

/* RadixSort.transformCountsToOffsets (RadixSort.java:149-169), in records not bytes */
static void counts_to_offsets(int64_t *counts, int64_t n, int desc, int sgn) {
  int start = sgn ? 128 : 0;
  if (desc) {
    int64_t pos = n;
    for (int i = start; i < start + 256; i++) { pos -= counts[i & 0xff]; counts[i & 0xff] = pos; }
  } else {
    int64_t pos = 0;
    for (int i = start; i < start + 256; i++) { int64_t t = counts[i & 0xff]; counts[i & 0xff] = pos; pos += t; }
  }
}

/* RadixSort.sortKeyPrefixArray (RadixSort.java:178-259): LSD over bytes [startByte,endByte] of the
 * prefix; bytes equal in all records are skipped (:213-236); the signed flag applies to the last
 * byte only (:200); desc = reversed bucket walk (:154-159).  (ptr,prefix) pairs move together.
 * ptr/prefix are sorted in place (n records each); returns number of passes executed. */
int32_t so_radix_sort_key_prefix(int64_t *ptr, int64_t *prefix, int64_t n, int32_t start_byte,
                                 int32_t end_byte, int32_t desc, int32_t sgn) {
  if (n <= 0) return 0;
  int64_t *ptr2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  int64_t *pre2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  int64_t *ip = ptr, *ik = prefix, *op = ptr2, *ok = pre2;
  uint64_t bmax = 0, bmin = ~0ULL;
  for (int64_t i = 0; i < n; i++) { bmax |= (uint64_t)prefix[i]; bmin &= (uint64_t)prefix[i]; }
  uint64_t changed = bmin ^ bmax;
  int passes = 0;
  for (int b = start_byte; b <= end_byte; b++) {
    if (((changed >> (b * 8)) & 0xff) == 0) continue;
    int64_t counts[256]; memset(counts, 0, sizeof(counts));
    for (int64_t i = 0; i < n; i++) counts[((uint64_t)ik[i] >> (b * 8)) & 0xff]++;
    counts_to_offsets(counts, n, desc, sgn && b == end_byte);
    for (int64_t i = 0; i < n; i++) {
      int bucket = (int)(((uint64_t)ik[i] >> (b * 8)) & 0xff);
      int64_t d = counts[bucket]++;
      op[d] = ip[i]; ok[d] = ik[i];
    }
    int64_t *t = ip; ip = op; op = t; t = ik; ik = ok; ok = t;
    passes++;
  }
  if (ip != ptr) { memcpy(ptr, ip, sizeof(int64_t) * (size_t)n); memcpy(prefix, ik, sizeof(int64_t) * (size_t)n); }
  free(ptr2); free(pre2);
  return passes;
}

/* Plain 8-byte-record variant: RadixSort.sort (RadixSort.java:43-67) */
int32_t so_radix_sort(int64_t *a, int64_t n, int32_t start_byte, int32_t end_byte, int32_t desc, int32_t sgn) {
  if (n <= 0) return 0;
  int64_t *dummy = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  memcpy(dummy, a, sizeof(int64_t) * (size_t)n);
  int32_t p = so_radix_sort_key_prefix(dummy, a, n, start_byte, end_byte, desc, sgn);
  free(dummy);
  return p;
}

/* UnsafeInMemorySorter with radix support: insertRecord (UnsafeInMemorySorter.java:241-262) keeps
 * NULL-prefix records in a block at the front by swapping the first non-null record to the end;
 * getSortedIterator (:348-390) radix-sorts only the non-null tail and chains the NULL block first
 * or last per nullsFirst, independent of ASC/DESC.  Output: permutation of row indices. */
void so_inmemory_sorter_radix(const int64_t *prefix, const uint8_t *isnull, int64_t n, int32_t desc,
                              int32_t sgn, int32_t nulls_first, int64_t *out_perm) {
  int64_t *ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
  int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
  int64_t pos = 0, nb = 0;
  for (int64_t i = 0; i < n; i++) {
    if (isnull[i]) {
      ptr[pos] = ptr[nb]; key[pos] = key[nb]; pos++;   /* swap a non-null record forward */
      ptr[nb] = i; key[nb] = prefix[i]; nb++;
    } else { ptr[pos] = i; key[pos] = prefix[i]; pos++; }
  }
  so_radix_sort_key_prefix(ptr + nb, key + nb, pos - nb, 0, 7, desc, sgn);
  int64_t o = 0;
  if (nulls_first) { for (int64_t i = 0; i < nb; i++) out_perm[o++] = ptr[i]; }
  for (int64_t i = nb; i < pos; i++) out_perm[o++] = ptr[i];
  if (!nulls_first) { for (int64_t i = 0; i < nb; i++) out_perm[o++] = ptr[i]; }
  free(ptr); free(key);
}

/* ---- full-row ordering (RowOrdering / InterpretedOrdering; used by TimSort when the sort is not
 * radix-eligible: multi-column, strings.  Stable.)  Spark double ordering: NaN == NaN, NaN is
 * largest, -0.0 == 0.0 (SQLOrderingUtil.compareDoubles). */
typedef struct so_sort_order { int32_t col; int32_t ascending; int32_t nulls_first; int32_t pad; } so_sort_order;

static int cmp_value(const so_column *c, int64_t a, int64_t b) {
  switch (c->type) {
    case SO_BOOL: { uint8_t x = ((const uint8_t *)c->data)[a] != 0, y = ((const uint8_t *)c->data)[b] != 0; return (x > y) - (x < y); }
    case SO_INT8: { int8_t x = ((const int8_t *)c->data)[a], y = ((const int8_t *)c->data)[b]; return (x > y) - (x < y); }
    case SO_INT16: { int16_t x = ((const int16_t *)c->data)[a], y = ((const int16_t *)c->data)[b]; return (x > y) - (x < y); }
    case SO_INT32: case SO_DATE32: { int32_t x = ((const int32_t *)c->data)[a], y = ((const int32_t *)c->data)[b]; return (x > y) - (x < y); }
    case SO_INT64: case SO_TIMESTAMP: case SO_DECIMAL64: { int64_t x = ((const int64_t *)c->data)[a], y = ((const int64_t *)c->data)[b]; return (x > y) - (x < y); }
    case SO_FLOAT32: case SO_FLOAT64: {
      double x = c->type == SO_FLOAT32 ? ((const float *)c->data)[a] : ((const double *)c->data)[a];
      double y = c->type == SO_FLOAT32 ? ((const float *)c->data)[b] : ((const double *)c->data)[b];
      if (x == y) return 0;
      int xn = x != x, yn = y != y;
      if (xn || yn) return xn - yn;           /* NaN largest, NaN == NaN */
      return (x > y) - (x < y);
    }
    case SO_STRING: {                           /* UTF8String.compareTo: unsigned bytewise, then length */
      int32_t oa = c->offsets[a], la = c->offsets[a + 1] - oa, ob = c->offsets[b], lb = c->offsets[b + 1] - ob;
      int m = la < lb ? la : lb;
      int r = memcmp((const uint8_t *)c->data + oa, (const uint8_t *)c->data + ob, (size_t)m);
      if (r) return r < 0 ? -1 : 1;
      return (la > lb) - (la < lb);
    }
  }
  return 0;
}

static int cmp_rows(const so_column *cols, const so_sort_order *ord, int32_t nord, int64_t a, int64_t b) {
  for (int k = 0; k < nord; k++) {
    const so_column *c = &cols[ord[k].col];
    int va = so_valid(c, a), vb = so_valid(c, b);
    if (!va && !vb) continue;
    if (!va) return ord[k].nulls_first ? -1 : 1;
    if (!vb) return ord[k].nulls_first ? 1 : -1;
    int r = cmp_value(c, a, b);
    if (r) return ord[k].ascending ? r : -r;
  }
  return 0;
}

/* stable merge sort of row indices by the full row ordering */
void so_sort_rows(const so_column *cols, const so_sort_order *ord, int32_t nord, int64_t n, int64_t *out_perm) {
  int64_t *a = out_perm, *b = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
  for (int64_t i = 0; i < n; i++) a[i] = i;
  for (int64_t w = 1; w < n; w *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * w) {
      int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int64_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) b[k++] = cmp_rows(cols, ord, nord, a[j], a[i]) < 0 ? a[j++] : a[i++];
      while (i < mid) b[k++] = a[i++];
      while (j < hi) b[k++] = a[j++];
    }
    int64_t *t = a; a = b; b = t;
  }
  if (a != out_perm) { memcpy(out_perm, a, sizeof(int64_t) * (size_t)n); free(a); } else free(b);
}

/* =====================================================================
 * Grouping: UnsafeFixedWidthAggregationMap / BytesToBytesMap restated
 * (BytesToBytesMap.java:604-643 safeLookup: pos = hash & mask, triangular probing pos += step++;
 *  append-only, doubling growAndRehash :1131; key equality = byte equality of the key row incl.
 *  null bits, so NULL == NULL for grouping; float keys are normalised first by
 *  NormalizeFloatingNumbers.scala:67 (-0.0 -> 0.0, NaN canonical)).
 * Keys here are fixed-width (<= 8 bytes per column) -- each row's key is packed to
 * nkeys x (int64 value, null flag) exactly like an UnsafeRow of fixed-width fields.
 * ===================================================================== */
static inline int64_t key_word(const so_column *c, int64_t i) {
  switch (c->type) {
    case SO_BOOL: return ((const uint8_t *)c->data)[i] ? 1 : 0;
    case SO_INT8: return ((const int8_t *)c->data)[i];
    case SO_INT16: return ((const int16_t *)c->data)[i];
    case SO_INT32: case SO_DATE32: return ((const int32_t *)c->data)[i];
    case SO_INT64: case SO_TIMESTAMP: case SO_DECIMAL64: return ((const int64_t *)c->data)[i];
    case SO_FLOAT32: { float f = ((const float *)c->data)[i]; if (f == 0.0f) f = 0.0f; return float_to_int_bits(f); }
    case SO_FLOAT64: { double d = ((const double *)c->data)[i]; if (d == 0.0) d = 0.0; return double_to_long_bits(d); }
  }
  return 0;
}

typedef struct { int64_t *keys; uint8_t *nulls; int64_t cap, size; int64_t *slot_gid; int32_t *slot_hash; int nk; } so_map;

static void map_init(so_map *m, int nk, int64_t cap) {
  m->nk = nk; m->cap = cap; m->size = 0;
  m->slot_gid = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
  m->slot_hash = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
  for (int64_t i = 0; i < cap; i++) m->slot_gid[i] = -1;
  m->keys = NULL; m->nulls = NULL;
}
static void map_grow(so_map *m) {
  int64_t ncap = m->cap * 2;
  int64_t *ng = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncap);
  int32_t *nh = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncap);
  for (int64_t i = 0; i < ncap; i++) ng[i] = -1;
  for (int64_t i = 0; i < m->cap; i++) {
    if (m->slot_gid[i] < 0) continue;
    int64_t pos = (uint32_t)m->slot_hash[i] & (ncap - 1), step = 1;
    while (ng[pos] >= 0) { pos = (pos + step) & (ncap - 1); step++; }
    ng[pos] = m->slot_gid[i]; nh[pos] = m->slot_hash[i];
  }
  free(m->slot_gid); free(m->slot_hash);
  m->slot_gid = ng; m->slot_hash = nh; m->cap = ncap;
}

/* Assigns each row a dense group id in first-seen order; out_first_row[g] = first row of group g.
 * Returns number of groups.  out_first_row must hold n entries. */
int64_t so_group_ids(const so_column *keys, int32_t nkeys, int64_t n, int64_t *out_gid, int64_t *out_first_row) {
  so_map m; map_init(&m, nkeys, 64);
  int64_t *gk = (int64_t *)malloc(sizeof(int64_t) * (size_t)((n ? n : 1) * (nkeys ? nkeys : 1)));
  uint8_t *gn = (uint8_t *)malloc((size_t)((n ? n : 1) * (nkeys ? nkeys : 1)));
  int64_t kw[64]; uint8_t kn[64];
  for (int64_t i = 0; i < n; i++) {
    int32_t h = 42;
    for (int k = 0; k < nkeys; k++) {
      kn[k] = !so_valid(&keys[k], i);
      kw[k] = kn[k] ? 0 : key_word(&keys[k], i);
      h = so_murmur3_long(kw[k] ^ (kn[k] ? 0x5bd1e995 : 0), h);   /* hash of the key row words */
    }
    int64_t pos = (uint32_t)h & (m.cap - 1), step = 1, g = -1;
    for (;;) {
      int64_t sg = m.slot_gid[pos];
      if (sg < 0) break;
      if (m.slot_hash[pos] == h &&
          memcmp(gk + sg * nkeys, kw, sizeof(int64_t) * (size_t)nkeys) == 0 &&
          memcmp(gn + sg * nkeys, kn, (size_t)nkeys) == 0) { g = sg; break; }
      pos = (pos + step) & (m.cap - 1); step++;
    }
    if (g < 0) {
      g = m.size++;
      memcpy(gk + g * nkeys, kw, sizeof(int64_t) * (size_t)nkeys);
      memcpy(gn + g * nkeys, kn, (size_t)nkeys);
      m.slot_gid[pos] = g; m.slot_hash[pos] = h;
      out_first_row[g] = i;
      if (m.size * 2 > m.cap) map_grow(&m);                   /* load factor 0.5 (BytesToBytesMap :880) */
    }
    out_gid[i] = g;
  }
  int64_t ng = m.size;
  free(m.slot_gid); free(m.slot_hash); free(gk); free(gn);
  return ng;
}

/* ---- aggregate buffer algebra, sequential in row order (one task's update loop).
 * Sum.scala:113-178 (NULL inputs skipped; all-NULL group -> NULL; long wraps in non-ANSI mode),
 * Average.scala:80-135 (sum as double, count of non-null), Count.scala:94-105, Min/Max. */
void so_agg_sum_i64(const int64_t *gid, const so_column *v, int64_t n, int64_t ng, int64_t *out, uint8_t *out_valid) {
  memset(out, 0, sizeof(int64_t) * (size_t)ng); memset(out_valid, 0, (size_t)ng);
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(v, i)) continue;
    int64_t x = key_word(v, i);
    out[gid[i]] = (int64_t)((uint64_t)out[gid[i]] + (uint64_t)x); out_valid[gid[i]] = 1;
  }
}
static inline double col_f64(const so_column *c, int64_t i) {
  switch (c->type) {
    case SO_FLOAT32: return ((const float *)c->data)[i];
    case SO_FLOAT64: return ((const double *)c->data)[i];
    default: return (double)key_word(c, i);
  }
}
void so_agg_sum_f64(const int64_t *gid, const so_column *v, int64_t n, int64_t ng, double *out, uint8_t *out_valid) {
  memset(out, 0, sizeof(double) * (size_t)ng); memset(out_valid, 0, (size_t)ng);
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(v, i)) continue;
    out[gid[i]] += col_f64(v, i); out_valid[gid[i]] = 1;
  }
}
/* count of non-null rows of v per group; v == NULL -> count(*) */
void so_agg_count(const int64_t *gid, const so_column *v, int64_t n, int64_t ng, int64_t *out) {
  memset(out, 0, sizeof(int64_t) * (size_t)ng);
  for (int64_t i = 0; i < n; i++) if (v == NULL || so_valid(v, i)) out[gid[i]]++;
}
/* min/max: integers exact; doubles with Spark ordering (NaN largest) */
void so_agg_minmax_i64(const int64_t *gid, const so_column *v, int64_t n, int64_t ng, int32_t is_max, int64_t *out, uint8_t *out_valid) {
  memset(out, 0, sizeof(int64_t) * (size_t)ng); memset(out_valid, 0, (size_t)ng);
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(v, i)) continue;
    int64_t x = key_word(v, i), g = gid[i];
    if (!out_valid[g] || (is_max ? x > out[g] : x < out[g])) out[g] = x;
    out_valid[g] = 1;
  }
}
static inline int dcmp(double x, double y) {
  if (x == y) return 0;
  int xn = x != x, yn = y != y;
  if (xn || yn) return xn - yn;
  return (x > y) - (x < y);
}
void so_agg_minmax_f64(const int64_t *gid, const so_column *v, int64_t n, int64_t ng, int32_t is_max, double *out, uint8_t *out_valid) {
  memset(out, 0, sizeof(double) * (size_t)ng); memset(out_valid, 0, (size_t)ng);
  for (int64_t i = 0; i < n; i++) {
    if (!so_valid(v, i)) continue;
    double x = col_f64(v, i); int64_t g = gid[i];
    if (!out_valid[g] || (is_max ? dcmp(x, out[g]) > 0 : dcmp(x, out[g]) < 0)) out[g] = x;
    out_valid[g] = 1;
  }
}

/* =====================================================================
 * Equi hash join (HashJoin.scala:184-400; HashedRelation.scala:136-168)
 * A row with any NULL key never matches (inner/outer/semi; HashJoin.scala:160-172 anyNull check).
 * Join types: 0 inner, 1 left outer (streamed side kept), 2 left semi, 3 left anti.
 * Output pairs in streamed(probe)-row order; matches of one probe row in build insertion order
 * (chain order is unspecified in the reference; results are compared as multisets).
 * Two-call protocol: out_probe == NULL -> returns the pair count only.
 * build_idx = -1 for unmatched rows of an outer/anti join; semi/anti emit build_idx = -1.
 * ===================================================================== */
int64_t so_hash_join(const so_column *bkeys, const so_column *pkeys, int32_t nkeys, int64_t nb, int64_t np,
                     int32_t join_type, int64_t *out_probe, int64_t *out_build) {
  /* build: group build rows by key (first-seen id), chain rows per group in insertion order */
  int64_t cap = 64; while (cap < 2 * nb) cap *= 2;
  int64_t *slot = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
  int32_t *shash = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
  for (int64_t i = 0; i < cap; i++) slot[i] = -1;
  int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb ? nb : 1));
  int64_t *tail = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nb ? nb : 1));
  int64_t kw[64], kw2[64];
  for (int64_t i = 0; i < nb; i++) {
    int anynull = 0; int32_t h = 42;
    for (int k = 0; k < nkeys; k++) { if (!so_valid(&bkeys[k], i)) { anynull = 1; break; } kw[k] = key_word(&bkeys[k], i); h = so_murmur3_long(kw[k], h); }
    next[i] = -1; tail[i] = i;
    if (anynull) continue;                                 /* never matches: not inserted */
    int64_t pos = (uint32_t)h & (cap - 1), step = 1;
    for (;;) {
      int64_t head = slot[pos];
      if (head < 0) { slot[pos] = i; shash[pos] = h; break; }
      if (shash[pos] == h) {
        int eq = 1;
        for (int k = 0; k < nkeys; k++) if (key_word(&bkeys[k], head) != kw[k]) { eq = 0; break; }
        if (eq) { next[tail[head]] = i; tail[head] = i; break; }
      }
      pos = (pos + step) & (cap - 1); step++;
    }
  }
  int64_t cnt = 0;
  for (int64_t j = 0; j < np; j++) {
    int anynull = 0; int32_t h = 42;
    for (int k = 0; k < nkeys; k++) { if (!so_valid(&pkeys[k], j)) { anynull = 1; break; } kw2[k] = key_word(&pkeys[k], j); h = so_murmur3_long(kw2[k], h); }
    int64_t head = -1;
    if (!anynull) {
      int64_t pos = (uint32_t)h & (cap - 1), step = 1;
      for (;;) {
        int64_t hd = slot[pos];
        if (hd < 0) break;
        if (shash[pos] == h) {
          int eq = 1;
          for (int k = 0; k < nkeys; k++) if (key_word(&bkeys[k], hd) != kw2[k]) { eq = 0; break; }
          if (eq) { head = hd; break; }
        }
        pos = (pos + step) & (cap - 1); step++;
      }
    }
    if (join_type == 0 || join_type == 1) {
      if (head >= 0) {
        for (int64_t b = head; b >= 0; b = next[b]) { if (out_probe) { out_probe[cnt] = j; out_build[cnt] = b; } cnt++; }
      } else if (join_type == 1) { if (out_probe) { out_probe[cnt] = j; out_build[cnt] = -1; } cnt++; }
    } else if (join_type == 2) {
      if (head >= 0) { if (out_probe) { out_probe[cnt] = j; out_build[cnt] = -1; } cnt++; }
    } else if (join_type == 3) {
      if (head < 0) { if (out_probe) { out_probe[cnt] = j; out_build[cnt] = -1; } cnt++; }
    }
  }
  free(slot); free(shash); free(next); free(tail);
  return cnt;
}

/* =====================================================================
 * Whole-stage "generated code" restatements used as the timed CPU baseline.
 *
 * TPC-H Q1 stage 1 = Scan -> Filter(l_shipdate <= cutoff) -> Project -> HashAggregate(Partial)
 * (plan: sql/core/src/test/resources/tpch-plan-stability/q1/simplified.txt), written the way
 * HashAggregateExec.doConsumeWithKeys emits it (HashAggregateExec.scala:907-1335): per row,
 * evaluate the predicate, build the key, hash it, probe the map (triangular probing as in
 * BytesToBytesMap.safeLookup), update the buffer in place.  One OpenMP thread = one Spark task
 * (local[N]); the per-task partial maps are merged at the end like the Final aggregate after the
 * Exchange.  Buffers per group: sum_qty, sum_price, sum_disc_price, sum_charge, sum_disc, count
 * (avg buffers (sum,count) alias the matching sums, which holds because NULL-free inputs make
 * Average's sum identical to Sum's).
 * out: for each of up to max_groups groups: key0,key1 and 6 buffer values.  Returns ngroups.
 * ===================================================================== */
typedef struct { int32_t used; int8_t k0, k1; double s[5]; int64_t cnt; } q1_slot;

int32_t so_q1_partial_final(const double *qty, const double *price, const double *disc, const double *tax,
                            const int8_t *rflag, const int8_t *lstatus, const int32_t *shipdate,
                            int64_t n, int32_t cutoff, int32_t max_groups,
                            int8_t *out_k0, int8_t *out_k1, double *out_sums /* [g][5] */, int64_t *out_cnt) {
  enum { CAP = 64 };
  int nt = so_threads();
  q1_slot *maps = (q1_slot *)calloc((size_t)nt * CAP, sizeof(q1_slot));
#pragma omp parallel
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    q1_slot *m = maps + (size_t)t * CAP;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      if (!(shipdate[i] <= cutoff)) continue;                       /* FilterExec */
      double p = price[i], d = disc[i];
      double dp = p * (1.0 - d);                                    /* ProjectExec expressions */
      double ch = dp * (1.0 + tax[i]);
      int8_t a = rflag[i], b = lstatus[i];
      int32_t h = so_murmur3_int(b, so_murmur3_int(a, 42));         /* key row hash */
      uint32_t pos = (uint32_t)h & (CAP - 1), step = 1;
      for (;;) {                                                    /* safeLookup: triangular probing */
        q1_slot *s = &m[pos];
        if (!s->used) { s->used = 1; s->k0 = a; s->k1 = b; }
        if (s->k0 == a && s->k1 == b) {
          s->s[0] += qty[i]; s->s[1] += p; s->s[2] += dp; s->s[3] += ch; s->s[4] += d; s->cnt++;
          break;
        }
        pos = (pos + step) & (CAP - 1); step++;
      }
    }
  }
  /* Exchange + Final: merge per-task partial buffers in task order */
  int32_t ng = 0;
  for (int t = 0; t < nt; t++) for (int s = 0; s < CAP; s++) {
    q1_slot *e = &maps[(size_t)t * CAP + s];
    if (!e->used) continue;
    int g = -1;
    for (int j = 0; j < ng; j++) if (out_k0[j] == e->k0 && out_k1[j] == e->k1) { g = j; break; }
    if (g < 0) {
      if (ng >= max_groups) continue;
      g = ng++; out_k0[g] = e->k0; out_k1[g] = e->k1; out_cnt[g] = 0;
      for (int k = 0; k < 5; k++) out_sums[g * 5 + k] = 0.0;
    }
    for (int k = 0; k < 5; k++) out_sums[g * 5 + k] += e->s[k];
    out_cnt[g] += e->cnt;
  }
  free(maps);
  return ng;
}

/* =====================================================================
 * Synthetic dataset on the host (include/sb_synth.h is the dataset's definition, shared with the
 * GPU generator as a header of pure functions; nothing of the product library is called).
 * The fill is an OpenMP static loop, so each page is first touched by the thread that will scan
 * it in the baseline loops below (NUMA placement follows the scan).
 * ===================================================================== */
#include "../include/sb_synth.h"

void so_synth_fill(int32_t table, int32_t col, int64_t n_orders, int64_t first, int64_t n, uint64_t seed, void *out) {
  const int w = sbs_width(table, col);
  const int is_f = table == SB_SYNTH_LINEITEM && sbs_lineitem_is_f64(col);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const int64_t row = first + i;
    if (is_f) { ((double *)out)[i] = sbs_lineitem_f64(seed, col, row); continue; }
    int64_t v;
    if (table == SB_SYNTH_LINEITEM) v = sbs_lineitem_i64(seed, col, row, n_orders);
    else if (table == SB_SYNTH_ORDERS) v = sbs_orders_i64(seed, col, row, n_orders);
    else if (table == SB_SYNTH_CUSTOMER) v = sbs_customer_i64(seed, col, row);
    else v = sbs_supplier_i64(seed, col, row);
    if (w == 1) ((int8_t *)out)[i] = (int8_t)v;
    else if (w == 4) ((int32_t *)out)[i] = (int32_t)v;
    else ((int64_t *)out)[i] = v;
  }
}
int64_t so_synth_rows(int32_t table, int64_t n_orders) {
  return table == SB_SYNTH_LINEITEM ? sbs_lineitem_rows(n_orders) : table == SB_SYNTH_ORDERS ? n_orders
       : table == SB_SYNTH_CUSTOMER ? sbs_customer_rows(n_orders) : sbs_supplier_rows(n_orders);
}

/* =====================================================================
 * TPC-H Q3 / Q5 as whole-stage restatements of the physical plans the GPU engine runs
 * (spark_b200/tpch.py q3_plan / q5_plan; shapes follow tpch-plan-stability/q3|q5/simplified.txt):
 * BroadcastHashJoin build sides become open-addressing relations keyed like LongToUnsafeRowMap
 * (HashedRelation.scala:594-622: h = k * 0x9E3779B9; slot = (h ^ h>>32) & mask, linear probing),
 * dense keys (customer, supplier) use its dense-array mode (:865-887); the streamed side is one
 * fused loop per stage: filter -> probe -> probe -> partial aggregate (HashAggregateExec.scala:907).
 * Every stage is an OpenMP loop (= local[N] tasks).  The aggregate buffer of a group lives next to the
 * build-side slot that determines it and is updated atomically, which is cheaper than Spark's
 * Partial -> Exchange -> Final and therefore generous to the CPU side of the comparison.
 * ===================================================================== */
typedef struct { int64_t key; int32_t a, b; } rel_slot;   /* key = -1: empty */
static inline uint64_t rel_hash(int64_t k) { uint64_t h = (uint64_t)k * 0x9E3779B9ull; return h ^ (h >> 32); }

static rel_slot *rel_alloc(int64_t n, int64_t *cap_out) {
  int64_t cap = 1024; while (cap < 2 * n) cap *= 2;
  rel_slot *t = (rel_slot *)malloc(sizeof(rel_slot) * (size_t)cap);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < cap; i++) t[i].key = -1;
  *cap_out = cap;
  return t;
}
static inline int64_t rel_insert(rel_slot *t, int64_t cap, int64_t key) {   /* unique keys; returns the slot */
  uint64_t pos = rel_hash(key) & (uint64_t)(cap - 1);
  for (;;) {
    int64_t cur = __atomic_load_n(&t[pos].key, __ATOMIC_RELAXED);
    if (cur == -1) {
      int64_t exp = -1;
      if (__atomic_compare_exchange_n(&t[pos].key, &exp, key, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return (int64_t)pos;
      cur = exp;
    }
    if (cur == key) return (int64_t)pos;
    pos = (pos + 1) & (uint64_t)(cap - 1);
  }
}
static inline int64_t rel_find(const rel_slot *t, int64_t cap, int64_t key) {
  uint64_t pos = rel_hash(key) & (uint64_t)(cap - 1);
  for (;;) {
    int64_t cur = t[pos].key;
    if (cur == key) return (int64_t)pos;
    if (cur == -1) return -1;
    pos = (pos + 1) & (uint64_t)(cap - 1);
  }
}
static inline void atomic_add_f64(double *p, double v) {
  uint64_t old = __atomic_load_n((uint64_t *)p, __ATOMIC_RELAXED), neu;
  do { double d; memcpy(&d, &old, 8); d += v; memcpy(&neu, &d, 8); }
  while (!__atomic_compare_exchange_n((uint64_t *)p, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

/* q3.sql: revenue per (l_orderkey, o_orderdate, o_shippriority) for segment customers, orders before `date`,
 * lines shipped after `date`; ORDER BY revenue DESC, o_orderdate LIMIT k.  Returns the number of result rows;
 * out_groups = number of groups before the limit. */
int32_t so_q3(const int64_t *c_custkey, const int8_t *c_mktsegment, int64_t n_cust,
              const int64_t *o_orderkey, const int64_t *o_custkey, const int32_t *o_orderdate, const int32_t *o_shippriority, int64_t n_ord,
              const int64_t *l_orderkey, const double *l_extendedprice, const double *l_discount, const int32_t *l_shipdate, int64_t n_li,
              int32_t segment, int32_t date, int32_t k,
              int64_t *out_orderkey, double *out_revenue, int32_t *out_orderdate, int32_t *out_shippriority, int64_t *out_groups) {
  /* build 1: filtered customer -> key set (dense: c_custkey is 1..n) */
  int64_t maxc = 0;
  for (int64_t i = 0; i < n_cust; i++) if (c_custkey[i] > maxc) maxc = c_custkey[i];
  uint8_t *cust_ok = (uint8_t *)calloc((size_t)maxc + 2, 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_cust; i++) if (c_mktsegment[i] == segment) cust_ok[c_custkey[i]] = 1;
  /* stage 2: orders: filter, probe customers, build relation 2 (orderkey -> date, priority) */
  int64_t cap;
  rel_slot *rel = rel_alloc(n_ord / 2 + 16, &cap);
  double *rev = (double *)calloc((size_t)cap, sizeof(double));
  uint8_t *touched = (uint8_t *)calloc((size_t)cap, 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_ord; i++) {
    if (!(o_orderdate[i] < date)) continue;
    int64_t c = o_custkey[i];
    if (c < 0 || c > maxc || !cust_ok[c]) continue;
    int64_t s = rel_insert(rel, cap, o_orderkey[i]);
    rel[s].a = o_orderdate[i]; rel[s].b = o_shippriority[i];
  }
  /* stage 3: lineitem: filter, probe, aggregate */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_li; i++) {
    if (!(l_shipdate[i] > date)) continue;
    int64_t s = rel_find(rel, cap, l_orderkey[i]);
    if (s < 0) continue;
    atomic_add_f64(&rev[s], l_extendedprice[i] * (1.0 - l_discount[i]));
    touched[s] = 1;
  }
  /* TakeOrderedAndProject (limit.scala:347-386): bounded selection of the k best */
  int32_t nres = 0; int64_t groups = 0;
  for (int64_t s = 0; s < cap; s++) {
    if (!touched[s]) continue;
    groups++;
    int32_t pos = nres;
    while (pos > 0 && (out_revenue[pos - 1] < rev[s] || (out_revenue[pos - 1] == rev[s] && out_orderdate[pos - 1] > rel[s].a))) pos--;
    if (pos >= k) continue;
    int32_t last = nres < k ? nres : k - 1;
    for (int32_t j = last; j > pos; j--) {
      out_orderkey[j] = out_orderkey[j - 1]; out_revenue[j] = out_revenue[j - 1];
      out_orderdate[j] = out_orderdate[j - 1]; out_shippriority[j] = out_shippriority[j - 1];
    }
    out_orderkey[pos] = rel[s].key; out_revenue[pos] = rev[s]; out_orderdate[pos] = rel[s].a; out_shippriority[pos] = rel[s].b;
    if (nres < k) nres++;
  }
  if (out_groups) *out_groups = groups;
  free(cust_ok); free(rel); free(rev); free(touched);
  return nres;
}

/* q5.sql: revenue per nation of `region` for orders in [date_lo, date_hi) where customer and supplier share the nation.
 * nation_region[25] maps n_nationkey -> n_regionkey.  out_revenue[25] indexed by nation key (0 where absent), out_seen[25]. */
void so_q5(const int64_t *c_custkey, const int64_t *c_nationkey, int64_t n_cust,
           const int64_t *o_orderkey, const int64_t *o_custkey, const int32_t *o_orderdate, int64_t n_ord,
           const int64_t *l_orderkey, const int64_t *l_suppkey, const double *l_extendedprice, const double *l_discount, int64_t n_li,
           const int64_t *s_suppkey, const int64_t *s_nationkey, int64_t n_supp,
           const int32_t *nation_region, int32_t region, int32_t date_lo, int32_t date_hi,
           double *out_revenue, uint8_t *out_seen) {
  int64_t maxc = 0, maxs = 0;
  for (int64_t i = 0; i < n_cust; i++) if (c_custkey[i] > maxc) maxc = c_custkey[i];
  for (int64_t i = 0; i < n_supp; i++) if (s_suppkey[i] > maxs) maxs = s_suppkey[i];
  int8_t *cnat = (int8_t *)malloc((size_t)maxc + 2), *snat = (int8_t *)malloc((size_t)maxs + 2);
  memset(cnat, -1, (size_t)maxc + 2); memset(snat, -1, (size_t)maxs + 2);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_cust; i++) cnat[c_custkey[i]] = (int8_t)c_nationkey[i];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_supp; i++)                      /* supplier |x| nation |x| region(filtered) */
    if (nation_region[s_nationkey[i]] == region) snat[s_suppkey[i]] = (int8_t)s_nationkey[i];
  int64_t cap;
  rel_slot *rel = rel_alloc(n_ord / 4 + 16, &cap);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_ord; i++) {
    if (!(o_orderdate[i] >= date_lo && o_orderdate[i] < date_hi)) continue;
    int64_t c = o_custkey[i];
    if (c < 0 || c > maxc || cnat[c] < 0) continue;
    int64_t s = rel_insert(rel, cap, o_orderkey[i]);
    rel[s].a = cnat[c];
  }
  int nt = so_threads();
  double *part = (double *)calloc((size_t)nt * 32, sizeof(double));
  uint8_t *pseen = (uint8_t *)calloc((size_t)nt * 32, 1);
#pragma omp parallel
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    double *p = part + (size_t)t * 32; uint8_t *ps = pseen + (size_t)t * 32;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n_li; i++) {
      int64_t s = rel_find(rel, cap, l_orderkey[i]);
      if (s < 0) continue;
      int64_t sk = l_suppkey[i];
      if (sk < 0 || sk > maxs || snat[sk] < 0) continue;
      if (snat[sk] != rel[s].a) continue;                     /* c_nationkey = s_nationkey */
      p[snat[sk]] += l_extendedprice[i] * (1.0 - l_discount[i]);
      ps[snat[sk]] = 1;
    }
  }
  for (int g = 0; g < 25; g++) { out_revenue[g] = 0; out_seen[g] = 0; }
  for (int t = 0; t < nt; t++) for (int g = 0; g < 25; g++) { out_revenue[g] += part[(size_t)t * 32 + g]; out_seen[g] |= pseen[(size_t)t * 32 + g]; }
  free(cnat); free(snat); free(rel); free(part); free(pseen);
}
