// aggregate.cu -- HashAggregateExec on the GPU, with the child FilterExec / ProjectExec fused in.
//
// Reference path replaced (citations relative to the reference tree, SQLX = sql/core/src/main/scala/
// org/apache/spark/sql/execution):
//   SQLX/aggregate/HashAggregateExec.scala:50 (doConsumeWithKeys :907-1335: per row project key ->
//   hash -> map probe/insert -> update buffer), UnsafeFixedWidthAggregationMap.java:124-181 over
//   BytesToBytesMap.java:604-643 (open addressing), buffer algebra Sum.scala:113-178,
//   Average.scala:80-135, Count.scala:94-105, Min/Max; modes AggUtils.scala:131-208;
//   FilterExec/ProjectExec SQLX/basicPhysicalOperators.scala:47,245 (fused by WholeStageCodegen there,
//   fused here by evaluating the predicate and the aggregate input expressions inside the update kernel).
//
// GPU design (HBM-bound; no tensor cores):
//   * one pass over the referenced input columns (coalesced loads, nothing materialised when the
//     predicate is a conjunction of column-vs-literal comparisons and every aggregate input is a
//     left-deep product of <= 3 factors of the form col | lit-col | lit+col | col-lit, which covers
//     TPC-H style price*(1-disc)*(1+tax) arithmetic; anything else is materialised first by expr.cu);
//   * group keys packed into one 64-bit word (fixed-width columns + null bits, or one raw 64-bit
//     column with reserved slots for NULL and for the EMPTY sentinel value);
//   * tier 1: a per-block dictionary of the first 8 distinct keys with LANE-PRIVATE shared-memory
//     accumulators (no atomics, no bank conflicts) -- this is what makes 4-group aggregates (Q1) run
//     at memory speed; tier 2: open-addressing table in HBM (linear probing, atomicCAS on the key word,
//     RED/ATOM on the accumulators);
//   * identical accumulator slots are de-duplicated (sum(x) and avg(x) share their sum; count(*) and the
//     counts of non-nullable averages share one counter).
// Floating SUM/AVG are order-dependent in the reference too (partials merge in fetch order); parity is
// 1e-6 relative for them, exact for keys/counts/integer sums/min/max.
#include <math.h>
#include "common.cuh"
#include "expr.cuh"
#include "primitives.cuh"

namespace sb {

constexpr int AGG_THREADS = 128;
constexpr int AGG_ITEMS = 8;
constexpr int AGG_DICT = 8;          // tier-1 dictionary entries per block
constexpr int AGG_MAX_SLOTS = 16;
constexpr int AGG_MAX_KEYS = 6;
constexpr int AGG_MAX_WORDS = 4;
constexpr int AGG_MAX_TERMS = 4;
constexpr int AGG_MAX_FACT = 3;
constexpr int AGG_PROBE_LIMIT = 64;
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;

enum SlotKind { K_ADD_I64 = 0, K_ADD_F64 = 1, K_MIN_U64 = 2, K_MAX_U64 = 3 };
enum ValXform { X_NONE = 0, X_SIGNED = 1, X_DOUBLE = 2 };   // value -> order-preserving u64 for min/max
enum FactorMode { F_COL = 0, F_LIT_MINUS_COL = 1, F_LIT_PLUS_COL = 2, F_COL_MINUS_LIT = 3 };
enum SlotClass { CLS_GENERIC = 0, CLS_ONE = 1, CLS_F64_PRODUCT = 2 };   // ONE: count(*) ; F64_PRODUCT: non-null double factors
enum TermOp { T_EQ = 0, T_NE, T_LT, T_LE, T_GT, T_GE, T_NOTNULL };

struct Factor {
  const void *data;
  const uint8_t *valid;
  int32_t type;
  int32_t mode;
  double lit;
};
struct SlotSrc {
  int32_t kind;      // SlotKind
  int32_t nf;        // number of factors (0 = constant one)
  int32_t is_one;    // value is 1 when every factor column is non-null (COUNT)
  int32_t xform;     // ValXform for min/max
  int32_t cls;       // SlotClass, precomputed on the host so the kernel takes ONE uniform branch per slot
  int32_t pad;
  Factor f[AGG_MAX_FACT];
};
struct KeySrc {
  const void *data;
  const uint8_t *valid;
  int32_t type;
  int32_t bits;      // value bits
  int32_t shift;     // position in the packed word
  int32_t null_shift;  // bit position of the null flag or -1
  int32_t word;      // which 64-bit key word holds the value (wide keys)
  int32_t null_word; // which word holds the null flag
};
struct FilterTerm {
  const void *data;
  const uint8_t *valid;
  int32_t type;
  int32_t op;
  int64_t lit;       // int64 value or double bits
  int32_t is_f64;
  int32_t pad;
};
struct AggArgs {
  int64_t n;
  int32_t nkeys, single64, nterms, nslots;
  KeySrc key[AGG_MAX_KEYS];
  FilterTerm term[AGG_MAX_TERMS];
  const uint8_t *mask;   // optional materialised predicate (1 byte / row)
  SlotSrc slot[AGG_MAX_SLOTS];
  int32_t nwords, pad0;  // 1 = packed single word; 2..AGG_MAX_WORDS = wide keys (lock-protocol table)
  uint32_t *tstate;      // wide keys only: [cap] 0 empty, 1 being written, 2 ready
  uint64_t *tkeys;       // [nwords][cap + 2]   slot cap = NULL key, slot cap+1 = key equal to the EMPTY sentinel
  uint64_t *tacc;        // [nslots][cap + 2]
  int32_t *flags;        // [0] abort (table too small), [1] NULL-key slot used, [2] sentinel-key slot used
  int64_t cap;
};

__device__ __forceinline__ uint64_t slot_identity(int kind) {
  return kind == K_MIN_U64 ? 0xFFFFFFFFFFFFFFFFull : 0ull;
}

__device__ __forceinline__ double factor_value(const Factor &f, int64_t row) {
  double x;
  switch (f.type) {
    case SB_FLOAT64: x = ((const double *)f.data)[row]; break;
    case SB_FLOAT32: x = (double)((const float *)f.data)[row]; break;
    default: x = (double)load_i64(f.data, f.type, row); break;
  }
  switch (f.mode) {
    case F_LIT_MINUS_COL: return __dsub_rn(f.lit, x);
    case F_LIT_PLUS_COL: return __dadd_rn(f.lit, x);
    case F_COL_MINUS_LIT: return __dsub_rn(x, f.lit);
    default: return x;
  }
}

// value of one accumulator slot for a row; returns false when the row does not contribute (NULL input)
__device__ __forceinline__ bool slot_value(const SlotSrc &s, int64_t row, uint64_t &out) {
#pragma unroll
  for (int k = 0; k < AGG_MAX_FACT; k++)
    if (k < s.nf && !bit_valid(s.f[k].valid, row)) return false;
  if (s.is_one) {
    out = 1;
    return true;
  }
  if (s.kind == K_ADD_F64 || s.xform == X_DOUBLE) {
    double v = factor_value(s.f[0], row);
    if (s.nf > 1) v = __dmul_rn(v, factor_value(s.f[1], row));
    if (s.nf > 2) v = __dmul_rn(v, factor_value(s.f[2], row));
    if (s.xform == X_DOUBLE) {   // order-preserving bits, NaN canonical (largest)
      int64_t b = double_bits_canonical(v);
      out = (uint64_t)b ^ ((uint64_t)(b >> 63) | 0x8000000000000000ull);
    } else {
      out = (uint64_t)__double_as_longlong(v);
    }
    return true;
  }
  int64_t v = load_i64(s.f[0].data, s.f[0].type, row);
  out = s.xform == X_SIGNED ? ((uint64_t)v ^ 0x8000000000000000ull) : (uint64_t)v;
  return true;
}

__device__ __forceinline__ uint64_t apply_op(int kind, uint64_t acc, uint64_t v) {
  switch (kind) {
    case K_ADD_I64: return acc + v;
    case K_ADD_F64: return (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((int64_t)acc), __longlong_as_double((int64_t)v)));
    case K_MIN_U64: return v < acc ? v : acc;
    default: return v > acc ? v : acc;
  }
}

__device__ __forceinline__ void global_op(int kind, uint64_t *addr, uint64_t v) {
  switch (kind) {
    case K_ADD_I64: atomicAdd((unsigned long long *)addr, (unsigned long long)v); break;
    case K_ADD_F64: atomicAdd((double *)addr, __longlong_as_double((int64_t)v)); break;
    case K_MIN_U64: atomicMin((unsigned long long *)addr, (unsigned long long)v); break;
    default: atomicMax((unsigned long long *)addr, (unsigned long long)v); break;
  }
}

__device__ __forceinline__ bool filter_row(const AggArgs &a, int64_t row) {
  if (a.mask && !a.mask[row]) return false;
#pragma unroll
  for (int t = 0; t < AGG_MAX_TERMS; t++) {
    if (t >= a.nterms) break;
    const FilterTerm &ft = a.term[t];
    if (!bit_valid(ft.valid, row)) return false;      // NULL comparison -> row dropped
    if (ft.op == T_NOTNULL) continue;
    int c;
    if (ft.is_f64) {
      double x = ft.type == SB_FLOAT32 ? (double)((const float *)ft.data)[row] : ((const double *)ft.data)[row];
      double y = __longlong_as_double(ft.lit);
      if (x == y) c = 0;
      else {
        bool xn = x != x, yn = y != y;
        c = (xn || yn) ? (int)xn - (int)yn : (x < y ? -1 : 1);
      }
    } else {
      int64_t x = load_i64(ft.data, ft.type, row);
      c = x == ft.lit ? 0 : (x < ft.lit ? -1 : 1);
    }
    bool ok = ft.op == T_EQ ? c == 0 : ft.op == T_NE ? c != 0 : ft.op == T_LT ? c < 0 : ft.op == T_LE ? c <= 0
              : ft.op == T_GT ? c > 0 : c >= 0;
    if (!ok) return false;
  }
  return true;
}

// group key of a row packed into one word.  special: 0 = regular key, 1 = NULL key of a single
// 64-bit column, 2 = a 64-bit key whose value equals the EMPTY sentinel.
__device__ __forceinline__ uint64_t pack_key(const AggArgs &a, int64_t row, int &special) {
  special = 0;
  if (a.single64) {
    const KeySrc &k = a.key[0];
    if (!bit_valid(k.valid, row)) { special = 1; return 0; }
    uint64_t v;
    if (k.type == SB_FLOAT64) {      // NormalizeFloatingNumbers: -0.0 -> 0.0, NaN canonical
      double d = ((const double *)k.data)[row];
      v = d == 0.0 ? 0ull : (uint64_t)double_bits_canonical(d);
    } else v = (uint64_t)((const int64_t *)k.data)[row];
    if (v == EMPTY_KEY) special = 2;
    return v;
  }
  uint64_t w = 0;
#pragma unroll
  for (int i = 0; i < AGG_MAX_KEYS; i++) {
    if (i >= a.nkeys) break;
    const KeySrc &k = a.key[i];
    if (!bit_valid(k.valid, row)) { w |= 1ull << k.null_shift; continue; }
    uint64_t v;
    if (k.type == SB_FLOAT32) {
      float f = ((const float *)k.data)[row];
      v = f == 0.0f ? 0u : (uint32_t)float_bits_canonical(f);
    } else {
      v = (uint64_t)load_i64(k.data, k.type, row);
      if (k.bits < 64) v &= (1ull << k.bits) - 1;
    }
    w |= v << k.shift;
  }
  return w;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // slot choice only; not contractual
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// find-or-insert in the HBM table; returns slot index or -1 (abort: table too small)
__device__ __forceinline__ int64_t table_slot(const AggArgs &a, uint64_t key, int special) {
  if (special == 1) { a.flags[1] = 1; return a.cap; }
  if (special == 2) { a.flags[2] = 1; return a.cap + 1; }
  uint64_t mask = (uint64_t)a.cap - 1;
  uint64_t h = mix64(key) & mask;
  for (int step = 0; step < AGG_PROBE_LIMIT; step++) {
    uint64_t cur = a.tkeys[h];
    if (cur == key) return (int64_t)h;
    if (cur == EMPTY_KEY) {
      uint64_t old = atomicCAS((unsigned long long *)&a.tkeys[h], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
      if (old == EMPTY_KEY || old == key) return (int64_t)h;
    }
    h = (h + 1) & mask;
  }
  a.flags[0] = 1;
  return -1;
}

constexpr int64_t DST_SKIP = INT64_MIN;

// ---------------------------------------------------------------------------------------------------------
// Tile processing.  A thread owns AGG_ITEMS rows of a tile (row = row0 + k*AGG_THREADS, coalesced per k).
// Rules that keep the instruction count per row low (first profile: 524 thread-instructions per row, 12% BRA,
// 12% BSSY/BSYNC -- see profiles/r01_agg_update.md):
//   * every descriptor (filter term, key column, accumulator slot) is read from the parameter bank once per
//     tile, and every warp-uniform decision (column type, factor mode, accumulator kind) is taken OUTSIDE the
//     unrolled row loop;
//   * FULL tiles use unclamped loads, so the AGG_ITEMS loads of a column are LDG [base + k*stride] with
//     immediate offsets and are all in flight together; only the last partial tile clamps the row index.
// ---------------------------------------------------------------------------------------------------------
template <bool FULL, typename T>
__device__ __forceinline__ void load_batch_as_i64(const void *__restrict__ data, int64_t row0, int64_t last, int64_t (&out)[AGG_ITEMS]) {
  const T *p = (const T *)data + row0;
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) {
    if (FULL) out[k] = (int64_t)p[k * AGG_THREADS];
    else {
      int64_t r = row0 + (int64_t)k * AGG_THREADS;
      out[k] = (int64_t)((const T *)data)[r < last ? r : last];
    }
  }
}
template <bool FULL>
__device__ __forceinline__ void load_i64_batch(const void *__restrict__ data, int32_t type, int64_t row0, int64_t last,
                                               int64_t (&out)[AGG_ITEMS]) {
  switch (type) {
    case SB_BOOL: load_batch_as_i64<FULL, uint8_t>(data, row0, last, out); break;
    case SB_INT8: load_batch_as_i64<FULL, int8_t>(data, row0, last, out); break;
    case SB_INT16: load_batch_as_i64<FULL, int16_t>(data, row0, last, out); break;
    case SB_INT32: case SB_DATE32: case SB_FLOAT32: load_batch_as_i64<FULL, int32_t>(data, row0, last, out); break;
    default: load_batch_as_i64<FULL, int64_t>(data, row0, last, out); break;
  }
}
template <bool FULL>
__device__ __forceinline__ void load_f64_batch(const void *__restrict__ data, int32_t type, int64_t row0, int64_t last,
                                               double (&out)[AGG_ITEMS]) {
  if (type == SB_FLOAT64) {
    const double *p = (const double *)data + row0;
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) {
      if (FULL) out[k] = p[k * AGG_THREADS];
      else {
        int64_t r = row0 + (int64_t)k * AGG_THREADS;
        out[k] = ((const double *)data)[r < last ? r : last];
      }
    }
  } else {
    int64_t t[AGG_ITEMS];
    load_i64_batch<FULL>(data, type, row0, last, t);
    if (type == SB_FLOAT32) {
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) out[k] = (double)__int_as_float((int32_t)t[k]);
    } else {
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) out[k] = (double)t[k];
    }
  }
}
// validity bits of the AGG_ITEMS rows; the caller skips this entirely when the column has no bitmap
template <bool FULL>
__device__ __forceinline__ void load_valid_batch(const uint8_t *__restrict__ valid, int64_t row0, int64_t last, bool (&out)[AGG_ITEMS]) {
  uint8_t b[AGG_ITEMS];
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) {
    int64_t r = row0 + (int64_t)k * AGG_THREADS;
    if (!FULL) r = r < last ? r : last;
    b[k] = valid[r >> 3];
  }
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) {
    int64_t r = row0 + (int64_t)k * AGG_THREADS;
    if (!FULL) r = r < last ? r : last;
    out[k] = (b[k] >> (r & 7)) & 1;
  }
}



// fused FilterExec: conjunction of column-vs-literal terms, applied to the thread's AGG_ITEMS rows
template <bool FULL>
__device__ __forceinline__ void apply_filter_terms(const AggArgs &a, int64_t row0, bool (&keep)[AGG_ITEMS]) {
  const int64_t last = a.n - 1;
  for (int t = 0; t < a.nterms; t++) {
    const void *data = a.term[t].data;
    const uint8_t *vptr = a.term[t].valid;
    const int32_t type = a.term[t].type, op = a.term[t].op, is_f64 = a.term[t].is_f64;
    const int64_t lit = a.term[t].lit;
    if (vptr) {   // a NULL comparison drops the row
      bool valid[AGG_ITEMS];
      load_valid_batch<FULL>(vptr, row0, last, valid);
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) keep[k] = keep[k] && valid[k];
    }
    if (op == T_NOTNULL) continue;
    int c[AGG_ITEMS];
    if (is_f64) {
      double x[AGG_ITEMS];
      load_f64_batch<FULL>(data, type, row0, last, x);
      const double y = __longlong_as_double(lit);
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) {   // SQLOrderingUtil.compareDoubles
        bool xn = x[k] != x[k], yn = y != y;
        c[k] = x[k] == y ? 0 : (xn || yn) ? (int)xn - (int)yn : (x[k] < y ? -1 : 1);
      }
    } else {
      int64_t x[AGG_ITEMS];
      load_i64_batch<FULL>(data, type, row0, last, x);
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) c[k] = x[k] == lit ? 0 : (x[k] < lit ? -1 : 1);
    }
    switch (op) {   // uniform: one specialised row loop per operator
#define SB_CMP(COND) _Pragma("unroll") for (int k = 0; k < AGG_ITEMS; k++) keep[k] = keep[k] && (COND);
      case T_EQ: SB_CMP(c[k] == 0) break;
      case T_NE: SB_CMP(c[k] != 0) break;
      case T_LT: SB_CMP(c[k] < 0) break;
      case T_LE: SB_CMP(c[k] <= 0) break;
      case T_GT: SB_CMP(c[k] > 0) break;
      default: SB_CMP(c[k] >= 0) break;
#undef SB_CMP
    }
  }
}

// value of one factor for the thread's rows: col | lit-col | lit+col | col-lit (mode switch outside the row loop)
template <bool FULL>
__device__ __forceinline__ void factor_batch(const Factor &f, int64_t row0, int64_t last, double (&y)[AGG_ITEMS]) {
  const int32_t mode = f.mode;
  const double lit = f.lit;
  load_f64_batch<FULL>(f.data, f.type, row0, last, y);
  switch (mode) {
    case F_LIT_MINUS_COL:
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dsub_rn(lit, y[k]);
      break;
    case F_LIT_PLUS_COL:
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dadd_rn(lit, y[k]);
      break;
    case F_COL_MINUS_LIT:
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dsub_rn(y[k], lit);
      break;
    default: break;
  }
}

// slot descriptor outer, rows inner.  dst[k] >= 0: HBM table slot; doff[k] >= 0: offset of the row's group in
// the lane-private shared-memory accumulators (dictionary hit); both negative: row filtered out.
template <bool FULL>
__device__ __forceinline__ void accumulate_slots(const AggArgs &a, int64_t row0, const int64_t (&dst)[AGG_ITEMS],
                                                 const int (&doff)[AGG_ITEMS], uint64_t *acc, int ns, int64_t stride) {
  const int64_t last = a.n - 1;
  for (int s = 0; s < ns; s++) {
    const int kind = a.slot[s].kind, nf = a.slot[s].nf, is_one = a.slot[s].is_one, xform = a.slot[s].xform;
    uint64_t v[AGG_ITEMS];
    bool ok[AGG_ITEMS];
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) ok[k] = dst[k] >= 0 || doff[k] >= 0;
    // NULL inputs do not contribute (Sum.scala:113, Count.scala:94)
    for (int f = 0; f < nf; f++) {
      const uint8_t *vptr = a.slot[s].f[f].valid;
      if (vptr) {
        bool valid[AGG_ITEMS];
        load_valid_batch<FULL>(vptr, row0, last, valid);
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) ok[k] = ok[k] && valid[k];
      }
    }
    if (is_one) {
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) v[k] = 1;
    } else if (kind == K_ADD_F64 || xform == X_DOUBLE) {
      double y[AGG_ITEMS];
      factor_batch<FULL>(a.slot[s].f[0], row0, last, y);
      for (int f = 1; f < nf; f++) {   // left-deep product, evaluated in the reference's order
        double z[AGG_ITEMS];
        factor_batch<FULL>(a.slot[s].f[f], row0, last, z);
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dmul_rn(y[k], z[k]);
      }
      if (xform == X_DOUBLE) {   // order-preserving bits for min/max, NaN canonical (largest)
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) {
          int64_t b = double_bits_canonical(y[k]);
          v[k] = (uint64_t)b ^ ((uint64_t)(b >> 63) | 0x8000000000000000ull);
        }
      } else {
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) v[k] = (uint64_t)__double_as_longlong(y[k]);
      }
    } else {
      int64_t x[AGG_ITEMS];
      load_i64_batch<FULL>(a.slot[s].f[0].data, a.slot[s].f[0].type, row0, last, x);
      const uint64_t flip = xform == X_SIGNED ? 0x8000000000000000ull : 0ull;
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) v[k] = (uint64_t)x[k] ^ flip;
    }
    uint64_t *sacc = acc + (size_t)s * AGG_THREADS;
    uint64_t *gacc = a.tacc + (int64_t)s * stride;
    switch (kind) {   // uniform: one specialised row loop per accumulator kind
#define SB_ACC(LOCAL, GLOBAL)                                          \
  _Pragma("unroll") for (int k = 0; k < AGG_ITEMS; k++) {              \
    if (!ok[k]) continue;                                              \
    if (doff[k] >= 0) { uint64_t *p = sacc + doff[k]; LOCAL; }         \
    else { uint64_t *p = gacc + dst[k]; GLOBAL; }                      \
  }
      case K_ADD_F64:
        SB_ACC(*(double *)p = __dadd_rn(*(double *)p, __longlong_as_double((int64_t)v[k])),
               atomicAdd((double *)p, __longlong_as_double((int64_t)v[k])))
        break;
      case K_ADD_I64:
        SB_ACC(*p += v[k], atomicAdd((unsigned long long *)p, (unsigned long long)v[k]))
        break;
      case K_MIN_U64:
        SB_ACC(*p = v[k] < *p ? v[k] : *p, atomicMin((unsigned long long *)p, (unsigned long long)v[k]))
        break;
      default:
        SB_ACC(*p = v[k] > *p ? v[k] : *p, atomicMax((unsigned long long *)p, (unsigned long long)v[k]))
        break;
#undef SB_ACC
    }
  }
}

// Branch-free variant used when every kept row of the warp resolved to a dictionary entry: filtered rows and
// NULL inputs are steered to a "trash" accumulator group (index AGG_DICT), so the row loop is LDS + op + STS.
template <bool FULL>
__device__ __forceinline__ void accumulate_slots_dict(const AggArgs &a, int64_t row0, const int (&doff)[AGG_ITEMS], uint64_t *acc,
                                                      int ns, int trash) {
  const int64_t last = a.n - 1;
  for (int s = 0; s < ns; s++) {
    const int cls = a.slot[s].cls;
    uint64_t *sacc = acc + (size_t)s * AGG_THREADS;
    if (cls == CLS_F64_PRODUCT) {
      const int nf = a.slot[s].nf;
      double y[AGG_ITEMS];
      factor_batch<FULL>(a.slot[s].f[0], row0, last, y);
      for (int f = 1; f < nf; f++) {
        double z[AGG_ITEMS];
        factor_batch<FULL>(a.slot[s].f[f], row0, last, z);
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dmul_rn(y[k], z[k]);
      }
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) {
        double *p = (double *)(sacc + doff[k]);
        *p = __dadd_rn(*p, y[k]);
      }
    } else if (cls == CLS_ONE) {
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) sacc[doff[k]] += 1;
    } else {
      const int kind = a.slot[s].kind, nf = a.slot[s].nf, is_one = a.slot[s].is_one, xform = a.slot[s].xform;
      uint64_t v[AGG_ITEMS];
      int tgt[AGG_ITEMS];
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) tgt[k] = doff[k];
      for (int f = 0; f < nf; f++) {
        const uint8_t *vptr = a.slot[s].f[f].valid;
        if (vptr) {
          bool valid[AGG_ITEMS];
          load_valid_batch<FULL>(vptr, row0, last, valid);
#pragma unroll
          for (int k = 0; k < AGG_ITEMS; k++) tgt[k] = valid[k] ? tgt[k] : trash;
        }
      }
      if (is_one) {
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) v[k] = 1;
      } else if (kind == K_ADD_F64 || xform == X_DOUBLE) {
        double y[AGG_ITEMS];
        factor_batch<FULL>(a.slot[s].f[0], row0, last, y);
        for (int f = 1; f < nf; f++) {
          double z[AGG_ITEMS];
          factor_batch<FULL>(a.slot[s].f[f], row0, last, z);
#pragma unroll
          for (int k = 0; k < AGG_ITEMS; k++) y[k] = __dmul_rn(y[k], z[k]);
        }
        if (xform == X_DOUBLE) {
#pragma unroll
          for (int k = 0; k < AGG_ITEMS; k++) {
            int64_t b = double_bits_canonical(y[k]);
            v[k] = (uint64_t)b ^ ((uint64_t)(b >> 63) | 0x8000000000000000ull);
          }
        } else {
#pragma unroll
          for (int k = 0; k < AGG_ITEMS; k++) v[k] = (uint64_t)__double_as_longlong(y[k]);
        }
      } else {
        int64_t x[AGG_ITEMS];
        load_i64_batch<FULL>(a.slot[s].f[0].data, a.slot[s].f[0].type, row0, last, x);
        const uint64_t flip = xform == X_SIGNED ? 0x8000000000000000ull : 0ull;
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) v[k] = (uint64_t)x[k] ^ flip;
      }
      switch (kind) {
#define SB_ACCD(EXPR) _Pragma("unroll") for (int k = 0; k < AGG_ITEMS; k++) { uint64_t *p = sacc + tgt[k]; EXPR; }
        case K_ADD_F64: SB_ACCD(*(double *)p = __dadd_rn(*(double *)p, __longlong_as_double((int64_t)v[k]))) break;
        case K_ADD_I64: SB_ACCD(*p += v[k]) break;
        case K_MIN_U64: SB_ACCD(*p = v[k] < *p ? v[k] : *p) break;
        default: SB_ACCD(*p = v[k] > *p ? v[k] : *p) break;
#undef SB_ACCD
      }
    }
  }
}

// Shared memory layout: uint64 dict_keys[AGG_DICT]; uint64 acc[(AGG_DICT + 1) * nslots][AGG_THREADS] (last group = trash)
template <bool FULL>
__device__ __forceinline__ void process_tile(const AggArgs &a, int64_t base, int use_dict, uint64_t *dict_keys, uint64_t *acc, int tid,
                                             int ns, int64_t stride) {
  const int64_t row0 = base + tid;
  const int64_t last = a.n - 1;
  bool keep[AGG_ITEMS];
  uint64_t key[AGG_ITEMS];
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) {
    keep[k] = FULL || row0 + (int64_t)k * AGG_THREADS < a.n;
    key[k] = 0;
  }
  if (a.mask) {
    int64_t m[AGG_ITEMS];
    load_batch_as_i64<FULL, uint8_t>(a.mask, row0, last, m);
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) keep[k] = keep[k] && m[k] != 0;
  }
  apply_filter_terms<FULL>(a, row0, keep);
  // ---- group key packing --------------------------------------------------------------------------------
  int special[AGG_ITEMS];
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) special[k] = 0;
  if (a.single64) {
    const void *kd = a.key[0].data;
    const uint8_t *kv = a.key[0].valid;
    const int32_t kt = a.key[0].type;
    int64_t x[AGG_ITEMS];
    load_batch_as_i64<FULL, int64_t>(kd, row0, last, x);      // raw 64-bit words (int64 or double bits)
    if (kt == SB_FLOAT64) {   // NormalizeFloatingNumbers: -0.0 -> 0.0, NaN canonical
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) {
        double d = __longlong_as_double(x[k]);
        x[k] = d == 0.0 ? 0ll : double_bits_canonical(d);
      }
    }
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) {
      key[k] = (uint64_t)x[k];
      special[k] = key[k] == EMPTY_KEY ? 2 : 0;
    }
    if (kv) {
      bool valid[AGG_ITEMS];
      load_valid_batch<FULL>(kv, row0, last, valid);
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) special[k] = valid[k] ? special[k] : 1;
    }
  } else {
    for (int i = 0; i < a.nkeys; i++) {
      const void *kd = a.key[i].data;
      const uint8_t *kv = a.key[i].valid;
      const int32_t kt = a.key[i].type, bits = a.key[i].bits, shift = a.key[i].shift, nshift = a.key[i].null_shift;
      int64_t x[AGG_ITEMS];
      load_i64_batch<FULL>(kd, kt, row0, last, x);
      if (kt == SB_FLOAT32) {
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) {
          float f = __int_as_float((int32_t)x[k]);
          x[k] = f == 0.0f ? 0 : (int64_t)(uint32_t)float_bits_canonical(f);
        }
      }
      const uint64_t vmask = bits < 64 ? (1ull << bits) - 1 : ~0ull;
      if (kv) {
        bool valid[AGG_ITEMS];
        load_valid_batch<FULL>(kv, row0, last, valid);
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) key[k] |= valid[k] ? ((uint64_t)x[k] & vmask) << shift : 1ull << nshift;
      } else {
#pragma unroll
        for (int k = 0; k < AGG_ITEMS; k++) key[k] |= ((uint64_t)x[k] & vmask) << shift;
      }
    }
  }
  // ---- where does each row accumulate? ---------------------------------------------------------------------
  int64_t dst[AGG_ITEMS];
  int doff[AGG_ITEMS];
  const int trash = AGG_DICT * ns * AGG_THREADS + tid;
  bool all_dict = true;
#pragma unroll
  for (int k = 0; k < AGG_ITEMS; k++) {
    dst[k] = -1;
    doff[k] = -1;
    if (!keep[k]) continue;
    int gid = -1;
    if (use_dict && special[k] == 0) {
      // linear probing over the AGG_DICT entries starting at a key-dependent entry: a resident key is
      // normally found by the first probe
      const uint32_t h0 = (uint32_t)(key[k] ^ (key[k] >> 7) ^ (key[k] >> 17) ^ (key[k] >> 32));
#pragma unroll 1
      for (int i = 0; i < AGG_DICT; i++) {
        const int g = (h0 + i) & (AGG_DICT - 1);
        uint64_t dk = dict_keys[g];
        if (dk == EMPTY_KEY) {
          uint64_t old = atomicCAS((unsigned long long *)&dict_keys[g], (unsigned long long)EMPTY_KEY, (unsigned long long)key[k]);
          dk = old == EMPTY_KEY ? key[k] : old;
        }
        if (dk == key[k]) { gid = g; break; }
      }
    }
    if (gid >= 0) doff[k] = gid * ns * AGG_THREADS + tid;
    else {
      dst[k] = table_slot(a, key[k], special[k]);
      all_dict = false;
    }
  }
  if (use_dict && __all_sync(0xffffffffu, all_dict)) {
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) doff[k] = doff[k] >= 0 ? doff[k] : trash;
    accumulate_slots_dict<FULL>(a, row0, doff, acc, ns, trash);
  } else {
    accumulate_slots<FULL>(a, row0, dst, doff, acc, ns, stride);
  }
}

__global__ void __launch_bounds__(AGG_THREADS) agg_update_kernel(const __grid_constant__ AggArgs a, int use_dict) {
  extern __shared__ uint64_t sm[];
  uint64_t *dict_keys = sm;
  uint64_t *acc = sm + AGG_DICT;
  const int tid = threadIdx.x;
  const int ns = a.nslots;
  if (use_dict) {
    if (tid < AGG_DICT) dict_keys[tid] = EMPTY_KEY;
    for (int s = 0; s < ns; s++) {
      uint64_t id = slot_identity(a.slot[s].kind);
      for (int g = 0; g <= AGG_DICT; g++) acc[(g * ns + s) * AGG_THREADS + tid] = id;
    }
    __syncthreads();
  }
  const int64_t tile = (int64_t)AGG_THREADS * AGG_ITEMS;
  const int64_t stride = a.cap + 2;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < a.n; base += (int64_t)gridDim.x * tile) {
    if (*(volatile int32_t *)a.flags) break;   // another block found the table too small: give up early
    if (base + tile <= a.n) process_tile<true>(a, base, use_dict, dict_keys, acc, tid, ns, stride);
    else process_tile<false>(a, base, use_dict, dict_keys, acc, tid, ns, stride);
  }
  if (!use_dict) return;
  __syncthreads();
  // merge the block dictionary into the HBM table: one warp per (group, slot) pair, shuffle tree
  const int lane = tid & 31, warp = tid >> 5, nwarps = AGG_THREADS / 32;
  for (int gs = warp; gs < AGG_DICT * ns; gs += nwarps) {
    int g = gs / ns, s = gs % ns;
    uint64_t key = dict_keys[g];
    if (key == EMPTY_KEY) continue;
    int kind = a.slot[s].kind;
    uint64_t v = slot_identity(kind);
    for (int t = lane; t < AGG_THREADS; t += 32) v = apply_op(kind, v, acc[(size_t)gs * AGG_THREADS + t]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v = apply_op(kind, v, __shfl_xor_sync(0xffffffffu, v, d));
    if (lane == 0) {
      int64_t slot = table_slot(a, key, 0);
      if (slot >= 0) global_op(kind, &a.tacc[(int64_t)s * stride + slot], v);
    }
  }
}

// ---- wide grouping keys (> 63 bits, e.g. Q3's (l_orderkey, o_orderdate, o_shippriority)) -----------------
// Same row loop, but the key is up to AGG_MAX_WORDS 64-bit words and the HBM table entry is published with a
// small state machine: 0 empty -> 1 (claimed by atomicCAS, key words being written) -> 2 ready.
template <int NW>
__device__ __forceinline__ int64_t table_slot_wide(const AggArgs &a, const uint64_t (&w)[NW]) {
  const int64_t stride = a.cap + 2;
  uint64_t mask = (uint64_t)a.cap - 1;
  uint64_t hh = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) hh = mix64(hh ^ w[i]);
  uint64_t h = hh & mask;
  volatile uint32_t *state = a.tstate;
  for (int step = 0; step < AGG_PROBE_LIMIT;) {
    uint32_t s = state[h];
    if (s == 0) {
      if (atomicCAS(&a.tstate[h], 0u, 1u) == 0u) {
#pragma unroll
        for (int i = 0; i < NW; i++) a.tkeys[(int64_t)i * stride + h] = w[i];
        __threadfence();
        state[h] = 2u;
        return (int64_t)h;
      }
      continue;   // somebody else claimed it: look again
    }
    if (s == 1) continue;   // being written
    __threadfence();
    bool eq = true;
#pragma unroll
    for (int i = 0; i < NW; i++) eq &= ((volatile uint64_t *)a.tkeys)[(int64_t)i * stride + h] == w[i];
    if (eq) return (int64_t)h;
    h = (h + 1) & mask;
    step++;
  }
  a.flags[0] = 1;
  return -1;
}

template <int NW>
__global__ void __launch_bounds__(AGG_THREADS) agg_update_wide_kernel(const __grid_constant__ AggArgs a) {
  const int tid = threadIdx.x;
  const int ns = a.nslots;
  const int64_t tile = (int64_t)AGG_THREADS * AGG_ITEMS;
  const int64_t stride = a.cap + 2;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < a.n; base += (int64_t)gridDim.x * tile) {
    if (a.flags[0]) break;
    const int64_t row0 = base + tid;
    bool keep[AGG_ITEMS];
    uint64_t key[AGG_ITEMS][NW];
    int64_t dst[AGG_ITEMS];
    int doff[AGG_ITEMS];
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) {
      int64_t row = row0 + (int64_t)k * AGG_THREADS;
      keep[k] = row < a.n && (a.mask == nullptr || a.mask[row]);
#pragma unroll
      for (int i = 0; i < NW; i++) key[k][i] = 0;
    }
    apply_filter_terms<false>(a, row0, keep);
    for (int i = 0; i < a.nkeys; i++) {
      const KeySrc ks = a.key[i];
#pragma unroll
      for (int k = 0; k < AGG_ITEMS; k++) {
        if (!keep[k]) continue;
        int64_t row = row0 + (int64_t)k * AGG_THREADS;
        uint64_t v;
        int target = ks.word;
        if (!bit_valid(ks.valid, row)) { v = 1ull << ks.null_shift; target = ks.null_word; }
        else {
          if (ks.type == SB_FLOAT32) {
            float f = ((const float *)ks.data)[row];
            v = f == 0.0f ? 0u : (uint32_t)float_bits_canonical(f);
          } else if (ks.type == SB_FLOAT64) {
            double d = ((const double *)ks.data)[row];
            v = d == 0.0 ? 0ull : (uint64_t)double_bits_canonical(d);
          } else {
            v = (uint64_t)load_i64(ks.data, ks.type, row);
            if (ks.bits < 64) v &= (1ull << ks.bits) - 1;
          }
          v <<= ks.shift;
        }
#pragma unroll
        for (int wi = 0; wi < NW; wi++)
          if (wi == target) key[k][wi] |= v;
      }
    }
#pragma unroll
    for (int k = 0; k < AGG_ITEMS; k++) {
      dst[k] = -1;
      doff[k] = -1;
      if (!keep[k]) continue;
      dst[k] = table_slot_wide<NW>(a, key[k]);
    }
    accumulate_slots<false>(a, row0, dst, doff, nullptr, ns, stride);
  }
}

__global__ void occupied_wide_kernel(const uint32_t *__restrict__ tstate, int64_t cap, uint8_t *__restrict__ occ) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) occ[i] = tstate[i] == 2u;
  else if (i < cap + 2) occ[i] = 0;
}

__global__ void fill_u64_kernel(uint64_t *p, int64_t n, uint64_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void occupied_kernel(const uint64_t *__restrict__ tkeys, int64_t cap, const int32_t *__restrict__ flags,
                                uint8_t *__restrict__ occ) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) occ[i] = tkeys[i] != EMPTY_KEY;
  else if (i == cap) occ[i] = flags[1] != 0;
  else if (i == cap + 1) occ[i] = flags[2] != 0;
}

// ---- emit: one thread per output group -------------------------------------------------------
struct EmitKey {
  void *out;
  uint32_t *out_valid;
  int32_t type, bits, shift, null_shift, word, null_word;
};
enum EmitOp { E_RAW = 0, E_SUM_NULLABLE, E_AVG, E_MINMAX };
struct EmitCol {
  void *out;
  uint32_t *out_valid;
  int32_t op;         // EmitOp
  int32_t out_type;   // SB_INT64 / SB_FLOAT64 / source type for min/max
  int32_t slot;       // main accumulator
  int32_t slot2;      // seen-count / avg count (or -1)
  int32_t xform;      // for min/max
  int32_t pad;
};
struct EmitArgs {
  int32_t nkeys, single64, ncols, pad;
  EmitKey key[AGG_MAX_KEYS];
  EmitCol col[2 * AGG_MAX_SLOTS];
  const uint64_t *tkeys;
  const uint64_t *tacc;
  const int64_t *slot_ids;
  int64_t cap, ngroups;
};

__device__ __forceinline__ void store_typed(void *out, int32_t type, int64_t i, int64_t v) {
  switch (type) {
    case SB_BOOL: case SB_INT8: ((int8_t *)out)[i] = (int8_t)v; break;
    case SB_INT16: ((int16_t *)out)[i] = (int16_t)v; break;
    case SB_INT32: case SB_DATE32: case SB_FLOAT32: ((int32_t *)out)[i] = (int32_t)v; break;
    default: ((int64_t *)out)[i] = v; break;
  }
}

__global__ void __launch_bounds__(256) agg_emit_kernel(const __grid_constant__ EmitArgs e) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in_range = r < e.ngroups;
  int64_t slot = in_range ? e.slot_ids[r] : 0;
  const int64_t stride = e.cap + 2;
#pragma unroll 1
  for (int k = 0; k < e.nkeys; k++) {
    const EmitKey &ek = e.key[k];
    bool valid = in_range;
    int64_t v = 0;
    if (in_range) {
      uint64_t key = e.tkeys[(int64_t)ek.word * stride + slot];
      if (e.single64) {
        if (slot == e.cap) valid = false;                       // NULL group key
        else if (slot == e.cap + 1) v = (int64_t)EMPTY_KEY;
        else v = (int64_t)key;
      } else {
        if (ek.null_shift >= 0 && ((e.tkeys[(int64_t)ek.null_word * stride + slot] >> ek.null_shift) & 1)) valid = false;
        else {
          uint64_t u = key >> ek.shift;
          if (ek.bits < 64) {
            u &= (1ull << ek.bits) - 1;
            // sign-extend integer types
            if (ek.type != SB_BOOL && ek.type != SB_FLOAT32 && (u >> (ek.bits - 1)) & 1) u |= ~((1ull << ek.bits) - 1);
          }
          v = (int64_t)u;
        }
      }
      store_typed(ek.out, ek.type, r, valid ? v : 0);
    }
    if (ek.out_valid) {
      uint32_t w = __ballot_sync(0xffffffffu, valid);
      if ((threadIdx.x & 31) == 0 && (r - (r & 31)) < e.ngroups) ek.out_valid[r >> 5] = w;
    }
  }
#pragma unroll 1
  for (int c = 0; c < e.ncols; c++) {
    const EmitCol &ec = e.col[c];
    bool valid = in_range;
    if (in_range) {
      uint64_t m = e.tacc[(int64_t)ec.slot * stride + slot];
      uint64_t m2 = ec.slot2 >= 0 ? e.tacc[(int64_t)ec.slot2 * stride + slot] : 1;
      int64_t v = (int64_t)m;
      switch (ec.op) {
        case E_SUM_NULLABLE: valid = m2 != 0; break;            // all-NULL group -> NULL (Sum.scala:180)
        case E_AVG: {                                           // Average.scala:109-127: count == 0 -> NULL
          valid = m2 != 0;
          double s = __longlong_as_double((int64_t)m);
          v = valid ? __double_as_longlong(__ddiv_rn(s, (double)(int64_t)m2)) : 0;
          break;
        }
        case E_MINMAX: {
          valid = m2 != 0;
          if (ec.xform == X_SIGNED) v = (int64_t)(m ^ 0x8000000000000000ull);
          else if (ec.xform == X_DOUBLE) {
            uint64_t b = (m >> 63) ? (m ^ 0x8000000000000000ull) : ~m;
            v = (int64_t)b;
            if (ec.out_type == SB_FLOAT32) v = (int64_t)__float_as_int((float)__longlong_as_double((int64_t)b));
          }
          break;
        }
        default: break;
      }
      store_typed(ec.out, ec.out_type, r, valid ? v : 0);
    }
    if (ec.out_valid) {
      uint32_t w = __ballot_sync(0xffffffffu, valid);
      if ((threadIdx.x & 31) == 0 && (r - (r & 31)) < e.ngroups) ec.out_valid[r >> 5] = w;
    }
  }
}

// ==============================================================================================
// host side: plan -> kernel arguments
// ==============================================================================================
struct ExprTree {   // tiny tree view of a postfix program
  int op, vtype, arg;
  int64_t lit;
  int l = -1, r = -1;
};
static int build_tree(const sb_expr &e, std::vector<ExprTree> &t) {
  std::vector<int> stack;
  for (int i = 0; i < e.n; i++) {
    ExprTree nd;
    nd.op = e.nodes[i].op;
    nd.vtype = e.nodes[i].vtype;
    nd.arg = e.nodes[i].arg;
    nd.lit = e.nodes[i].lit.i;
    switch (nd.op) {
      case SB_OP_COL: case SB_OP_LIT_I64: case SB_OP_LIT_F64: case SB_OP_LIT_NULL: break;
      case SB_OP_NEG: case SB_OP_NOT: case SB_OP_ISNULL: case SB_OP_ISNOTNULL:
      case SB_OP_CAST_F64: case SB_OP_CAST_I64: case SB_OP_CAST_I32:
        nd.l = stack.back(); stack.pop_back(); break;
      default:
        nd.r = stack.back(); stack.pop_back();
        nd.l = stack.back(); stack.pop_back();
    }
    t.push_back(nd);
    stack.push_back((int)t.size() - 1);
  }
  return stack.back();
}

static bool is_f64_col(const sb_table *in, const ExprTree &n) {
  return n.op == SB_OP_COL && in->cols[n.arg].type == SB_FLOAT64;
}
static double lit_as_double(const ExprTree &n) {
  if (n.op == SB_OP_LIT_F64) { double d; memcpy(&d, &n.lit, 8); return d; }
  return (double)n.lit;
}
static bool is_lit(const ExprTree &n) { return n.op == SB_OP_LIT_F64 || n.op == SB_OP_LIT_I64; }

static bool match_factor(const sb_table *in, const std::vector<ExprTree> &t, int i, Factor &f) {
  const ExprTree &n = t[i];
  auto set_col = [&](const ExprTree &c) {
    const Column &col = in->cols[c.arg];
    f.data = col.d(); f.valid = col.v(); f.type = col.type;
  };
  if (is_f64_col(in, n)) { set_col(n); f.mode = F_COL; f.lit = 0; return true; }
  if (n.vtype != SB_VT_F64) return false;
  if (n.op == SB_OP_SUB && is_lit(t[n.l]) && is_f64_col(in, t[n.r])) { set_col(t[n.r]); f.mode = F_LIT_MINUS_COL; f.lit = lit_as_double(t[n.l]); return true; }
  if (n.op == SB_OP_SUB && is_f64_col(in, t[n.l]) && is_lit(t[n.r])) { set_col(t[n.l]); f.mode = F_COL_MINUS_LIT; f.lit = lit_as_double(t[n.r]); return true; }
  if (n.op == SB_OP_ADD && is_lit(t[n.l]) && is_f64_col(in, t[n.r])) { set_col(t[n.r]); f.mode = F_LIT_PLUS_COL; f.lit = lit_as_double(t[n.l]); return true; }
  if (n.op == SB_OP_ADD && is_f64_col(in, t[n.l]) && is_lit(t[n.r])) { set_col(t[n.l]); f.mode = F_LIT_PLUS_COL; f.lit = lit_as_double(t[n.r]); return true; }
  return false;
}

// left-deep product of <= 3 factors, evaluated in the reference's order ((f0*f1)*f2)
static bool match_product(const sb_table *in, const sb_expr &e, SlotSrc &s) {
  std::vector<ExprTree> t;
  int root = build_tree(e, t);
  Factor f[AGG_MAX_FACT];
  int chain[AGG_MAX_FACT];
  int nf = 0, cur = root;
  // peel right factors: Mul(Mul(a,b),c) -> [a,b,c]
  int rights[AGG_MAX_FACT];
  int nr = 0;
  while (t[cur].op == SB_OP_MUL && t[cur].vtype == SB_VT_F64 && nr < AGG_MAX_FACT - 1) {
    rights[nr++] = t[cur].r;
    cur = t[cur].l;
  }
  chain[nf++] = cur;
  for (int k = nr - 1; k >= 0; k--) chain[nf++] = rights[k];
  for (int k = 0; k < nf; k++)
    if (!match_factor(in, t, chain[k], f[k])) return false;
  s.nf = nf;
  for (int k = 0; k < nf; k++) s.f[k] = f[k];
  return true;
}

static bool match_filter(const sb_table *in, const sb_expr &e, AggArgs &a) {
  std::vector<ExprTree> t;
  int root = build_tree(e, t);
  std::vector<int> work{root};
  a.nterms = 0;
  while (!work.empty()) {
    int i = work.back();
    work.pop_back();
    const ExprTree &n = t[i];
    if (n.op == SB_OP_AND) { work.push_back(n.l); work.push_back(n.r); continue; }
    if (a.nterms >= AGG_MAX_TERMS) return false;
    FilterTerm &ft = a.term[a.nterms];
    memset(&ft, 0, sizeof(ft));
    if (n.op == SB_OP_ISNOTNULL && t[n.l].op == SB_OP_COL && in->cols[t[n.l].arg].type != SB_STRING) {
      const Column &c = in->cols[t[n.l].arg];
      ft.data = c.d(); ft.valid = c.v(); ft.type = c.type; ft.op = T_NOTNULL;
      a.nterms++;
      continue;
    }
    if (n.op < SB_OP_EQ || n.op > SB_OP_GE) return false;
    int op = n.op - SB_OP_EQ;   // T_EQ..T_GE share the order of SB_OP_EQ..SB_OP_GE
    const ExprTree *col = &t[n.l], *lit = &t[n.r];
    if (is_lit(*col) && lit->op == SB_OP_COL) {   // lit cmp col -> flip
      std::swap(col, lit);
      static const int flip[6] = {T_EQ, T_NE, T_GT, T_GE, T_LT, T_LE};
      op = flip[op];
    }
    if (col->op != SB_OP_COL || !is_lit(*lit)) return false;
    const Column &c = in->cols[col->arg];
    if (c.type == SB_STRING) return false;
    ft.data = c.d(); ft.valid = c.v(); ft.type = c.type; ft.op = op;
    ft.is_f64 = n.arg == SB_VT_F64;
    if (ft.is_f64) {
      if (c.type != SB_FLOAT64 && c.type != SB_FLOAT32) return false;
      double d = lit_as_double(*lit);
      memcpy(&ft.lit, &d, 8);
    } else {
      if (c.type == SB_FLOAT64 || c.type == SB_FLOAT32 || lit->op != SB_OP_LIT_I64) return false;
      ft.lit = lit->lit;
    }
    a.nterms++;
  }
  return true;
}

static bool same_slot(const SlotSrc &x, const SlotSrc &y) {
  if (x.kind != y.kind || x.nf != y.nf || x.is_one != y.is_one || x.xform != y.xform) return false;
  for (int k = 0; k < x.nf; k++) {
    const Factor &a = x.f[k], &b = y.f[k];
    if (a.data != b.data || a.valid != b.valid || a.type != b.type || a.mode != b.mode) return false;
    if (memcmp(&a.lit, &b.lit, 8) != 0) return false;
    // is_one slots only look at validity: columns without validity are interchangeable
  }
  return true;
}

struct AggBuilder {
  const sb_table *in;
  cudaStream_t st;
  AggArgs args;
  std::vector<Column> temps;   // materialised expression results, released at the end

  int add_slot(SlotSrc s) {
    if (s.is_one) {   // drop factors whose column can never be NULL: count(x) over a non-nullable x == count(*)
      int k2 = 0;
      for (int k = 0; k < s.nf; k++)
        if (s.f[k].valid) s.f[k2++] = s.f[k];
      s.nf = k2;
      for (int k = 0; k < s.nf; k++) { s.f[k].mode = F_COL; s.f[k].lit = 0; s.f[k].data = nullptr; s.f[k].type = 0; }
    }
    for (int i = 0; i < args.nslots; i++)
      if (same_slot(args.slot[i], s)) return i;
    if (args.nslots >= AGG_MAX_SLOTS) fail(SB_ERR_UNSUPPORTED, "aggregate needs more than %d accumulator slots", AGG_MAX_SLOTS);
    args.slot[args.nslots] = s;
    return args.nslots++;
  }

  // value source for an aggregate input expression; want_f64 = accumulate as double
  SlotSrc source_for(const sb_expr &e, bool want_f64, int32_t *src_type) {
    SlotSrc s;
    memset(&s, 0, sizeof(s));
    int col;
    if (expr_is_column(e, &col)) {
      SB_REQUIRE(col >= 0 && col < (int)in->cols.size(), "aggregate input column %d out of range", col);
      const Column &c = in->cols[col];
      if (c.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "aggregates over string columns are not supported");
      s.nf = 1;
      s.f[0].data = c.d(); s.f[0].valid = c.v(); s.f[0].type = c.type; s.f[0].mode = F_COL;
      *src_type = c.type;
      return s;
    }
    expr_validate(in, e);
    (void)want_f64;
    if (e.nodes[e.n - 1].vtype == SB_VT_F64 && match_product(in, e, s)) {
      *src_type = SB_FLOAT64;
      return s;
    }
    Column tmp = eval_projection(in, e, nullptr, in->nrows, st);   // general path: materialise
    temps.push_back(tmp);
    s.nf = 1;
    s.f[0].data = tmp.d(); s.f[0].valid = tmp.v(); s.f[0].type = tmp.type; s.f[0].mode = F_COL;
    *src_type = tmp.type;
    return s;
  }
};

static bool is_float_type(int32_t t) { return t == SB_FLOAT32 || t == SB_FLOAT64; }

struct OutPlan {   // one output column
  int op, out_type, slot, slot2, xform;
  bool nullable;
};

static int64_t next_pow2(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

static void hash_aggregate_impl(const sb_table *in, const sb_agg_plan *plan, cudaStream_t st, sb_table **out) {
  SB_REQUIRE(in && plan && out, "null argument");
  SB_REQUIRE(plan->mode >= SB_AGG_MODE_PARTIAL && plan->mode <= SB_AGG_MODE_COMPLETE, "bad aggregate mode %d", plan->mode);
  SB_REQUIRE(plan->nkeys >= 0 && plan->naggs >= 0, "bad plan");
  const int64_t n = in->nrows;
  AggBuilder b;
  b.in = in;
  b.st = st;
  memset(&b.args, 0, sizeof(b.args));
  AggArgs &a = b.args;
  a.n = n;
  std::vector<OutPlan> outs;
  Scratch mask_buf(0, st);
  void *mask_ptr = nullptr;

  struct Cleanup {
    AggBuilder &b;
    void *&mask;
    cudaStream_t st;
    ~Cleanup() {
      for (auto &c : b.temps) column_release(c);
      if (mask) cudaFreeAsync(mask, st);
    }
  } cleanup{b, mask_ptr, st};

  // ---- keys --------------------------------------------------------------------------------
  SB_REQUIRE(plan->nkeys <= AGG_MAX_KEYS, "at most %d grouping keys are supported (got %d)", AGG_MAX_KEYS, plan->nkeys);
  a.nkeys = plan->nkeys;
  int total_bits = 0;
  for (int k = 0; k < plan->nkeys; k++) {
    int ci = plan->key_cols[k];
    SB_REQUIRE(ci >= 0 && ci < (int)in->cols.size(), "key column %d out of range", ci);
    const Column &c = in->cols[ci];
    if (c.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "string grouping keys are not supported (dictionary-encode them)");
    KeySrc &ks = a.key[k];
    ks.data = c.d(); ks.valid = c.v(); ks.type = c.type;
    ks.bits = type_width(c.type) * 8;
    total_bits += ks.bits + (c.validity ? 1 : 0);
  }
  a.nwords = 1;
  if (plan->nkeys == 1 && a.key[0].bits == 64) {
    a.single64 = 1;
    a.key[0].shift = 0;
    a.key[0].null_shift = -1;
  } else if (total_bits <= 63) {
    int pos = 0;
    for (int k = 0; k < plan->nkeys; k++) {
      a.key[k].shift = pos;
      pos += a.key[k].bits;
      if (a.key[k].valid) a.key[k].null_shift = pos++;
      else a.key[k].null_shift = -1;
    }
  } else {
    // wide key: value fields first (a field never straddles words), then the null flags in the free bits
    int used[AGG_MAX_WORDS + 1] = {0};
    int nw = 0;
    for (int k = 0; k < plan->nkeys; k++) {
      int w = 0;
      while (w < AGG_MAX_WORDS && used[w] + a.key[k].bits > 64) w++;
      if (w >= AGG_MAX_WORDS) fail(SB_ERR_UNSUPPORTED, "grouping keys do not fit in %d 64-bit words", AGG_MAX_WORDS);
      a.key[k].word = w;
      a.key[k].shift = used[w];
      used[w] += a.key[k].bits;
      if (w + 1 > nw) nw = w + 1;
    }
    for (int k = 0; k < plan->nkeys; k++) {
      a.key[k].null_shift = -1;
      if (!a.key[k].valid) continue;
      int w = 0;
      while (w < AGG_MAX_WORDS && used[w] + 1 > 64) w++;
      if (w >= AGG_MAX_WORDS) fail(SB_ERR_UNSUPPORTED, "grouping keys do not fit in %d 64-bit words", AGG_MAX_WORDS);
      a.key[k].null_word = w;
      a.key[k].null_shift = used[w]++;
      if (w + 1 > nw) nw = w + 1;
    }
    a.nwords = nw < 2 ? 2 : nw;
  }

  // ---- fused filter ---------------------------------------------------------------------------
  if (plan->filter) {
    expr_validate(in, *plan->filter);
    if (!match_filter(in, *plan->filter, a)) {
      a.nterms = 0;
      SB_CUDA(cudaMallocAsync(&mask_ptr, (size_t)n + 16, st));
      eval_predicate(in, *plan->filter, (uint8_t *)mask_ptr, st);
      a.mask = (const uint8_t *)mask_ptr;
    }
  }

  // ---- aggregates -> accumulator slots + output plan -------------------------------------------
  const bool final_mode = plan->mode == SB_AGG_MODE_FINAL;
  const bool emit_buffers = plan->mode == SB_AGG_MODE_PARTIAL;
  int next_buf_col = plan->nkeys;   // Final: buffers follow the keys positionally
  for (int i = 0; i < plan->naggs; i++) {
    const sb_agg_spec &sp = plan->aggs[i];
    auto buffer_source = [&](int col) {
      SB_REQUIRE(col < (int)in->cols.size(), "Final aggregate expects buffer column %d but the input has %zu columns", col, in->cols.size());
      const Column &c = in->cols[col];
      SlotSrc s;
      memset(&s, 0, sizeof(s));
      s.nf = 1;
      s.f[0].data = c.d(); s.f[0].valid = c.v(); s.f[0].type = c.type; s.f[0].mode = F_COL;
      return s;
    };
    switch (sp.func) {
      case SB_AGG_SUM: {
        int32_t src_type;
        SlotSrc s = final_mode ? buffer_source(next_buf_col) : b.source_for(sp.input, false, &src_type);
        if (final_mode) { src_type = in->cols[next_buf_col].type; next_buf_col++; }
        else if (s.nf >= 1 && is_float_type(s.f[0].type)) src_type = SB_FLOAT64;
        bool f64 = is_float_type(src_type) || s.nf > 1 || s.f[0].mode != F_COL;
        s.kind = f64 ? K_ADD_F64 : K_ADD_I64;
        // a global aggregate (no keys) over zero rows yields NULL, so it always tracks "seen"
        bool nullable = plan->nkeys == 0;
        for (int k = 0; k < s.nf; k++) nullable |= s.f[k].valid != nullptr;
        int main = b.add_slot(s);
        int seen = -1;
        if (nullable) {
          SlotSrc c = s;
          c.kind = K_ADD_I64; c.is_one = 1;
          seen = b.add_slot(c);
        }
        outs.push_back({nullable ? E_SUM_NULLABLE : E_RAW, f64 ? SB_FLOAT64 : SB_INT64, main, seen, X_NONE, nullable});
        break;
      }
      case SB_AGG_AVG: {
        int sum_slot, cnt_slot;
        if (final_mode) {
          SlotSrc s = buffer_source(next_buf_col); s.kind = K_ADD_F64;
          SB_REQUIRE(in->cols[next_buf_col].type == SB_FLOAT64, "avg buffer sum must be float64");
          SlotSrc c = buffer_source(next_buf_col + 1); c.kind = K_ADD_I64;
          SB_REQUIRE(in->cols[next_buf_col + 1].type == SB_INT64, "avg buffer count must be int64");
          next_buf_col += 2;
          sum_slot = b.add_slot(s);
          cnt_slot = b.add_slot(c);
        } else {
          int32_t src_type;
          SlotSrc s = b.source_for(sp.input, true, &src_type);
          s.kind = K_ADD_F64;
          sum_slot = b.add_slot(s);
          SlotSrc c = s;
          c.kind = K_ADD_I64; c.is_one = 1;
          cnt_slot = b.add_slot(c);
        }
        if (emit_buffers) {
          outs.push_back({E_RAW, SB_FLOAT64, sum_slot, -1, X_NONE, false});
          outs.push_back({E_RAW, SB_INT64, cnt_slot, -1, X_NONE, false});
        } else {
          outs.push_back({E_AVG, SB_FLOAT64, sum_slot, cnt_slot, X_NONE, true});
        }
        break;
      }
      case SB_AGG_COUNT: case SB_AGG_COUNT_STAR: {
        SlotSrc s;
        memset(&s, 0, sizeof(s));
        if (final_mode) {
          s = buffer_source(next_buf_col);
          SB_REQUIRE(in->cols[next_buf_col].type == SB_INT64, "count buffer must be int64");
          next_buf_col++;
          s.kind = K_ADD_I64;
        } else if (sp.func == SB_AGG_COUNT) {
          int32_t src_type;
          s = b.source_for(sp.input, false, &src_type);
          s.kind = K_ADD_I64; s.is_one = 1;
        } else {
          s.kind = K_ADD_I64; s.is_one = 1; s.nf = 0;
        }
        outs.push_back({E_RAW, SB_INT64, b.add_slot(s), -1, X_NONE, false});
        break;
      }
      case SB_AGG_MIN: case SB_AGG_MAX: {
        int32_t src_type;
        SlotSrc s = final_mode ? buffer_source(next_buf_col) : b.source_for(sp.input, false, &src_type);
        if (final_mode) { src_type = in->cols[next_buf_col].type; next_buf_col++; }
        if (s.nf != 1 || s.f[0].mode != F_COL) {   // computed double expression
          src_type = SB_FLOAT64;
        }
        s.kind = sp.func == SB_AGG_MIN ? K_MIN_U64 : K_MAX_U64;
        s.xform = is_float_type(src_type) ? X_DOUBLE : X_SIGNED;
        int main = b.add_slot(s);
        SlotSrc c = s;
        c.kind = K_ADD_I64; c.is_one = 1; c.xform = X_NONE;
        int seen = b.add_slot(c);
        outs.push_back({E_MINMAX, src_type, main, seen, s.xform, true});
        break;
      }
      default: fail(SB_ERR_INVALID, "unknown aggregate function %d", sp.func);
    }
  }
  a.nslots = b.args.nslots;
  for (int i = 0; i < a.nslots; i++) {
    SlotSrc &sl = a.slot[i];
    sl.cls = CLS_GENERIC;
    if (sl.is_one && sl.nf == 0 && sl.kind == K_ADD_I64) sl.cls = CLS_ONE;
    if (!sl.is_one && sl.kind == K_ADD_F64 && sl.xform == X_NONE && sl.nf >= 1) {
      bool plain = true;
      for (int k = 0; k < sl.nf; k++) plain &= sl.f[k].type == SB_FLOAT64 && sl.f[k].valid == nullptr;
      if (plain) sl.cls = CLS_F64_PRODUCT;
    }
  }

  // ---- table sizing + update, retrying with a larger table when probing gives up ---------------
  int64_t cap_max = next_pow2(n > 512 ? 2 * n : 1024);
  int64_t cap = plan->expected_groups > 0 ? next_pow2(2 * plan->expected_groups) : (1 << 16);
  if (cap < 1024) cap = 1024;
  if (cap > cap_max) cap = cap_max;
  const bool use_dict = plan->nkeys > 0 || true;
  size_t smem = use_dict ? (size_t)(AGG_DICT + (size_t)(AGG_DICT + 1) * a.nslots * AGG_THREADS) * 8 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA(cudaFuncSetAttribute(agg_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  SB_REQUIRE(smem <= 200 * 1024, "aggregate needs %zu bytes of shared memory", smem);
  int blocks_per_sm = smem ? (int)((227 * 1024) / (smem + 1024)) : 8;
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  if (blocks_per_sm > 8) blocks_per_sm = 8;
  int grid = grid_for(n, AGG_THREADS * AGG_ITEMS, rt().num_sms * blocks_per_sm);

  Scratch flags(16, st);
  void *tkeys = nullptr, *tacc = nullptr, *tstate = nullptr;
  struct TableFree {
    void *&k, *&v, *&s;
    cudaStream_t st;
    ~TableFree() {
      if (k) cudaFreeAsync(k, st);
      if (v) cudaFreeAsync(v, st);
      if (s) cudaFreeAsync(s, st);
    }
  } table_free_guard{tkeys, tacc, tstate, st};

  for (;;) {
    int64_t slots = cap + 2;
    SB_CUDA(cudaMallocAsync(&tkeys, (size_t)slots * 8 * a.nwords, st));
    SB_CUDA(cudaMallocAsync(&tacc, (size_t)slots * 8 * (a.nslots ? a.nslots : 1), st));
    if (a.nwords > 1) {
      SB_CUDA(cudaMallocAsync(&tstate, (size_t)slots * 4, st));
      SB_CUDA(cudaMemsetAsync(tstate, 0, (size_t)slots * 4, st));
    } else {
      SB_CUDA(cudaMemsetAsync(tkeys, 0xff, (size_t)slots * 8, st));
    }
    SB_CUDA(cudaMemsetAsync(tacc, 0, (size_t)slots * 8 * (a.nslots ? a.nslots : 1), st));
    for (int s = 0; s < a.nslots; s++)
      if (a.slot[s].kind == K_MIN_U64) SB_CUDA(cudaMemsetAsync((uint64_t *)tacc + (int64_t)s * slots, 0xff, (size_t)slots * 8, st));
    SB_CUDA(cudaMemsetAsync(flags.ptr, 0, 16, st));
    a.tkeys = (uint64_t *)tkeys;
    a.tstate = (uint32_t *)tstate;
    a.tacc = (uint64_t *)tacc;
    a.flags = flags.as<int32_t>();
    a.cap = cap;
    if (n > 0) {
      KernelTimer kt(final_mode ? "agg_update_final" : "agg_update", st);
      if (a.nwords == 1) agg_update_kernel<<<grid, AGG_THREADS, smem, st>>>(a, use_dict ? 1 : 0);
      else if (a.nwords == 2) agg_update_wide_kernel<2><<<grid, AGG_THREADS, 0, st>>>(a);
      else if (a.nwords == 3) agg_update_wide_kernel<3><<<grid, AGG_THREADS, 0, st>>>(a);
      else agg_update_wide_kernel<4><<<grid, AGG_THREADS, 0, st>>>(a);
      SB_LAUNCH_CHECK();
    }
    int32_t hflags[4];
    SB_CUDA(cudaMemcpyAsync(hflags, flags.ptr, 16, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (!hflags[0]) break;
    if (cap >= cap_max) fail(SB_ERR_CUDA, "hash aggregate table overflow at maximum capacity %lld", (long long)cap);
    cudaFreeAsync(tkeys, st); tkeys = nullptr;
    cudaFreeAsync(tacc, st); tacc = nullptr;
    if (tstate) { cudaFreeAsync(tstate, st); tstate = nullptr; }
    cap = cap * 16 > cap_max ? cap_max : cap * 16;
  }

  // ---- collect occupied slots -------------------------------------------------------------------
  int64_t slots = cap + 2;
  int64_t ngroups;
  Scratch slot_ids(slots * 8, st);
  if (plan->nkeys == 0) {
    // no grouping keys: exactly one output row, even for empty input (AggregateCodegenSupport.scala:131)
    // every row used key 0 -> find its slot, or use slot 0 of an untouched table
    Scratch occ(slots + 16, st);
    occupied_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>((uint64_t *)tkeys, cap, flags.as<int32_t>(), occ.as<uint8_t>());
    SB_LAUNCH_CHECK();
    ngroups = compact_mask(occ.as<uint8_t>(), slots, slot_ids.as<int64_t>(), st);
    if (ngroups == 0) {
      SB_CUDA(cudaMemsetAsync(slot_ids.ptr, 0, 8, st));
      ngroups = 1;
    }
  } else {
    Scratch occ(slots + 16, st);
    if (a.nwords > 1) occupied_wide_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>((uint32_t *)tstate, cap, occ.as<uint8_t>());
    else occupied_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>((uint64_t *)tkeys, cap, flags.as<int32_t>(), occ.as<uint8_t>());
    SB_LAUNCH_CHECK();
    ngroups = compact_mask(occ.as<uint8_t>(), slots, slot_ids.as<int64_t>(), st);
  }

  // ---- emit -------------------------------------------------------------------------------------
  sb_table *t = table_new(ngroups);
  try {
    EmitArgs e;
    memset(&e, 0, sizeof(e));
    e.nkeys = plan->nkeys;
    e.single64 = a.single64;
    e.tkeys = (const uint64_t *)tkeys;
    e.tacc = (const uint64_t *)tacc;
    e.slot_ids = slot_ids.as<int64_t>();
    e.cap = cap;
    e.ngroups = ngroups;
    for (int k = 0; k < plan->nkeys; k++) {
      const Column &src = in->cols[plan->key_cols[k]];
      Column c = column_alloc(src.type, src.scale, ngroups, src.validity != nullptr, st);
      t->cols.push_back(c);
      e.key[k].out = c.data->ptr;
      e.key[k].out_valid = c.validity ? (uint32_t *)c.validity->ptr : nullptr;
      e.key[k].type = src.type;
      e.key[k].bits = a.key[k].bits;
      e.key[k].shift = a.key[k].shift;
      e.key[k].null_shift = a.key[k].null_shift;
      e.key[k].word = a.key[k].word;
      e.key[k].null_word = a.key[k].null_word;
    }
    SB_REQUIRE(outs.size() <= 2 * AGG_MAX_SLOTS, "too many aggregate output columns");
    e.ncols = (int)outs.size();
    for (size_t i = 0; i < outs.size(); i++) {
      Column c = column_alloc(outs[i].out_type, 0, ngroups, outs[i].nullable, st);
      t->cols.push_back(c);
      e.col[i].out = c.data->ptr;
      e.col[i].out_valid = c.validity ? (uint32_t *)c.validity->ptr : nullptr;
      e.col[i].op = outs[i].op;
      e.col[i].out_type = outs[i].out_type;
      e.col[i].slot = outs[i].slot;
      e.col[i].slot2 = outs[i].slot2;
      e.col[i].xform = outs[i].xform;
    }
    if (ngroups > 0) {
      agg_emit_kernel<<<(unsigned)((ngroups + 255) / 256), 256, 0, st>>>(e);
      SB_LAUNCH_CHECK();
    }
    SB_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_hash_aggregate(const sb_table *in, const sb_agg_plan *plan, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  hash_aggregate_impl(in, plan, stream_of(s), out);
  SB_API_END
}
