// simple_pred.cuh -- conjunctions of column-vs-literal comparisons (most pushed-down filters: Q1, Q3, Q5, Q6 ...), shared by the
// FilterExec fast path (expr.cu) and the operators that fuse the filter below them into their own first pass (join.cu).
#pragma once
#include "common.cuh"

namespace sb {

constexpr int SP_MAX_TERMS = 4;
constexpr int SP_ROWS = 16;
enum SpOp { SP_EQ = 0, SP_NE, SP_LT, SP_LE, SP_GT, SP_GE, SP_NOTNULL };
struct SimplePred {
  int32_t nterms;
  int32_t type[SP_MAX_TERMS], op[SP_MAX_TERMS], f64[SP_MAX_TERMS];
  const void *data[SP_MAX_TERMS];
  const uint8_t *valid[SP_MAX_TERMS];
  int64_t lit[SP_MAX_TERMS];   // int64 value or double bits
};

// host: does `e` have that shape over table `in`?  Fills sp when it does (expr.cu)
bool match_simple_predicate(const sb_table *in, const sb_expr &e, SimplePred &sp);

#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ void sp_load16(const void *__restrict__ data, int64_t row0, int64_t n, int64_t (&x)[SP_ROWS]) {
  const T *p = (const T *)data + row0;
  if (row0 + SP_ROWS <= n && ((uintptr_t)p & 15) == 0) {
    constexpr int PER = 16 / sizeof(T);
#pragma unroll
    for (int v = 0; v < SP_ROWS / PER; v++) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + v);
      T tmp[PER];
      memcpy(tmp, &q, 16);
#pragma unroll
      for (int j = 0; j < PER; j++) x[v * PER + j] = (int64_t)tmp[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < SP_ROWS; j++) x[j] = row0 + j < n ? (int64_t)p[j] : 0;
  }
}

// bit j of the result: the predicate is TRUE for row row0 + j (row0 a multiple of 16; bits of rows >= n are unspecified)
__device__ __forceinline__ uint32_t simple_pred_eval16(const SimplePred &sp, int64_t row0, int64_t n) {
  uint32_t keep = 0xFFFFu;
#pragma unroll 1
  for (int t = 0; t < sp.nterms; t++) {
    if (sp.valid[t]) {
      // rows row0..row0+15 start on a byte boundary of the bitmap (row0 is a multiple of 16)
      const uint8_t *v = sp.valid[t] + (row0 >> 3);
      uint32_t bits = v[0];
      if (row0 + 8 < n) bits |= (uint32_t)v[1] << 8;
      keep &= bits;
    }
    const int op = sp.op[t];
    if (op == SP_NOTNULL) continue;
    int64_t x[SP_ROWS];
    switch (sp.type[t]) {
      case SB_BOOL: sp_load16<uint8_t>(sp.data[t], row0, n, x); break;
      case SB_INT8: sp_load16<int8_t>(sp.data[t], row0, n, x); break;
      case SB_INT16: sp_load16<int16_t>(sp.data[t], row0, n, x); break;
      case SB_INT32: case SB_DATE32: case SB_FLOAT32: sp_load16<int32_t>(sp.data[t], row0, n, x); break;
      default: sp_load16<int64_t>(sp.data[t], row0, n, x); break;
    }
    uint32_t lt = 0, eq = 0;
    if (sp.f64[t]) {
      const double y = __longlong_as_double(sp.lit[t]);
      const bool yn = y != y;
#pragma unroll
      for (int j = 0; j < SP_ROWS; j++) {   // SQLOrderingUtil.compareDoubles: NaN equals NaN and is larger than anything else
        const double d = sp.type[t] == SB_FLOAT32 ? (double)__int_as_float((int32_t)x[j]) : __longlong_as_double(x[j]);
        const bool dn = d != d;
        const int c = d == y ? 0 : (dn || yn) ? (int)dn - (int)yn : (d < y ? -1 : 1);
        lt |= (c < 0 ? 1u : 0u) << j;
        eq |= (c == 0 ? 1u : 0u) << j;
      }
    } else {
      const int64_t lit = sp.lit[t];
#pragma unroll
      for (int j = 0; j < SP_ROWS; j++) {
        lt |= (x[j] < lit ? 1u : 0u) << j;
        eq |= (x[j] == lit ? 1u : 0u) << j;
      }
    }
    uint32_t ok;
    switch (op) {
      case SP_EQ: ok = eq; break;
      case SP_NE: ok = ~eq; break;
      case SP_LT: ok = lt; break;
      case SP_LE: ok = lt | eq; break;
      case SP_GT: ok = ~(lt | eq); break;
      default: ok = ~lt; break;
    }
    keep &= ok;
  }
  return keep & 0xFFFFu;
}
// the same predicate for ONE row (lane-strided kernels: the 32 lanes of a warp read 32 consecutive rows, so plain loads coalesce)
__device__ __forceinline__ bool simple_pred_row(const SimplePred &sp, int64_t row) {
  bool keep = true;
#pragma unroll 1
  for (int t = 0; t < sp.nterms; t++) {
    if (sp.valid[t]) keep = keep && ((sp.valid[t][row >> 3] >> (row & 7)) & 1);
    const int op = sp.op[t];
    if (op == SP_NOTNULL) continue;
    int64_t x;
    switch (sp.type[t]) {
      case SB_BOOL: x = ((const uint8_t *)sp.data[t])[row]; break;
      case SB_INT8: x = ((const int8_t *)sp.data[t])[row]; break;
      case SB_INT16: x = ((const int16_t *)sp.data[t])[row]; break;
      case SB_INT32: case SB_DATE32: case SB_FLOAT32: x = ((const int32_t *)sp.data[t])[row]; break;
      default: x = ((const int64_t *)sp.data[t])[row]; break;
    }
    int c;
    if (sp.f64[t]) {   // SQLOrderingUtil.compareDoubles: NaN equals NaN and is larger than anything else
      const double y = __longlong_as_double(sp.lit[t]);
      const double d = sp.type[t] == SB_FLOAT32 ? (double)__int_as_float((int32_t)x) : __longlong_as_double(x);
      const bool dn = d != d, yn = y != y;
      c = d == y ? 0 : (dn || yn) ? (int)dn - (int)yn : (d < y ? -1 : 1);
    } else {
      c = x == sp.lit[t] ? 0 : (x < sp.lit[t] ? -1 : 1);
    }
    bool ok;
    switch (op) {
      case SP_EQ: ok = c == 0; break;
      case SP_NE: ok = c != 0; break;
      case SP_LT: ok = c < 0; break;
      case SP_LE: ok = c <= 0; break;
      case SP_GT: ok = c > 0; break;
      default: ok = c >= 0; break;
    }
    keep = keep && ok;
  }
  return keep;
}
#endif

}  // namespace sb
