// simple_pred.cuh -- conjunctions of column-vs-literal comparisons (most pushed-down filters: Q1, Q3, Q5, Q6 ...), shared by the
// FilterExec fast path (expr.cu) and the operators that fuse the filter below them into their own first pass (join.cu).
#pragma once
#include "common.cuh"
#include <limits.h>

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
// keep[j] &&= x[j] <op> lit, the operator chosen once outside the row loop (uniform switch, one compare per row)
template <typename T, int R>
__device__ __forceinline__ void sp_compare_rows(const T (&x)[R], T lit, int op, bool (&keep)[R]) {
  switch (op) {
#define SB_SP_ROWS(COND) _Pragma("unroll") for (int j = 0; j < R; j++) keep[j] = keep[j] && (COND);
    case SP_EQ: SB_SP_ROWS(x[j] == lit) break;
    case SP_NE: SB_SP_ROWS(x[j] != lit) break;
    case SP_LT: SB_SP_ROWS(x[j] < lit) break;
    case SP_LE: SB_SP_ROWS(x[j] <= lit) break;
    case SP_GT: SB_SP_ROWS(x[j] > lit) break;
    default: SB_SP_ROWS(x[j] >= lit) break;
#undef SB_SP_ROWS
  }
}
// the same predicate for R rows row0, row0 + stride, ... (lane-strided kernels: the 32 lanes of a warp read 32 consecutive rows,
// so plain loads coalesce).  Term by term: all R loads of a term are issued before the first comparison uses one, so they overlap;
// type and operator switches are outside the row loops (uniform).  FULL: every row is < n (no clamping); otherwise rows >= n are
// clamped for the loads and masked by the caller.  32-bit columns against a literal that fits compare at 32 bits.
template <int R, bool FULL>
__device__ __forceinline__ void simple_pred_rows(const SimplePred &sp, int64_t row0, int64_t stride, int64_t n, bool (&keep)[R]) {
  int64_t rr[R];
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int64_t r = row0 + j * stride;
    rr[j] = FULL || r < n ? r : n - 1;
  }
#pragma unroll 1
  for (int t = 0; t < sp.nterms; t++) {
    if (sp.valid[t]) {
      uint8_t vb[R];
#pragma unroll
      for (int j = 0; j < R; j++) vb[j] = sp.valid[t][rr[j] >> 3];
#pragma unroll
      for (int j = 0; j < R; j++) keep[j] = keep[j] && ((vb[j] >> (rr[j] & 7)) & 1);
    }
    const int op = sp.op[t];
    if (op == SP_NOTNULL) continue;
    const int64_t lit = sp.lit[t];
    const bool f64 = sp.f64[t] != 0;
    const int32_t ty = sp.type[t];
    if (!f64 && (ty == SB_INT32 || ty == SB_DATE32)) {
      int32_t x32[R];
      const int32_t *p = (const int32_t *)sp.data[t];
#pragma unroll
      for (int j = 0; j < R; j++) x32[j] = p[rr[j]];
      if (lit >= INT32_MIN && lit <= INT32_MAX) {
        sp_compare_rows<int32_t, R>(x32, (int32_t)lit, op, keep);
      } else {
        int64_t x[R];
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = x32[j];
        sp_compare_rows<int64_t, R>(x, lit, op, keep);
      }
      continue;
    }
    int64_t x[R];
    switch (ty) {
      case SB_BOOL: case SB_INT8:
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = ty == SB_BOOL ? (int64_t)((const uint8_t *)sp.data[t])[rr[j]] : (int64_t)((const int8_t *)sp.data[t])[rr[j]];
        break;
      case SB_INT16:
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = ((const int16_t *)sp.data[t])[rr[j]];
        break;
      case SB_INT32: case SB_DATE32: case SB_FLOAT32:
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = ((const int32_t *)sp.data[t])[rr[j]];
        break;
      default:
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = ((const int64_t *)sp.data[t])[rr[j]];
        break;
    }
    if (!f64) {
      sp_compare_rows<int64_t, R>(x, lit, op, keep);
      continue;
    }
    // doubles: SQLOrderingUtil.compareDoubles (NaN equals NaN and is larger than anything else) as a three-way result
    const double y = __longlong_as_double(lit);
    const bool f32 = ty == SB_FLOAT32, yn = y != y;
    int c[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      const double d = f32 ? (double)__int_as_float((int32_t)x[j]) : __longlong_as_double(x[j]);
      const bool dn = d != d;
      c[j] = d == y ? 0 : (dn || yn) ? (int)dn - (int)yn : (d < y ? -1 : 1);
    }
    sp_compare_rows<int, R>(c, 0, op, keep);
  }
}
#endif

}  // namespace sb
