/*
 * sb_synth.h -- the synthetic TPC-H-shaped dataset of bench.py and the parity tests, as a pure function of
 * (seed, table, column, row).  The reference ships the TPC-H queries, schema and plan goldens
 * (sql/core/src/test/resources/tpch/q{1,3,5}.sql, sql/core/src/test/scala/org/apache/spark/sql/TPCHBase.scala:36-92)
 * but no data generator, so the value distributions follow the TPC-H specification:
 *   - every 7 consecutive orders carry 1,2,...,7 line items (4 per order on average, like dbgen);
 *   - o_orderkey is sparse (8 of every 32 keys), a third of the customers have no orders;
 *   - l_shipdate = o_orderdate + [1,121], l_commitdate = o_orderdate + [30,90], l_receiptdate = l_shipdate + [1,30];
 *   - l_returnflag = 'R'|'A' when the receipt date is past, else 'N'; l_linestatus = 'O' when shipped after
 *     CURRENTDATE (1995-06-17), else 'F';
 *   - l_extendedprice = l_quantity * part retail price (cents / 100.0), discount 0.00..0.10, tax 0.00..0.08.
 * Money / quantity columns are DOUBLE (north_star's 1e-6 tolerance for floating SUM/AVG), keys BIGINT, dates
 * DATE (int32 days), one-character flags and the market segment are 1-byte codes.
 *
 * Because a value depends only on its coordinates (counter-based generator), the GPU (csrc/synth.cu) and the
 * CPU baseline (oracle/spark_oracle.c, OpenMP, parallel first touch) fill identical columns independently and
 * in any chunking; tests compare the two bit for bit.  Plain C, usable from CUDA (SBS_FN = __host__ __device__).
 */
#ifndef SB_SYNTH_H
#define SB_SYNTH_H

#include <stdint.h>

#ifndef SBS_FN
#ifdef __CUDACC__
#define SBS_FN static __host__ __device__ __forceinline__
#else
#define SBS_FN static inline
#endif
#endif

/* tables */
#define SB_SYNTH_LINEITEM 1
#define SB_SYNTH_ORDERS 2
#define SB_SYNTH_CUSTOMER 3
#define SB_SYNTH_SUPPLIER 4

/* lineitem columns (type, bytes): the 74 B/row fixed-width row of BASELINE.json configs[3] */
#define SB_L_ORDERKEY 0      /* int64 */
#define SB_L_PARTKEY 1       /* int64 */
#define SB_L_SUPPKEY 2       /* int64 */
#define SB_L_LINENUMBER 3    /* int32 */
#define SB_L_QUANTITY 4      /* float64 */
#define SB_L_EXTENDEDPRICE 5 /* float64 */
#define SB_L_DISCOUNT 6      /* float64 */
#define SB_L_TAX 7           /* float64 */
#define SB_L_RETURNFLAG 8    /* int8 */
#define SB_L_LINESTATUS 9    /* int8 */
#define SB_L_SHIPDATE 10     /* date32 */
#define SB_L_COMMITDATE 11   /* date32 */
#define SB_L_RECEIPTDATE 12  /* date32 */
#define SB_L_NCOLS 13
/* orders */
#define SB_O_ORDERKEY 0      /* int64 */
#define SB_O_CUSTKEY 1       /* int64 */
#define SB_O_ORDERDATE 2     /* date32 */
#define SB_O_SHIPPRIORITY 3  /* int32 */
#define SB_O_NCOLS 4
/* customer */
#define SB_C_CUSTKEY 0       /* int64 */
#define SB_C_MKTSEGMENT 1    /* int8 code 0..4 */
#define SB_C_NATIONKEY 2     /* int64 */
#define SB_C_NCOLS 3
/* supplier */
#define SB_S_SUPPKEY 0       /* int64 */
#define SB_S_NATIONKEY 1     /* int64 */
#define SB_S_NCOLS 2

#define SBS_ORDERDATE_MIN 8035   /* 1992-01-01 */
#define SBS_ORDERDATE_MAX 10440  /* 1998-08-02 */
#define SBS_CURRENT_DATE 9298    /* 1995-06-17 */

/* table cardinalities from the number of orders (TPC-H: orders = 1.5 M x SF) */
SBS_FN int64_t sbs_lineitem_rows(int64_t n_orders) {
  const int cum[7] = {0, 1, 3, 6, 10, 15, 21};
  return (n_orders / 7) * 28 + cum[n_orders % 7];
}
SBS_FN int64_t sbs_customer_rows(int64_t n_orders) { int64_t n = n_orders / 10; return n < 1 ? 1 : n; }
SBS_FN int64_t sbs_supplier_rows(int64_t n_orders) { int64_t n = n_orders / 150; return n < 1 ? 1 : n; }

/* counter-based generator: splitmix64 finaliser of (seed, stream, index) */
SBS_FN uint64_t sbs_mix(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
SBS_FN uint32_t sbs_draw(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t m) {   /* uniform in [0, m) */
  uint64_t x = sbs_mix(seed + 0x9e3779b97f4a7c15ull * (idx + 1) + ((uint64_t)stream << 56) + stream);
  return (uint32_t)(((x >> 32) * (uint64_t)m) >> 32);
}

/* lineitem row -> (order index, line number) */
SBS_FN int64_t sbs_line_order(int64_t row, int32_t *linenumber) {
  int64_t b = row / 28;
  int w = (int)(row % 28);
  int j = w >= 21 ? 6 : w >= 15 ? 5 : w >= 10 ? 4 : w >= 6 ? 3 : w >= 3 ? 2 : w >= 1 ? 1 : 0;
  const int cum[7] = {0, 1, 3, 6, 10, 15, 21};
  *linenumber = w - cum[j] + 1;
  return b * 7 + j;
}
SBS_FN int64_t sbs_orderkey(int64_t o) { return (o / 8) * 32 + (o % 8) + 1; }
SBS_FN int32_t sbs_orderdate(uint64_t seed, int64_t o) {
  return SBS_ORDERDATE_MIN + (int32_t)sbs_draw(seed, 20, (uint64_t)o, SBS_ORDERDATE_MAX - SBS_ORDERDATE_MIN + 1);
}
SBS_FN int64_t sbs_custkey(uint64_t seed, int64_t o, int64_t n_cust) {
  int64_t c = 1 + (int64_t)(sbs_mix(seed + 0x51ed270b7f4a7c15ull * (uint64_t)(o + 1) + 21) % (uint64_t)n_cust);
  if (c % 3 == 0) c = c > 1 ? c - 1 : 1;   /* a third of the customers never order */
  return c;
}

/* One value, widened: integers / dates / codes as int64, doubles as their value.  `n_orders` fixes the key domains. */
SBS_FN int64_t sbs_lineitem_i64(uint64_t seed, int col, int64_t row, int64_t n_orders) {
  int32_t ln;
  int64_t o = sbs_line_order(row, &ln);
  switch (col) {
    case SB_L_ORDERKEY: return sbs_orderkey(o);
    case SB_L_PARTKEY: {
      int64_t np = n_orders * 2 / 15; if (np < 1) np = 1;   /* part = 200 K x SF */
      return 1 + (int64_t)(sbs_mix(seed + 0x2545f4914f6cdd1dull * (uint64_t)(row + 1) + 1) % (uint64_t)np);
    }
    case SB_L_SUPPKEY:
      return 1 + (int64_t)(sbs_mix(seed + 0x2545f4914f6cdd1dull * (uint64_t)(row + 1) + 2) % (uint64_t)sbs_supplier_rows(n_orders));
    case SB_L_LINENUMBER: return ln;
    case SB_L_SHIPDATE: return sbs_orderdate(seed, o) + 1 + (int32_t)sbs_draw(seed, 10, (uint64_t)row, 121);
    case SB_L_COMMITDATE: return sbs_orderdate(seed, o) + 30 + (int32_t)sbs_draw(seed, 11, (uint64_t)row, 61);
    case SB_L_RECEIPTDATE:
      return sbs_orderdate(seed, o) + 1 + (int32_t)sbs_draw(seed, 10, (uint64_t)row, 121) + 1 + (int32_t)sbs_draw(seed, 12, (uint64_t)row, 30);
    case SB_L_RETURNFLAG: {
      int32_t receipt = sbs_orderdate(seed, o) + 1 + (int32_t)sbs_draw(seed, 10, (uint64_t)row, 121) + 1 + (int32_t)sbs_draw(seed, 12, (uint64_t)row, 30);
      if (receipt > SBS_CURRENT_DATE) return 'N';
      return sbs_draw(seed, 8, (uint64_t)row, 2) ? 'A' : 'R';
    }
    case SB_L_LINESTATUS: {
      int32_t ship = sbs_orderdate(seed, o) + 1 + (int32_t)sbs_draw(seed, 10, (uint64_t)row, 121);
      return ship > SBS_CURRENT_DATE ? 'O' : 'F';
    }
    default: return 0;
  }
}
SBS_FN double sbs_lineitem_f64(uint64_t seed, int col, int64_t row) {
  switch (col) {
    case SB_L_QUANTITY: return (double)(1 + sbs_draw(seed, 4, (uint64_t)row, 50));
    case SB_L_EXTENDEDPRICE: {
      int64_t q = 1 + sbs_draw(seed, 4, (uint64_t)row, 50);
      int64_t retail = 90000 + sbs_draw(seed, 5, (uint64_t)row, 110001);   /* cents */
      return (double)(q * retail) / 100.0;
    }
    case SB_L_DISCOUNT: return (double)sbs_draw(seed, 6, (uint64_t)row, 11) / 100.0;
    case SB_L_TAX: return (double)sbs_draw(seed, 7, (uint64_t)row, 9) / 100.0;
    default: return 0.0;
  }
}
SBS_FN int sbs_lineitem_is_f64(int col) { return col >= SB_L_QUANTITY && col <= SB_L_TAX; }

SBS_FN int64_t sbs_orders_i64(uint64_t seed, int col, int64_t o, int64_t n_orders) {
  switch (col) {
    case SB_O_ORDERKEY: return sbs_orderkey(o);
    case SB_O_CUSTKEY: return sbs_custkey(seed, o, sbs_customer_rows(n_orders));
    case SB_O_ORDERDATE: return sbs_orderdate(seed, o);
    default: return 0;   /* o_shippriority is 0 in TPC-H */
  }
}
SBS_FN int64_t sbs_customer_i64(uint64_t seed, int col, int64_t c) {
  switch (col) {
    case SB_C_CUSTKEY: return c + 1;
    case SB_C_MKTSEGMENT: return sbs_draw(seed, 30, (uint64_t)c, 5);
    default: return sbs_draw(seed, 31, (uint64_t)c, 25);
  }
}
SBS_FN int64_t sbs_supplier_i64(uint64_t seed, int col, int64_t s) {
  return col == SB_S_SUPPKEY ? s + 1 : (int64_t)sbs_draw(seed, 40, (uint64_t)s, 25);
}

/* byte width of a synthetic column */
SBS_FN int sbs_width(int table, int col) {
  if (table == SB_SYNTH_LINEITEM) {
    if (col <= SB_L_SUPPKEY) return 8;
    if (col == SB_L_LINENUMBER) return 4;
    if (col <= SB_L_TAX) return 8;
    if (col <= SB_L_LINESTATUS) return 1;
    return 4;
  }
  if (table == SB_SYNTH_ORDERS) return col <= SB_O_CUSTKEY ? 8 : 4;
  if (table == SB_SYNTH_CUSTOMER) return col == SB_C_MKTSEGMENT ? 1 : 8;
  return 8;
}

#endif /* SB_SYNTH_H */
