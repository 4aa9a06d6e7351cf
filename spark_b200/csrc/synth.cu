// synth.cu -- fills HBM columns with the synthetic TPC-H-shaped dataset of include/sb_synth.h (a pure function of
// (seed, table, column, row), so the CPU baseline regenerates the very same values on the host).  SURVEY.md 8d:
// the reference has no data generator; BASELINE.json configs[3] asks for rows "generated on-device per GPU shard
// from a counter-based RNG".  This is workload plumbing for bench.py / tests, not an operator of the path.
#include "common.cuh"
#include "../../include/sb_synth.h"

namespace sb {

template <typename T>
__global__ void synth_kernel(int table, int col, uint64_t seed, int64_t n_orders, int64_t first, int64_t n, T *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = first + i;
    if (table == SB_SYNTH_LINEITEM) {
      if (sbs_lineitem_is_f64(col)) out[i] = (T)sbs_lineitem_f64(seed, col, row);
      else out[i] = (T)sbs_lineitem_i64(seed, col, row, n_orders);
    } else if (table == SB_SYNTH_ORDERS) out[i] = (T)sbs_orders_i64(seed, col, row, n_orders);
    else if (table == SB_SYNTH_CUSTOMER) out[i] = (T)sbs_customer_i64(seed, col, row);
    else out[i] = (T)sbs_supplier_i64(seed, col, row);
  }
}

static int32_t synth_type(int table, int col) {
  if (table == SB_SYNTH_LINEITEM) {
    if (col <= SB_L_SUPPKEY) return SB_INT64;
    if (col == SB_L_LINENUMBER) return SB_INT32;
    if (col <= SB_L_TAX) return SB_FLOAT64;
    if (col <= SB_L_LINESTATUS) return SB_INT8;
    return SB_DATE32;
  }
  if (table == SB_SYNTH_ORDERS) return col <= SB_O_CUSTKEY ? SB_INT64 : (col == SB_O_ORDERDATE ? SB_DATE32 : SB_INT32);
  if (table == SB_SYNTH_CUSTOMER) return col == SB_C_MKTSEGMENT ? SB_INT8 : SB_INT64;
  return SB_INT64;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_synth_table(int32_t table, const int32_t *columns, int32_t ncols, int64_t n_orders, int64_t first_row,
                              int64_t nrows, uint64_t seed, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(out && columns && ncols > 0, "null argument");
  SB_REQUIRE(table >= SB_SYNTH_LINEITEM && table <= SB_SYNTH_SUPPLIER, "unknown synthetic table %d", table);
  const int maxcol = table == SB_SYNTH_LINEITEM ? SB_L_NCOLS : table == SB_SYNTH_ORDERS ? SB_O_NCOLS : table == SB_SYNTH_CUSTOMER ? SB_C_NCOLS : SB_S_NCOLS;
  const int64_t total = table == SB_SYNTH_LINEITEM ? sbs_lineitem_rows(n_orders) : table == SB_SYNTH_ORDERS ? n_orders
                        : table == SB_SYNTH_CUSTOMER ? sbs_customer_rows(n_orders) : sbs_supplier_rows(n_orders);
  SB_REQUIRE(n_orders > 0 && first_row >= 0 && nrows >= 0 && first_row + nrows <= total, "rows [%lld, %lld) outside the table (%lld rows)",
             (long long)first_row, (long long)(first_row + nrows), (long long)total);
  cudaStream_t st = stream_of(s);
  sb_table *t = table_new(nrows);
  try {
    for (int i = 0; i < ncols; i++) {
      const int col = columns[i];
      SB_REQUIRE(col >= 0 && col < maxcol, "column %d out of range for synthetic table %d", col, table);
      Column c = column_alloc(synth_type(table, col), 0, nrows, false, st);
      t->cols.push_back(c);
      if (nrows == 0) continue;
      const int grid = grid_for(nrows, 256 * 4, rt().num_sms * 8);
      switch (type_width(c.type)) {
        case 1: synth_kernel<int8_t><<<grid, 256, 0, st>>>(table, col, seed, n_orders, first_row, nrows, (int8_t *)c.data->ptr); break;
        case 4: synth_kernel<int32_t><<<grid, 256, 0, st>>>(table, col, seed, n_orders, first_row, nrows, (int32_t *)c.data->ptr); break;
        default:
          if (c.type == SB_FLOAT64) synth_kernel<double><<<grid, 256, 0, st>>>(table, col, seed, n_orders, first_row, nrows, (double *)c.data->ptr);
          else synth_kernel<int64_t><<<grid, 256, 0, st>>>(table, col, seed, n_orders, first_row, nrows, (int64_t *)c.data->ptr);
      }
      SB_LAUNCH_CHECK();
    }
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}
