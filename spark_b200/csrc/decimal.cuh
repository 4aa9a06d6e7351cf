// decimal.cuh -- SUM / AVG over decimal columns (Sum.scala:80-178, Average.scala:80-135 for DecimalType), csrc/decimal.cu
#pragma once
#include "common.cuh"

namespace sb {

// does the plan aggregate a decimal column with SUM / AVG (update modes) or carry decimal sum buffers (merge modes)?
bool plan_has_decimal_sums(const sb_table *in, const sb_agg_plan *plan);
// Runs such a plan: the decimal aggregates are rewritten into int64 limb sums + a count, `run` (the fixed-width aggregate) does
// the grouping, and the limbs are composed back into decimal(p + 10, s) sums / decimal(p + 4, s + 4) averages.
typedef void (*AggregateFn)(const sb_table *in, const sb_agg_plan *plan, cudaStream_t st, sb_table **out);
void hash_aggregate_decimals(const sb_table *in, const sb_agg_plan *plan, cudaStream_t st, AggregateFn run, sb_table **out);

}  // namespace sb
