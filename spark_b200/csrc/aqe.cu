// aqe.cu -- the runtime statistics adaptive query execution reads from an exchange, and the partition coalescing it drives.
//
// Reference: ShuffleExchangeExec.mapOutputStatisticsFuture / runtimeStatistics (SQLX/exchange/ShuffleExchangeExec.scala:235-262:
// MapOutputStatistics.bytesByPartitionId summed over the map tasks), CoalesceShufflePartitions -> ShufflePartitionsUtil
// (SQLX/adaptive/ShufflePartitionsUtil.scala:45-90 coalescePartitions, :91-126 coalescePartitionsWithoutSkew, :263-349
// coalescePartitionsAndGetSpecs, :351-369 attachDataSize), read back through AQEShuffleReadExec (CoalescedPartitionSpec).
// The statistics come from the per-partition row counts the exchange already all-gathers; the coalescing walk is host code
// (it is a few hundred iterations over partition sizes), restated line by line and pinned by ShufflePartitionsUtilSuite's vectors.
#include <algorithm>
#include <cmath>
#include "common.cuh"

using namespace sb;

extern "C" int sb_coalesce_partitions(const int64_t *const *bytes_by_partition, int32_t nshuffles, int32_t num_partitions,
                                      int64_t advisory_target_size, int32_t min_num_partitions, int64_t min_partition_size,
                                      int32_t max_reducer_partitions_per_task, int32_t *out_start, int32_t *out_end, int64_t *out_data_size,
                                      int32_t *out_nspecs) {
  SB_API_BEGIN
  SB_REQUIRE(out_start && out_end && out_nspecs && nshuffles >= 0 && num_partitions >= 0, "bad argument");
  *out_nspecs = 0;
  if (nshuffles == 0) return SB_OK;
  SB_REQUIRE(bytes_by_partition && min_num_partitions >= 1, "bad argument");
  const int maxr = max_reducer_partitions_per_task > 0 ? max_reducer_partitions_per_task : INT32_MAX;
  // coalescePartitions :60-66: the target shrinks when minNumPartitions asks for more parallelism, never below minPartitionSize
  long double total = 0;
  for (int s = 0; s < nshuffles; s++)
    for (int p = 0; p < num_partitions; p++) total += (long double)bytes_by_partition[s][p];
  const int64_t max_target = (int64_t)std::ceil((double)total / (double)min_num_partitions);
  const int64_t target = std::max(std::min(max_target, advisory_target_size), min_partition_size);
  // coalescePartitionsAndGetSpecs(0, numPartitions, ...) :263-349
  std::vector<int32_t> st, en;
  int64_t coalesced = 0, latest_size = 0;
  int i = 0, latest_split = 0;
  auto create = [&](bool force) {
    if (coalesced > 0 || force) { st.push_back(latest_split); en.push_back(i); }
  };
  auto within = [&](int a, int b) { return b - a <= maxr; };
  while (i < num_partitions) {
    int64_t cur = 0;
    for (int s = 0; s < nshuffles; s++) cur += bytes_by_partition[s][i];
    if (i > latest_split && i - latest_split >= maxr) {
      create(false);
      latest_split = i;
      latest_size = coalesced;
      coalesced = cur;
    } else if (i > latest_split && coalesced + cur > target) {
      if (coalesced < min_partition_size) {
        if (latest_size > 0 && latest_size < cur && !st.empty() && within(st.back(), i)) {   // pack with the partition before
          en.back() = i;
          latest_split = i;
          latest_size += coalesced;
          coalesced = cur;
        } else coalesced += cur;                                                              // pack with the one after
      } else {
        create(false);
        latest_split = i;
        latest_size = coalesced;
        coalesced = cur;
      }
    } else coalesced += cur;
    i++;
  }
  if (coalesced < min_partition_size && latest_size > 0 && !st.empty() && within(st.back(), num_partitions)) en.back() = num_partitions;
  else create(st.empty());
  // coalescePartitionsWithoutSkew :121-125: nothing to do when the layout did not shrink
  if ((int)st.size() >= num_partitions) return SB_OK;
  for (size_t k = 0; k < st.size(); k++) {
    out_start[k] = st[k];
    out_end[k] = en[k];
    if (out_data_size)                                    // attachDataSize :351-369: per shuffle, the bytes of the range
      for (int s = 0; s < nshuffles; s++) {
        int64_t sz = 0;
        for (int p = st[k]; p < en[k]; p++) sz += bytes_by_partition[s][p];
        out_data_size[(size_t)s * st.size() + k] = sz;
      }
  }
  *out_nspecs = (int32_t)st.size();
  SB_API_END
}
