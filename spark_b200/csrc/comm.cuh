// comm.cuh -- what the fused exchange (partition.cu: sb_shuffle_exchange) needs from the communicator in comm.cu
#pragma once
#include <vector>
#include "common.cuh"

namespace sb {

struct CommInfo {
  int rank = 0, nranks = 1;
  bool up = false;   // sb_comm_init done
};
CommInfo comm_info();
// every rank contributes `count` int64 values (host); all[r * count + i] on return.  Collective; synchronises the stream.
void comm_allgather_host(const int64_t *mine, int64_t count, int64_t *all, cudaStream_t st);
// enqueues an 8-byte all-gather on `st` (no host synchronisation): work enqueued after it starts only when every rank has
// reached it -- the "every remote store has landed" edge of the fused exchange
void comm_barrier_enqueue(cudaStream_t st);
// Collective (same `need` on every rank): makes sure every rank owns a receive window of at least `need` bytes mapped by all
// peers; bases[r] = rank r's window as addressable from THIS process (own window: the local pointer).  false: peer mapping is
// unavailable in this environment (decided collectively) -- the caller falls back to sb_hash_partition + sb_all_to_all.
bool comm_window(size_t need, cudaStream_t st, std::vector<void *> &bases);

}  // namespace sb
