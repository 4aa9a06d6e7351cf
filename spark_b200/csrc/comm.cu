// comm.cu -- the exchange between executors (one process per GPU): NCCL over NVLink 5 / NVSwitch.
//
// Replaces, for this path only, the reference's shuffle transport: SortShuffleManager writers ->
// local disk -> Netty fetch (core/src/main/scala/org/apache/spark/shuffle/sort/SortShuffleManager.scala:70,
// core/src/main/scala/org/apache/spark/storage/ShuffleBlockFetcherIterator.scala) and TorrentBroadcast
// (core/src/main/scala/org/apache/spark/broadcast/TorrentBroadcast.scala:60) for the broadcast build side.
// Nothing is serialized, nothing touches disk: column buffers go HBM -> NVLink -> HBM.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that a host process which already loaded
// a copy (e.g. the one bundled with PyTorch) shares it; the unique id travels host-side (driver plugin
// in Spark, torch.distributed/gloo in the tests).
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"
#include "primitives.cuh"

namespace sb {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define SB_SYM(field, sym) api.field = (decltype(api.field))dlsym(api.handle, sym)
    SB_SYM(GetUniqueId, "ncclGetUniqueId");
    SB_SYM(CommInitRank, "ncclCommInitRank");
    SB_SYM(CommDestroy, "ncclCommDestroy");
    SB_SYM(Send, "ncclSend");
    SB_SYM(Recv, "ncclRecv");
    SB_SYM(GroupStart, "ncclGroupStart");
    SB_SYM(GroupEnd, "ncclGroupEnd");
    SB_SYM(AllGather, "ncclAllGather");
    SB_SYM(GetErrorString, "ncclGetErrorString");
#undef SB_SYM
  });
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd ||
      !api.AllGather)
    fail(SB_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
  return api;
}

#define SB_NCCL(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess)                                                                                 \
      ::sb::fail(SB_ERR_NCCL, "%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "?"); \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};
static Comm &comm() {
  static Comm c;
  return c;
}

// contiguous ownership: rank r owns partitions [part_lo(r), part_lo(r+1))
static inline int32_t part_lo(int32_t r, int32_t nparts, int32_t nranks) {
  return (int32_t)(((int64_t)r * nparts + nranks - 1) / nranks);
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_comm_get_unique_id(uint8_t out_id[SB_UNIQUE_ID_BYTES]) {
  SB_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == SB_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  SB_NCCL(nccl().GetUniqueId(&id));
  memcpy(out_id, &id, SB_UNIQUE_ID_BYTES);
  SB_API_END
}

int sb_comm_init(int32_t rank, int32_t nranks, const uint8_t id_bytes[SB_UNIQUE_ID_BYTES]) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d of %d", rank, nranks);
  Comm &c = comm();
  SB_REQUIRE(c.comm == nullptr, "communicator already initialised");
  if (nranks > 1) {
    ncclUniqueId id;
    memcpy(&id, id_bytes, SB_UNIQUE_ID_BYTES);
    SB_NCCL(nccl().CommInitRank(&c.comm, nranks, id, rank));
  }
  c.rank = rank;
  c.nranks = nranks;
  SB_API_END
}

int sb_comm_destroy(void) {
  SB_API_BEGIN
  Comm &c = comm();
  if (c.comm) {
    cudaDeviceSynchronize();
    nccl().CommDestroy(c.comm);
    c.comm = nullptr;
  }
  c.rank = 0;
  c.nranks = 1;
  SB_API_END
}

int sb_comm_rank(int32_t *rank, int32_t *nranks) {
  SB_API_BEGIN
  *rank = comm().rank;
  *nranks = comm().nranks;
  SB_API_END
}

// Host-only: rows this rank sends to every destination rank, given its partition boundaries.
int sb_exchange_plan(const int64_t *part_offsets, int32_t num_partitions, int32_t nranks, int64_t *out_send_rows) {
  SB_API_BEGIN
  SB_REQUIRE(part_offsets && out_send_rows && num_partitions >= 1 && nranks >= 1, "bad argument");
  for (int r = 0; r < nranks; r++) {
    int32_t lo = part_lo(r, num_partitions, nranks), hi = part_lo(r + 1, num_partitions, nranks);
    if (lo > num_partitions) lo = num_partitions;
    if (hi > num_partitions) hi = num_partitions;
    out_send_rows[r] = part_offsets[hi] - part_offsets[lo];
  }
  SB_API_END
}

int sb_all_to_all(const sb_table *in, const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s, sb_table **out,
                  int64_t *out_part_offsets_host) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && part_offsets_host && out && out_part_offsets_host, "null argument");
  Comm &c = comm();
  SB_REQUIRE(c.nranks > 1 && c.comm, "sb_all_to_all needs an initialised communicator with more than one rank");
  cudaStream_t st = stream_of(s);
  const int R = c.nranks;
  const int P = num_partitions;
  // 1. every rank learns every rank's per-partition row counts
  std::vector<int64_t> my_counts(P), all_counts((size_t)R * P);
  for (int p = 0; p < P; p++) my_counts[p] = part_offsets_host[p + 1] - part_offsets_host[p];
  Scratch d_my(P * 8, st), d_all((int64_t)R * P * 8, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, my_counts.data(), (size_t)P * 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, (size_t)P, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(all_counts.data(), d_all.ptr, (size_t)R * P * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  // 2. layout of what this rank receives: [source rank][owned partitions]
  const int lo = part_lo(c.rank, P, R), hi = part_lo(c.rank + 1, P, R);
  std::vector<int64_t> recv_rows(R), recv_off(R + 1, 0), send_rows(R), send_off(R);
  for (int src = 0; src < R; src++) {
    int64_t rows = 0;
    for (int p = lo; p < hi; p++) rows += all_counts[(size_t)src * P + p];
    recv_rows[src] = rows;
    recv_off[src + 1] = recv_off[src] + rows;
  }
  for (int d = 0; d < R; d++) {
    int dl = part_lo(d, P, R), dh = part_lo(d + 1, P, R);
    send_off[d] = part_offsets_host[dl];
    send_rows[d] = part_offsets_host[dh] - part_offsets_host[dl];
  }
  const int64_t nrecv = recv_off[R];
  // received partition sizes (summed over sources) for the caller
  out_part_offsets_host[0] = 0;
  for (int p = 0; p < P; p++) {
    int64_t rows = 0;
    if (p >= lo && p < hi)
      for (int src = 0; src < R; src++) rows += all_counts[(size_t)src * P + p];
    out_part_offsets_host[p + 1] = out_part_offsets_host[p] + rows;
  }
  // 3. one grouped send/recv per (column buffer, peer)
  sb_table *t = table_new(nrecv);
  try {
    std::vector<Scratch *> temps;
    struct Pending { Column *col; uint8_t *recv_bytes; };
    std::vector<Pending> pend;
    t->cols.reserve(in->cols.size());
    for (auto &col : in->cols) {
      if (col.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "all-to-all of string columns is not implemented (dictionary-encode them)");
      t->cols.push_back(column_alloc(col.type, col.scale, nrecv, col.validity != nullptr, st));
    }
    SB_NCCL(nccl().GroupStart());
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      const Column &src = in->cols[ci];
      Column &dst = t->cols[ci];
      const int w = type_width(src.type);
      for (int peer = 0; peer < R; peer++) {
        if (send_rows[peer] > 0)
          SB_NCCL(nccl().Send((const char *)src.d() + send_off[peer] * w, (size_t)(send_rows[peer] * w), ncclUint8, peer, c.comm, st));
        if (recv_rows[peer] > 0)
          SB_NCCL(nccl().Recv((char *)dst.data->ptr + recv_off[peer] * w, (size_t)(recv_rows[peer] * w), ncclUint8, peer, c.comm, st));
      }
      if (src.validity) {
        Scratch *sb = new Scratch(in->nrows + 16, st), *rb = new Scratch(nrecv + 16, st);
        temps.push_back(sb);
        temps.push_back(rb);
        bitmap_to_bytes(src.v(), in->nrows, sb->as<uint8_t>(), st);
        for (int peer = 0; peer < R; peer++) {
          if (send_rows[peer] > 0) SB_NCCL(nccl().Send(sb->as<uint8_t>() + send_off[peer], (size_t)send_rows[peer], ncclUint8, peer, c.comm, st));
          if (recv_rows[peer] > 0) SB_NCCL(nccl().Recv(rb->as<uint8_t>() + recv_off[peer], (size_t)recv_rows[peer], ncclUint8, peer, c.comm, st));
        }
        pend.push_back({&dst, rb->as<uint8_t>()});
      }
    }
    SB_NCCL(nccl().GroupEnd());
    count_launch();   // the grouped NCCL kernel
    for (auto &p : pend) {
      bytes_to_bitmap(p.recv_bytes, nrecv, (uint32_t *)p.col->validity->ptr, st);
    }
    SB_CUDA(cudaStreamSynchronize(st));
    for (auto *x : temps) delete x;
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

int sb_all_gather(const sb_table *in, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out, "null argument");
  Comm &c = comm();
  cudaStream_t st = stream_of(s);
  if (c.nranks <= 1 || !c.comm) {   // single executor: the broadcast is the table itself
    sb_table *t = table_new(in->nrows);
    for (auto &col : in->cols) t->cols.push_back(column_share(col));
    *out = t;
    return SB_OK;
  }
  const int R = c.nranks;
  int64_t my = in->nrows;
  std::vector<int64_t> rows(R), off(R + 1, 0);
  Scratch d_my(8, st), d_all(R * 8, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, &my, 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, 1, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(rows.data(), d_all.ptr, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < R; r++) off[r + 1] = off[r] + rows[r];
  const int64_t total = off[R];
  sb_table *t = table_new(total);
  try {
    std::vector<Scratch *> temps;
    struct Pending { Column *col; uint8_t *bytes; };
    std::vector<Pending> pend;
    t->cols.reserve(in->cols.size());
    for (auto &col : in->cols) {
      if (col.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "all-gather of string columns is not implemented (dictionary-encode them)");
      t->cols.push_back(column_alloc(col.type, col.scale, total, col.validity != nullptr, st));
    }
    SB_NCCL(nccl().GroupStart());
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      const Column &src = in->cols[ci];
      Column &dst = t->cols[ci];
      const int w = type_width(src.type);
      for (int peer = 0; peer < R; peer++) {
        if (my > 0) SB_NCCL(nccl().Send(src.d(), (size_t)(my * w), ncclUint8, peer, c.comm, st));
        if (rows[peer] > 0) SB_NCCL(nccl().Recv((char *)dst.data->ptr + off[peer] * w, (size_t)(rows[peer] * w), ncclUint8, peer, c.comm, st));
      }
      if (src.validity) {
        Scratch *sb = new Scratch(my + 16, st), *rb = new Scratch(total + 16, st);
        temps.push_back(sb);
        temps.push_back(rb);
        bitmap_to_bytes(src.v(), my, sb->as<uint8_t>(), st);
        for (int peer = 0; peer < R; peer++) {
          if (my > 0) SB_NCCL(nccl().Send(sb->as<uint8_t>(), (size_t)my, ncclUint8, peer, c.comm, st));
          if (rows[peer] > 0) SB_NCCL(nccl().Recv(rb->as<uint8_t>() + off[peer], (size_t)rows[peer], ncclUint8, peer, c.comm, st));
        }
        pend.push_back({&dst, rb->as<uint8_t>()});
      }
    }
    SB_NCCL(nccl().GroupEnd());
    count_launch();
    for (auto &p : pend) {
      bytes_to_bitmap(p.bytes, total, (uint32_t *)p.col->validity->ptr, st);
    }
    SB_CUDA(cudaStreamSynchronize(st));
    for (auto *x : temps) delete x;
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

}  // extern "C"
