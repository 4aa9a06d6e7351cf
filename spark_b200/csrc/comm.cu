// comm.cu -- the exchange between executors (one process per GPU) over NVLink 5 / NVSwitch: copy engines push bucket slices
// into peer-mapped receive windows (CUDA IPC), NCCL carries the control messages and is the fallback data path.
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
#include <algorithm>
#include <memory>
#include <nccl.h>
#include "common.cuh"
#include "primitives.cuh"
#include "rtc.cuh"
#include "comm.cuh"
#include "strings.cuh"

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

// ---- peer windows: the exchange's data path over NVLink / NVSwitch ---------------------------------------------------------------
// NCCL send/recv moves unregistered buffers through its staging FIFOs (measured here: 124 GB/s per direction between two B200s
// with 32 channels).  The buckets are large contiguous slices, so the copy engines can push them straight into the destination
// GPU's memory: every rank owns one receive WINDOW (cudaMalloc + CUDA IPC, mapped once by every peer), a sender writes each
// destination's slice of every column at the offset the receiver's layout dictates (all ranks know all counts, so all layouts
// are computed locally), and NCCL is only the control plane: the counts all-gather that opens an exchange doubles as "every
// rank is done with the previous window", a one-word all-gather after the pushes as "every push has landed".
struct PeerWindow {
  void *local = nullptr;
  size_t bytes = 0;
  std::vector<void *> remote;   // remote[r]: rank r's window mapped into this process (nullptr for self)
  std::vector<cudaStream_t> push_streams;
  cudaEvent_t ev_ready = nullptr;
  std::vector<cudaEvent_t> ev_pushed;
  bool disabled = false;        // IPC not available in this environment: the NCCL data path is used
};
static PeerWindow &window() {
  static PeerWindow w;
  return w;
}

// collective: an 8-byte all-gather every rank has to join
static void nccl_barrier(cudaStream_t st) {
  Comm &c = comm();
  Scratch a(8, st), b(8 * c.nranks, st);
  SB_CUDA(cudaMemsetAsync(a.ptr, 0, 8, st));
  SB_NCCL(nccl().AllGather(a.ptr, b.ptr, 1, ncclInt64, c.comm, st));
  SB_CUDA(cudaStreamSynchronize(st));
}

// collective: nobody frees a window a peer still has mapped
static void window_release(cudaStream_t st) {
  PeerWindow &w = window();
  const bool had = w.local != nullptr;
  for (void *p : w.remote)
    if (p) cudaIpcCloseMemHandle(p);
  w.remote.clear();
  if (had && comm().comm) nccl_barrier(st);
  if (w.local) cudaFree(w.local);
  w.local = nullptr;
  w.bytes = 0;
}

// Collective: every rank calls it with the same `need` (derived from the all-gathered counts).  Returns false when the peer
// path is unavailable on ANY rank (decided collectively, so all ranks take the same branch).
static bool window_ensure(size_t need, cudaStream_t st) {
  PeerWindow &w = window();
  Comm &c = comm();
  if (w.disabled) return false;
  if (w.bytes >= need && !w.remote.empty()) return true;
  const int R = c.nranks;
  SB_CUDA(cudaStreamSynchronize(st));
  window_release(st);
  size_t want = need + need / 2 + (64u << 20);
  want = (want + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  struct Msg { cudaIpcMemHandle_t h; int64_t ok; };
  Msg mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = cudaMalloc(&w.local, want) == cudaSuccess && cudaIpcGetMemHandle(&mine.h, w.local) == cudaSuccess;
  if (!mine.ok) cudaGetLastError();
  std::vector<Msg> all(R);
  Scratch d_my(sizeof(Msg), st), d_all(sizeof(Msg) * R, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, &mine, sizeof(Msg), cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, sizeof(Msg), ncclUint8, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(all.data(), d_all.ptr, sizeof(Msg) * R, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  int64_t ok = 1;
  for (int r = 0; r < R; r++) ok &= all[r].ok;
  w.remote.assign(R, nullptr);
  if (ok)
    for (int r = 0; r < R && ok; r++) {
      if (r == c.rank) continue;
      if (cudaIpcOpenMemHandle(&w.remote[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        w.remote[r] = nullptr;
        ok = 0;
      }
    }
  // second round: did every rank manage to map every peer?
  int64_t mapped = ok;
  std::vector<int64_t> all_mapped(R);
  Scratch d_m(8, st), d_ma(8 * R, st);
  SB_CUDA(cudaMemcpyAsync(d_m.ptr, &mapped, 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_m.ptr, d_ma.ptr, 1, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(all_mapped.data(), d_ma.ptr, 8 * R, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < R; r++) ok &= all_mapped[r];
  if (!ok) {
    window_release(st);
    w.disabled = true;
    return false;
  }
  w.bytes = want;
  if (w.push_streams.empty()) {
    w.push_streams.resize(R, nullptr);
    w.ev_pushed.resize(R, nullptr);
    SB_CUDA(cudaEventCreateWithFlags(&w.ev_ready, cudaEventDisableTiming));
    for (int r = 0; r < R; r++) {
      SB_CUDA(cudaStreamCreateWithFlags(&w.push_streams[r], cudaStreamNonBlocking));
      SB_CUDA(cudaEventCreateWithFlags(&w.ev_pushed[r], cudaEventDisableTiming));
    }
  }
  return true;
}

// ---- small-table all-gather: one fixed-size packed message per rank ------------------------------------------------------------
// The broadcast of a small relation or of partial-aggregate rows (Q1: 4 rows x 12 columns per rank) is latency, not bandwidth:
// count exchange + one send/recv per (column, peer) cost ~150 us.  Instead every rank packs [row count | every column padded to
// AG_PACK_ROWS rows | one validity byte per row and column] into one buffer, ONE ncclAllGather moves it, and one kernel unpacks.
// The protocol is optimistic and deterministic: all ranks read all headers; if any rank had more rows than fit, all of them take
// the general path afterwards.
constexpr int64_t AG_PACK_ROWS = 256;
constexpr int AG_MAX_COLS = 64;
struct AgPackArgs {
  int32_t ncols, nranks;
  int64_t my_rows, stride;                  // bytes per rank message
  uint64_t my_valid_mask;                   // bit c: this rank's column c carries a validity bitmap
  const void *src[AG_MAX_COLS];             // pack: this rank's columns
  const uint8_t *src_valid[AG_MAX_COLS];
  void *dst[AG_MAX_COLS];                   // unpack: output columns
  uint8_t *dst_valid_bytes[AG_MAX_COLS];    // unpack: one byte per output row (nullable columns) or nullptr
  int32_t width[AG_MAX_COLS];
  int64_t data_off[AG_MAX_COLS], valid_off[AG_MAX_COLS];
  int64_t rows[16], off[16];                // unpack: rows of every rank and their first output row (nranks <= 16)
};
__global__ void ag_pack_kernel(const __grid_constant__ AgPackArgs a, uint8_t *__restrict__ msg) {
  const int c = blockIdx.x;
  if (c == a.ncols) {
    if (threadIdx.x == 0) {
      ((int64_t *)msg)[0] = a.my_rows;
      ((uint64_t *)msg)[1] = a.my_valid_mask;
    }
    return;
  }
  const int64_t rows = a.my_rows < AG_PACK_ROWS ? a.my_rows : AG_PACK_ROWS;
  const int64_t bytes = rows * a.width[c];
  for (int64_t i = threadIdx.x; i < bytes; i += blockDim.x) msg[a.data_off[c] + i] = ((const uint8_t *)a.src[c])[i];
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) msg[a.valid_off[c] + i] = bit_valid(a.src_valid[c], i) ? 1 : 0;
}
__global__ void ag_unpack_kernel(const __grid_constant__ AgPackArgs a, const uint8_t *__restrict__ all) {
  const int c = blockIdx.x, r = blockIdx.y;
  const uint8_t *msg = all + (int64_t)r * a.stride;
  const int64_t bytes = a.rows[r] * a.width[c];
  uint8_t *out = (uint8_t *)a.dst[c] + a.off[r] * a.width[c];
  for (int64_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = msg[a.data_off[c] + i];
  if (a.dst_valid_bytes[c])
    for (int64_t i = threadIdx.x; i < a.rows[r]; i += blockDim.x) a.dst_valid_bytes[c][a.off[r] + i] = msg[a.valid_off[c] + i];
}

CommInfo comm_info() {
  Comm &c = comm();
  CommInfo i;
  i.rank = c.rank;
  i.nranks = c.nranks;
  i.up = c.comm != nullptr;
  return i;
}
void comm_allgather_host(const int64_t *mine, int64_t count, int64_t *all, cudaStream_t st) {
  Comm &c = comm();
  if (!c.comm) {
    memcpy(all, mine, (size_t)count * 8);
    return;
  }
  Scratch d_my(count * 8 + 8, st), d_all((int64_t)c.nranks * count * 8 + 8, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, mine, (size_t)count * 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, (size_t)count, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(all, d_all.ptr, (size_t)c.nranks * count * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
}
void comm_barrier_enqueue(cudaStream_t st) {
  Comm &c = comm();
  if (!c.comm) return;
  Scratch a(8, st), b(8 * c.nranks, st);
  SB_CUDA(cudaMemsetAsync(a.ptr, 0, 8, st));
  SB_NCCL(nccl().AllGather(a.ptr, b.ptr, 1, ncclInt64, c.comm, st));
  count_launch();
}
bool comm_window(size_t need, cudaStream_t st, std::vector<void *> &bases) {
  Comm &c = comm();
  if (!c.comm) return false;
  if (!window_ensure(need, st)) return false;
  PeerWindow &w = window();
  bases.assign(c.nranks, nullptr);
  for (int r = 0; r < c.nranks; r++) bases[r] = r == c.rank ? w.local : w.remote[r];
  return true;
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
  if (nranks > 1 || id_bytes) {   // a one-rank communicator is a real NCCL communicator too (self-exchange tests)
    SB_REQUIRE(id_bytes, "sb_comm_init needs the unique id");
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
    window_release(nullptr);
    window().disabled = false;
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

struct TableHold {
  sb_table *t = nullptr;
  ~TableHold() { if (t) table_free(t); }
};

// rows per (source rank, partition) of the calling thread's last exchange: [R][P + 1] as all-gathered (all_to_all_by_source fills it)
static thread_local std::vector<int64_t> g_last_counts;

// the transport: received rows grouped by SOURCE rank, every source block partition-contiguous
static int all_to_all_by_source(const sb_table *in, const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s, sb_table **out,
                                int64_t *out_part_offsets_host) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && part_offsets_host && out && out_part_offsets_host, "null argument");
  Comm &c = comm();
  SB_REQUIRE(c.comm, "sb_all_to_all needs an initialised communicator (sb_comm_init)");
  cudaStream_t st = stream_of(s);
  const int R = c.nranks;
  const int P = num_partitions;
  // 1. every rank learns every rank's per-partition row counts
  // (the last word of every rank's message says which of its columns carry a validity bitmap: a column is nullable in the
  // exchange if it is on ANY rank -- a rank whose rows happen to hold no NULL may have dropped the bitmap)
  SB_REQUIRE(in->cols.size() <= 64, "the exchange supports up to 64 columns");
  const int PS = P + 1;
  std::vector<int64_t> my_counts(PS), all_counts((size_t)R * PS);
  for (int p = 0; p < P; p++) my_counts[p] = part_offsets_host[p + 1] - part_offsets_host[p];
  uint64_t my_mask = 0;
  for (size_t ci = 0; ci < in->cols.size(); ci++)
    if (in->cols[ci].validity) my_mask |= 1ull << ci;
  my_counts[P] = (int64_t)my_mask;
  Scratch d_my(PS * 8, st), d_all((int64_t)R * PS * 8, st);
  {
    KernelTimer kt("a2a_counts", st);
    SB_CUDA(cudaMemcpyAsync(d_my.ptr, my_counts.data(), (size_t)PS * 8, cudaMemcpyHostToDevice, st));
    SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, (size_t)PS, ncclInt64, c.comm, st));
    SB_CUDA(cudaMemcpyAsync(all_counts.data(), d_all.ptr, (size_t)R * PS * 8, cudaMemcpyDeviceToHost, st));
  }
  SB_CUDA(cudaStreamSynchronize(st));
  g_last_counts = all_counts;
  uint64_t any_mask = 0;
  for (int r = 0; r < R; r++) any_mask |= (uint64_t)all_counts[(size_t)r * PS + P];
  auto nullable = [&](size_t ci) { return ((any_mask >> ci) & 1) != 0; };
  // 2. layout of what this rank receives: [source rank][owned partitions]
  const int lo = part_lo(c.rank, P, R), hi = part_lo(c.rank + 1, P, R);
  std::vector<int64_t> recv_rows(R), recv_off(R + 1, 0), send_rows(R), send_off(R);
  for (int src = 0; src < R; src++) {
    int64_t rows = 0;
    for (int p = lo; p < hi; p++) rows += all_counts[(size_t)src * PS + p];
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
      for (int src = 0; src < R; src++) rows += all_counts[(size_t)src * PS + p];
    out_part_offsets_host[p + 1] = out_part_offsets_host[p] + rows;
  }
  // 3. data path.  Window layout of a receiver d: the columns one after the other (256-byte aligned), each laid out like the
  //    output column ([source rank][owned partitions]); a nullable column is followed by its validity as one byte per row.
  bool has_string = false;
  for (auto &col : in->cols) has_string |= col.type == SB_STRING;
  auto recv_total_of = [&](int d) {
    int64_t rows = 0;
    const int dl = part_lo(d, P, R), dh = part_lo(d + 1, P, R);
    for (int src = 0; src < R; src++)
      for (int p = dl; p < dh; p++) rows += all_counts[(size_t)src * PS + p];
    return rows;
  };
  auto recv_off_of = [&](int d, int src_rank) {   // first row of src_rank's segment in d's columns
    int64_t rows = 0;
    const int dl = part_lo(d, P, R), dh = part_lo(d + 1, P, R);
    for (int sr = 0; sr < src_rank; sr++)
      for (int p = dl; p < dh; p++) rows += all_counts[(size_t)sr * PS + p];
    return rows;
  };
  auto window_layout = [&](int64_t rows, std::vector<size_t> &col_off, std::vector<size_t> &val_off) {
    size_t cur = 0;
    col_off.assign(in->cols.size(), 0);
    val_off.assign(in->cols.size(), 0);
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      col_off[ci] = cur;
      cur += ((size_t)rows * type_width(in->cols[ci].type) + 255) / 256 * 256;
      if (nullable(ci)) {
        val_off[ci] = cur;
        cur += ((size_t)rows + 255) / 256 * 256;
      }
    }
    return cur;
  };
  const bool force_nccl = config().exchange_nccl != 0;
  bool peer_path = !force_nccl && !has_string;
  if (peer_path) {
    size_t need = 0;
    std::vector<size_t> co, vo;
    for (int d = 0; d < R; d++) need = std::max(need, window_layout(recv_total_of(d), co, vo));
    peer_path = window_ensure(need, st);
  }
  if (peer_path) {
    PeerWindow &w = window();
    sb_table *t = table_new(nrecv);
    try {
      KernelTimer kt_transfer("a2a_transfer", st);
      const int me = c.rank;
      std::vector<std::unique_ptr<Scratch>> temps;   // released (stream-ordered) on every exit, an exception included
      std::vector<uint8_t *> send_bytes(in->cols.size(), nullptr);
      t->cols.reserve(in->cols.size());
      for (size_t ci = 0; ci < in->cols.size(); ci++) {
        const Column &col = in->cols[ci];
        t->cols.push_back(column_alloc(col.type, col.scale, nrecv, nullable(ci), st));
        if (nullable(ci)) {
          Scratch *sb = new Scratch(in->nrows + 16, st);
          temps.emplace_back(sb);
          bitmap_to_bytes(col.v(), in->nrows, sb->as<uint8_t>(), st);
          send_bytes[ci] = sb->as<uint8_t>();
        }
      }
      SB_CUDA(cudaEventRecord(w.ev_ready, st));
      // pushes: one stream per destination so every NVLink direction is busy at once
      for (int d = 0; d < R; d++) {
        if (send_rows[d] == 0) continue;
        cudaStream_t ps = w.push_streams[d];
        SB_CUDA(cudaStreamWaitEvent(ps, w.ev_ready, 0));
        std::vector<size_t> co, vo;
        window_layout(recv_total_of(d), co, vo);
        const int64_t roff = recv_off_of(d, me);
        for (size_t ci = 0; ci < in->cols.size(); ci++) {
          const Column &col = in->cols[ci];
          const int wd = type_width(col.type);
          char *dst_base = d == me ? (char *)t->cols[ci].data->ptr : (char *)w.remote[d] + co[ci];
          SB_CUDA(cudaMemcpyAsync(dst_base + roff * wd, (const char *)col.d() + send_off[d] * wd, (size_t)(send_rows[d] * wd),
                                  cudaMemcpyDeviceToDevice, ps));
          if (nullable(ci) && d != me)
            SB_CUDA(cudaMemcpyAsync((char *)w.remote[d] + vo[ci] + roff, send_bytes[ci] + send_off[d], (size_t)send_rows[d],
                                    cudaMemcpyDeviceToDevice, ps));
        }
        SB_CUDA(cudaEventRecord(w.ev_pushed[d], ps));
        SB_CUDA(cudaStreamWaitEvent(st, w.ev_pushed[d], 0));
      }
      // "every push has landed": a one-word all-gather ordered after this rank's pushes on st
      {
        Scratch d_one(8, st), d_allone(8 * R, st);
        SB_CUDA(cudaMemsetAsync(d_one.ptr, 0, 8, st));
        SB_NCCL(nccl().AllGather(d_one.ptr, d_allone.ptr, 1, ncclInt64, c.comm, st));
        count_launch();
        // copy out of the window (everything except this rank's own segment, which went straight to the columns)
        std::vector<size_t> co, vo;
        window_layout(nrecv, co, vo);
        std::vector<Scratch *> vbytes;
        for (size_t ci = 0; ci < in->cols.size(); ci++) {
          const Column &col = in->cols[ci];
          const int wd = type_width(col.type);
          char *outp = (char *)t->cols[ci].data->ptr;
          const char *win = (const char *)w.local + co[ci];
          const int64_t a0 = recv_off[me], a1 = recv_off[me + 1];
          if (a0 > 0) SB_CUDA(cudaMemcpyAsync(outp, win, (size_t)(a0 * wd), cudaMemcpyDeviceToDevice, st));
          if (nrecv > a1) SB_CUDA(cudaMemcpyAsync(outp + a1 * wd, win + a1 * wd, (size_t)((nrecv - a1) * wd), cudaMemcpyDeviceToDevice, st));
          if (nullable(ci)) {
            // own rows' validity bytes join the others in the window, then the whole column is re-packed into a bitmap
            uint8_t *vwin = (uint8_t *)w.local + vo[ci];
            if (a1 > a0) SB_CUDA(cudaMemcpyAsync(vwin + a0, send_bytes[ci] + send_off[me], (size_t)(a1 - a0), cudaMemcpyDeviceToDevice, st));
            bytes_to_bitmap(vwin, nrecv, (uint32_t *)t->cols[ci].validity->ptr, st);
          }
        }
        SB_CUDA(cudaStreamSynchronize(st));
      }
      temps.clear();
    } catch (...) {
      table_free(t);
      throw;
    }
    *out = t;
    return SB_OK;
  }
  // NCCL data path (string-free tables only as well; used when CUDA IPC is unavailable or SB_EXCHANGE=nccl):
  // one grouped send/recv per (column buffer, peer)
  sb_table *t = table_new(nrecv);
  try {
    std::vector<std::unique_ptr<Scratch>> temps;   // released (stream-ordered) on every exit, an exception included
    struct Pending { Column *col; uint8_t *recv_bytes; };
    std::vector<Pending> pend;
    std::vector<std::pair<const uint8_t *, uint8_t *>> self_valid;   // (send bytes, recv bytes) of every nullable column
    t->cols.reserve(in->cols.size());
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      const Column &col = in->cols[ci];
      if (col.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "all-to-all of string columns is not implemented (dictionary-encode them)");
      t->cols.push_back(column_alloc(col.type, col.scale, nrecv, nullable(ci), st));
    }
    KernelTimer kt_transfer("a2a_transfer", st);
    SB_NCCL(nccl().GroupStart());
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      const Column &src = in->cols[ci];
      Column &dst = t->cols[ci];
      const int w = type_width(src.type);
      for (int peer = 0; peer < R; peer++) {
        if (peer == c.rank) continue;   // the rank's own partitions never leave HBM (copied below, outside the NCCL group)
        if (send_rows[peer] > 0)
          SB_NCCL(nccl().Send((const char *)src.d() + send_off[peer] * w, (size_t)(send_rows[peer] * w), ncclUint8, peer, c.comm, st));
        if (recv_rows[peer] > 0)
          SB_NCCL(nccl().Recv((char *)dst.data->ptr + recv_off[peer] * w, (size_t)(recv_rows[peer] * w), ncclUint8, peer, c.comm, st));
      }
      if (nullable(ci)) {
        Scratch *sb = new Scratch(in->nrows + 16, st), *rb = new Scratch(nrecv + 16, st);
        temps.emplace_back(sb);
        temps.emplace_back(rb);
        bitmap_to_bytes(src.v(), in->nrows, sb->as<uint8_t>(), st);
        for (int peer = 0; peer < R; peer++) {
          if (peer == c.rank) continue;
          if (send_rows[peer] > 0) SB_NCCL(nccl().Send(sb->as<uint8_t>() + send_off[peer], (size_t)send_rows[peer], ncclUint8, peer, c.comm, st));
          if (recv_rows[peer] > 0) SB_NCCL(nccl().Recv(rb->as<uint8_t>() + recv_off[peer], (size_t)recv_rows[peer], ncclUint8, peer, c.comm, st));
        }
        pend.push_back({&dst, rb->as<uint8_t>()});
        self_valid.push_back({sb->as<uint8_t>(), rb->as<uint8_t>()});
      }
    }
    static cudaStream_t self_st = nullptr;
    static cudaEvent_t ev_begin = nullptr, ev_done = nullptr;
    if (!self_st) {
      SB_CUDA(cudaStreamCreateWithFlags(&self_st, cudaStreamNonBlocking));
      SB_CUDA(cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    }
    SB_CUDA(cudaEventRecord(ev_begin, st));   // inputs and byte-expanded validity are ready here; the NCCL kernel is enqueued by GroupEnd
    SB_NCCL(nccl().GroupEnd());
    count_launch();   // the grouped NCCL kernel
    // own partitions: plain device-to-device copies, on a second stream so they overlap the NVLink traffic
    {
      const int me = c.rank;
      if (send_rows[me] > 0) {
        SB_CUDA(cudaStreamWaitEvent(self_st, ev_begin, 0));
        for (size_t ci = 0; ci < in->cols.size(); ci++) {
          const int w = type_width(in->cols[ci].type);
          SB_CUDA(cudaMemcpyAsync((char *)t->cols[ci].data->ptr + recv_off[me] * w, (const char *)in->cols[ci].d() + send_off[me] * w,
                                  (size_t)(send_rows[me] * w), cudaMemcpyDeviceToDevice, self_st));
        }
        for (auto &sv : self_valid)
          SB_CUDA(cudaMemcpyAsync(sv.second + recv_off[me], sv.first + send_off[me], (size_t)send_rows[me], cudaMemcpyDeviceToDevice, self_st));
        SB_CUDA(cudaEventRecord(ev_done, self_st));
        SB_CUDA(cudaStreamWaitEvent(st, ev_done, 0));
      }
    }
    for (auto &p : pend) {
      bytes_to_bitmap(p.recv_bytes, nrecv, (uint32_t *)p.col->validity->ptr, st);
    }
    SB_CUDA(cudaStreamSynchronize(st));
    temps.clear();
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// source-grouped row i -> its place in the partition-contiguous result: rows of partition p from source s start at
// part_offset[p] + sum_{s' < s} count[s'][p].  One thread per row finds its (source, partition) segment by binary search.
__global__ void regroup_index_kernel(const int64_t *__restrict__ seg_src, const int64_t *__restrict__ seg_dst, int32_t nseg, int64_t n,
                                     int64_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nseg - 1;   // last segment with seg_src <= i (empty segments share a start: any of them maps i correctly only
  while (lo < hi) {            // if it is the LAST one with that start, which this search returns)
    const int mid = (lo + hi + 1) >> 1;
    if (seg_src[mid] <= i) lo = mid; else hi = mid - 1;
  }
  idx[seg_dst[lo] + (i - seg_src[lo])] = i;
}

// ShuffleExchangeExec's reduce side as its readers expect it: partition-contiguous over the owned partitions (a coalesced read of
// AQEShuffleReadExec is a slice by partition offsets), rows of a partition ordered by source rank, arrival order inside a source.
static int all_to_all_fixed(const sb_table *in, const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s, sb_table **out,
                            int64_t *out_part_offsets_host) {
  sb_table *by_source = nullptr;
  int rc = all_to_all_by_source(in, part_offsets_host, num_partitions, s, &by_source, out_part_offsets_host);
  if (rc != SB_OK) return rc;
  const int R = comm().nranks, P = num_partitions, PS = P + 1, me = comm().rank;
  if (R == 1) {
    *out = by_source;
    return SB_OK;
  }
  SB_API_BEGIN
  TableHold hold;
  hold.t = by_source;
  cudaStream_t st = stream_of(s);
  const int64_t n = by_source->nrows;
  const int lo = part_lo(me, P, R), hi = part_lo(me + 1, P, R);
  const int nseg = R * (hi - lo);
  if (n == 0 || nseg == 0) {
    *out = by_source;
    hold.t = nullptr;
    return SB_OK;
  }
  const std::vector<int64_t> &cnt = g_last_counts;
  std::vector<int64_t> seg(2 * (size_t)nseg), before((size_t)(hi - lo), 0);
  int64_t src = 0;
  for (int r = 0; r < R; r++)
    for (int p = lo; p < hi; p++) {
      const size_t k = (size_t)r * (hi - lo) + (p - lo);
      seg[k] = src;
      seg[nseg + k] = out_part_offsets_host[p] + before[p - lo];
      src += cnt[(size_t)r * PS + p];
      before[p - lo] += cnt[(size_t)r * PS + p];
    }
  Scratch dseg((int64_t)seg.size() * 8, st), idx(n * 8 + 16, st);
  SB_CUDA(cudaMemcpyAsync(dseg.ptr, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, st));
  regroup_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dseg.as<int64_t>(), dseg.as<int64_t>() + nseg, nseg, n, idx.as<int64_t>());
  SB_LAUNCH_CHECK();
  *out = gather_table(by_source, idx.as<int64_t>(), n, false, st);
  SB_CUDA(cudaStreamSynchronize(st));   // `seg` (host) is read by the copy above
  SB_API_END
}

// MapOutputStatistics.bytesByPartitionId of the exchange whose map side produced `partitioned` (ShuffleExchangeExec.scala:235-262):
// rows x fixed row width (+ string arenas are not attributed to partitions), summed over the ranks with one all-gather.
int sb_map_output_statistics(const sb_table *partitioned, const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s,
                             int64_t *out_bytes_by_partition) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(partitioned && part_offsets_host && out_bytes_by_partition && num_partitions >= 1, "bad argument");
  Comm &c = comm();
  cudaStream_t st = stream_of(s);
  int64_t row_bytes = 0;
  for (auto &col : partitioned->cols) row_bytes += col.type == SB_STRING ? 4 : type_width(col.type);
  std::vector<int64_t> mine(num_partitions), all;
  for (int p = 0; p < num_partitions; p++) mine[p] = part_offsets_host[p + 1] - part_offsets_host[p];
  if (c.comm) {
    const int R = c.nranks;
    all.resize((size_t)R * num_partitions);
    Scratch d_my((int64_t)num_partitions * 8, st), d_all((int64_t)R * num_partitions * 8, st);
    SB_CUDA(cudaMemcpyAsync(d_my.ptr, mine.data(), (size_t)num_partitions * 8, cudaMemcpyHostToDevice, st));
    SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, (size_t)num_partitions, ncclInt64, c.comm, st));
    SB_CUDA(cudaMemcpyAsync(all.data(), d_all.ptr, all.size() * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < num_partitions; p++) {
      int64_t rows = 0;
      for (int r = 0; r < R; r++) rows += all[(size_t)r * num_partitions + p];
      out_bytes_by_partition[p] = rows * row_bytes;
    }
  } else {
    for (int p = 0; p < num_partitions; p++) out_bytes_by_partition[p] = mine[p] * row_bytes;
  }
  SB_API_END
}

// Rows every rank holds for every partition (the all-gather sb_all_to_all starts with): out_counts[rank * num_partitions + p].
// The reduce side needs it to cut one reducer partition out of a received table, whose rows are grouped by SOURCE rank.
int sb_exchange_counts(const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s, int64_t *out_counts) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(part_offsets_host && out_counts && num_partitions >= 1, "bad argument");
  Comm &c = comm();
  cudaStream_t st = stream_of(s);
  std::vector<int64_t> mine(num_partitions);
  for (int p = 0; p < num_partitions; p++) mine[p] = part_offsets_host[p + 1] - part_offsets_host[p];
  if (!c.comm) {
    memcpy(out_counts, mine.data(), (size_t)num_partitions * 8);
    return SB_OK;
  }
  Scratch d_my((int64_t)num_partitions * 8, st), d_all((int64_t)c.nranks * num_partitions * 8, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, mine.data(), (size_t)num_partitions * 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, (size_t)num_partitions, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(out_counts, d_all.ptr, (size_t)c.nranks * num_partitions * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_API_END
}

static int all_gather_fixed(const sb_table *in, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out, "null argument");
  Comm &c = comm();
  cudaStream_t st = stream_of(s);
  if (!c.comm) {   // no communicator: single executor, the broadcast is the table itself
    sb_table *t = table_new(in->nrows);
    for (auto &col : in->cols) t->cols.push_back(column_share(col));
    *out = t;
    return SB_OK;
  }
  const int R = c.nranks;
  int64_t my = in->nrows;
  // ---- optimistic packed path (see ag_pack_kernel) ----
  bool packable = R <= 16 && (int)in->cols.size() <= AG_MAX_COLS;
  for (auto &col : in->cols) packable = packable && col.type != SB_STRING;
  if (packable) {
    AgPackArgs a;
    memset(&a, 0, sizeof(a));
    a.ncols = (int)in->cols.size();
    a.nranks = R;
    a.my_rows = my;
    int64_t cur = 16;
    for (int ci = 0; ci < a.ncols; ci++) {
      const Column &col = in->cols[ci];
      a.width[ci] = type_width(col.type);
      a.src[ci] = col.d();
      a.src_valid[ci] = col.v();
      if (col.validity) a.my_valid_mask |= 1ull << ci;
      a.data_off[ci] = cur;
      cur += (AG_PACK_ROWS * a.width[ci] + 15) / 16 * 16;
      a.valid_off[ci] = cur;
      cur += AG_PACK_ROWS;
    }
    a.stride = (cur + 255) / 256 * 256;
    Scratch msg(a.stride, st), all(a.stride * R, st);
    ag_pack_kernel<<<a.ncols + 1, 128, 0, st>>>(a, msg.as<uint8_t>());
    SB_LAUNCH_CHECK();
    SB_NCCL(nccl().AllGather(msg.ptr, all.ptr, (size_t)a.stride, ncclUint8, c.comm, st));
    count_launch();
    std::vector<int64_t> hdr(2 * R);
    SB_CUDA(cudaMemcpy2DAsync(hdr.data(), 16, all.ptr, (size_t)a.stride, 16, (size_t)R, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    bool fits = true;
    int64_t total = 0;
    uint64_t any_valid = 0;   // a column is nullable in the result if it is on ANY rank (a rank whose rows happen to hold no NULL
                              // may have dropped the bitmap)
    for (int r = 0; r < R; r++) {
      fits = fits && hdr[2 * r] <= AG_PACK_ROWS;
      a.rows[r] = hdr[2 * r];
      a.off[r] = total;
      total += hdr[2 * r];
      any_valid |= (uint64_t)hdr[2 * r + 1];
    }
    if (fits) {
      sb_table *t = table_new(total);
      try {
        std::vector<std::unique_ptr<Scratch>> temps;   // released (stream-ordered) on every exit, an exception included
        for (int ci = 0; ci < a.ncols; ci++) {
          const Column &col = in->cols[ci];
          const bool nullable = (any_valid >> ci) & 1;
          t->cols.push_back(column_alloc(col.type, col.scale, total, nullable, st));
          a.dst[ci] = t->cols[ci].data->ptr;
          a.dst_valid_bytes[ci] = nullptr;
          if (nullable) {
            Scratch *vb = new Scratch(total + 16, st);
            temps.emplace_back(vb);
            a.dst_valid_bytes[ci] = vb->as<uint8_t>();
          }
        }
        if (total > 0) {
          ag_unpack_kernel<<<dim3(a.ncols, R), 128, 0, st>>>(a, all.as<uint8_t>());
          SB_LAUNCH_CHECK();
          for (int ci = 0; ci < a.ncols; ci++)
            if (a.dst_valid_bytes[ci]) bytes_to_bitmap(a.dst_valid_bytes[ci], total, (uint32_t *)t->cols[ci].validity->ptr, st);
        }
        temps.clear();   // stream-ordered frees
      } catch (...) {
        table_free(t);
        throw;
      }
      *out = t;
      return SB_OK;
    }
    // some rank's table did not fit: every rank saw that and continues with the general protocol
  }
  std::vector<int64_t> rows(R), off(R + 1, 0), hdr2(2 * R);
  uint64_t my_mask = 0, any_mask = 0;   // nullable in the result = nullable on any rank (see the packed path)
  for (size_t ci = 0; ci < in->cols.size() && ci < 64; ci++)
    if (in->cols[ci].validity) my_mask |= 1ull << ci;
  int64_t my_hdr[2] = {my, (int64_t)my_mask};
  Scratch d_my(16, st), d_all(R * 16, st);
  SB_CUDA(cudaMemcpyAsync(d_my.ptr, my_hdr, 16, cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_my.ptr, d_all.ptr, 2, ncclInt64, c.comm, st));
  SB_CUDA(cudaMemcpyAsync(hdr2.data(), d_all.ptr, (size_t)R * 16, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < R; r++) {
    rows[r] = hdr2[2 * r];
    any_mask |= (uint64_t)hdr2[2 * r + 1];
    off[r + 1] = off[r] + rows[r];
  }
  SB_REQUIRE(in->cols.size() <= 64, "all-gather supports up to 64 columns");
  const int64_t total = off[R];
  sb_table *t = table_new(total);
  try {
    std::vector<std::unique_ptr<Scratch>> temps;   // released (stream-ordered) on every exit, an exception included
    struct Pending { Column *col; uint8_t *bytes; };
    std::vector<Pending> pend;
    t->cols.reserve(in->cols.size());
    for (size_t ci = 0; ci < in->cols.size(); ci++) {
      const Column &col = in->cols[ci];
      if (col.type == SB_STRING) fail(SB_ERR_UNSUPPORTED, "all-gather of string columns is not implemented (dictionary-encode them)");
      t->cols.push_back(column_alloc(col.type, col.scale, total, (any_mask >> ci) & 1, st));
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
      if ((any_mask >> ci) & 1) {
        Scratch *sb = new Scratch(my + 16, st), *rb = new Scratch(total + 16, st);
        temps.emplace_back(sb);
        temps.emplace_back(rb);
        bitmap_to_bytes(src.v(), my, sb->as<uint8_t>(), st);   // no bitmap on this rank: all ones
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
    temps.clear();
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// ---- string columns in the collectives ----------------------------------------------------------------------------------------------
// The transports above move fixed-width columns.  Strings ride on them:
//   all-gather : a string column travels as its int32 LENGTHS (a fixed-width column that keeps the validity) plus one more
//                all-gather of the byte arena; rank order is row order, so the concatenated arenas are the result's arena and an
//                exclusive scan of the lengths gives its offsets;
//   all-to-all : the column is dictionary-encoded locally (csrc/strings.cu), the int32 codes are exchanged together with a hidden
//                "source rank" column, the (small) dictionaries are all-gathered, every received code is rebased by its source's
//                offset into the concatenation of the dictionaries and decoded.  Low-cardinality columns (flags, names, dates as
//                text) move 4 bytes per row; a column of unique strings moves its dictionary to every rank -- correct, not cheap.
__global__ void lengths_kernel(const int32_t *__restrict__ off, const uint8_t *__restrict__ valid, int64_t n, int32_t *__restrict__ len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = bit_valid(valid, i) ? off[i + 1] - off[i] : 0;
}
__global__ void fill_i32_kernel(int32_t *out, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
__global__ void rebase_codes_kernel(int32_t *__restrict__ codes, const uint8_t *__restrict__ valid, const int32_t *__restrict__ src_rank,
                                    const int64_t *__restrict__ base, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && bit_valid(valid, i)) codes[i] += (int32_t)base[src_rank[i]];
}

static bool has_string_column(const sb_table *t) {
  for (auto &c : t->cols)
    if (c.type == SB_STRING) return true;
  return false;
}
static void check_rc(int rc) {
  if (rc != SB_OK) fail(rc, "%s", sb_last_error());
}

int sb_all_gather(const sb_table *in, sb_stream *s, sb_table **out) {
  if (!in || !has_string_column(in) || !comm().comm) return all_gather_fixed(in, s, out);
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(out, "null argument");
  cudaStream_t st = stream_of(s);
  const int64_t n = in->nrows;
  TableHold fixed, moved;
  fixed.t = table_new(n);
  for (auto &c : in->cols) {
    if (c.type != SB_STRING) {
      fixed.t->cols.push_back(column_share(c));
      continue;
    }
    Column len = column_alloc(SB_INT32, 0, n, false, st);
    if (n > 0) {
      lengths_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c.o(), c.v(), n, (int32_t *)len.data->ptr);
      SB_LAUNCH_CHECK();
    }
    if (c.validity) {
      buffer_retain(c.validity);
      len.validity = c.validity;
      len.null_count = c.null_count;
    }
    fixed.t->cols.push_back(len);
  }
  check_rc(all_gather_fixed(fixed.t, s, &moved.t));
  const int64_t total = moved.t->nrows;
  SB_REQUIRE(total < (1ll << 31), "all-gather of a string column: too many rows");
  for (size_t ci = 0; ci < in->cols.size(); ci++) {
    const Column &c = in->cols[ci];
    if (c.type != SB_STRING) continue;
    TableHold bytes, bytes_all;
    int64_t nbytes = 0;
    if (n > 0) {   // bytes in use = offsets[n] (NULL rows own no bytes in a compact arena; a sparse one is shipped as it is)
      int32_t last = 0;
      SB_CUDA(cudaMemcpyAsync(&last, c.o() + n, 4, cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      nbytes = last;
    }
    bytes.t = table_new(nbytes);
    {
      Column b;
      b.type = SB_INT8;
      b.length = nbytes;
      b.null_count = 0;
      b.data = c.data;
      buffer_retain(b.data);
      bytes.t->cols.push_back(b);
    }
    check_rc(all_gather_fixed(bytes.t, s, &bytes_all.t));
    SB_REQUIRE(bytes_all.t->nrows < (1ll << 31), "all-gather of a string column: more than 2 GiB of characters");
    Column &lens = moved.t->cols[ci];
    Column r;
    r.type = SB_STRING;
    r.length = total;
    r.null_count = lens.null_count;
    r.string_bytes = bytes_all.t->nrows;
    r.offsets = buffer_alloc((total + 1) * 4 + 16, st);
    Scratch tot(4, st);
    // NULL rows were given length 0 above, but their arena bytes (if any) travelled: offsets must follow the ARENA, so a sparse
    // arena is not supported here -- producers in this library (gather, decode, import) write compact arenas
    exclusive_scan_i32((const int32_t *)lens.d(), (int32_t *)r.offsets->ptr, total, tot.as<int32_t>(), st);
    SB_CUDA(cudaMemcpyAsync((int32_t *)r.offsets->ptr + total, tot.ptr, 4, cudaMemcpyDeviceToDevice, st));
    r.data = bytes_all.t->cols[0].data;
    buffer_retain(r.data);
    r.validity = lens.validity;
    if (r.validity) buffer_retain(r.validity);
    column_release(lens);
    lens = r;
  }
  *out = moved.t;
  moved.t = nullptr;
  SB_CUDA(cudaStreamSynchronize(st));
  SB_API_END
}

int sb_all_to_all(const sb_table *in, const int64_t *part_offsets_host, int32_t num_partitions, sb_stream *s, sb_table **out,
                  int64_t *out_part_offsets_host) {
  if (!in || !has_string_column(in)) return all_to_all_fixed(in, part_offsets_host, num_partitions, s, out, out_part_offsets_host);
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(out && part_offsets_host && out_part_offsets_host, "null argument");
  Comm &c = comm();
  SB_REQUIRE(c.comm, "sb_all_to_all needs an initialised communicator (sb_comm_init)");
  cudaStream_t st = stream_of(s);
  const int64_t n = in->nrows;
  std::vector<int> scols;
  for (size_t ci = 0; ci < in->cols.size(); ci++)
    if (in->cols[ci].type == SB_STRING) scols.push_back((int)ci);
  EncodedView ev;
  encode_string_columns(in, scols, nullptr, st, ev);
  TableHold fixed, moved;
  fixed.t = table_new(n);
  for (auto &col : ev.view->cols) fixed.t->cols.push_back(column_share(col));
  {
    Column src = column_alloc(SB_INT32, 0, n, false, st);
    if (n > 0) {
      fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int32_t *)src.data->ptr, n, c.rank);
      SB_LAUNCH_CHECK();
    }
    fixed.t->cols.push_back(src);
  }
  check_rc(all_to_all_fixed(fixed.t, part_offsets_host, num_partitions, s, &moved.t, out_part_offsets_host));
  const int64_t m = moved.t->nrows;
  const Column src_rank = moved.t->cols.back();
  for (int ci : scols) {
    const Column *dict = ev.dictionary_of(ci);
    TableHold dt, dall;
    dt.t = table_new(dict->length);
    dt.t->cols.push_back(column_share(*dict));
    check_rc(sb_all_gather(dt.t, s, &dall.t));
    std::vector<int64_t> sizes(c.nranks), base(c.nranks);
    const int64_t mine = dict->length;
    comm_allgather_host(&mine, 1, sizes.data(), st);
    int64_t acc = 0;
    for (int r = 0; r < c.nranks; r++) {
      base[r] = acc;
      acc += sizes[r];
    }
    SB_REQUIRE(acc < (1ll << 31), "exchange of a string column: the dictionaries of all ranks exceed 2^31 entries");
    Scratch dbase((int64_t)c.nranks * 8, st);
    SB_CUDA(cudaMemcpyAsync(dbase.ptr, base.data(), (size_t)c.nranks * 8, cudaMemcpyHostToDevice, st));
    Column &codes = moved.t->cols[ci];
    if (m > 0) {
      rebase_codes_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>((int32_t *)codes.data->ptr, codes.v(), (const int32_t *)src_rank.d(), dbase.as<int64_t>(), m);
      SB_LAUNCH_CHECK();
    }
    Column decoded = dictionary_decode(codes, dall.t->cols[0], st);
    SB_CUDA(cudaStreamSynchronize(st));   // `base` (host) and `dbase` are done with
    column_release(codes);
    codes = decoded;
  }
  column_release(moved.t->cols.back());
  moved.t->cols.pop_back();
  *out = moved.t;
  moved.t = nullptr;
  SB_API_END
}

}  // extern "C"
