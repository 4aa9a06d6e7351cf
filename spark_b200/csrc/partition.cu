// partition.cu -- ShuffleExchangeExec map side on the GPU.
//
// Reference path replaced (citations relative to the reference tree):
//   partition id  = Pmod(Murmur3Hash(keys, 42), n)      sql/catalyst/.../plans/physical/partitioning.scala:339-341
//   Murmur3       = common/unsafe/.../hash/Murmur3_x86_32.java:47-150, type dispatch hash.scala:707-760
//   grouping rows by partition id = UnsafeShuffleWriter + ShuffleInMemorySorter radix on the partition
//                   bytes of PackedRecordPointer (core/src/main/java/org/apache/spark/shuffle/sort/
//                   ShuffleInMemorySorter.java:202-225) / BypassMergeSortShuffleWriter for n <= 200.
//
// GPU design (HBM-bound integer work, no tensor cores): a stable single-pass multisplit.
//   K1 pid_hist   : one coalesced read of the key columns, Murmur3 chain in registers, pid stored as
//                   int32, per-block shared-memory histogram -> hist[p][block]
//   K2 scan       : exclusive scan over hist in (partition, block) order -> every block's first output
//                   slot for every partition, and the partition boundaries
//   K3 scatter    : each block re-walks its contiguous chunk; every warp owns a contiguous sub-chunk and
//                   keeps running per-partition cursors in shared memory; ranks inside a 32-row group come
//                   from __match_any_sync, so arrival order inside a partition is preserved (stable);
//                   all fixed-width payload columns are moved with one coalesced read and one write each.
// Algorithmic bytes: 2 x rowbytes per row; overhead: keys re-read once + 12 B/row of pid
// traffic + the (partitions x blocks) histogram.
#include "common.cuh"
#include "primitives.cuh"
#include "multisplit.cuh"

namespace sb {

constexpr int MAX_KEYS = 8;
struct KeyCols {
  int n;
  int type[MAX_KEYS];
  const void *data[MAX_KEYS];
  const uint8_t *valid[MAX_KEYS];
  const int32_t *offs[MAX_KEYS];
};

// Murmur3Hash over the key columns of row i; NULL leaves the running hash unchanged (hash.scala:714),
// the running hash seeds the next column (hash.scala:400-409).
__device__ __forceinline__ uint32_t row_hash(const KeyCols &k, int64_t i, uint32_t seed) {
  uint32_t h = seed;
#pragma unroll 1
  for (int c = 0; c < k.n; c++) {
    if (!bit_valid(k.valid[c], i)) continue;
    switch (k.type[c]) {
      case SB_BOOL: h = mm3_int(((const uint8_t *)k.data[c])[i] ? 1u : 0u, h); break;
      case SB_INT8: h = mm3_int((uint32_t)(int32_t)((const int8_t *)k.data[c])[i], h); break;
      case SB_INT16: h = mm3_int((uint32_t)(int32_t)((const int16_t *)k.data[c])[i], h); break;
      case SB_INT32: case SB_DATE32: h = mm3_int(((const uint32_t *)k.data[c])[i], h); break;
      case SB_INT64: case SB_TIMESTAMP: case SB_DECIMAL64: h = mm3_long(((const uint64_t *)k.data[c])[i], h); break;
      case SB_FLOAT32: {
        float f = ((const float *)k.data[c])[i];
        h = mm3_int(f == 0.0f ? 0u : (uint32_t)float_bits_canonical(f), h);   // -0.0f -> 0 (hash.scala:718)
        break;
      }
      case SB_FLOAT64: {
        double d = ((const double *)k.data[c])[i];
        h = mm3_long(d == 0.0 ? 0ull : (uint64_t)double_bits_canonical(d), h);  // -0.0 -> 0 (hash.scala:720)
        break;
      }
      case SB_STRING: {
        int32_t o = k.offs[c][i], len = k.offs[c][i + 1] - o;
        h = mm3_bytes((const uint8_t *)k.data[c] + o, len, h);
        break;
      }
    }
  }
  return h;
}

// pmod (sql/api/.../catalyst/util/MathUtils.scala:96-99)
__device__ __forceinline__ int32_t pmod(int32_t a, int32_t n) {
  int32_t r = a % n;
  return r < 0 ? (r + n) % n : r;
}

constexpr int PART_THREADS = 256;
constexpr int PART_WARPS = PART_THREADS / 32;

// K1: pid + block histogram.  hist layout [partition][block].  key_mode: 0 = murmur3 hash of keys,
// 1 = round robin (pid = (start + 1 + i) mod n).
__global__ void __launch_bounds__(PART_THREADS) pid_hist_kernel(KeyCols keys, int64_t n, int32_t nparts, int64_t chunk,
                                                                int32_t *__restrict__ pid_out, uint32_t *__restrict__ hist,
                                                                int key_mode, int32_t rr_start) {
  extern __shared__ uint32_t sh_hist[];
  if (hist) {
    for (int p = threadIdx.x; p < nparts; p += PART_THREADS) sh_hist[p] = 0;
    __syncthreads();
  }
  int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t i = begin + threadIdx.x; i < end; i += PART_THREADS) {
    int32_t p;
    if (key_mode == 0) p = pmod((int32_t)row_hash(keys, i, 42u), nparts);
    else p = (int32_t)(((int64_t)rr_start + 1 + i) % nparts);
    pid_out[i] = p;
    if (hist) atomicAdd(&sh_hist[p], 1u);
  }
  if (!hist) return;
  __syncthreads();
  for (int p = threadIdx.x; p < nparts; p += PART_THREADS) hist[(int64_t)p * gridDim.x + blockIdx.x] = sh_hist[p];
}

constexpr int SCATTER_MAX_COLS = 16;
struct ScatterCols {
  int ncols;
  int width[SCATTER_MAX_COLS];
  const void *src[SCATTER_MAX_COLS];
  void *dst[SCATTER_MAX_COLS];
  const uint8_t *src_valid[SCATTER_MAX_COLS];
  uint32_t *dst_valid[SCATTER_MAX_COLS];   // pre-set to all ones; NULL rows clear their bit
};

// K3: stable scatter.  base[p][block] = first output row of partition p for this block (exclusive scan
// of hist).  Shared memory: PART_WARPS x nparts cursors.
__global__ void __launch_bounds__(PART_THREADS) scatter_kernel(ScatterCols cols, const int32_t *__restrict__ pid,
                                                               const uint32_t *__restrict__ base, int64_t n, int32_t nparts,
                                                               int64_t chunk, int64_t *__restrict__ perm_out) {
  extern __shared__ uint32_t cursors[];   // [warp][nparts]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t begin = (int64_t)blockIdx.x * chunk;
  int64_t end = begin + chunk < n ? begin + chunk : n;
  // every warp owns a contiguous, 32-aligned sub-chunk
  int64_t len = end > begin ? end - begin : 0;
  int64_t sub = ((len + PART_WARPS - 1) / PART_WARPS + 31) / 32 * 32;
  int64_t wbeg = begin + warp * sub;
  int64_t wend = wbeg + sub < end ? wbeg + sub : end;

  // phase 1: per-warp histogram of its sub-chunk
  uint32_t *mine = cursors + (size_t)warp * nparts;
  for (int p = lane; p < nparts; p += 32) mine[p] = 0;
  __syncwarp();
  for (int64_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&mine[pid[i]], 1u);
  __syncthreads();
  // phase 2: exclusive prefix over warps + global block base -> cursors
  for (int p = threadIdx.x; p < nparts; p += PART_THREADS) {
    uint32_t run = base[(int64_t)p * gridDim.x + blockIdx.x];
#pragma unroll
    for (int w = 0; w < PART_WARPS; w++) {
      uint32_t c = cursors[(size_t)w * nparts + p];
      cursors[(size_t)w * nparts + p] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase 3: each warp walks its sub-chunk in order, 32 rows at a time
  for (int64_t i0 = wbeg; i0 < wend; i0 += 32) {
    int64_t i = i0 + lane;
    bool active = i < wend;
    int32_t p = active ? pid[i] : -1 - lane;             // inactive lanes get unique dummy keys
    uint32_t grp = __match_any_sync(0xffffffffu, p);
    uint32_t rank = __popc(grp & lanemask_lt());
    uint32_t dest = 0;
    if (active) {
      uint32_t b = mine[p];
      dest = b + rank;
      __syncwarp(grp);
      if (rank == 0) mine[p] = b + __popc(grp);
    }
    __syncwarp();
    if (!active) continue;
    if (perm_out) perm_out[dest] = i;
#pragma unroll 1
    for (int c = 0; c < cols.ncols; c++) {
      switch (cols.width[c]) {
        case 1: ((uint8_t *)cols.dst[c])[dest] = ((const uint8_t *)cols.src[c])[i]; break;
        case 2: ((uint16_t *)cols.dst[c])[dest] = ((const uint16_t *)cols.src[c])[i]; break;
        case 4: ((uint32_t *)cols.dst[c])[dest] = ((const uint32_t *)cols.src[c])[i]; break;
        default: ((uint64_t *)cols.dst[c])[dest] = ((const uint64_t *)cols.src[c])[i]; break;
      }
      if (cols.dst_valid[c] && !bit_valid(cols.src_valid[c], i))
        atomicAnd(&cols.dst_valid[c][dest >> 5], ~(1u << (dest & 31)));
    }
  }
}

__global__ void part_offsets_kernel(const uint32_t *__restrict__ base, int32_t nparts, int nblocks, int64_t n,
                                    int64_t *__restrict__ out_offsets) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nparts) out_offsets[p] = base[(int64_t)p * nblocks];
  if (p == nparts) out_offsets[p] = n;
}

static KeyCols make_keys(const sb_table *in, const int32_t *key_cols, int32_t nkeys) {
  SB_REQUIRE(nkeys >= 0 && nkeys <= MAX_KEYS, "at most %d partitioning keys are supported (got %d)", MAX_KEYS, nkeys);
  KeyCols k;
  k.n = nkeys;
  for (int i = 0; i < nkeys; i++) {
    SB_REQUIRE(key_cols[i] >= 0 && key_cols[i] < (int)in->cols.size(), "key column %d out of range", key_cols[i]);
    const Column &c = in->cols[key_cols[i]];
    k.type[i] = c.type;
    k.data[i] = c.d();
    k.valid[i] = c.v();
    k.offs[i] = c.o();
  }
  return k;
}

PartGeometry part_geometry(int64_t n) {
  PartGeometry g;
  int64_t want = (n + PART_THREADS * 16 - 1) / (PART_THREADS * 16);
  int maxb = rt().num_sms * 4;
  g.nblocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
  g.chunk = ((n + g.nblocks - 1) / g.nblocks + 31) / 32 * 32;
  if (g.chunk < 32) g.chunk = 32;
  return g;
}

void multisplit_scatter(const int32_t *bucket_dev, uint32_t *hist_dev, int32_t nbuckets, const PartGeometry &g,
                        const SplitCol *cols, int ncols, int64_t n, int64_t *perm_out, int64_t *offsets_dev,
                        cudaStream_t st) {
  SB_REQUIRE(nbuckets >= 1 && nbuckets <= MULTISPLIT_MAX_BUCKETS, "multisplit supports up to %d buckets (got %d)",
             MULTISPLIT_MAX_BUCKETS, nbuckets);
  SB_REQUIRE(n < (1ll << 32), "tables of 2^32 rows or more must be split in chunks");
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  size_t smem = (size_t)PART_WARPS * nbuckets * 4;
  exclusive_scan_i32((const int32_t *)hist_dev, (int32_t *)hist_dev, (int64_t)nbuckets * g.nblocks, nullptr, st);
  if (offsets_dev) {
    part_offsets_kernel<<<(nbuckets + 1 + 255) / 256, 256, 0, st>>>(hist_dev, nbuckets, g.nblocks, n, offsets_dev);
    SB_LAUNCH_CHECK();
  }
  if (n == 0) return;
  int done = 0;
  bool first = true;
  while (first || done < ncols) {
    ScatterCols sc;
    sc.ncols = 0;
    while (done < ncols && sc.ncols < SCATTER_MAX_COLS) {
      const SplitCol &c = cols[done++];
      int k = sc.ncols++;
      sc.width[k] = c.width;
      sc.src[k] = c.src;
      sc.dst[k] = c.dst;
      sc.src_valid[k] = c.src_valid;
      sc.dst_valid[k] = c.dst_valid;
    }
    if (sc.ncols > 0 || (first && perm_out)) {
      KernelTimer kt("partition_scatter", st);
      scatter_kernel<<<g.nblocks, PART_THREADS, smem, st>>>(sc, bucket_dev, hist_dev, n, nbuckets, g.chunk,
                                                            first ? perm_out : nullptr);
      SB_LAUNCH_CHECK();
    }
    first = false;
  }
}

// mode 0 = hash, 1 = round robin
static void partition_impl(const sb_table *in, const int32_t *key_cols, int32_t nkeys, int32_t nparts, int mode,
                           int32_t rr_start, cudaStream_t st, sb_table **out, int64_t *out_offsets_host) {
  SB_REQUIRE(in && out && out_offsets_host, "null argument");
  SB_REQUIRE(nparts >= 1, "num_partitions must be >= 1");
  int64_t n = in->nrows;
  if (nparts > MULTISPLIT_MAX_BUCKETS)
    fail(SB_ERR_UNSUPPORTED, "num_partitions %d exceeds the single-pass multisplit limit (%d)", nparts, MULTISPLIT_MAX_BUCKETS);
  KeyCols keys;
  keys.n = 0;
  if (mode == 0) keys = make_keys(in, key_cols, nkeys);
  PartGeometry g = part_geometry(n);
  Scratch pid(n * 4 + 16, st);
  Scratch hist((int64_t)nparts * g.nblocks * 4 + 16, st);
  Scratch offs_dev((nparts + 1) * 8, st);
  static bool attr_set = false;
  if (!attr_set) {
    SB_CUDA(cudaFuncSetAttribute(pid_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  pid_hist_kernel<<<g.nblocks, PART_THREADS, (size_t)nparts * 4, st>>>(keys, n, nparts, g.chunk, pid.as<int32_t>(),
                                                                        hist.as<uint32_t>(), mode, rr_start);
  SB_LAUNCH_CHECK();

  sb_table *t = table_new(n);
  try {
    bool need_perm = false;
    for (auto &c : in->cols) need_perm |= c.type == SB_STRING;
    Scratch perm(need_perm ? n * 8 + 8 : 0, st);
    t->cols.resize(in->cols.size());
    std::vector<SplitCol> sc;
    for (size_t i = 0; i < in->cols.size(); i++) {
      const Column &c = in->cols[i];
      if (c.type == SB_STRING) continue;
      Column r = column_alloc(c.type, c.scale, n, c.validity != nullptr, st);
      if (r.validity) SB_CUDA(cudaMemsetAsync(r.validity->ptr, 0xff, (size_t)bitmap_alloc_bytes(n), st));
      r.null_count = c.null_count;
      t->cols[i] = r;
      sc.push_back({type_width(c.type), c.d(), r.data->ptr, c.v(), r.validity ? (uint32_t *)r.validity->ptr : nullptr});
    }
    multisplit_scatter(pid.as<int32_t>(), hist.as<uint32_t>(), nparts, g, sc.data(), (int)sc.size(), n,
                       need_perm ? perm.as<int64_t>() : nullptr, offs_dev.as<int64_t>(), st);
    if (need_perm) {
      for (size_t i = 0; i < in->cols.size(); i++)
        if (in->cols[i].type == SB_STRING) t->cols[i] = gather_column(in->cols[i], perm.as<int64_t>(), n, false, st);
    }
    SB_CUDA(cudaMemcpyAsync(out_offsets_host, offs_dev.ptr, (size_t)(nparts + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    table_free(t);
    throw;
  }
  *out = t;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_partition_ids(const sb_table *in, const int32_t *key_cols, int32_t nkeys, int32_t num_partitions, sb_stream *s,
                     int32_t *out_ids_device) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(in && out_ids_device, "null argument");
  SB_REQUIRE(num_partitions >= 1, "num_partitions must be >= 1");
  cudaStream_t st = stream_of(s);
  KeyCols keys = make_keys(in, key_cols, nkeys);
  int64_t n = in->nrows;
  if (n > 0) {
    PartGeometry g = part_geometry(n);
    pid_hist_kernel<<<g.nblocks, PART_THREADS, 0, st>>>(keys, n, num_partitions, g.chunk, out_ids_device, nullptr, 0, 0);
    SB_LAUNCH_CHECK();
  }
  SB_API_END
}

int sb_hash_partition(const sb_table *in, const int32_t *key_cols, int32_t nkeys, int32_t num_partitions, sb_stream *s,
                      sb_table **out, int64_t *out_offsets_host) {
  SB_API_BEGIN
  require_init();
  partition_impl(in, key_cols, nkeys, num_partitions, 0, 0, stream_of(s), out, out_offsets_host);
  SB_API_END
}

int sb_round_robin_partition(const sb_table *in, int32_t start, int32_t num_partitions, sb_stream *s, sb_table **out,
                             int64_t *out_offsets_host) {
  SB_API_BEGIN
  require_init();
  partition_impl(in, nullptr, 0, num_partitions, 1, start, stream_of(s), out, out_offsets_host);
  SB_API_END
}

}  // extern "C"
