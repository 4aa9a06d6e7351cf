// radix.cu -- the key sort under SortExec: stable LSD radix sort of (64-bit prefix, row id) pairs.
//
// Reference: RadixSort.sortKeyPrefixArray (core/src/main/java/org/apache/spark/util/collection/unsafe/sort/RadixSort.java:178-259):
// LSD over 8-bit digits of the 64-bit prefix, (prefix, record pointer) pairs moved together, one counting pre-pass that also
// tells which bytes are identical in every record so their passes are skipped (:213-236).  A stable sort's result does not
// depend on how the passes are organised, so the GPU version keeps those rules and changes the machinery:
//   * ONE read of the keys builds all eight 256-bin histograms (and so decides which passes run);
//   * every pass is ONE kernel ("onesweep"): a block takes the next 4096-pair tile (ticket counter, so tiles run in memory
//     order), ranks its keys stably per warp (eight ballots per key over per-warp digit counters), publishes the tile's 256 digit counts
//     and obtains its exclusive prefix over all earlier tiles by DECOUPLED LOOK-BACK (64-bit status words: aggregate / inclusive
//     prefix) -- no separate histogram + scan + scatter launches and no second read of the keys -- then stages the tile in
//     shared memory in digit order and writes every digit's run contiguously (coalesced 8 + 4 byte stores).
// HBM traffic per pass: pairs read once, written once (24 B per pair); algorithmic bytes of the whole sort: SURVEY.md 8d
// counts 2 x 12 B per pair, passes are reported separately by the bench.
#include "radix.cuh"
#include "rtc.cuh"

namespace sb {

constexpr uint64_t RS_FLAG_AGG = 1ull << 62, RS_FLAG_INCL = 2ull << 62, RS_VALUE_MASK = (1ull << 62) - 1;

// all eight digit histograms in one read of the keys; also used to skip constant bytes
__global__ void __launch_bounds__(256) rs_histogram_kernel(const uint64_t *__restrict__ keys, int64_t n, unsigned long long *__restrict__ counts) {
  __shared__ uint32_t sh[8 * 256];
  for (int i = threadIdx.x; i < 8 * 256; i += 256) sh[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const uint64_t k = keys[i];
#pragma unroll
    for (int b = 0; b < 8; b++) atomicAdd(&sh[b * 256 + ((k >> (8 * b)) & 0xff)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 256; i += 256)
    if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
}

// counts[8][256] -> exclusive bases per byte (in place)
__global__ void __launch_bounds__(256) rs_scan_kernel(unsigned long long *__restrict__ counts) {
  __shared__ unsigned long long s[256];
  const int b = blockIdx.x, t = threadIdx.x;
  const unsigned long long c = counts[b * 256 + t];
  s[t] = c;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {
    unsigned long long v = t >= d ? s[t - d] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  counts[b * 256 + t] = s[t] - c;
}

__device__ __forceinline__ uint64_t ld_status(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Lanes of the warp that hold the same 8-bit digit: eight ballots, one per digit bit.  One bit test feeds the vote AND the select
// (LOP3.P, VOTE, SEL, LOP3 = 4 instructions per bit); written in PTX because the C form comes back from the optimiser as shift + mask +
// compare per use (7 per bit; ncu, round 2: the pass is issue-bound at 173 thread-instructions per key).
template <int B>
__device__ __forceinline__ void peer_bit(uint32_t d, uint32_t &m) {
  uint32_t bal, sel;
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
               "selp.b32 %1, 0, 0xffffffff, p;\n\t}" : "=r"(bal), "=r"(sel) : "r"(d), "n"(1 << B));
  m &= bal ^ sel;
}
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, uint32_t m) {
  peer_bit<0>(d, m); peer_bit<1>(d, m); peer_bit<2>(d, m); peer_bit<3>(d, m);
  peer_bit<4>(d, m); peer_bit<5>(d, m); peer_bit<6>(d, m); peer_bit<7>(d, m);
  return m;
}

// One LSD pass over byte `byte`.  status: [tiles][256] zero-initialised; ticket: zero-initialised tile counter.
// RS_THREADS x RS_ITEMS pairs per tile (a warp owns 32 x RS_ITEMS consecutive rows); digit d is owned by thread d (RS_THREADS >= 256).
// FULL = the tile holds RS_TILE pairs (all but the last one): no bounds predicates anywhere in the unrolled per-key code.
// n < 2^32 (SortExec's row ids are 32-bit), so output positions are 32-bit numbers.
template <int RS_THREADS, int RS_ITEMS, bool FULL>
__device__ __forceinline__ void rs_onesweep_tile(const uint64_t *__restrict__ in_keys, const uint32_t *__restrict__ in_vals,
                                                 uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals, int64_t tile, int tile_n, int shift,
                                                 const unsigned long long *__restrict__ gbase, uint64_t *__restrict__ status, uint8_t *rs_smem) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;
  uint64_t *s_keys = (uint64_t *)rs_smem;                                   // [RS_TILE]
  uint32_t *s_dst_off = (uint32_t *)(s_keys + RS_TILE);                     // [256] global position of sorted tile position p with digit d: s_dst_off[d] + p (mod 2^32)
  uint32_t *s_bin_start = s_dst_off + 256;                                  // [256] first tile-local position of every digit
  uint32_t *s_vals = s_bin_start + 256;                                     // [RS_TILE]
  uint32_t(*s_whist)[256] = (uint32_t(*)[256])(s_vals + RS_TILE);           // [RS_WARPS][256] per-warp digit counters, then exclusive prefixes over the warps
  uint32_t(*s_wrun)[256] = s_whist + RS_WARPS;                              // [RS_WARPS][256] running per-warp counters of the ranking phase
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile_base = tile * RS_TILE;
  // ---- load (warp-striped: warp w owns rows [w * 32 * RS_ITEMS, ...) of the tile; item k of lane l is row seg + k * 32 + l) ----
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint16_t rank[RS_ITEMS];
  const int seg = warp * (32 * RS_ITEMS);
  {
    const uint64_t *kp = in_keys + tile_base + seg + lane;
    const uint32_t *vp = in_vals + tile_base + seg + lane;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      if (FULL || seg + k * 32 + lane < tile_n) {
        key[k] = kp[k * 32];
        val[k] = vp[k * 32];
      } else {
        key[k] = ~0ull;
        val[k] = 0;
      }
    }
  }
  // ---- early counts: per-warp digit histograms by shared-memory atomics, so the tile's 256 counts can be PUBLISHED before the
  // (long) ranking phase -- successors then find an aggregate, or already an inclusive prefix, when they look back -----------------
  uint32_t *wh = s_whist[warp];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++)
    if (FULL || seg + k * 32 + lane < tile_n) atomicAdd(&wh[(uint32_t)(key[k] >> shift) & 0xffu], 1u);
  __syncthreads();
  const bool digit_owner = tid < 256;
  uint32_t count = 0;
  uint64_t *my_status = status + tile * 256 + (digit_owner ? tid : 0);
  if (digit_owner) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {   // per-warp counts -> exclusive prefix over the warps (order of the warps = order of the rows)
      const uint32_t c = s_whist[w][tid];
      s_whist[w][tid] = count;
      count += c;
    }
    st_status(my_status, RS_FLAG_AGG | count);
  }
  // ---- stable rank inside the warp's segment: (k, lane) order is memory order -----------------------------------------------------
  // Which lanes hold the same digit?  match.any answers in one instruction but with a long, serialising latency (ncu, round 2:
  // 17 of 27 warp-stall samples per issue were the instruction after MATCH waiting for it).  Eight ballots -- one per digit
  // bit, all independent across bits AND across the thread's RS_ITEMS keys -- give the same mask and pipeline freely.
  // (Measured, 25 M keys, 8 passes: ballots 1.82 ms; RS_ITEMS independent MATCHes issued back to back 2.69 ms.)
  uint32_t *wr = s_wrun[warp];
  uint32_t peers[RS_ITEMS];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const uint32_t d = (uint32_t)(key[k] >> shift) & 0xffu;
    if (FULL) peers[k] = digit_peers(d, 0xffffffffu);
    else {
      const bool valid = seg + k * 32 + lane < tile_n;
      const uint32_t m = digit_peers(d, __ballot_sync(0xffffffffu, valid));
      peers[k] = valid ? m : 0u;
    }
  }
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const uint32_t p = peers[k];
    const uint32_t d = (uint32_t)(key[k] >> shift) & 0xffu;
    const int leader = __ffs(p) - 1;                    // -1 for rows past the end of the tile
    uint32_t base = 0;
    if (lane == leader) {
      base = wr[d];
      wr[d] = base + __popc(p);
    }
    base = __shfl_sync(0xffffffffu, base, FULL ? leader : (leader < 0 ? lane : leader));
    rank[k] = (uint16_t)(base + __popc(p & lt));
    __syncwarp();
  }
  // ---- tile-local exclusive scan of the 256 digit counts: warp scans + a scan of the eight warp totals ----------------------------
  __shared__ uint32_t s_wsum[8];
  uint32_t incl = count;
  if (digit_owner) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
  }
  __syncthreads();
  uint32_t bin_start = 0;
  if (digit_owner) {
    for (int w = 0; w < warp; w++) bin_start += s_wsum[w];
    bin_start += incl - count;
    s_bin_start[tid] = bin_start;
    // decoupled look-back: sum the aggregates of earlier tiles until one carries an inclusive prefix.  Four predecessors are
    // fetched per round trip: with hundreds of tiles in flight the walk is long, and one L2 latency per step made it the
    // critical path of a pass.
    uint64_t excl = 0;
    int64_t t = tile - 1;
    bool done = t < 0;
    while (!done) {
      uint64_t sv[4];
#pragma unroll
      for (int j = 0; j < 4; j++) sv[j] = t - j >= 0 ? ld_status(status + (t - j) * 256 + tid) : RS_FLAG_INCL;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint64_t flag = sv[j] & ~RS_VALUE_MASK;
        if (flag == 0) break;             // not published yet (the tile holding that ticket is running): fetch again from here
        excl += sv[j] & RS_VALUE_MASK;
        t--;
        if (flag == RS_FLAG_INCL) { done = true; break; }
      }
    }
    st_status(my_status, RS_FLAG_INCL | (excl + count));
    s_dst_off[tid] = (uint32_t)gbase[tid] + (uint32_t)excl - bin_start;
  }
  __syncthreads();
  // ---- stage the tile in digit order ------------------------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    if (FULL || seg + k * 32 + lane < tile_n) {
      const uint32_t d = (uint32_t)(key[k] >> shift) & 0xffu;
      const uint32_t p = s_bin_start[d] + s_whist[warp][d] + rank[k];
      s_keys[p] = key[k];
      s_vals[p] = val[k];
    }
  }
  __syncthreads();
  // ---- every digit's run goes out contiguously ------------------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const int p = k * RS_THREADS + tid;
    if (FULL || p < tile_n) {
      const uint64_t kk = s_keys[p];
      const uint32_t dst = s_dst_off[(uint32_t)(kk >> shift) & 0xffu] + (uint32_t)p;
      out_keys[dst] = kk;
      out_vals[dst] = s_vals[p];
    }
  }
}

template <int RS_THREADS, int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, (RS_THREADS * RS_ITEMS <= 4096 ? 768 : 1024) / RS_THREADS) rs_onesweep_kernel(const uint64_t *__restrict__ in_keys, const uint32_t *__restrict__ in_vals,
                                                                 uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals, int64_t n, int byte,
                                                                 const unsigned long long *__restrict__ gbase /* [256] of this byte */,
                                                                 uint64_t *__restrict__ status, uint32_t *__restrict__ ticket) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;
  static_assert(RS_THREADS >= 256 && RS_THREADS % 32 == 0, "one thread per digit");
  extern __shared__ __align__(16) uint8_t rs_smem[];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  {   // the two counter arrays ([2][RS_WARPS][256] words behind keys, offsets and values) start at zero: 16-byte stores
    uint4 *z = (uint4 *)(rs_smem + (size_t)RS_TILE * 12 + 2 * 256 * 4);
    for (int i = tid; i < 2 * RS_WARPS * 256 / 4; i += RS_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t tile_base = tile * RS_TILE;
  const int tile_n = (int)(n - tile_base < RS_TILE ? n - tile_base : RS_TILE);
  if (tile_n == RS_TILE) rs_onesweep_tile<RS_THREADS, RS_ITEMS, true>(in_keys, in_vals, out_keys, out_vals, tile, tile_n, 8 * byte, gbase, status, rs_smem);
  else rs_onesweep_tile<RS_THREADS, RS_ITEMS, false>(in_keys, in_vals, out_keys, out_vals, tile, tile_n, 8 * byte, gbase, status, rs_smem);
}

// Tiny inputs (the 4-row result of Q1, top-N candidates): one block ranks every element by counting -- rank = #keys smaller +
// #equal keys with a smaller position -- which is a stable sort in one launch with no host round trip.
constexpr int SMALL_SORT_MAX = 2048;
__global__ void __launch_bounds__(256) small_sort_kernel(uint64_t *keys, uint32_t *vals, int n) {
  __shared__ uint64_t sk[SMALL_SORT_MAX];
  __shared__ uint32_t sv[SMALL_SORT_MAX];
  for (int i = threadIdx.x; i < n; i += 256) { sk[i] = keys[i]; sv[i] = vals[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const uint64_t k = sk[i];
    int rank = 0;
    for (int j = 0; j < n; j++) rank += (sk[j] < k) || (sk[j] == k && j < i);
    keys[rank] = k;
    vals[rank] = sv[i];
  }
}

int radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, cudaStream_t st, uint64_t *keys_alt, uint64_t **sorted_keys) {
  if (sorted_keys) *sorted_keys = keys;
  if (n <= 1) return 0;
  SB_REQUIRE(n < (1ll << 32), "radix_sort_pairs: row ids and output positions are 32-bit");
  if (n <= SMALL_SORT_MAX) {
    small_sort_kernel<<<1, 256, 0, st>>>(keys, vals, (int)n);
    SB_LAUNCH_CHECK();
    return 1;
  }
  Scratch counts(8 * 256 * 8, st);
  SB_CUDA(cudaMemsetAsync(counts.ptr, 0, 8 * 256 * 8, st));
  {
    KernelTimer kt("sort_histogram", st);
    const int grid = grid_for(n, 256 * 16, rt().num_sms * 8);
    rs_histogram_kernel<<<grid, 256, 0, st>>>(keys, n, counts.as<unsigned long long>());
    SB_LAUNCH_CHECK();
  }
  std::vector<unsigned long long> h(8 * 256);
  SB_CUDA(cudaMemcpyAsync(h.data(), counts.ptr, 8 * 256 * 8, cudaMemcpyDeviceToHost, st));
  rs_scan_kernel<<<8, 256, 0, st>>>(counts.as<unsigned long long>());   // runs while the host looks at the counts
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaStreamSynchronize(st));
  int bytes[8], passes = 0;
  for (int b = 0; b < 8; b++) {
    bool varies = true;
    for (int d = 0; d < 256; d++)
      if (h[b * 256 + d] == (unsigned long long)n) varies = false;   // every record shares this byte: skip the pass (RadixSort.java:213-236)
    if (varies) bytes[passes++] = b;
  }
  if (passes == 0) return 0;
  // tile geometry (sb_config_set("sort_variant", v) for experiments): threads x items
  struct Variant { const void *fn; int threads, items; };
  static const Variant variants[] = {
      {(const void *)rs_onesweep_kernel<256, 16>, 256, 16}, {(const void *)rs_onesweep_kernel<512, 8>, 512, 8},
      {(const void *)rs_onesweep_kernel<256, 8>, 256, 8},   {(const void *)rs_onesweep_kernel<512, 16>, 512, 16},
      {(const void *)rs_onesweep_kernel<384, 12>, 384, 12}, {(const void *)rs_onesweep_kernel<1024, 8>, 1024, 8}};
  int vi = config().sort_variant;
  if (vi < 0 || vi >= (int)(sizeof(variants) / sizeof(variants[0]))) vi = 0;
  const Variant &V = variants[vi];
  const int tile_rows = V.threads * V.items;
  const size_t smem = (size_t)tile_rows * 12 + 2 * 256 * 4 + (size_t)(V.threads / 32) * 256 * 4 * 2;
  SB_CUDA(cudaFuncSetAttribute(V.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (n + tile_rows - 1) / tile_rows;
  Scratch keys2(keys_alt ? 16 : n * 8 + 16, st), vals2(n * 4 + 16, st), status(tiles * 256 * 8 + 16, st), tickets(8 * 4, st);
  SB_CUDA(cudaMemsetAsync(tickets.ptr, 0, 32, st));
  uint64_t *ik = keys, *ok = keys_alt ? keys_alt : keys2.as<uint64_t>();
  uint32_t *iv = vals, *ov = vals2.as<uint32_t>();
  if (passes & 1) {   // the sorted values must land in `vals`: an odd number of passes starts from the scratch copy
    SB_CUDA(cudaMemcpyAsync(vals2.ptr, vals, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    std::swap(iv, ov);
  }
  KernelTimer kt("sort_passes", st);
  for (int p = 0; p < passes; p++) {
    SB_CUDA(cudaMemsetAsync(status.ptr, 0, (size_t)tiles * 256 * 8, st));   // one status array, re-armed per pass (passes are stream-ordered)
    const unsigned long long *gb = counts.as<unsigned long long>() + bytes[p] * 256;
    uint64_t *stp = status.as<uint64_t>();
    uint32_t *tk = tickets.as<uint32_t>() + p;
    int byte = bytes[p];
    void *args[] = {&ik, &iv, &ok, &ov, (void *)&n, &byte, &gb, &stp, &tk};
    SB_CUDA(cudaLaunchKernel(V.fn, dim3((unsigned)tiles), dim3((unsigned)V.threads), args, smem, st));
    count_launch();
    std::swap(ik, ok);
    std::swap(iv, ov);
  }
  if (sorted_keys) *sorted_keys = ik;   // after the last swap `ik` is where the last pass wrote
  return passes;   // scratch buffers are freed in stream order
}

}  // namespace sb
