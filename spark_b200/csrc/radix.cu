// radix.cu -- the key sort under SortExec: stable LSD radix sort of (64-bit prefix, row id) pairs.
//
// Reference: RadixSort.sortKeyPrefixArray (core/src/main/java/org/apache/spark/util/collection/unsafe/sort/RadixSort.java:178-259):
// LSD over 8-bit digits of the 64-bit prefix, (prefix, record pointer) pairs moved together, one counting pre-pass that also
// tells which bytes are identical in every record so their passes are skipped (:213-236).  A stable sort's result does not
// depend on how the passes are organised, so the GPU version keeps those rules and changes the machinery:
//   * ONE read of the keys builds all eight 256-bin histograms (and so decides which passes run);
//   * every pass is ONE kernel ("onesweep"): a block takes the next 4096-pair tile (ticket counter, so tiles run in memory
//     order), ranks its keys stably per warp with match.any over per-warp digit counters, publishes the tile's 256 digit counts
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

// One LSD pass over byte `byte`.  status: [tiles][256] zero-initialised; ticket: zero-initialised tile counter.
// RS_THREADS x RS_ITEMS pairs per tile (a warp owns 32 x RS_ITEMS consecutive rows); digit d is owned by thread d (RS_THREADS >= 256).
template <int RS_THREADS, int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, (RS_THREADS * RS_ITEMS <= 4096 ? 768 : 1024) / RS_THREADS) rs_onesweep_kernel(const uint64_t *__restrict__ in_keys, const uint32_t *__restrict__ in_vals,
                                                                 uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals, int64_t n, int byte,
                                                                 const unsigned long long *__restrict__ gbase /* [256] of this byte */,
                                                                 uint64_t *__restrict__ status, uint32_t *__restrict__ ticket) {
  constexpr int RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;
  static_assert(RS_THREADS >= 256 && RS_THREADS % 32 == 0, "one thread per digit");
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint64_t *s_keys = (uint64_t *)rs_smem;                                   // [RS_TILE]
  int64_t *s_dst_off = (int64_t *)(s_keys + RS_TILE);                       // [256] global position of sorted tile position p with digit d: s_dst_off[d] + p
  uint32_t *s_vals = (uint32_t *)(s_dst_off + 256);                         // [RS_TILE]
  uint32_t(*s_whist)[256] = (uint32_t(*)[256])(s_vals + RS_TILE);           // [RS_WARPS][256] per-warp digit counters, then exclusive prefixes over the warps
  uint32_t(*s_wrun)[256] = s_whist + RS_WARPS;                              // [RS_WARPS][256] running per-warp counters of the ranking phase
  uint32_t *s_bin_start = (uint32_t *)(s_wrun + RS_WARPS);                  // [256] first tile-local position of every digit
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < 2 * RS_WARPS * 256; i += RS_THREADS) (&s_whist[0][0])[i] = 0;   // s_whist and s_wrun
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t tile_base = tile * RS_TILE;
  const int tile_n = (int)(n - tile_base < RS_TILE ? n - tile_base : RS_TILE);
  const int shift = 8 * byte;
  // ---- load (warp-striped: warp w owns rows [w * 512, (w + 1) * 512) of the tile; item k of lane l is row w * 512 + k * 32 + l) ----
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint16_t rank[RS_ITEMS];
  const int seg = warp * (32 * RS_ITEMS);
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const int r = seg + k * 32 + lane;
    if (r < tile_n) {
      key[k] = in_keys[tile_base + r];
      val[k] = in_vals[tile_base + r];
    } else {
      key[k] = ~0ull;
      val[k] = 0;
    }
  }
  // ---- early counts: per-warp digit histograms by shared-memory atomics, so the tile's 256 counts can be PUBLISHED before the
  // (long) ranking phase -- successors then find an aggregate, or already an inclusive prefix, when they look back -----------------
  uint32_t *wh = s_whist[warp];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++)
    if (seg + k * 32 + lane < tile_n) atomicAdd(&wh[(uint32_t)((key[k] >> shift) & 0xff)], 1u);
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
    const bool valid = seg + k * 32 + lane < tile_n;
    const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
    uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
      m &= ((d >> b) & 1) ? bal : ~bal;
    }
    peers[k] = valid ? m : 0u;
  }
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const uint32_t p = peers[k];
    const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
    const int leader = __ffs(p) - 1;                    // -1 for rows past the end of the tile
    uint32_t base = 0;
    if (lane == leader) {
      base = wr[d];
      wr[d] = base + __popc(p);
    }
    base = __shfl_sync(0xffffffffu, base, leader < 0 ? lane : leader);
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
    s_dst_off[tid] = (int64_t)gbase[tid] + (int64_t)excl - (int64_t)bin_start;
  }
  __syncthreads();
  // ---- stage the tile in digit order ------------------------------------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    if (seg + k * 32 + lane < tile_n) {
      const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
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
    if (p < tile_n) {
      const uint64_t kk = s_keys[p];
      const int64_t dst = s_dst_off[(kk >> shift) & 0xff] + p;
      out_keys[dst] = kk;
      out_vals[dst] = s_vals[p];
    }
  }
}

// ---- onesweep, second form: a dedicated LOOK-BACK WARP ----------------------------------------------------------------------------------
// ncu of the first form showed the digit-owner threads spinning in the look-back after their ranking while the other warps sat at the
// barrier.  Here the last warp of the block does nothing but the look-back (lane l owns digits l, l + 32, ... l + 224: eight coalesced
// 256-byte status reads per predecessor, four predecessors per round trip), concurrently with the data warps' ranking and staging.  The
// tile-local scan runs BEFORE the ranking, so the running per-warp counters start at the final tile positions and the ranking yields the
// staging position directly (one LDS + one STS by the leader lane, no second lookup).  Status words carry the pass number (epoch), so
// the status array is zeroed once per sort instead of once per pass.
//   word = flag(2) | epoch(6) | value(56);  flag 1 = aggregate of the tile, 2 = inclusive prefix, 0 / other epoch = not published
// EARLY = per-warp counts by shared-memory atomics before the ranking (published early, look-back overlaps the ranking); otherwise the
// counts fall out of the ranking and the look-back overlaps only the staging.
constexpr int RS2_EPOCH_SHIFT = 56;
constexpr uint64_t RS2_VALUE_MASK = (1ull << RS2_EPOCH_SHIFT) - 1;
template <int RS_THREADS, int RS_ITEMS, bool EARLY, int LB_W>
__global__ void __launch_bounds__(RS_THREADS, ((RS_THREADS - 32) * RS_ITEMS <= 4608 ? 2 : 1)) rs_onesweep2_kernel(
    const uint64_t *__restrict__ in_keys, const uint32_t *__restrict__ in_vals, uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_vals,
    int64_t n, int byte, const unsigned long long *__restrict__ gbase, uint64_t *__restrict__ status, uint32_t *__restrict__ ticket, uint32_t epoch) {
  constexpr int DT = RS_THREADS - 32, DW = DT / 32, RS_TILE = DT * RS_ITEMS;
  static_assert(DT >= 256, "one data thread per digit");
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint64_t *s_keys = (uint64_t *)rs_smem;                           // [RS_TILE]
  int64_t *s_dst_off = (int64_t *)(s_keys + RS_TILE);               // [256]
  uint32_t *s_vals = (uint32_t *)(s_dst_off + 256);                 // [RS_TILE]
  uint32_t(*s_wpos)[256] = (uint32_t(*)[256])(s_vals + RS_TILE);    // [DW][256] counts -> running tile positions of (warp, digit)
  uint32_t *s_bin_start = (uint32_t *)(s_wpos + DW);                // [256]
  uint32_t *s_count = s_bin_start + 256;                            // [256]
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_wsum[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool data = warp < DW;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < DW * 256; i += RS_THREADS) (&s_wpos[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t tile_base = tile * RS_TILE;
  const int tile_n = (int)(n - tile_base < RS_TILE ? n - tile_base : RS_TILE);
  const bool full = tile_n == RS_TILE;
  const int shift = 8 * byte;
  const uint64_t tag_agg = (1ull << 62) | ((uint64_t)epoch << RS2_EPOCH_SHIFT), tag_incl = (2ull << 62) | ((uint64_t)epoch << RS2_EPOCH_SHIFT);
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint16_t pos[RS_ITEMS];
  const int seg = warp * (32 * RS_ITEMS);
  if (data) {
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      const int r = seg + k * 32 + lane;
      if (full || r < tile_n) {
        key[k] = in_keys[tile_base + r];
        val[k] = in_vals[tile_base + r];
      } else {
        key[k] = ~0ull;
        val[k] = 0;
      }
    }
  }
  uint32_t *wp = s_wpos[data ? warp : 0];
  const uint32_t lt = (1u << lane) - 1;
  // peers of item k: lanes of the warp holding the same digit (eight independent ballots, see the first form)
  auto peers_of = [&](int k, uint32_t d) -> uint32_t {
    uint32_t m = full ? 0xffffffffu : __ballot_sync(0xffffffffu, seg + k * 32 + lane < tile_n);
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
      m &= ((d >> b) & 1) ? bal : ~bal;
    }
    return (full || seg + k * 32 + lane < tile_n) ? m : 0u;
  };
  if (EARLY) {
    if (data) {
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++)
        if (full || seg + k * 32 + lane < tile_n) atomicAdd(&wp[(uint32_t)((key[k] >> shift) & 0xff)], 1u);
    }
  } else {
    if (data) {   // ranking first: pos = rank inside the warp's segment, the counters end up as the warp's digit counts
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
        const uint32_t p = peers_of(k, d);
        const int leader = __ffs(p) - 1;
        uint32_t base = 0;
        if (lane == leader) {
          base = wp[d];
          wp[d] = base + __popc(p);
        }
        base = __shfl_sync(0xffffffffu, base, leader < 0 ? lane : leader);
        pos[k] = (uint16_t)(base + __popc(p & lt));
        __syncwarp();
      }
    }
  }
  __syncthreads();
  // ---- digit owners: counts over the warps, publish the aggregate, tile-local scan ----------------------------------------------------
  const bool digit_owner = tid < 256;
  uint32_t count = 0, incl = 0;
  uint32_t wc[DW];
  if (digit_owner) {
#pragma unroll
    for (int w = 0; w < DW; w++) {
      wc[w] = s_wpos[w][tid];
      count += wc[w];
    }
    st_status(status + tile * 256 + tid, tag_agg | count);
    s_count[tid] = count;
    incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
  }
  __syncthreads();
  if (digit_owner) {
    uint32_t bin_start = incl - count;
    for (int w = 0; w < warp; w++) bin_start += s_wsum[w];
    s_bin_start[tid] = bin_start;
    uint32_t run = bin_start;
#pragma unroll
    for (int w = 0; w < DW; w++) {   // (warp, digit) -> first tile position of that warp's keys with that digit
      s_wpos[w][tid] = run;
      run += wc[w];
    }
  }
  __syncthreads();
  if (!data) {
    // ---- the look-back warp: exclusive prefix of every digit over the earlier tiles --------------------------------------------------
    uint64_t excl[8];
    bool done[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { excl[j] = 0; done[j] = false; }
    int64_t t = tile - 1;
    while (t >= 0) {
      uint64_t sv[LB_W][8];
#pragma unroll
      for (int i = 0; i < LB_W; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) sv[i][j] = t - i >= 0 ? ld_status(status + (t - i) * 256 + j * 32 + lane) : tag_incl;
      // how many of the LB_W tiles can every digit of every lane consume?  (stop after an inclusive prefix, or at an unpublished word)
      int take = LB_W;
      bool all_done_after = true;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (done[j]) continue;
        int c = 0;
        bool fin = false;
#pragma unroll
        for (int i = 0; i < LB_W; i++) {
          if (fin || c < i) continue;
          const uint64_t tag = sv[i][j] & ~RS2_VALUE_MASK;
          if (tag == tag_incl) { c = i + 1; fin = true; }
          else if (tag == tag_agg) c = i + 1;
        }
        if (!fin) { take = c < take ? c : take; all_done_after = false; }
      }
      // lanes advance together by the smallest count (the words of one tile are published together, so little is lost)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int ot = __shfl_xor_sync(0xffffffffu, take, o);
        take = ot < take ? ot : take;
      }
      const bool warp_done = __all_sync(0xffffffffu, all_done_after);
      const int use = warp_done ? LB_W : take;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (done[j]) continue;
#pragma unroll
        for (int i = 0; i < LB_W; i++) {
          if (i < use && !done[j]) {
            excl[j] += sv[i][j] & RS2_VALUE_MASK;
            if ((sv[i][j] & ~RS2_VALUE_MASK) == tag_incl) done[j] = true;
          }
        }
      }
      if (warp_done) break;
      t -= use;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int d = j * 32 + lane;
      st_status(status + tile * 256 + d, tag_incl | (excl[j] + s_count[d]));
      s_dst_off[d] = (int64_t)gbase[d] + (int64_t)excl[j] - (int64_t)s_bin_start[d];
    }
  } else {
    // ---- data warps: rank (EARLY) and stage the tile in digit order ------------------------------------------------------------------
    if (EARLY) {
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t d = (uint32_t)((key[k] >> shift) & 0xff);
        const uint32_t p = peers_of(k, d);
        const int leader = __ffs(p) - 1;
        uint32_t base = 0;
        if (lane == leader) {
          base = wp[d];
          wp[d] = base + __popc(p);
        }
        base = __shfl_sync(0xffffffffu, base, leader < 0 ? lane : leader);
        pos[k] = (uint16_t)(base + __popc(p & lt));
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        if (full || seg + k * 32 + lane < tile_n) {
          s_keys[pos[k]] = key[k];
          s_vals[pos[k]] = val[k];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < RS_ITEMS; k++) {
        if (full || seg + k * 32 + lane < tile_n) {
          const uint32_t p = wp[(uint32_t)((key[k] >> shift) & 0xff)] + pos[k];
          s_keys[p] = key[k];
          s_vals[p] = val[k];
        }
      }
    }
  }
  __syncthreads();
  if (data) {
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      const int p = k * DT + tid;
      if (full || p < tile_n) {
        const uint64_t kk = s_keys[p];
        const int64_t dst = s_dst_off[(kk >> shift) & 0xff] + p;
        out_keys[dst] = kk;
        out_vals[dst] = s_vals[p];
      }
    }
  }
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
  // tile geometry (sb_config_set("sort_variant", v) for experiments): threads x items; form 2 = dedicated look-back warp
  struct Variant { const void *fn; int threads, items, form; };
  static const Variant variants[] = {
      {(const void *)rs_onesweep_kernel<256, 16>, 256, 16, 1}, {(const void *)rs_onesweep_kernel<512, 8>, 512, 8, 1},
      {(const void *)rs_onesweep_kernel<256, 8>, 256, 8, 1},   {(const void *)rs_onesweep_kernel<512, 16>, 512, 16, 1},
      {(const void *)rs_onesweep_kernel<384, 12>, 384, 12, 1}, {(const void *)rs_onesweep_kernel<1024, 8>, 1024, 8, 1},
      {(const void *)rs_onesweep2_kernel<384, 12, true, 2>, 384, 12, 2},    // 6
      {(const void *)rs_onesweep2_kernel<384, 12, false, 2>, 384, 12, 2},   // 7
      {(const void *)rs_onesweep2_kernel<416, 12, true, 2>, 416, 12, 2},    // 8
      {(const void *)rs_onesweep2_kernel<320, 16, true, 2>, 320, 16, 2},    // 9
      {(const void *)rs_onesweep2_kernel<544, 8, true, 2>, 544, 8, 2},      // 10
      {(const void *)rs_onesweep2_kernel<384, 12, true, 4>, 384, 12, 2},    // 11
      {(const void *)rs_onesweep2_kernel<288, 16, true, 2>, 288, 16, 2},    // 12
      {(const void *)rs_onesweep2_kernel<1024, 8, true, 4>, 1024, 8, 2},    // 13: one block per SM
  };
  int vi = config().sort_variant;
  if (vi < 0 || vi >= (int)(sizeof(variants) / sizeof(variants[0]))) vi = 0;
  const Variant &V = variants[vi];
  const int data_threads = V.form == 2 ? V.threads - 32 : V.threads;
  const int tile_rows = data_threads * V.items;
  const size_t smem = V.form == 2 ? (size_t)tile_rows * 12 + 256 * 8 + (size_t)(data_threads / 32) * 256 * 4 + 2 * 256 * 4
                                  : (size_t)tile_rows * 12 + 256 * 8 + (size_t)(V.threads / 32) * 256 * 4 * 2 + 256 * 4;
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
  if (V.form == 2) SB_CUDA(cudaMemsetAsync(status.ptr, 0, (size_t)tiles * 256 * 8, st));   // once: the words carry the pass number
  for (int p = 0; p < passes; p++) {
    if (V.form == 1) SB_CUDA(cudaMemsetAsync(status.ptr, 0, (size_t)tiles * 256 * 8, st));   // re-armed per pass (passes are stream-ordered)
    const unsigned long long *gb = counts.as<unsigned long long>() + bytes[p] * 256;
    uint64_t *stp = status.as<uint64_t>();
    uint32_t *tk = tickets.as<uint32_t>() + p;
    int byte = bytes[p];
    uint32_t epoch = (uint32_t)p + 1;
    void *args[] = {&ik, &iv, &ok, &ov, (void *)&n, &byte, &gb, &stp, &tk, &epoch};   // form 1 ignores the last one
    SB_CUDA(cudaLaunchKernel(V.fn, dim3((unsigned)tiles), dim3((unsigned)V.threads), args, smem, st));
    count_launch();
    std::swap(ik, ok);
    std::swap(iv, ov);
  }
  if (sorted_keys) *sorted_keys = ik;   // after the last swap `ik` is where the last pass wrote
  return passes;   // scratch buffers are freed in stream order
}

}  // namespace sb
