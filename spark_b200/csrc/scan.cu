// scan.cu -- the columnar scan boundary: Parquet column-chunk pages (host memory) -> Arrow columns in HBM.
//
// Reference path replaced (SQLJ = sql/core/src/main/java/org/apache/spark/sql/execution/datasources/parquet):
//   FileSourceScanExec.doExecuteColumnar (SQLX/DataSourceScanExec.scala:735-760) pulls ColumnarBatches from
//   VectorizedParquetRecordReader.nextBatch (SQLJ/VectorizedParquetRecordReader.java), whose per-column work is
//   VectorizedColumnReader.readBatch (SQLJ/VectorizedColumnReader.java:321-460: readPageV1 / readPageV2, definition levels,
//   initDataReader), VectorizedRleValuesReader (SQLJ/VectorizedRleValuesReader.java:95-117 initFromPage, :981-1020
//   readNextGroup: the RLE / bit-packed hybrid) and VectorizedPlainValuesReader, decoding on the CPU into
//   OffHeapColumnVectors that RowToColumnar / the plugin would then copy to the device.
// Here the ENCODED page bytes cross PCIe (one copy per column chunk) and the GPU decodes them: for TPC-H lineitem the
// dictionary-encoded columns are 1-12 bits per value instead of 8-64, which is what the end-to-end rate is bound by.
//
// Supported: physical types BOOLEAN / INT32 / INT64 / FLOAT / DOUBLE; encodings PLAIN and RLE_DICTIONARY (= PLAIN_DICTIONARY:
// bit-width byte + hybrid runs over a PLAIN dictionary page); data pages V1 (levels length-prefixed in front of the values)
// and V2 (levels in their own section); flat schemas (max definition level <= 1, no repetition); uncompressed pages (the
// codec is the file writer's choice; decompression is not part of this path).  BYTE_ARRAY and DELTA_* are rejected with
// SB_ERR_UNSUPPORTED rather than decoded on the CPU.
//
// Kernel: one 256-thread block per page.  (1) definition levels -> one validity byte per row; (2) values -> dense value
// space (page-local), PLAIN by a widening/narrowing copy, RLE_DICTIONARY by walking the hybrid stream: thread 0 parses up
// to 64 run headers into shared memory, then the whole block decodes those runs' values in parallel (value i of the batch
// finds its run by binary search in the batch's prefix sums -- balanced whatever the run lengths are) and gathers from the
// dictionary; (3) NULLable pages: block scan of the validity bytes maps rows to value indices (out[row] = dense[idx]).
// Algorithmic bytes: encoded bytes read once + decoded column written once.
#include <algorithm>
#include <memory>
#include "common.cuh"
#include "primitives.cuh"

namespace sb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_RUNS = 64;

struct PageTask {
  const uint8_t *values;     // device: encoded values of the page
  const uint8_t *def;        // device: definition-level hybrid stream (bit width 1) or nullptr
  int64_t values_bytes, def_bytes;
  int64_t row_start;         // first row of the page in the column
  int32_t num_values;        // rows of the page (NULLs included)
  int32_t encoding;          // SB_ENC_*
  int32_t col;
  int32_t pad;
};
struct ColTask {
  void *out;                 // decoded column (row space)
  uint8_t *valid_bytes;      // one byte per row, or nullptr when no page of the column carries definition levels
  uint8_t *dense;            // value-space staging for NULLable columns (same size as out), else nullptr
  const uint8_t *dict;       // device: PLAIN dictionary values
  int32_t dict_count;
  int32_t phys_width;        // 0 BOOLEAN (bit-packed), 4, 8
  int32_t out_width;         // 1, 2, 4, 8
  int32_t pad;
};

__device__ __forceinline__ uint64_t load_le(const uint8_t *p, int nbytes) {
  uint64_t w = 0;
  for (int b = 0; b < nbytes; b++) w |= (uint64_t)p[b] << (8 * b);
  return w;
}
__device__ __forceinline__ void store_value(void *out, int width, int64_t i, uint64_t v) {
  switch (width) {
    case 1: ((uint8_t *)out)[i] = (uint8_t)v; break;
    case 2: ((uint16_t *)out)[i] = (uint16_t)v; break;
    case 4: ((uint32_t *)out)[i] = (uint32_t)v; break;
    default: ((uint64_t *)out)[i] = v; break;
  }
}

// unsigned LEB128 at p[pos] (bounded by nbytes); returns the value, *len = bytes consumed
__device__ __forceinline__ uint32_t read_uvarint(const uint8_t *__restrict__ p, int64_t pos, int64_t nbytes, int *len) {
  uint32_t v = 0;
  int shift = 0, l = 0;
  while (pos + l < nbytes) {
    const uint8_t b = p[pos + l++];
    v |= (uint32_t)(b & 0x7f) << shift;
    shift += 7;
    if (!(b & 0x80) || shift > 28) break;
  }
  *len = l;
  return v;
}

// Walks an RLE / bit-packed hybrid stream (VectorizedRleValuesReader.readNextGroup) and calls emit(index, value) for the first
// max_values values.  Block-cooperative; returns the number of values produced (same in every thread).
//   * Run headers form a chain (a header says where the next one is), and every link costs a memory round trip.  Writers emit
//     long stretches of identical runs (parquet-mr: bit-packed runs of 504 values, header 0x7F), so warp 0 walks the chain
//     SPECULATIVELY: it reads the header at the cursor, then lane l checks that the same header sits l run-lengths further
//     on; the leading lanes that agree are 32 links resolved in two round trips instead of 32.  Irregular streams degrade to
//     one link per step, never to a wrong answer.
//   * Up to 64 runs form a batch; the batch's values are decoded by the whole block (value i finds its run by binary search
//     in the batch's prefix sums).  part / nparts stripes the BATCHES of one page over several blocks: every block walks the
//     whole chain (cheap), but decodes only its own batches, so a page of a million values is not one block's job.
template <class Emit>
__device__ int64_t hybrid_decode(const uint8_t *__restrict__ p, int64_t nbytes, int bw, int64_t max_values, int part, int nparts, Emit emit) {
  __shared__ uint64_t s_arg[SCAN_RUNS];        // RLE: the value; packed: byte offset of the run's first group
  __shared__ uint8_t s_rle[SCAN_RUNS];
  __shared__ uint32_t s_pref[SCAN_RUNS + 1];
  __shared__ int s_nruns;
  __shared__ int64_t s_pos;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_pos = 0;
  __syncthreads();
  int64_t produced = 0;
  if (bw == 0) {   // a dictionary of one entry: every value is 0 (initFromPage :109-113)
    if (part == 0)
      for (int64_t i = tid; i < max_values; i += SCAN_THREADS) emit(i, 0ull);
    return max_values;
  }
  const int vbytes = (bw + 7) >> 3;
  for (int batch = 0; produced < max_values; batch++) {
    if (tid < 32) {
      int64_t pos = s_pos;
      int n = 0;
      int64_t acc = 0;
      const int64_t want = max_values - produced;
      if (lane == 0) s_pref[0] = 0;
      while (n < SCAN_RUNS && pos < nbytes && acc < want) {
        int hlen;
        const uint32_t header = read_uvarint(p, pos, nbytes, &hlen);
        const bool packed = header & 1;
        const int64_t c = packed ? (int64_t)(header >> 1) * 8 : (int64_t)(header >> 1);
        const int64_t run_bytes = hlen + (packed ? (int64_t)(header >> 1) * bw : vbytes);
        if (c <= 0) {   // empty run: nothing to decode, step over it
          pos += run_bytes;
          continue;
        }
        const int64_t q = pos + (int64_t)lane * run_bytes;
        bool same = lane == 0;
        if (lane > 0 && q + hlen <= nbytes) {
          int hl2;
          same = read_uvarint(p, q, nbytes, &hl2) == header && hl2 == hlen;
        }
        const uint32_t ball = __ballot_sync(0xffffffffu, same);
        int cnt = ball == 0xffffffffu ? 32 : __ffs(~ball) - 1;
        if (cnt > SCAN_RUNS - n) cnt = SCAN_RUNS - n;
        const int64_t needed = (want - acc + c - 1) / c;   // runs until the page's value count is reached
        if (cnt > needed) cnt = (int)needed;
        if (lane < cnt) {
          s_rle[n + lane] = packed ? 0 : 1;
          s_arg[n + lane] = packed ? (uint64_t)(q + hlen) : (q + hlen + vbytes <= nbytes ? load_le(p + q + hlen, vbytes) : 0);
          int64_t upto = acc + (int64_t)(lane + 1) * c;
          if (upto > want) upto = want;                    // the last packed group may be padded
          s_pref[n + lane + 1] = (uint32_t)upto;
        }
        n += cnt;
        pos += (int64_t)cnt * run_bytes;
        acc += (int64_t)cnt * c;
        if (acc > want) acc = want;
        if (acc >= (1ll << 30)) break;                     // keep the batch's prefix sums in 32 bits
      }
      if (lane == 0) {
        s_nruns = n;
        s_pos = pos;
      }
    }
    __syncthreads();
    const int nruns = s_nruns;
    const uint32_t total = nruns ? s_pref[nruns] : 0;
    if (nruns == 0 || total == 0) break;   // stream exhausted (malformed page: fewer values than announced)
    if (batch % nparts == part) {
      for (uint32_t i = tid; i < total; i += SCAN_THREADS) {
        int lo = 0, hi = nruns;               // last run with pref <= i
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_pref[mid] <= i) lo = mid; else hi = mid;
        }
        uint64_t v;
        if (s_rle[lo]) v = s_arg[lo];
        else {
          const uint64_t bit = (uint64_t)(i - s_pref[lo]) * (uint64_t)bw;
          const int64_t byte = (int64_t)s_arg[lo] + (int64_t)(bit >> 3);
          const int need = (int)(((bit & 7) + bw + 7) >> 3);
          const int avail = (int)(nbytes - byte < need ? (nbytes - byte > 0 ? nbytes - byte : 0) : need);
          v = (load_le(p + byte, avail) >> (bit & 7)) & ((bw >= 64) ? ~0ull : ((1ull << bw) - 1));
        }
        emit(produced + i, v);
      }
    }
    produced += total;
    __syncthreads();
  }
  return produced;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_decode_kernel(const PageTask *__restrict__ pages, const ColTask *__restrict__ cols, int nparts) {
  const PageTask pg = pages[blockIdx.x / nparts];
  int part = blockIdx.x % nparts;
  const ColTask c = cols[pg.col];
  const int tid = threadIdx.x;
  const int64_t n = pg.num_values;
  // ---- (1) definition levels -> validity bytes -------------------------------------------------------------------------
  uint8_t *vb = c.valid_bytes ? c.valid_bytes + pg.row_start : nullptr;
  const bool page_nullable = pg.def != nullptr && pg.def_bytes > 0;
  if (page_nullable || pg.encoding != SB_ENC_RLE_DICTIONARY) {   // only hybrid-encoded, NULL-free pages are striped over several blocks
    if (part != 0) return;
    nparts = 1;
  }
  if (vb && part == 0) {
    if (page_nullable) {
      const int64_t got = hybrid_decode(pg.def, pg.def_bytes, 1, n, 0, 1, [&](int64_t i, uint64_t v) { vb[i] = (uint8_t)(v & 1); });
      for (int64_t i = got + tid; i < n; i += SCAN_THREADS) vb[i] = 0;
    } else {
      for (int64_t i = tid; i < n; i += SCAN_THREADS) vb[i] = 1;
    }
    __syncthreads();
  }
  // ---- (2) values -> dense value space (or straight to the output when the page has no NULLs) -----------------------------
  void *dst_page = (uint8_t *)c.out + pg.row_start * c.out_width;
  void *dense = page_nullable ? (void *)(c.dense + pg.row_start * c.out_width) : dst_page;
  const int ow = c.out_width;
  if (pg.encoding == SB_ENC_PLAIN) {
    if (c.phys_width == 0) {   // BOOLEAN: bit-packed, LSB first
      const int64_t avail = pg.values_bytes * 8 < n ? pg.values_bytes * 8 : n;
      for (int64_t i = tid; i < avail; i += SCAN_THREADS) ((uint8_t *)dense)[i] = (pg.values[i >> 3] >> (i & 7)) & 1;
    } else {
      const int pw = c.phys_width;
      const int64_t avail = pg.values_bytes / pw < n ? pg.values_bytes / pw : n;
      if (pw == ow && ((((uintptr_t)pg.values) | ((uintptr_t)dense)) & 15) == 0) {   // straight copy, 16 bytes per thread
        const int64_t bytes = avail * pw, vec = bytes >> 4;
        const uint4 *s4 = (const uint4 *)pg.values;
        uint4 *d4 = (uint4 *)dense;
        for (int64_t i = tid; i < vec; i += SCAN_THREADS) d4[i] = s4[i];
        for (int64_t i = (vec << 4) + tid; i < bytes; i += SCAN_THREADS) ((uint8_t *)dense)[i] = pg.values[i];
      } else {
        for (int64_t i = tid; i < avail; i += SCAN_THREADS) {
          uint64_t v = load_le(pg.values + i * pw, pw);
          if (pw == 4 && ow == 8) v = (uint64_t)(int64_t)(int32_t)v;   // decimal(p <= 9) stored as INT32
          store_value(dense, ow, i, v);
        }
      }
    }
  } else if (pg.encoding == SB_ENC_RLE_BOOLEAN) {   // BOOLEAN values as RLE (data page V2 writers): [4-byte length][hybrid, bit width 1]
    if (pg.values_bytes > 4)
      hybrid_decode(pg.values + 4, pg.values_bytes - 4, 1, n, 0, 1, [&](int64_t i, uint64_t v) { ((uint8_t *)dense)[i] = (uint8_t)(v & 1); });
  } else {   // RLE_DICTIONARY: [bit width][hybrid runs of dictionary indices]
    const int bw = pg.values_bytes > 0 ? pg.values[0] : 0;
    const uint8_t *dict = c.dict;
    const int pw = c.phys_width, dcount = c.dict_count;
    hybrid_decode(pg.values + 1, pg.values_bytes - 1, bw, n, part, nparts, [&](int64_t i, uint64_t idx) {
      uint64_t v = 0;
      if ((int64_t)idx < dcount) v = pw == 0 ? dict[idx] : load_le(dict + idx * pw, pw);
      if (pw == 4 && ow == 8) v = (uint64_t)(int64_t)(int32_t)v;
      store_value(dense, ow, i, v);
    });
  }
  if (!page_nullable) return;
  __syncthreads();
  // ---- (3) expand value space -> row space: row r takes dense[#valid rows before r] --------------------------------------
  __shared__ int s_warp[SCAN_THREADS / 32];
  __shared__ int64_t s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  constexpr int PER = 4;
  for (int64_t tile = 0; tile < n; tile += SCAN_THREADS * PER) {
    const int64_t r0 = tile + (int64_t)tid * PER;
    int cnt = 0;
    uint8_t v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      v[k] = r0 + k < n ? vb[r0 + k] : 0;
      cnt += v[k];
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (tid >> 5); w++) wbase += s_warp[w];
    int total = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) total += s_warp[w];
    int64_t idx = s_base + wbase + incl - cnt;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      if (r0 + k >= n) break;
      uint64_t val = 0;
      if (v[k]) {
        switch (ow) {
          case 1: val = ((const uint8_t *)dense)[idx]; break;
          case 2: val = ((const uint16_t *)dense)[idx]; break;
          case 4: val = ((const uint32_t *)dense)[idx]; break;
          default: val = ((const uint64_t *)dense)[idx]; break;
        }
        idx++;
      }
      store_value(dst_page, ow, r0 + k, val);
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
}

// ---- write side (tests, bench.py): dictionary indices of a sorted dictionary + bit packing into hybrid runs --------------
// One block per page of page_rows rows.  A page is [bit width byte] + bit-packed runs of 504 values (header 0x7F, the run
// length parquet-mr's RunLengthBitPackingHybridEncoder emits: 63 groups of 8) + one shorter last run, padded to 8 values.
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) encode_dict_pages_kernel(const T *__restrict__ col, int64_t n, const T *__restrict__ dict, int32_t dcount,
                                                                         int bw, int64_t page_rows, int64_t page_stride, uint8_t *__restrict__ out) {
  const int64_t page = blockIdx.x;
  const int64_t r_lo = page * page_rows, r_hi = r_lo + page_rows < n ? r_lo + page_rows : n;
  uint8_t *dst = out + page * page_stride;
  if (threadIdx.x == 0) dst[0] = (uint8_t)bw;
  const int64_t rows = r_hi - r_lo;
  const int64_t full_runs = rows / 504, tail = rows - full_runs * 504;
  const int64_t run_bytes = 1 + 63 * (int64_t)bw;
  // one thread per group of 8 values: writes bw bytes
  const int64_t groups = (rows + 7) / 8;
  for (int64_t g = threadIdx.x; g < groups; g += SCAN_THREADS) {
    const int64_t run = g / 63, gin = g - run * 63;
    uint8_t *rp = dst + 1 + run * run_bytes;
    if (gin == 0) {
      const int64_t groups_in_run = run < full_runs ? 63 : (tail + 7) / 8;
      rp[0] = (uint8_t)((groups_in_run << 1) | 1);
    }
    uint8_t *gp = rp + 1 + gin * bw;
    uint64_t acc = 0;
    int nbits = 0, ob = 0;
    for (int k = 0; k < 8; k++) {
      const int64_t r = r_lo + g * 8 + k;
      uint64_t idx = 0;
      if (r < r_hi) {
        const T v = col[r];
        int lo = 0, hi = dcount;   // first dict entry >= v (dict is sorted, v is present)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (dict[mid] < v) lo = mid + 1; else hi = mid;
        }
        idx = (uint64_t)lo;
      }
      acc |= idx << nbits;           // nbits < 8 here and bw <= 24: fits 64 bits
      nbits += bw;
      while (nbits >= 8) {
        gp[ob++] = (uint8_t)acc;
        acc >>= 8;
        nbits -= 8;
      }
    }
  }
}

}  // namespace sb

using namespace sb;

// ---- host: Thrift compact protocol reader for Parquet PageHeader (parquet-format/src/main/thrift/parquet.thrift) ------------
namespace {
struct Thrift {
  const uint8_t *p, *end;
  bool ok = true;
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end) {
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift > 63) break;
    }
    ok = false;
    return 0;
  }
  int64_t zigzag() {
    const uint64_t u = varint();
    return (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
  }
  void skip(int type);
  void skip_struct() {
    int16_t last = 0;
    while (ok && p < end) {
      const uint8_t b = *p++;
      if (b == 0) return;
      const int type = b & 0x0f, delta = b >> 4;
      if (delta) last = (int16_t)(last + delta); else last = (int16_t)zigzag();
      skip(type);
    }
    ok = false;
  }
};
void Thrift::skip(int type) {
  switch (type) {
    case 1: case 2: break;                                  // BOOLEAN_TRUE / FALSE (value in the type nibble)
    case 3: p += 1; break;                                  // BYTE
    case 4: case 5: case 6: varint(); break;                // I16 / I32 / I64 (zigzag varints)
    case 7: p += 8; break;                                  // DOUBLE
    case 8: { const uint64_t n = varint(); p += n; break; } // BINARY
    case 9: case 10: {                                      // LIST / SET
      if (p >= end) { ok = false; return; }
      const uint8_t h = *p++;
      uint64_t n = h >> 4;
      const int et = h & 0x0f;
      if (n == 15) n = varint();
      for (uint64_t i = 0; i < n && ok; i++) {
        if (et == 1 || et == 2) p += 1; else skip(et);      // booleans inside lists take one byte each
      }
      break;
    }
    case 11: {                                              // MAP
      const uint64_t n = varint();
      if (n) {
        if (p >= end) { ok = false; return; }
        const uint8_t kv = *p++;
        for (uint64_t i = 0; i < n && ok; i++) { skip(kv >> 4); skip(kv & 0x0f); }
      }
      break;
    }
    case 12: skip_struct(); break;
    default: ok = false;
  }
  if (p > end) ok = false;
}

struct PageHeader {
  int32_t type = -1, uncompressed = 0, compressed = 0;
  int32_t num_values = 0, encoding = -1, def_enc = -1, rep_enc = -1;     // data page v1 / v2 / dictionary
  int32_t num_nulls = 0, num_rows = 0, def_len = 0, rep_len = 0;
  bool v2_compressed = true, is_v2 = false;
};
// reads the fields of one nested header struct we care about
void read_inner(Thrift &t, PageHeader &h, int which) {   // which: 5 data v1, 7 dictionary, 8 data v2
  int16_t last = 0;
  while (t.ok && t.p < t.end) {
    const uint8_t b = *t.p++;
    if (b == 0) return;
    const int type = b & 0x0f, delta = b >> 4;
    if (delta) last = (int16_t)(last + delta); else last = (int16_t)t.zigzag();
    if (type == 5 || type == 4 || type == 6) {
      const int32_t v = (int32_t)t.zigzag();
      if (which == 5) { if (last == 1) h.num_values = v; else if (last == 2) h.encoding = v; else if (last == 3) h.def_enc = v; else if (last == 4) h.rep_enc = v; }
      else if (which == 7) { if (last == 1) h.num_values = v; else if (last == 2) h.encoding = v; }
      else { if (last == 1) h.num_values = v; else if (last == 2) h.num_nulls = v; else if (last == 3) h.num_rows = v; else if (last == 4) h.encoding = v;
             else if (last == 5) h.def_len = v; else if (last == 6) h.rep_len = v; }
    } else if ((type == 1 || type == 2) && which == 8 && last == 7) {
      h.v2_compressed = type == 1;
    } else t.skip(type);
  }
  t.ok = false;
}
bool read_page_header(Thrift &t, PageHeader &h) {
  int16_t last = 0;
  while (t.ok && t.p < t.end) {
    const uint8_t b = *t.p++;
    if (b == 0) return t.ok;
    const int type = b & 0x0f, delta = b >> 4;
    if (delta) last = (int16_t)(last + delta); else last = (int16_t)t.zigzag();
    if (type == 5 && last <= 3) {
      const int32_t v = (int32_t)t.zigzag();
      if (last == 1) h.type = v; else if (last == 2) h.uncompressed = v; else h.compressed = v;
    } else if (type == 12 && (last == 5 || last == 7 || last == 8)) {
      if (last == 8) h.is_v2 = true;
      read_inner(t, h, last);
    } else t.skip(type);
  }
  return false;
}
}  // namespace

// Host only (no device needed): walks the pages of one column chunk as it lies in a Parquet file (the byte range
// [dictionary_page_offset or data_page_offset, + total_compressed_size) of the column's metadata) and fills the descriptors
// sb_scan_decode takes.  max_def_level: 0 for a required column, 1 for an optional one.
extern "C" int sb_parquet_chunk_pages(const uint8_t *chunk, int64_t nbytes, int32_t max_def_level, sb_page *out_pages, int32_t pages_cap,
                                      int32_t *out_npages, int64_t *out_dict_offset, int32_t *out_dict_count) {
  SB_API_BEGIN
  SB_REQUIRE(chunk && out_pages && out_npages && out_dict_offset && out_dict_count && nbytes >= 0, "null argument");
  SB_REQUIRE(max_def_level == 0 || max_def_level == 1, "nested schemas are not supported (max definition level %d)", max_def_level);
  int32_t np = 0;
  *out_dict_offset = -1;
  *out_dict_count = 0;
  int64_t pos = 0;
  while (pos < nbytes) {
    Thrift t{chunk + pos, chunk + nbytes};
    PageHeader h;
    if (!read_page_header(t, h)) fail(SB_ERR_INVALID, "malformed Parquet page header at byte %lld of the column chunk", (long long)pos);
    const int64_t body = t.p - chunk;
    if (h.compressed != h.uncompressed && !(h.is_v2 && !h.v2_compressed))
      fail(SB_ERR_UNSUPPORTED, "compressed Parquet pages are not decoded on this path (page at byte %lld: %d -> %d bytes)", (long long)pos,
           h.compressed, h.uncompressed);
    SB_REQUIRE(body + h.compressed <= nbytes, "Parquet page at byte %lld overruns the column chunk", (long long)pos);
    if (h.type == 2) {   // DICTIONARY_PAGE
      SB_REQUIRE(h.encoding == 0 || h.encoding == 2, "dictionary page encoding %d is not PLAIN", h.encoding);
      *out_dict_offset = body;
      *out_dict_count = h.num_values;
    } else if (h.type == 0 || h.type == 3) {   // DATA_PAGE / DATA_PAGE_V2
      SB_REQUIRE(np < pages_cap, "more than %d pages in the column chunk", pages_cap);
      sb_page &pg = out_pages[np++];
      memset(&pg, 0, sizeof(pg));
      if (h.encoding == 0) pg.encoding = SB_ENC_PLAIN;
      else if (h.encoding == 2 || h.encoding == 8) pg.encoding = SB_ENC_RLE_DICTIONARY;
      else if (h.encoding == 3) pg.encoding = SB_ENC_RLE_BOOLEAN;   // only BOOLEAN columns may use it (checked in sb_scan_decode)
      else fail(SB_ERR_UNSUPPORTED, "Parquet value encoding %d is not supported (PLAIN and RLE_DICTIONARY are)", h.encoding);
      pg.num_values = h.num_values;
      int64_t vpos = body, vend = body + h.compressed;
      if (h.type == 0) {
        if (max_def_level > 0) {   // [4-byte length][RLE definition levels]
          SB_REQUIRE(h.def_enc == 3, "definition levels must be RLE encoded (got %d)", h.def_enc);
          SB_REQUIRE(vpos + 4 <= vend, "truncated definition levels");
          const uint32_t len = (uint32_t)chunk[vpos] | ((uint32_t)chunk[vpos + 1] << 8) | ((uint32_t)chunk[vpos + 2] << 16) | ((uint32_t)chunk[vpos + 3] << 24);
          pg.def_offset = vpos + 4;
          pg.def_bytes = len;
          vpos += 4 + (int64_t)len;
        }
      } else {
        SB_REQUIRE(h.rep_len == 0, "repeated fields are not supported");
        if (h.def_len > 0) {
          pg.def_offset = vpos;
          pg.def_bytes = h.def_len;
        }
        vpos += h.def_len;
      }
      SB_REQUIRE(vpos <= vend, "Parquet page levels overrun the page");
      pg.values_offset = vpos;
      pg.values_bytes = vend - vpos;
    }   // INDEX_PAGE and unknown page types are skipped
    pos = body + h.compressed;
  }
  *out_npages = np;
  SB_API_END
}

static int phys_width_of(int32_t physical) {
  switch (physical) {
    case SB_PHYS_BOOLEAN: return 0;
    case SB_PHYS_INT32: case SB_PHYS_FLOAT: return 4;
    case SB_PHYS_INT64: case SB_PHYS_DOUBLE: return 8;
  }
  fail(SB_ERR_UNSUPPORTED, "Parquet physical type %d is not supported by the GPU scan (BOOLEAN, INT32, INT64, FLOAT, DOUBLE are)", physical);
}

extern "C" int sb_scan_decode(const sb_column_chunk *chunks, int32_t ncols, sb_stream *s, sb_table **out) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(chunks && out && ncols > 0, "null argument");
  cudaStream_t st = stream_of(s);
  int64_t nrows = -1;
  // the descriptor arrays are copied to the device asynchronously: they live on the heap until the stream has consumed them
  // (cudaLaunchHostFunc below), so the call returns without a host synchronisation and the next row group's copy can be queued
  // on another stream while this one decodes
  struct Keep {
    std::vector<PageTask> tasks;
    std::vector<ColTask> ctasks;
  };
  std::unique_ptr<Keep> keep(new Keep());
  std::vector<PageTask> &tasks = keep->tasks;
  std::vector<ColTask> &ctasks = keep->ctasks;
  ctasks.resize(ncols);
  std::vector<std::unique_ptr<Scratch>> temps;
  sb_table *t = nullptr;
  try {
    for (int ci = 0; ci < ncols; ci++) {
      const sb_column_chunk &ch = chunks[ci];
      SB_REQUIRE(ch.data && ch.pages && ch.npages >= 0, "column chunk %d: null buffers", ci);
      const int pw = phys_width_of(ch.physical_type);
      const int ow = type_width(ch.type);
      SB_REQUIRE(ow > 0, "column chunk %d: string columns are not supported by the GPU scan", ci);
      SB_REQUIRE(pw == 0 ? ch.type == SB_BOOL : (pw == 8 ? ow == 8 : (ow <= 4 || ch.type == SB_DECIMAL64)), "column chunk %d: physical type %d cannot feed column type %d", ci,
                 ch.physical_type, ch.type);
      int64_t rows = 0;
      bool nullable = false;
      for (int p = 0; p < ch.npages; p++) {
        rows += ch.pages[p].num_values;
        nullable |= ch.pages[p].def_bytes > 0;
      }
      if (nrows < 0) {
        nrows = rows;
        t = table_new(nrows);
      }
      SB_REQUIRE(rows == nrows, "column chunk %d has %lld rows, chunk 0 has %lld", ci, (long long)rows, (long long)nrows);
      // the encoded bytes cross PCIe once, as they lie in the file
      Scratch *dev = new Scratch(ch.data_bytes + 32, st);
      temps.emplace_back(dev);
      if (ch.data_bytes > 0) SB_CUDA(cudaMemcpyAsync(dev->ptr, ch.data, (size_t)ch.data_bytes, cudaMemcpyHostToDevice, st));
      Column col = column_alloc(ch.type, ch.scale, nrows, nullable, st);
      t->cols.push_back(col);
      ColTask &ct = ctasks[ci];
      ct.out = col.data->ptr;
      ct.valid_bytes = nullptr;
      ct.dense = nullptr;
      if (nullable) {
        Scratch *vb = new Scratch(nrows + 16, st), *dn = new Scratch(nrows * ow + 16, st);
        temps.emplace_back(vb);
        temps.emplace_back(dn);
        ct.valid_bytes = vb->as<uint8_t>();
        ct.dense = dn->as<uint8_t>();
      }
      ct.dict = nullptr;
      ct.dict_count = 0;
      if (ch.dict_offset >= 0 && ch.dict_count >= 0) {   // an all-NULL chunk has an empty dictionary page
        SB_REQUIRE(ch.dict_offset + (int64_t)ch.dict_count * (pw ? pw : 1) <= ch.data_bytes, "column chunk %d: dictionary overruns the chunk", ci);
        ct.dict = dev->as<uint8_t>() + ch.dict_offset;
        ct.dict_count = ch.dict_count;
      }
      ct.phys_width = pw;
      ct.out_width = ow;
      int64_t row = 0;
      for (int p = 0; p < ch.npages; p++) {
        const sb_page &pg = ch.pages[p];
        SB_REQUIRE(pg.values_offset >= 0 && pg.values_offset + pg.values_bytes <= ch.data_bytes && pg.def_offset + pg.def_bytes <= ch.data_bytes,
                   "column chunk %d page %d: offsets outside the chunk", ci, p);
        SB_REQUIRE(pg.encoding == SB_ENC_PLAIN || (pg.encoding == SB_ENC_RLE_DICTIONARY && ct.dict) || (pg.encoding == SB_ENC_RLE_BOOLEAN && pw == 0),
                   "column chunk %d page %d: encoding %d%s", ci, p, pg.encoding,
                   pg.encoding == SB_ENC_RLE_DICTIONARY ? " without a dictionary page" : " is not supported for this column");
        PageTask pt;
        pt.values = dev->as<uint8_t>() + pg.values_offset;
        pt.values_bytes = pg.values_bytes;
        pt.def = pg.def_bytes > 0 ? dev->as<uint8_t>() + pg.def_offset : nullptr;
        pt.def_bytes = pg.def_bytes;
        pt.row_start = row;
        pt.num_values = pg.num_values;
        pt.encoding = pg.encoding;
        pt.col = ci;
        pt.pad = 0;
        tasks.push_back(pt);
        row += pg.num_values;
      }
    }
    if (!tasks.empty() && nrows > 0) {
      Scratch d_pages((int64_t)tasks.size() * sizeof(PageTask), st), d_cols((int64_t)ncols * sizeof(ColTask), st);
      SB_CUDA(cudaMemcpyAsync(d_pages.ptr, tasks.data(), tasks.size() * sizeof(PageTask), cudaMemcpyHostToDevice, st));
      SB_CUDA(cudaMemcpyAsync(d_cols.ptr, ctasks.data(), (size_t)ncols * sizeof(ColTask), cudaMemcpyHostToDevice, st));
      {
        KernelTimer kt("scan_decode", st);
        // enough blocks to fill the machine: pages of hybrid-encoded values are striped over `nparts` blocks each
        int nparts = (int)((rt().num_sms * 6 + (int64_t)tasks.size() - 1) / (int64_t)tasks.size());
        nparts = nparts < 1 ? 1 : (nparts > 16 ? 16 : nparts);
        scan_decode_kernel<<<(unsigned)(tasks.size() * nparts), SCAN_THREADS, 0, st>>>(d_pages.as<PageTask>(), d_cols.as<ColTask>(), nparts);
        SB_LAUNCH_CHECK();
      }
      for (int ci = 0; ci < ncols; ci++)
        if (ctasks[ci].valid_bytes) bytes_to_bitmap(ctasks[ci].valid_bytes, nrows, (uint32_t *)t->cols[ci].validity->ptr, st);
      Keep *raw = keep.release();
      if (cudaLaunchHostFunc(st, [](void *p) { delete (Keep *)p; }, raw) != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(st);
        delete raw;
      }
    }
  } catch (...) {
    if (t) table_free(t);
    throw;
  }
  *out = t;
  SB_API_END
}

// Write-side twin for tests and bench.py: encodes one fixed-width column of `t` as a column chunk in the layout
// sb_scan_decode reads.  dictionary == nullptr: PLAIN pages.  Otherwise `dictionary` is a one-column table holding the
// column's distinct values in ascending order (e.g. sb_hash_aggregate with no aggregates + sb_sort): the chunk is
// [PLAIN dictionary values][pages of bit-packed indices].  The chunk stays on the device (out_chunk: one SB_INT8 column of
// bytes); page descriptors come back on the host.  No NULLs (definition levels are not written).
extern "C" int sb_scan_encode(const sb_table *t, int32_t col, const sb_table *dictionary, int64_t page_rows, sb_stream *s, sb_table **out_chunk,
                              sb_page *out_pages, int32_t pages_cap, int32_t *out_npages, int64_t *out_dict_offset, int32_t *out_dict_count) {
  SB_API_BEGIN
  require_init();
  SB_REQUIRE(t && out_chunk && out_pages && out_npages && out_dict_offset && out_dict_count, "null argument");
  SB_REQUIRE(col >= 0 && col < (int)t->cols.size(), "column %d out of range", col);
  const Column &c = t->cols[col];
  SB_REQUIRE(c.type != SB_STRING && c.type != SB_BOOL && !c.validity, "sb_scan_encode takes NULL-free fixed-width numeric columns");
  const int ow = type_width(c.type), pw = ow == 8 ? 8 : 4;
  SB_REQUIRE(page_rows > 0 && page_rows % 8 == 0, "page_rows must be a positive multiple of 8");
  cudaStream_t st = stream_of(s);
  const int64_t n = t->nrows;
  const int64_t npages = n == 0 ? 0 : (n + page_rows - 1) / page_rows;
  SB_REQUIRE(npages <= pages_cap, "%lld pages do not fit the descriptor array (%d)", (long long)npages, pages_cap);
  sb_table *chunk = nullptr;
  if (!dictionary) {
    SB_REQUIRE(pw == ow, "PLAIN pages of %d-byte columns need a widening writer (dictionary-encode them)", ow);
    chunk = table_new(n * ow);
    Column bytes = column_alloc(SB_INT8, 0, n * ow, false, st);
    chunk->cols.push_back(bytes);
    if (n > 0) SB_CUDA(cudaMemcpyAsync(bytes.data->ptr, c.d(), (size_t)(n * ow), cudaMemcpyDeviceToDevice, st));
    for (int64_t p = 0; p < npages; p++) {
      sb_page &pg = out_pages[p];
      memset(&pg, 0, sizeof(pg));
      const int64_t lo = p * page_rows, hi = std::min(n, lo + page_rows);
      pg.encoding = SB_ENC_PLAIN;
      pg.num_values = (int32_t)(hi - lo);
      pg.values_offset = lo * ow;
      pg.values_bytes = (hi - lo) * ow;
    }
    *out_dict_offset = -1;
    *out_dict_count = 0;
  } else {
    SB_REQUIRE(dictionary->cols.size() == 1 && dictionary->cols[0].type == c.type && !dictionary->cols[0].validity, "dictionary must be one NULL-free column of the column's type");
    const int64_t dcount = dictionary->nrows;
    SB_REQUIRE(dcount >= 1 && dcount <= (1 << 24), "dictionary of %lld entries", (long long)dcount);
    int bw = 0;
    while ((1ll << bw) < dcount) bw++;
    if (bw == 0) bw = 1;   // keep one bit per value so every page has real runs
    const int64_t dict_bytes = (dcount * pw + 15) / 16 * 16;
    const int64_t full = page_rows / 504, tail = page_rows - full * 504;
    const int64_t page_stride = (1 + full * (1 + 63 * (int64_t)bw) + (tail ? 1 + ((tail + 7) / 8) * bw : 0) + 15) / 16 * 16;
    const int64_t total = dict_bytes + npages * page_stride;
    chunk = table_new(total);
    Column bytes = column_alloc(SB_INT8, 0, total, false, st);
    chunk->cols.push_back(bytes);
    uint8_t *base = (uint8_t *)bytes.data->ptr;
    SB_CUDA(cudaMemsetAsync(base, 0, (size_t)total, st));
    // dictionary page payload: PLAIN values at the physical width (INT32 for 1/2/4-byte columns)
    if (pw == ow) SB_CUDA(cudaMemcpyAsync(base, dictionary->cols[0].d(), (size_t)(dcount * ow), cudaMemcpyDeviceToDevice, st));
    else {
      std::vector<uint8_t> hv((size_t)(dcount * ow));
      SB_CUDA(cudaMemcpyAsync(hv.data(), dictionary->cols[0].d(), hv.size(), cudaMemcpyDeviceToHost, st));
      SB_CUDA(cudaStreamSynchronize(st));
      std::vector<int32_t> wide((size_t)dcount);
      for (int64_t i = 0; i < dcount; i++) wide[i] = ow == 1 ? (int32_t)((int8_t *)hv.data())[i] : (int32_t)((int16_t *)hv.data())[i];
      SB_CUDA(cudaMemcpyAsync(base, wide.data(), (size_t)dcount * 4, cudaMemcpyHostToDevice, st));
      SB_CUDA(cudaStreamSynchronize(st));
    }
    if (npages > 0) {
      uint8_t *pages_base = base + dict_bytes;
      const void *dv = dictionary->cols[0].d();
      switch (c.type) {
        case SB_INT8: encode_dict_pages_kernel<int8_t><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const int8_t *)c.d(), n, (const int8_t *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
        case SB_INT16: encode_dict_pages_kernel<int16_t><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const int16_t *)c.d(), n, (const int16_t *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
        case SB_INT32: case SB_DATE32: encode_dict_pages_kernel<int32_t><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const int32_t *)c.d(), n, (const int32_t *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
        case SB_FLOAT32: encode_dict_pages_kernel<float><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const float *)c.d(), n, (const float *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
        case SB_FLOAT64: encode_dict_pages_kernel<double><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const double *)c.d(), n, (const double *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
        default: encode_dict_pages_kernel<int64_t><<<(unsigned)npages, SCAN_THREADS, 0, st>>>((const int64_t *)c.d(), n, (const int64_t *)dv, (int32_t)dcount, bw, page_rows, page_stride, pages_base); break;
      }
      SB_LAUNCH_CHECK();
    }
    for (int64_t p = 0; p < npages; p++) {
      sb_page &pg = out_pages[p];
      memset(&pg, 0, sizeof(pg));
      const int64_t lo = p * page_rows, hi = std::min(n, lo + page_rows), rows = hi - lo;
      const int64_t fr = rows / 504, tl = rows - fr * 504;
      pg.encoding = SB_ENC_RLE_DICTIONARY;
      pg.num_values = (int32_t)rows;
      pg.values_offset = dict_bytes + p * page_stride;
      pg.values_bytes = 1 + fr * (1 + 63 * (int64_t)bw) + (tl ? 1 + ((tl + 7) / 8) * bw : 0);
    }
    *out_dict_offset = 0;
    *out_dict_count = (int32_t)dcount;
  }
  *out_npages = (int32_t)npages;
  *out_chunk = chunk;
  SB_API_END
}
