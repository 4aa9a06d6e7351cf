// common.cuh -- shared host/device plumbing for libsparkb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>
#include <mutex>
#include "../../include/spark_b200.h"

namespace sb {

// ----------------------------------------------------------------------------------------------
// errors: every extern "C" entry point catches sb::Error and maps it to a code + thread-local text
// ----------------------------------------------------------------------------------------------
struct Error {
  int code;
  std::string msg;
};
void set_last_error(const std::string &m);
[[noreturn]] void fail(int code, const char *fmt, ...);

#define SB_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      cudaGetLastError();                                                                          \
      ::sb::fail(_e == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA, "%s failed: %s (%s:%d)", \
                 #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                               \
    }                                                                                              \
  } while (0)

#define SB_API_BEGIN try {
#define SB_API_END                                                                                 \
  return SB_OK;                                                                                    \
  }                                                                                                \
  catch (const ::sb::Error &e) {                                                                   \
    ::sb::set_last_error(e.msg);                                                                   \
    return e.code;                                                                                 \
  }                                                                                                \
  catch (const std::exception &e) {                                                                \
    ::sb::set_last_error(e.what());                                                                \
    return SB_ERR_INVALID;                                                                         \
  }

#define SB_REQUIRE(cond, ...)                              \
  do {                                                     \
    if (!(cond)) ::sb::fail(SB_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// ----------------------------------------------------------------------------------------------
// runtime state
// ----------------------------------------------------------------------------------------------
struct Runtime {
  bool initialized = false;
  int device = -1;
  int num_sms = 0;
  int cc = 0;
  std::atomic<int64_t> launches{0};
};
Runtime &rt();
void require_init();

inline int type_width(int32_t t) {
  switch (t) {
    case SB_BOOL: case SB_INT8: return 1;
    case SB_INT16: return 2;
    case SB_INT32: case SB_FLOAT32: case SB_DATE32: return 4;
    case SB_INT64: case SB_FLOAT64: case SB_TIMESTAMP: case SB_DECIMAL64: return 8;
    case SB_STRING: return 0;
  }
  fail(SB_ERR_INVALID, "unknown column type %d", t);
}

inline int64_t bitmap_bytes(int64_t n) { return (n + 7) / 8; }
// device bitmaps are padded to whole 32-bit words so kernels can write them with one ballot per warp
inline int64_t bitmap_alloc_bytes(int64_t n) { return ((n + 31) / 32) * 4 + 4; }

}  // namespace sb

// ----------------------------------------------------------------------------------------------
// opaque handle types
// ----------------------------------------------------------------------------------------------
struct sb_stream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
};

namespace sb {

struct Buffer {  // a device allocation shared between tables (select/zip share buffers)
  void *ptr = nullptr;
  int64_t bytes = 0;
  bool owned = true;
  cudaStream_t st = nullptr;   // allocation stream: the buffer is freed in this stream's order (or stream 0 once it is gone)
  std::atomic<int> refs{1};
};
Buffer *buffer_alloc(int64_t bytes, cudaStream_t st);   // stream-ordered pool allocation
Buffer *buffer_borrow(const void *p);
void buffer_retain(Buffer *b);
void buffer_release(Buffer *b);

struct Column {
  int32_t type = 0;
  int32_t scale = 0;
  int64_t length = 0;
  int64_t null_count = 0;   // -1 unknown
  Buffer *data = nullptr;
  Buffer *validity = nullptr;
  Buffer *offsets = nullptr;
  int64_t string_bytes = 0;

  const void *d() const { return data ? data->ptr : nullptr; }
  const uint8_t *v() const { return validity ? (const uint8_t *)validity->ptr : nullptr; }
  const int32_t *o() const { return offsets ? (const int32_t *)offsets->ptr : nullptr; }
};

Column column_alloc(int32_t type, int32_t scale, int64_t n, bool with_validity, cudaStream_t st);
Column column_share(const Column &c);
void column_release(Column &c);

// RAII scratch allocation on a stream
struct Scratch {
  void *ptr = nullptr;
  cudaStream_t st;
  Scratch(int64_t bytes, cudaStream_t s) : st(s) {
    if (bytes > 0) SB_CUDA(cudaMallocAsync(&ptr, (size_t)bytes, s));
  }
  ~Scratch() {
    if (ptr) cudaFreeAsync(ptr, st);
  }
  template <typename T> T *as() { return (T *)ptr; }
  Scratch(const Scratch &) = delete;
  Scratch &operator=(const Scratch &) = delete;
};

inline cudaStream_t stream_of(sb_stream *s) { return s ? s->stream : (cudaStream_t)0; }

inline void count_launch(int n = 1) { rt().launches.fetch_add(n, std::memory_order_relaxed); }
#define SB_LAUNCH_CHECK()            \
  do {                               \
    ::sb::count_launch();            \
    SB_CUDA(cudaGetLastError());     \
  } while (0)

// RAII device timer around one kernel launch (active only after sb_profile_enable(1))
struct KernelTimer {
  const char *name;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(const char *name, cudaStream_t st);
  ~KernelTimer();
};

inline int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace sb

struct sb_table {
  std::vector<sb::Column> cols;
  int64_t nrows = 0;
  std::atomic<int> refs{1};
};

namespace sb {
sb_table *table_new(int64_t nrows);
void table_free(sb_table *t);
}  // namespace sb

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace sb {

// ---- TMA (cp.async.bulk) + mbarrier: the copy engine moves contiguous tiles global -> shared while the warps compute ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// same, tagged evict-first in L2: a stream that is read exactly once should not push out lines other kernels' stores are
// still completing
__device__ __forceinline__ void tma_load_1d_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t smem_add_acq_rel(uint32_t *p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool bit_valid(const uint8_t *__restrict__ bm, int64_t i) {
  return bm == nullptr || ((bm[i >> 3] >> (i & 7)) & 1);
}

// Murmur3_x86_32 (common/unsafe/src/main/java/org/apache/spark/unsafe/hash/Murmur3_x86_32.java:47-150)
__device__ __forceinline__ uint32_t mm3_mixK1(uint32_t k1) {
  k1 *= 0xcc9e2d51u;
  k1 = __funnelshift_l(k1, k1, 15);
  k1 *= 0x1b873593u;
  return k1;
}
__device__ __forceinline__ uint32_t mm3_mixH1(uint32_t h1, uint32_t k1) {
  h1 ^= k1;
  h1 = __funnelshift_l(h1, h1, 13);
  return h1 * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t mm3_fmix(uint32_t h1, uint32_t len) {
  h1 ^= len;
  h1 ^= h1 >> 16;
  h1 *= 0x85ebca6bu;
  h1 ^= h1 >> 13;
  h1 *= 0xc2b2ae35u;
  h1 ^= h1 >> 16;
  return h1;
}
__device__ __forceinline__ uint32_t mm3_int(uint32_t v, uint32_t seed) {
  return mm3_fmix(mm3_mixH1(seed, mm3_mixK1(v)), 4);
}
__device__ __forceinline__ uint32_t mm3_long(uint64_t v, uint32_t seed) {
  uint32_t h1 = mm3_mixH1(seed, mm3_mixK1((uint32_t)v));
  h1 = mm3_mixH1(h1, mm3_mixK1((uint32_t)(v >> 32)));
  return mm3_fmix(h1, 8);
}
// hashUnsafeBytes: 4-byte little-endian words, then every tail byte sign-extended as its own block
__device__ __forceinline__ uint32_t mm3_bytes(const uint8_t *__restrict__ p, int len, uint32_t seed) {
  uint32_t h1 = seed;
  int aligned = len & ~3;
  for (int i = 0; i < aligned; i += 4) {
    uint32_t w = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    h1 = mm3_mixH1(h1, mm3_mixK1(w));
  }
  for (int i = aligned; i < len; i++) h1 = mm3_mixH1(h1, mm3_mixK1((uint32_t)(int32_t)(int8_t)p[i]));
  return mm3_fmix(h1, (uint32_t)len);
}

// Double.doubleToLongBits / Float.floatToIntBits: every NaN becomes the canonical quiet NaN
__device__ __forceinline__ int64_t double_bits_canonical(double d) {
  return d != d ? 0x7ff8000000000000LL : __double_as_longlong(d);
}
__device__ __forceinline__ int32_t float_bits_canonical(float f) {
  return f != f ? 0x7fc00000 : __float_as_int(f);
}

// value of a fixed-width column widened to 64 bits the way an UnsafeRow field / hash input sees it
__device__ __forceinline__ int64_t load_i64(const void *__restrict__ data, int32_t type, int64_t i) {
  switch (type) {
    case SB_BOOL: return ((const uint8_t *)data)[i] ? 1 : 0;
    case SB_INT8: return ((const int8_t *)data)[i];
    case SB_INT16: return ((const int16_t *)data)[i];
    case SB_INT32: case SB_DATE32: return ((const int32_t *)data)[i];
    case SB_FLOAT32: return ((const int32_t *)data)[i];
    default: return ((const int64_t *)data)[i];
  }
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace sb
#endif
