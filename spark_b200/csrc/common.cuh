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
    case SB_DECIMAL128: return 16;
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
// device helpers (their own header: the runtime-compiled aggregate kernels include it without the host side)
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#include "device_helpers.cuh"
#endif
