// device_helpers.cuh -- device-only helpers shared by every kernel: TMA / mbarrier wrappers, Murmur3, canonical float
// bits.  No host headers: this file is also compiled at run time by NVRTC (csrc/rtc.cu) together with agg_kernels.cuh.
#pragma once
#ifdef __CUDACC_RTC__
#include "spark_b200.h"
#else
#include <stdint.h>
#include "../../include/spark_b200.h"
#endif

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
