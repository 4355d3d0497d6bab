// Shared helpers for libonerf_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/onerf.h"

struct onerf_ctx {
  int device;
  int num_sms;
  int64_t launches;
  void* pack_tables;   // pack.cu: per-layout job tables in device memory (created on first use)
};
void onerf_free_pack_tables(onerf_ctx* ctx);

void onerf_set_error(const char* fmt, ...);

#define ONERF_CHECK_ARG(cond, msg)                       \
  do {                                                   \
    if (!(cond)) {                                       \
      onerf_set_error("%s: %s", __func__, msg);          \
      return ONERF_ERR_BAD_ARG;                          \
    }                                                    \
  } while (0)

#define ONERF_UNSUPPORTED(cond, msg)                     \
  do {                                                   \
    if (cond) {                                          \
      onerf_set_error("%s: unsupported: %s", __func__, msg); \
      return ONERF_ERR_UNSUPPORTED;                      \
    }                                                    \
  } while (0)

#define ONERF_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      onerf_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
      return ONERF_ERR_CUDA;                                                      \
    }                                                                             \
  } while (0)

#define ONERF_LAUNCH_CHECK(ctx)                                                   \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      onerf_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return ONERF_ERR_CUDA;                                                      \
    }                                                                             \
    (ctx)->launches++;                                                            \
  } while (0)

static inline bool onerf_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (used only when the caller does not inject its own random buffers)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// uniform in [0,1) with 24 random bits (same granularity as torch.rand for fp32)
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }

// one U[0,1) for element `idx` of random stream `stream_id`
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream_id, uint64_t idx) {
  uint4 r = philox4x32(make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), stream_id, 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return u01(v);
}

// one N(0,1) for element idx (Box-Muller on two uniforms of the same counter)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream_id, uint64_t idx) {
  uint4 r = philox4x32(make_uint4((uint32_t)(idx >> 1), (uint32_t)(idx >> 33), stream_id, 1u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t a = (idx & 1) ? r.z : r.x, b = (idx & 1) ? r.w : r.y;
  float u1 = ((a >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = u01(b);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
