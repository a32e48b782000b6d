// Shared device helpers for libapr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/apr_b200.h"

namespace apr {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);
#define APR_CUDA_CHECK(expr)                          \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) {                          \
      ::apr::set_cuda_error(_e, #expr);               \
      return APR_E_CUDA;                              \
    }                                                 \
  } while (0)
#define APR_LAUNCH_CHECK() APR_CUDA_CHECK(cudaGetLastError())

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool valid_dim(int d) { return d >= 4 && d <= 512 && (d % 4) == 0; }
int sm_count();

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 and the Feistel permutation -- bit-identical to oracle/apr_oracle.py
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kStreamPerm = 0x50455231u;
constexpr uint32_t kStreamNeg = 0x4E454731u;

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = 0xD2511F53ull * c0;
    const uint64_t p1 = 0xCD9E8D57ull * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}

// tf.truncated_normal element `e` of table `table_id` (oracle.truncated_normal): Philox counter (e_lo, e_hi, attempt,
// table_id), key (seed, tag) -> four Box-Muller normals, the first with |z| <= 2 wins, else the next attempt
__device__ inline float truncated_normal_elem(uint64_t e, uint32_t seed, uint32_t table_id, uint32_t tag) {
  const float two_pi = 6.283185307179586f;
  for (uint32_t attempt = 0;; ++attempt) {
    uint32_t w[4];
    philox4x32_10(uint32_t(e), uint32_t(e >> 32), attempt, table_id, seed, tag, w);
    float uf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) uf[k] = __ull2float_rn((unsigned long long)w[k] + 1ull) * 2.3283064365386963e-10f;
    float z[4];
    const float r0 = sqrtf(-2.0f * logf(uf[0])), r1 = sqrtf(-2.0f * logf(uf[2]));
    z[0] = r0 * cosf(two_pi * uf[1]); z[1] = r0 * sinf(two_pi * uf[1]);
    z[2] = r1 * cosf(two_pi * uf[3]); z[3] = r1 * sinf(two_pi * uf[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (fabsf(z[k]) <= 2.0f) return z[k];
  }
}
constexpr uint32_t kStreamAdv = 0x41445631u;   // --adv random noise (engine.STREAM_ADV)

struct PermKeys { uint32_t k[8]; };

__host__ __device__ inline uint32_t feistel_perm(uint32_t x, uint32_t n, int half_bits, const PermKeys& keys) {
  const uint32_t mask = (1u << half_bits) - 1u;
  do {
    uint32_t L = x >> half_bits, R = x & mask;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t F = fmix32(R ^ keys.k[r]) & mask;
      const uint32_t nl = R;
      R = L ^ F;
      L = nl;
    }
    x = (L << half_bits) | R;
  } while (x >= n);
  return x;
}

// ------------------------------------------------------------------------------------------------
// 128-bit row access.  Tables change between launches / grid barriers, so bypass L1 (.cg) for rows.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
// fire-and-forget vector reduction at L2 (sm_90+): one 16-byte RED instead of four scalar ones
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// sum over the G lanes of an aligned lane group (G power of two <= 32); `mask` names exactly those lanes
template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

}  // namespace apr
