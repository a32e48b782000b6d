// Leave-one-out ranking evaluation for sm_100a: exact fp32 paths.
//
// Replaces _eval_by_user (utils.py:244-261), evaluation.py:54-76,114-135 and Recommender.rank.  Every score is the
// chain acc = fmaf(p[t], q[t], acc), t ascending, acc0 = 0 (oracle.score_pairs), so rank positions and top-K ids are
// bit-exact against the oracle.  The full-rank kernel is a register-tiled fp32 GEMM on the CUDA cores whose k loop
// preserves that order per accumulator; its epilogue counts score >= score(held-out) and streams candidates that beat
// a running per-user threshold into a small top-K buffer, so the [users, items] score matrix is never written.
// (The tcgen05 bf16x3 filter in eval_tc.cu produces the same outputs by re-scoring its ambiguous candidates here.)
#include <math_constants.h>

#include "common.cuh"

namespace apr {

__device__ __forceinline__ float score_chain(const float* __restrict__ p, const float* __restrict__ q, int d) {
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* q4 = reinterpret_cast<const float4*>(q);
  float acc = 0.f;
  for (int e = 0; e < d / 4; ++e) {
    const float4 a = __ldg(p4 + e), b = __ldg(q4 + e);
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256)
score_pairs_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ users,
                   const int32_t* __restrict__ items, int64_t n, float* __restrict__ scores) {
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += int64_t(gridDim.x) * blockDim.x)
    scores[t] = score_chain(P + int64_t(users[t]) * d, Q + int64_t(items[t]) * d, d);
}

// N4: the all-item scorer `all_rating = u . Q^T` of IRGAN.py:36-39 / APL.py:205-211 for a batch of users:
// out[k, c - item_lo] = score(users[k], c) with the pinned fma chain.  One CTA per (user, 1024-item slab); the user row
// sits in shared memory, every thread walks the rows of its items with 128-bit loads.  Dense output: meant for the
// per-user `rank` of the Recommender plug-ins (a few users at a time), not for full evaluation (that never writes scores).
__global__ void __launch_bounds__(256)
score_all_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ users, int item_lo,
                 int item_hi, float* __restrict__ out) {
  extern __shared__ float sp[];
  const int k = blockIdx.y;
  for (int e = threadIdx.x; e < d; e += blockDim.x) sp[e] = P[int64_t(users[k]) * d + e];
  __syncthreads();
  const int n_items = item_hi - item_lo;
  for (int c = blockIdx.x * 1024 + threadIdx.x; c < min(n_items, (int(blockIdx.x) + 1) * 1024); c += blockDim.x) {
    const float4* q4 = reinterpret_cast<const float4*>(Q + int64_t(item_lo + c) * d);
    float acc = 0.f;
    for (int e = 0; e < d / 4; ++e) {
      const float4 b = __ldg(q4 + e);
      acc = fmaf(sp[4 * e], b.x, acc); acc = fmaf(sp[4 * e + 1], b.y, acc);
      acc = fmaf(sp[4 * e + 2], b.z, acc); acc = fmaf(sp[4 * e + 3], b.w, acc);
    }
    out[int64_t(k) * n_items + c] = acc;
  }
}

// one CTA per user; candidate rows are read by 128-bit loads, the user row is staged in shared memory
__global__ void __launch_bounds__(128)
eval_candidates_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ users,
                       const int64_t* __restrict__ cand_ptr, const int32_t* __restrict__ cand_idx,
                       int32_t* __restrict__ position, float* __restrict__ scores) {
  extern __shared__ float sp[];  // d floats
  __shared__ float s_pos;
  __shared__ int s_cnt;
  const int k = blockIdx.x;
  const int user = users[k];
  const int64_t c0 = cand_ptr[k], c1 = cand_ptr[k + 1];
  for (int e = threadIdx.x; e < d; e += blockDim.x) sp[e] = P[int64_t(user) * d + e];
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (c1 <= c0) { if (threadIdx.x == 0) position[k] = 0; return; }
  auto score = [&](int item) {
    const float4* q4 = reinterpret_cast<const float4*>(Q + int64_t(item) * d);
    float acc = 0.f;
    for (int e = 0; e < d / 4; ++e) {
      const float4 b = __ldg(q4 + e);
      acc = fmaf(sp[4 * e], b.x, acc); acc = fmaf(sp[4 * e + 1], b.y, acc);
      acc = fmaf(sp[4 * e + 2], b.z, acc); acc = fmaf(sp[4 * e + 3], b.w, acc);
    }
    return acc;
  };
  if (threadIdx.x == 0) {
    s_pos = score(cand_idx[c1 - 1]);
    if (scores) scores[c1 - 1] = s_pos;
  }
  __syncthreads();
  const float pos_score = s_pos;
  int cnt = 0;
  for (int64_t c = c0 + threadIdx.x; c < c1 - 1; c += blockDim.x) {
    const float s = score(cand_idx[c]);
    if (scores) scores[c] = s;
    cnt += (s >= pos_score);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) position[k] = s_cnt;
}

// ------------------------------------------------------------------------------------------------
// full-rank: exact fp32 tile kernel
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, KC = 16, BNP = BN + 4;
constexpr int TOPK_CAP = 256;  // per-user candidate buffer; supports k_top <= 128

__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int32_t key) {
  int64_t a = lo, b = hi;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (idx[m] < key) a = m + 1; else b = m;
  }
  return a < hi && idx[a] == key;
}

template <bool TOPK>
__global__ void __launch_bounds__(256)
fullrank_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ users,
                int n_users, const float* __restrict__ spos, int item_lo, int item_hi, int items_per_cta,
                const int64_t* __restrict__ excl_ptr, const int32_t* __restrict__ excl_idx, int k_top,
                int32_t* __restrict__ position, float* __restrict__ buf_score, int32_t* __restrict__ buf_id,
                int32_t* __restrict__ topk_ids, float* __restrict__ topk_scores) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                 // [d][BM]
  float* Bs = As + size_t(d) * BM;  // [KC][BNP]
  float* sp = Bs + KC * BNP;        // [BM] score of the held-out item
  float* thr = sp + BM;             // [BM] running top-K threshold
  int* n_u = reinterpret_cast<int*>(thr + BM);            // [BM]
  float* sc_s = reinterpret_cast<float*>(n_u + BM);       // [8][TOPK_CAP]   (TOPK only)
  int* sc_i = reinterpret_cast<int*>(sc_s + 8 * TOPK_CAP);  // [8][TOPK_CAP]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  for (int f = tid; f < BM * (d / 4); f += 256) {
    const int m = f % BM, kq = f / BM;
    float4 v = f4_zero();
    if (m0 + m < n_users) v = __ldg(reinterpret_cast<const float4*>(P + int64_t(users[m0 + m]) * d) + kq);
    As[(kq * 4 + 0) * BM + m] = v.x; As[(kq * 4 + 1) * BM + m] = v.y;
    As[(kq * 4 + 2) * BM + m] = v.z; As[(kq * 4 + 3) * BM + m] = v.w;
  }
  if (tid < BM) {
    sp[tid] = (m0 + tid < n_users) ? spos[m0 + tid] : CUDART_INF_F;
    thr[tid] = -CUDART_INF_F;
    n_u[tid] = 0;
  }
  __syncthreads();

  const int n_begin = item_lo + blockIdx.y * items_per_cta;
  const int n_end = min(item_hi, n_begin + items_per_cta);
  int cnt[4] = {0, 0, 0, 0};
  const int ld_item = tid >> 2, ld_kq = tid & 3;

  auto compact = [&](bool final_pass) {
    // warp w serves users w, w+8, ...; rank-by-counting keeps the best k_top entries sorted
    const int w = tid >> 5, lane = tid & 31;
    float* ss = sc_s + w * TOPK_CAP;
    int* si = sc_i + w * TOPK_CAP;
    for (int m = w; m < BM; m += 8) {
      const int n = n_u[m];
      if (m0 + m >= n_users) continue;
      if (!final_pass && n <= TOPK_CAP - BN) continue;
      float* bs = buf_score + int64_t(m0 + m) * TOPK_CAP;
      int32_t* bi = buf_id + int64_t(m0 + m) * TOPK_CAP;
      for (int e = lane; e < n; e += 32) { ss[e] = bs[e]; si[e] = bi[e]; }
      __syncwarp();
      for (int e = lane; e < n; e += 32) {
        const float se = ss[e];
        const int ie = si[e];
        int rank = 0;
        for (int f = 0; f < n; ++f) {
          const float sf = ss[f];
          rank += (sf > se) || (sf == se && si[f] < ie);
        }
        if (rank < k_top) { bs[rank] = se; bi[rank] = ie; }
        if (rank == k_top - 1) thr[m] = se;
      }
      __syncwarp();
      if (lane == 0) n_u[m] = min(n, k_top);
    }
  };

  for (int n0 = n_begin; n0 < n_end; n0 += BN) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < d; k0 += KC) {
      {
        const int gi = n0 + ld_item, kk = k0 + ld_kq * 4;
        float4 v = f4_zero();
        if (gi < n_end && kk < d) v = __ldg(reinterpret_cast<const float4*>(Q + int64_t(gi) * d + kk));
        Bs[(ld_kq * 4 + 0) * BNP + ld_item] = v.x; Bs[(ld_kq * 4 + 1) * BNP + ld_item] = v.y;
        Bs[(ld_kq * 4 + 2) * BNP + ld_item] = v.z; Bs[(ld_kq * 4 + 3) * BNP + ld_item] = v.w;
      }
      __syncthreads();
      const int kmax = min(KC, d - k0);
#pragma unroll 4
      for (int k = 0; k < kmax; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[(k0 + k) * BM + ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k * BNP + tx * 4]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int m = ty * 4 + a;
      const float s_held = sp[m];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int gi = n0 + tx * 4 + b;
        if (gi < n_end) {
          const float s = acc[a][b];
          cnt[a] += (s >= s_held);
          if (TOPK) {
            if (m0 + m < n_users && s >= thr[m]) {
              const int64_t e0 = excl_ptr[m0 + m], e1 = excl_ptr[m0 + m + 1];
              if (!csr_contains(excl_idx, e0, e1, gi)) {
                const int pos = atomicAdd(&n_u[m], 1);
                buf_score[int64_t(m0 + m) * TOPK_CAP + pos] = s;
                buf_id[int64_t(m0 + m) * TOPK_CAP + pos] = gi;
              }
            }
          }
        }
      }
    }
    if (TOPK) {
      __syncthreads();
      compact(false);
      __syncthreads();
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int c = cnt[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (tx == 0 && c && m0 + ty * 4 + a < n_users) atomicAdd(&position[m0 + ty * 4 + a], c);
  }
  if (TOPK) {
    __syncthreads();
    compact(true);
    __syncthreads();
    for (int f = tid; f < BM * k_top; f += 256) {
      const int m = f / k_top, r = f - m * k_top;
      if (m0 + m < n_users) {
        const bool ok = r < n_u[m];
        topk_ids[int64_t(m0 + m) * k_top + r] = ok ? buf_id[int64_t(m0 + m) * TOPK_CAP + r] : -1;
        topk_scores[int64_t(m0 + m) * k_top + r] = ok ? buf_score[int64_t(m0 + m) * TOPK_CAP + r] : -CUDART_INF_F;
      }
    }
  }
}

// subtract the excluded items (train items and the held-out item itself) that the dense count included
__global__ void __launch_bounds__(256)
excl_correction_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ users,
                       int n_users, const float* __restrict__ spos, int item_lo, int item_hi,
                       const int64_t* __restrict__ excl_ptr, const int32_t* __restrict__ excl_idx,
                       int32_t* __restrict__ position) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int k = warp; k < n_users; k += nwarps) {
    const float s_held = spos[k];
    const float* p = P + int64_t(users[k]) * d;
    int c = 0;
    for (int64_t e = excl_ptr[k] + lane; e < excl_ptr[k + 1]; e += 32) {
      const int item = excl_idx[e];
      if (item >= item_lo && item < item_hi) c += (score_chain(p, Q + int64_t(item) * d, d) >= s_held);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && c) atomicSub(&position[k], c);
  }
}

__global__ void __launch_bounds__(256) sum_squares_kernel(const float* __restrict__ x, int64_t n, double* out) {
  double acc = 0.0;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) {
    const double v = x[e];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k];
    atomicAdd(out, t);
  }
}

struct EvalLayout { int64_t off_spos, off_buf_score, off_buf_id, total; };
static EvalLayout eval_layout(int n_users, int k_top) {
  EvalLayout L;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += (bytes + 255) & ~int64_t(255); return r; };
  L.off_spos = take(int64_t(n_users) * 4);
  L.off_buf_score = take(k_top > 0 ? int64_t(n_users) * TOPK_CAP * 4 : 0);
  L.off_buf_id = take(k_top > 0 ? int64_t(n_users) * TOPK_CAP * 4 : 0);
  L.total = o;
  return L;
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t need = (n + threads - 1) / threads;
  const int64_t cap = int64_t(sm_count()) * 16;
  return int(need < 1 ? 1 : (need < cap ? need : cap));
}

int eval_fullrank_exact(const float* P, const float* Q, int d, const int32_t* users, const int32_t* test_item, int n_users,
                        int item_lo, int item_hi, const int64_t* excl_ptr, const int32_t* excl_idx, int k_top,
                        int32_t* position, int32_t* topk_ids, float* topk_scores, void* ws, int64_t ws_bytes,
                        cudaStream_t st) {
  const EvalLayout L = eval_layout(n_users, k_top);
  if (ws_bytes < L.total) return APR_E_WORKSPACE;
  float* spos = reinterpret_cast<float*>(static_cast<char*>(ws) + L.off_spos);
  float* buf_score = reinterpret_cast<float*>(static_cast<char*>(ws) + L.off_buf_score);
  int32_t* buf_id = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + L.off_buf_id);
  score_pairs_kernel<<<grid_for(n_users, 256), 256, 0, st>>>(P, Q, d, users, test_item, n_users, spos);
  const int tiles_m = (n_users + BM - 1) / BM;
  const int n_items = item_hi - item_lo;
  size_t smem = (size_t(d) * BM + KC * BNP + 3 * BM) * 4;
  if (k_top > 0) {
    smem += size_t(8) * TOPK_CAP * 8;
    APR_CUDA_CHECK(cudaFuncSetAttribute(fullrank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    fullrank_kernel<true><<<dim3(tiles_m, 1), 256, smem, st>>>(P, Q, d, users, n_users, spos, item_lo, item_hi, n_items,
                                                              excl_ptr, excl_idx, k_top, position, buf_score, buf_id,
                                                              topk_ids, topk_scores);
  } else {
    // split the item range so that the grid has >= ~4 CTAs per SM
    const int target = sm_count() * 4;
    int splits = max(1, min((target + tiles_m - 1) / tiles_m, (n_items + BN - 1) / BN));
    int per = ((n_items + splits - 1) / splits + BN - 1) / BN * BN;
    splits = (n_items + per - 1) / per;
    APR_CUDA_CHECK(cudaFuncSetAttribute(fullrank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    fullrank_kernel<false><<<dim3(tiles_m, splits), 256, smem, st>>>(P, Q, d, users, n_users, spos, item_lo, item_hi, per,
                                                                     excl_ptr, excl_idx, 0, position, nullptr, nullptr,
                                                                     nullptr, nullptr);
  }
  APR_LAUNCH_CHECK();
  excl_correction_kernel<<<grid_for(int64_t(n_users) * 32, 256), 256, 0, st>>>(P, Q, d, users, n_users, spos, item_lo,
                                                                               item_hi, excl_ptr, excl_idx, position);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int launch_score_pairs(const float* P, const float* Q, int d, const int32_t* users, const int32_t* items, int64_t n,
                       float* scores, cudaStream_t st) {
  score_pairs_kernel<<<grid_for(n, 256), 256, 0, st>>>(P, Q, d, users, items, n, scores);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int launch_excl_correction(const float* P, const float* Q, int d, const int32_t* users, int n_users, const float* spos,
                           int item_lo, int item_hi, const int64_t* excl_ptr, const int32_t* excl_idx, int32_t* position,
                           cudaStream_t st) {
  excl_correction_kernel<<<grid_for(int64_t(n_users) * 32, 256), 256, 0, st>>>(P, Q, d, users, n_users, spos, item_lo,
                                                                               item_hi, excl_ptr, excl_idx, position);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // namespace apr

using namespace apr;

extern "C" {

int apr_score_pairs(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* items, int64_t n,
                    float* scores, apr_stream_t stream) {
  if (!P || !Q || !users || !items || !scores || n < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  score_pairs_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, Q, d, users, items, n, scores);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_score_all_items(const float* P, const float* Q, int32_t d, const int32_t* users, int32_t n_users, int32_t item_lo,
                        int32_t item_hi, float* scores, apr_stream_t stream) {
  if (!P || !Q || !users || !scores || n_users < 1 || n_users > 65535 || item_hi <= item_lo || item_lo < 0 || !valid_dim(d))
    return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  const int slabs = (item_hi - item_lo + 1023) / 1024;
  score_all_kernel<<<dim3(slabs, n_users), 256, size_t(d) * 4, static_cast<cudaStream_t>(stream)>>>(P, Q, d, users, item_lo, item_hi,
                                                                                                   scores);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_eval_candidates(const float* P, const float* Q, int32_t d, const int32_t* users, const int64_t* cand_ptr,
                        const int32_t* cand_idx, int32_t n_users, int32_t* position, float* scores, apr_stream_t stream) {
  if (!P || !Q || !users || !cand_ptr || !cand_idx || !position || n_users < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  eval_candidates_kernel<<<n_users, 128, size_t(d) * 4, static_cast<cudaStream_t>(stream)>>>(P, Q, d, users, cand_ptr,
                                                                                             cand_idx, position, scores);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int64_t apr_eval_workspace_bytes(int32_t n_users, int32_t k_top, int32_t d) {
  if (n_users < 1 || k_top < 0 || k_top > 128 || !valid_dim(d)) return -1;
  return eval_layout(n_users, k_top).total;
}

int apr_eval_fullrank(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* test_item,
                      int32_t n_users, int32_t item_lo, int32_t item_hi, const int64_t* excl_ptr, const int32_t* excl_idx,
                      int32_t k_top, int32_t* position, int32_t* topk_ids, float* topk_scores, int32_t exact, void* ws,
                      int64_t ws_bytes, apr_stream_t stream) {
  if (!P || !Q || !users || !test_item || !excl_ptr || !position || !ws) return APR_E_ARG;
  if (n_users < 1 || item_hi <= item_lo || item_lo < 0 || k_top < 0 || k_top > 128 || !valid_dim(d)) return APR_E_ARG;
  if (k_top > 0 && (!topk_ids || !topk_scores)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  if (!exact && (d % 8) == 0 && d <= 256 && (reinterpret_cast<uintptr_t>(ws) & 1023u) == 0) {
    // tensor-core route (same results): taken when the caller's workspace is large enough for it; the pipeline error
    // flag of that path (never set in a healthy run) is the int32 right behind the tensor-core layout
    const int64_t need = apr_eval_tc_topk_workspace_bytes(n_users, item_hi - item_lo, d, k_top);
    if (need > 0 && ws_bytes >= need + 1024) {
      int32_t* err = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + need);
      APR_CUDA_CHECK(cudaMemsetAsync(err, 0, 4, static_cast<cudaStream_t>(stream)));
      return apr_eval_fullrank_tc_topk(P, Q, d, users, test_item, n_users, item_lo, item_hi, excl_ptr, excl_idx, k_top, position,
                                       topk_ids, topk_scores, 0, nullptr, ws, need, err, stream);
    }
  }
  return eval_fullrank_exact(P, Q, d, users, test_item, n_users, item_lo, item_hi, excl_ptr, excl_idx, k_top, position,
                             topk_ids, topk_scores, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int apr_sum_squares(const float* x, int64_t n, double* out, apr_stream_t stream) {
  if (!x || !out || n < 1) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  APR_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double), st));
  sum_squares_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, n, out);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // extern "C"
