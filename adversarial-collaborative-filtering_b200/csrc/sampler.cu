// Counter-based epoch sampler, DNS selection and table initialisation for sm_100a.
//
// Replaces shuffle/_get_train_batch (APR.py:39-81), the dns>1 branch of training_batch (utils.py:121-139) and
// MF._create_variables (APR.py:105-119).  Integer outputs are bit-identical to oracle/apr_oracle.py.
#include <algorithm>

#include "common.cuh"

namespace apr {

constexpr int kMaxNegAttempts = 1 << 16;

__global__ void __launch_bounds__(256)
sample_epoch_kernel(const int32_t* __restrict__ pairs_u, const int32_t* __restrict__ pairs_i, uint32_t n_pairs,
                    int64_t n_draws, int dns, uint32_t num_items, const int64_t* __restrict__ csr_ptr,
                    const int32_t* __restrict__ csr_idx, int csr_rows, uint32_t seed, uint32_t epoch, PermKeys keys,
                    int half_bits, int batch, int batch_lo, int batch_local, int64_t fork_chunk_draws, int fork_workers,
                    int32_t* __restrict__ out_u,
                    int32_t* __restrict__ out_i, int32_t* __restrict__ out_udns, int32_t* __restrict__ out_j,
                    int32_t* err_flag) {
  // a rank of a sharded run draws the triples [batch_lo, batch_lo + batch_local) of every batch: `fl` numbers ITS draws,
  // `f` is the draw's index in the whole epoch -- the counter of every random number -- so the shards are bit-identical
  // slices of the unsharded epoch ("counter-based RNG keyed by global triple index; no collective")
  for (int64_t fl = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; fl < n_draws; fl += int64_t(gridDim.x) * blockDim.x) {
    const int64_t tl = fl / dns;
    const int k = int(fl - tl * dns);
    const int64_t step = tl / batch_local;
    const int64_t t = step * batch + batch_lo + (tl - step * batch_local);
    int64_t f = t * dns + k;
    const uint32_t pair = feistel_perm(uint32_t(t), n_pairs, half_bits, keys);
    const int32_t user = pairs_u[pair];
    if (k == 0) { out_u[tl] = user; out_i[tl] = pairs_i[pair]; }
    out_udns[fl] = user;
    int64_t lo = 0, hi = 0;
    if (user < csr_rows) { lo = csr_ptr[user]; hi = csr_ptr[user + 1]; }
    // legacy mode (oracle.fork_counter): the W chunks of one pool.map round share one random stream
    if (fork_workers > 0) f = ((f / fork_chunk_draws) / fork_workers) * fork_chunk_draws + (f % fork_chunk_draws);
    uint32_t w[4];
    int32_t chosen = -1;
    for (int attempt = 0; attempt < kMaxNegAttempts; ++attempt) {
      if ((attempt & 3) == 0)
        philox4x32_10(uint32_t(uint64_t(f)), uint32_t(uint64_t(f) >> 32), uint32_t(attempt >> 2), epoch, seed, kStreamNeg, w);
      const int32_t cand = int32_t((uint64_t(w[attempt & 3]) * uint64_t(num_items)) >> 32);
      // binary search in the user's sorted train row: `while j in trainList[u]` (APR.py:77)
      int64_t a = lo, b = hi;
      while (a < b) {
        const int64_t m = (a + b) >> 1;
        if (csr_idx[m] < cand) a = m + 1; else b = m;
      }
      if (!(a < hi && csr_idx[a] == cand)) { chosen = cand; break; }
    }
    if (chosen < 0) { atomicOr(err_flag, 1); chosen = 0; }
    out_j[fl] = chosen;
  }
}

// first-maximum argmax over the dns sampled negatives of each positive, scores in the pinned fma order
__global__ void __launch_bounds__(256)
select_dns_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ u_dns,
                  const int32_t* __restrict__ j_dns, int64_t n_pos, int dns, int32_t* __restrict__ out_j) {
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n_pos; t += int64_t(gridDim.x) * blockDim.x) {
    float best = 0.f;
    int32_t best_j = -1;
    for (int k = 0; k < dns; ++k) {
      const int32_t user = u_dns[t * dns + k], item = j_dns[t * dns + k];
      const float4* p = reinterpret_cast<const float4*>(P + int64_t(user) * d);
      const float4* q = reinterpret_cast<const float4*>(Q + int64_t(item) * d);
      float acc = 0.f;
      for (int e = 0; e < d / 4; ++e) {
        const float4 a = __ldcg(p + e), b = __ldcg(q + e);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      if (best_j < 0 || acc > best) { best = acc; best_j = item; }
    }
    out_j[t] = best_j;
  }
}

__global__ void __launch_bounds__(256)
truncated_normal_kernel(float* __restrict__ W, int64_t n, float stddev, uint32_t seed, uint32_t table_id, uint32_t tag) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x)
    W[e] = truncated_normal_elem(uint64_t(e), seed, table_id, tag) * stddev;
}

__global__ void fill_kernel(float* __restrict__ x, int64_t n, float v) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += int64_t(gridDim.x) * blockDim.x) x[e] = v;
}

static inline int grid_for(int64_t n, int threads) {
  const int64_t need = (n + threads - 1) / threads;
  const int64_t cap = int64_t(sm_count()) * 16;
  return int(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace apr

using namespace apr;

extern "C" {

int apr_sample_epoch_shard(const int32_t* pairs_u, const int32_t* pairs_i, int64_t n_pairs, int32_t batch,
                           int32_t num_items, const int64_t* csr_ptr, const int32_t* csr_idx, int32_t csr_rows,
                           uint32_t seed, uint32_t epoch, int32_t dns, int32_t batch_lo, int32_t batch_local,
                           int32_t legacy_fork_workers, int32_t* out_u, int32_t* out_i, int32_t* out_udns, int32_t* out_j,
                           int32_t* err_flag, apr_stream_t stream) {
  if (!pairs_u || !pairs_i || !csr_ptr || !out_u || !out_i || !out_udns || !out_j || !err_flag) return APR_E_ARG;
  if (n_pairs < 1 || n_pairs > 0x7fffffffLL || batch < 1 || num_items < 1 || dns < 1 || csr_rows < 0) return APR_E_ARG;
  if (batch_lo < 0 || batch_local < 1 || batch_lo + batch_local > batch || legacy_fork_workers < 0) return APR_E_ARG;
  const int64_t S = n_pairs / batch;
  if (S < 1) return APR_E_ARG;
  const int64_t n_draws = S * batch_local * dns;
  PermKeys keys;
  uint32_t a[4], b[4];
  philox4x32_10(0, 0, 0, epoch, seed, kStreamPerm, a);
  philox4x32_10(1, 0, 0, epoch, seed, kStreamPerm, b);
  for (int k = 0; k < 4; ++k) { keys.k[k] = a[k]; keys.k[4 + k] = b[k]; }
  int bits = 0;
  while ((int64_t(1) << bits) < n_pairs) ++bits;  // bit_length(n-1)
  if (bits < 2) bits = 2;
  bits += bits & 1;
  int64_t fork_chunk_draws = 1;
  if (legacy_fork_workers > 0)
    fork_chunk_draws = std::max<int64_t>(1, (S + 4 * int64_t(legacy_fork_workers) - 1) / (4 * int64_t(legacy_fork_workers))) * batch * dns;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sample_epoch_kernel<<<grid_for(n_draws, 256), 256, 0, st>>>(pairs_u, pairs_i, uint32_t(n_pairs), n_draws, dns,
                                                              uint32_t(num_items), csr_ptr, csr_idx, csr_rows, seed, epoch,
                                                              keys, bits / 2, batch, batch_lo, batch_local, fork_chunk_draws,
                                                              legacy_fork_workers, out_u, out_i, out_udns, out_j, err_flag);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_sample_epoch(const int32_t* pairs_u, const int32_t* pairs_i, int64_t n_pairs, int32_t batch, int32_t num_items,
                     const int64_t* csr_ptr, const int32_t* csr_idx, int32_t csr_rows, uint32_t seed, uint32_t epoch,
                     int32_t dns, int32_t* out_u, int32_t* out_i, int32_t* out_udns, int32_t* out_j, int32_t* err_flag,
                     apr_stream_t stream) {
  return apr_sample_epoch_shard(pairs_u, pairs_i, n_pairs, batch, num_items, csr_ptr, csr_idx, csr_rows, seed, epoch, dns, 0,
                                batch, 0, out_u, out_i, out_udns, out_j, err_flag, stream);
}

int apr_select_dns(const float* P, const float* Q, int32_t d, const int32_t* u_dns, const int32_t* j_dns, int64_t n_pos,
                   int32_t dns, int32_t* out_j, apr_stream_t stream) {
  if (!P || !Q || !u_dns || !j_dns || !out_j || n_pos < 1 || dns < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  select_dns_kernel<<<grid_for(n_pos, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, Q, d, u_dns, j_dns, n_pos, dns,
                                                                                        out_j);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_init_truncated_normal(float* W, int64_t rows, int32_t d, float stddev, uint32_t seed, uint32_t table_id,
                              uint32_t stream_tag, apr_stream_t stream) {
  if (!W || rows < 1 || d < 1) return APR_E_ARG;
  const int64_t n = rows * d;
  truncated_normal_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(W, n, stddev, seed, table_id,
                                                                                          stream_tag);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_fill_f32(float* x, int64_t n, float value, apr_stream_t stream) {
  if (!x || n < 1) return APR_E_ARG;
  fill_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, value);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // extern "C"
