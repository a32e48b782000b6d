// N3: the Keras `MatrixFactorization` (MF.py:7-59) and `BPR` (BPR.py:23-102) Recommender models on the same
// gather / dot machinery as the APR step.
//
//   MF.py   pred = <U[u], V[i]>;  loss = mean_b binary_crossentropy(y_b, pred_b)   (Keras: the RAW dot product is clipped
//           to [1e-7, 1 - 1e-7] and treated as a probability -- MF.py:21-24);  optimizer "adam"
//   BPR.py  pred = 1 - log sigmoid(<U[u],V[i]> - <U[u],V[j]>);  loss = mean_b pred_b   (BPR.py:11-21,44-56);  "adam"
//
// Keras' Adam (standalone Keras 2.2 on TF1, BPR.py:1) multiplies the Embedding gradients -- IndexedSlices -- by scalars,
// which densifies them: EVERY row's moments decay and every row moves at every batch, touched or not.  That is restated
// here as written: kmf_grad_kernel scatter-adds the batch gradient into dense gradient tables (vector RED),
// kmf_adam_kernel is one dense pass over a table (reads g and zeroes it, updates m, v, w).  The cost per batch is
// O(table), as in the reference; these models exist for the small Recommender datasets (run.py), not for configs[3].
// Parity status: Keras / TF are not installable here: arithmetic restated from the published
// definitions in oracle/apr_oracle.py (keras_*), kernels tested against that.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace apr {

constexpr float kKerasEps = 1e-7f;   // K.epsilon(): the clip of binary_crossentropy and Adam's epsilon

template <int G>
__device__ __forceinline__ float gsum(float v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// one lane group of G lanes per instance; rows of up to 4 * G * 4 floats (d <= 512 with G = 32)
template <int G>
__global__ void __launch_bounds__(256)
kmf_grad_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d, const int32_t* __restrict__ u,
                const int32_t* __restrict__ i, const int32_t* __restrict__ j, const float* __restrict__ y, int n, int loss_kind,
                float* __restrict__ gP, float* __restrict__ gQ, double* __restrict__ loss_out) {
  const int lane = threadIdx.x % G;
  const unsigned wl = threadIdx.x & 31;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G, ngroups = gridDim.x * (blockDim.x / G);
  const float inv_n = 1.0f / float(n);
  float loss = 0.f;
  for (int b = gid; b < n; b += ngroups) {
    const float* p = P + int64_t(u[b]) * d;
    const float* q = Q + int64_t(i[b]) * d;
    const float* r = loss_kind ? Q + int64_t(j[b]) * d : nullptr;
    float s_pq = 0.f, s_pr = 0.f;
    for (int e = lane * 4; e < d; e += G * 4) {
      const float4 a = ldcg4(p + e), c = ldcg4(q + e);
      s_pq += f4_dot(a, c);
      if (loss_kind) s_pr += f4_dot(a, ldcg4(r + e));
    }
    s_pq = gsum<G>(s_pq, mask);
    float cq, cr = 0.f;     // dL/d<p,q> and dL/d<p,r>, already divided by the batch size (Keras: mean over the batch)
    if (loss_kind == 0) {
      // K.binary_crossentropy(y, pred): pred clipped to [eps, 1 - eps]; the clip passes gradient inside the interval only
      const float yy = y[b];
      const float ph = fminf(fmaxf(s_pq, kKerasEps), 1.0f - kKerasEps);
      loss += -(yy * logf(ph) + (1.0f - yy) * logf(1.0f - ph));
      const float inside = (s_pq >= kKerasEps && s_pq <= 1.0f - kKerasEps) ? 1.0f : 0.0f;
      cq = inside * (-yy / ph + (1.0f - yy) / (1.0f - ph)) * inv_n;
    } else {
      s_pr = gsum<G>(s_pr, mask);
      const float x = s_pq - s_pr;
      // 1 - log sigmoid(x) = 1 + softplus(-x);  d/dx = -sigmoid(-x)
      loss += 1.0f + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
      const float c = -1.0f / (1.0f + expf(x)) * inv_n;
      cq = c; cr = -c;
    }
    for (int e = lane * 4; e < d; e += G * 4) {
      const float4 a = ldcg4(p + e), c = ldcg4(q + e);
      float4 gp = f4_scale(c, cq);
      red_add4(gQ + int64_t(i[b]) * d + e, f4_scale(a, cq));
      if (loss_kind) {
        gp = f4_fma(cr, ldcg4(r + e), gp);
        red_add4(gQ + int64_t(j[b]) * d + e, f4_scale(a, cr));
      }
      red_add4(gP + int64_t(u[b]) * d + e, gp);
    }
  }
  __shared__ double sh;
  if (threadIdx.x == 0) sh = 0.0;
  __syncthreads();
  if (lane == 0 && loss != 0.f) atomicAdd(&sh, double(loss));
  __syncthreads();
  if (threadIdx.x == 0 && sh != 0.0) atomicAdd(loss_out, sh);
}

// Keras 2.2 Adam.get_updates:  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
// w -= lr_t m / (sqrt(v) + eps).  Dense over the whole table; g is consumed (zeroed) here.
__global__ void __launch_bounds__(256)
kmf_adam_kernel(float* __restrict__ W, float* __restrict__ M, float* __restrict__ V, float* __restrict__ Gd, int64_t n4,
                float lr_t, float b1, float b2, float eps) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n4; e += int64_t(gridDim.x) * blockDim.x) {
    float4 g = reinterpret_cast<float4*>(Gd)[e];
    float4 m = reinterpret_cast<float4*>(M)[e], v = reinterpret_cast<float4*>(V)[e], w = reinterpret_cast<float4*>(W)[e];
#define APR_ADAM_LANE(f)                                  \
  m.f = b1 * m.f + (1.0f - b1) * g.f;                     \
  v.f = b2 * v.f + (1.0f - b2) * g.f * g.f;               \
  w.f = w.f - lr_t * m.f / (sqrtf(v.f) + eps);
    APR_ADAM_LANE(x) APR_ADAM_LANE(y) APR_ADAM_LANE(z) APR_ADAM_LANE(w)
#undef APR_ADAM_LANE
    reinterpret_cast<float4*>(M)[e] = m;
    reinterpret_cast<float4*>(V)[e] = v;
    reinterpret_cast<float4*>(W)[e] = w;
    reinterpret_cast<float4*>(Gd)[e] = f4_zero();
  }
}

}  // namespace apr

using namespace apr;

extern "C" {

/* One Keras training batch (model.fit's inner step) of MF.py (loss_kind 0: users, items, labels y in {0,1}) or BPR.py
 * (loss_kind 1: users, positive items, negative items j).  mP/vP/mQ/vQ = Adam moments (zeros initially), gP/gQ = dense
 * gradient scratch of the tables' shapes, all zero on entry and left zero.  t = 1-based Adam iteration.
 * *loss_sum (device double) += the batch's summed loss (mean = / n). */
int apr_keras_step(float* P, float* Q, float* mP, float* vP, float* mQ, float* vQ, float* gP, float* gQ, int64_t rows_p,
                   int64_t rows_q, int32_t d, const int32_t* u, const int32_t* i, const int32_t* j, const float* y,
                   int32_t n, int32_t loss_kind, float lr, float beta1, float beta2, int64_t t, double* loss_sum,
                   apr_stream_t stream) {
  if (!P || !Q || !mP || !vP || !mQ || !vQ || !gP || !gQ || !u || !i || !loss_sum) return APR_E_ARG;
  if (rows_p < 1 || rows_q < 1 || !valid_dim(d) || n < 1 || t < 1) return APR_E_ARG;
  if ((loss_kind == 0 && !y) || (loss_kind == 1 && !j) || loss_kind < 0 || loss_kind > 1) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q) || !aligned16(gP) || !aligned16(gQ) || !aligned16(mP) || !aligned16(vP) || !aligned16(mQ) ||
      !aligned16(vQ))
    return APR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count();
  const int q = d / 4;
  const int grid = std::max(1, std::min((n + 7) / 8, sms * 8));
  if (q <= 8) kmf_grad_kernel<8><<<grid, 256, 0, st>>>(P, Q, d, u, i, j, y, n, loss_kind, gP, gQ, loss_sum);
  else if (q <= 16) kmf_grad_kernel<16><<<grid, 256, 0, st>>>(P, Q, d, u, i, j, y, n, loss_kind, gP, gQ, loss_sum);
  else kmf_grad_kernel<32><<<grid, 256, 0, st>>>(P, Q, d, u, i, j, y, n, loss_kind, gP, gQ, loss_sum);
  APR_LAUNCH_CHECK();
  const double lr_t = double(lr) * sqrt(1.0 - pow(double(beta2), double(t))) / (1.0 - pow(double(beta1), double(t)));
  for (int tb = 0; tb < 2; ++tb) {
    const int64_t n4 = (tb ? rows_q : rows_p) * int64_t(d) / 4;
    const int g2 = int(std::max<int64_t>(1, std::min<int64_t>((n4 + 255) / 256, int64_t(sms) * 16)));
    kmf_adam_kernel<<<g2, 256, 0, st>>>(tb ? Q : P, tb ? mQ : mP, tb ? vQ : vP, tb ? gQ : gP, n4, float(lr_t), beta1, beta2, kKerasEps);
  }
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // extern "C"
