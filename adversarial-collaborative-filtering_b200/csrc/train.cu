// APR / BPR-MF training step for sm_100a.
//
// Replaces, per batch (utils.py:113-119):
//     sess.run([update_P, update_Q])   APR.py:180-191   Delta_r = eps * G_r / ||G_r||, G = plain BPR gradient,
//                                                       duplicates of a row inside the batch summed
//     sess.run(optimizer)              APR.py:143-165,193-195   Adagrad on L(Theta) + reg_adv L(Theta+Delta) + reg
//
// Design (DESIGN.md "Training step"):
//   prepare   hash de-duplication of the rows a batch touches.  Users: every distinct user of the batch becomes a
//             SEGMENT (its triples made contiguous by a counting sort) owned by one lane group, so user-side sums
//             (G_P, g_P) live in registers and the user's Adagrad update needs no atomics and no extra pass.
//             Items: every distinct item gets a compact SLOT in an L2-resident workspace [<=2B, d].
//   phase 1   per segment: gather p, per triple gather q,n (128-bit loads), x, c=-sigma(-x); G_P in registers;
//             G_Q[slot] += +-c p by 16-byte vector RED at L2; c kept per triple.
//   phase 2   per segment: Delta_P from G_P (group shuffle norm); per triple Delta_Q from G_Q[slot] rows, adversarial
//             forward, c'; g_P in registers -> Adagrad on the user row; H_Q[slot] += total item gradient (RED).
//   phase 3   per item slot: Adagrad on the item row with H_Q[slot]; G_Q, H_Q slots re-zeroed.
//   Delta and per-triple gradients never exist in HBM as tables; only the touched-row workspace (L2 resident for
//   B <= ~32k at d=128) is written.  BPR (adver=0) fuses phase 1+2.
//   mode 0 launches one kernel per phase; mode 1 runs all steps in ONE persistent cooperative kernel with grid
//   barriers between phases.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace apr {

static inline int pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
struct TrainLayout {
  int S, B, d, Tu, Ti;
  int64_t off_hdr, off_GQ, off_HQ, off_GP, off_cbuf, off_ucnt, off_icnt, off_useg_user, off_useg_off, off_ucursor,
      off_uentry, off_strip, off_slot_i, off_slot_j, off_iu_item, off_tkey_u, off_tval_u, off_tkey_i, off_tval_i, total;
  int64_t zero_bytes;  // [0, zero_bytes) must be zero between steps (hdr, GQ, HQ)
};

static TrainLayout make_layout(int S, int B, int d) {
  TrainLayout L;
  L.S = S; L.B = B; L.d = d;
  L.Tu = max(32, 2 * pow2ceil(B));
  L.Ti = max(32, 4 * pow2ceil(B));
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += (bytes + 255) & ~int64_t(255); return r; };
  L.off_hdr = take(256);
  L.off_GQ = take(int64_t(2) * B * d * 4);
  L.off_HQ = take(int64_t(2) * B * d * 4);
  L.zero_bytes = o;
  L.off_GP = take(int64_t(B) * d * 4);
  L.off_cbuf = take(int64_t(B) * 4);
  L.off_ucnt = take(int64_t(S) * 4);
  L.off_icnt = take(int64_t(S) * 4);
  L.off_useg_user = take(int64_t(S) * B * 4);
  L.off_useg_off = take(int64_t(S) * (B + 1) * 4);
  L.off_ucursor = take(int64_t(S) * B * 4);
  L.off_uentry = take(int64_t(S) * B * 4);
  L.off_strip = take(int64_t(S) * B * 4);
  L.off_slot_i = take(int64_t(S) * B * 4);
  L.off_slot_j = take(int64_t(S) * B * 4);
  L.off_iu_item = take(int64_t(S) * 2 * B * 4);
  L.off_tkey_u = take(int64_t(S) * L.Tu * 4);
  L.off_tval_u = take(int64_t(S) * L.Tu * 4);
  L.off_tkey_i = take(int64_t(S) * L.Ti * 4);
  L.off_tval_i = take(int64_t(S) * L.Ti * 4);
  L.total = o;
  return L;
}

template <typename T>
static inline T* at(void* ws, int64_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }

// header word 0: id-out-of-range flag
// ------------------------------------------------------------------------------------------------
// prepare kernels
// ------------------------------------------------------------------------------------------------
__global__ void prep_insert_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                   const int32_t* __restrict__ j, int S, int B, int64_t rows_p, int64_t rows_q, int Tu,
                                   int Ti, int32_t* tkey_u, int32_t* tval_u, int32_t* tkey_i, int32_t* uentry,
                                   int32_t* slot_i, int32_t* slot_j, int32_t* hdr) {
  const int64_t total = int64_t(S) * 3 * B;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int s = int(t / (3 * int64_t(B)));
    const int n = int(t - int64_t(s) * 3 * B);
    const int which = n / B, b = n - which * B;
    int32_t key;
    int32_t* table;
    int T;
    int64_t rows;
    if (which == 0) { key = u[int64_t(s) * B + b]; table = tkey_u + int64_t(s) * Tu; T = Tu; rows = rows_p; }
    else if (which == 1) { key = i[int64_t(s) * B + b]; table = tkey_i + int64_t(s) * Ti; T = Ti; rows = rows_q; }
    else { key = j[int64_t(s) * B + b]; table = tkey_i + int64_t(s) * Ti; T = Ti; rows = rows_q; }
    if (key < 0 || key >= rows) { atomicOr(hdr, 1); key = 0; }
    uint32_t h = fmix32(uint32_t(key)) & uint32_t(T - 1);
    while (true) {
      const int32_t prev = atomicCAS(&table[h], -1, key);
      if (prev == -1 || prev == key) break;
      h = (h + 1) & uint32_t(T - 1);
    }
    if (which == 0) { atomicAdd(&tval_u[int64_t(s) * Tu + h], 1); uentry[int64_t(s) * B + b] = int32_t(h); }
    else if (which == 1) slot_i[int64_t(s) * B + b] = int32_t(h);
    else slot_j[int64_t(s) * B + b] = int32_t(h);
  }
}

// one thread per hash-table entry; warps never straddle tables (Tu, Ti are multiples of 32)
__global__ void prep_compact_kernel(int S, int B, int Tu, int Ti, const int32_t* __restrict__ tkey_u, int32_t* tval_u,
                                    const int32_t* __restrict__ tkey_i, int32_t* tval_i, int32_t* ucnt, int32_t* icnt,
                                    int32_t* useg_user, int32_t* ucursor, int32_t* iu_item) {
  const int64_t nU = int64_t(S) * Tu, nI = int64_t(S) * Ti;
  const unsigned lane = threadIdx.x & 31;
  for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) & ~int64_t(31); base < nU + nI;
       base += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = base + lane;
    const bool is_user = e < nU;
    int32_t key = -1;
    int s = 0;
    int64_t ei = 0;
    if (is_user) { s = int(e / Tu); key = tkey_u[e]; }
    else { ei = e - nU; s = int(ei / Ti); key = tkey_i[ei]; }
    const bool valid = key != -1;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (m == 0) continue;
    int basev = 0;
    const int leader = __ffs(m) - 1;
    if (int(lane) == leader) basev = atomicAdd(is_user ? &ucnt[s] : &icnt[s], __popc(m));
    basev = __shfl_sync(0xffffffffu, basev, leader);
    if (valid) {
      const int slot = basev + __popc(m & ((1u << lane) - 1u));
      if (is_user) {
        useg_user[int64_t(s) * B + slot] = key;
        ucursor[int64_t(s) * B + slot] = tval_u[e];
        tval_u[e] = slot;
      } else {
        iu_item[int64_t(s) * 2 * B + slot] = key;
        tval_i[ei] = slot;
      }
    }
  }
}

// one CTA per step: exclusive scan of the per-segment counts -> segment offsets; counts reset to 0 (cursors)
__global__ void __launch_bounds__(256) prep_scan_kernel(int B, const int32_t* __restrict__ ucnt, int32_t* ucursor,
                                                        int32_t* useg_off) {
  const int s = blockIdx.x;
  const int nu = ucnt[s];
  int32_t* cnt = ucursor + int64_t(s) * B;
  int32_t* off = useg_off + int64_t(s) * (B + 1);
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nu; base += 256) {
    const int idx = base + threadIdx.x;
    const int v = idx < nu ? cnt[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    int wbase = 0;
    for (int k = 0; k < w; ++k) wbase += warp_tot[k];
    const int carry = carry_s;
    if (idx < nu) { off[idx] = carry + wbase + x - v; cnt[idx] = 0; }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wbase + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[nu] = carry_s;
}

__global__ void prep_scatter_kernel(int S, int B, int Tu, int Ti, const int32_t* __restrict__ tval_u,
                                    const int32_t* __restrict__ tval_i, const int32_t* __restrict__ uentry,
                                    const int32_t* __restrict__ useg_off, int32_t* ucursor, int32_t* strip,
                                    int32_t* slot_i, int32_t* slot_j) {
  const int64_t total = int64_t(S) * 3 * B;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int s = int(t / (3 * int64_t(B)));
    const int n = int(t - int64_t(s) * 3 * B);
    const int which = n / B, b = n - which * B;
    const int64_t sb = int64_t(s) * B + b;
    if (which == 0) {
      const int slot = tval_u[int64_t(s) * Tu + uentry[sb]];
      const int pos = useg_off[int64_t(s) * (B + 1) + slot] + atomicAdd(&ucursor[int64_t(s) * B + slot], 1);
      strip[int64_t(s) * B + pos] = b;
    } else if (which == 1) {
      slot_i[sb] = tval_i[int64_t(s) * Ti + slot_i[sb]];
    } else {
      slot_j[sb] = tval_i[int64_t(s) * Ti + slot_j[sb]];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// step kernels
// ------------------------------------------------------------------------------------------------
struct StepCtx {
  float* P; float* Q; float* accP; float* accQ;
  const int32_t* u; const int32_t* i; const int32_t* j;  // whole chunk [S*B]
  int d, B, S;
  float lr, kreg, reg_adv, eps;
  int adver;
  // workspace
  const int32_t* ucnt; const int32_t* icnt; const int32_t* useg_user; const int32_t* useg_off; const int32_t* strip;
  const int32_t* slot_i; const int32_t* slot_j; const int32_t* iu_item;
  float* GQ; float* HQ; float* GP; float* cbuf;
  float* stats;  // nullable [S,2]
};

template <int G, int V>
struct Row {
  float4 v[V];
};

template <int G, int V>
__device__ __forceinline__ void row_load(Row<G, V>& r, const float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    r.v[k] = (e < d) ? ldcg4(base + e) : f4_zero();
  }
}
template <int G, int V>
__device__ __forceinline__ void row_store(const Row<G, V>& r, float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) stcg4(base + e, r.v[k]);
  }
}
template <int G, int V>
__device__ __forceinline__ void row_red(const Row<G, V>& r, float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) red_add4(base + e, r.v[k]);
  }
}
template <int G, int V>
__device__ __forceinline__ float row_dot(const Row<G, V>& a, const Row<G, V>& b, unsigned mask) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) s += f4_dot(a.v[k], b.v[k]);
  return group_sum<G>(s, mask);
}

__device__ __forceinline__ float softplus_neg(float r) {  // softplus(-r)
  return fmaxf(-r, 0.f) + log1pf(expf(-fabsf(r)));
}
// c = d softplus(-clip(x,-80,1e8)) / dx  (APR.py:148-150; the clip passes gradient on the closed interval)
__device__ __forceinline__ float bpr_coeff(float x, float& r_out) {
  const float r = fminf(fmaxf(x, -80.0f), 1e8f);
  r_out = r;
  const float m = (x >= -80.0f && x <= 1e8f) ? 1.0f : 0.0f;
  return -m / (1.0f + expf(r));
}

template <int G, int V>
__device__ __forceinline__ void adagrad_row(float* W, float* A, const Row<G, V>& w, const Row<G, V>& g, int lane, int d,
                                            float lr) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) {
      float4 a = ldcg4(A + e);
      const float4 gg = g.v[k];
      a.x = fmaf(gg.x, gg.x, a.x); a.y = fmaf(gg.y, gg.y, a.y); a.z = fmaf(gg.z, gg.z, a.z); a.w = fmaf(gg.w, gg.w, a.w);
      float4 nw = w.v[k];
      nw.x -= lr * gg.x / sqrtf(a.x); nw.y -= lr * gg.y / sqrtf(a.y);
      nw.z -= lr * gg.z / sqrtf(a.z); nw.w -= lr * gg.w / sqrtf(a.w);
      stcg4(A + e, a);
      stcg4(W + e, nw);
    }
  }
}

// block-level accumulation of {loss, correct} into stats[2*s .. 2*s+1]
__device__ __forceinline__ void stats_flush(float* stats, int s, float loss, float correct, bool group_leader) {
  __shared__ float sh[2];
  if (threadIdx.x == 0) { sh[0] = 0.f; sh[1] = 0.f; }
  __syncthreads();
  if (group_leader && (loss != 0.f || correct != 0.f)) { atomicAdd(&sh[0], loss); atomicAdd(&sh[1], correct); }
  __syncthreads();
  if (threadIdx.x == 0 && (sh[0] != 0.f || sh[1] != 0.f)) {
    atomicAdd(&stats[2 * s], sh[0]);
    atomicAdd(&stats[2 * s + 1], sh[1]);
  }
  __syncthreads();
}

// phase 1 (adver=1): plain forward/backward.  fused=true (adver=0): complete BPR step for the user side.
template <int G, int V, bool FUSED_BPR>
__device__ __forceinline__ void phase_plain(const StepCtx& c, int s, int gid, int ngroups, int lane, unsigned mask) {
  const int d = c.d, B = c.B;
  const int nu = c.ucnt[s];
  const int32_t* useg_user = c.useg_user + int64_t(s) * B;
  const int32_t* useg_off = c.useg_off + int64_t(s) * (B + 1);
  const int32_t* strip = c.strip + int64_t(s) * B;
  const int32_t* ii = c.i + int64_t(s) * B;
  const int32_t* jj = c.j + int64_t(s) * B;
  const int32_t* slot_i = c.slot_i + int64_t(s) * B;
  const int32_t* slot_j = c.slot_j + int64_t(s) * B;
  float* acc_ws = FUSED_BPR ? c.HQ : c.GQ;
  float loss = 0.f, correct = 0.f;
  for (int seg = gid; seg < nu; seg += ngroups) {
    const int user = useg_user[seg];
    const int b0 = useg_off[seg], b1 = useg_off[seg + 1];
    Row<G, V> p, g;
    row_load<G, V>(p, c.P + int64_t(user) * d, lane, d);
#pragma unroll
    for (int k = 0; k < V; ++k) g.v[k] = f4_zero();
    for (int pos = b0; pos < b1; ++pos) {
      const int b = strip[pos];
      const int it = ii[b], jt = jj[b], si = slot_i[b], sj = slot_j[b];
      Row<G, V> q, n;
      row_load<G, V>(q, c.Q + int64_t(it) * d, lane, d);
      row_load<G, V>(n, c.Q + int64_t(jt) * d, lane, d);
      const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
      float r;
      const float cf = bpr_coeff(x, r);
      if (c.stats) { loss += softplus_neg(r); correct += (x > 0.f) ? 1.f : 0.f; }
      Row<G, V> hi, hj;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        g.v[k] = f4_fma(cf, f4_sub(q.v[k], n.v[k]), g.v[k]);
        hi.v[k] = f4_scale(p.v[k], cf);
        hj.v[k] = f4_scale(p.v[k], -cf);
        if (FUSED_BPR) {
          g.v[k] = f4_fma(c.kreg, p.v[k], g.v[k]);
          hi.v[k] = f4_fma(c.kreg, q.v[k], hi.v[k]);
          hj.v[k] = f4_fma(c.kreg, n.v[k], hj.v[k]);
        }
      }
      row_red<G, V>(hi, acc_ws + int64_t(si) * d, lane, d);
      row_red<G, V>(hj, acc_ws + int64_t(sj) * d, lane, d);
      if (!FUSED_BPR && lane == 0) c.cbuf[pos] = cf;
    }
    if (FUSED_BPR) adagrad_row<G, V>(c.P + int64_t(user) * d, c.accP + int64_t(user) * d, p, g, lane, d, c.lr);
    else row_store<G, V>(g, c.GP + int64_t(seg) * d, lane, d);
  }
  if (c.stats) stats_flush(c.stats, s, loss, correct, lane == 0);
}

// eps * rsqrt(max(||G||^2, 1e-12))  (tf.nn.l2_normalize epsilon, APR.py:190-191)
template <int G, int V>
__device__ __forceinline__ float delta_scale(const Row<G, V>& g, float eps, unsigned mask) {
  const float ss = row_dot<G, V>(g, g, mask);
  return eps / sqrtf(fmaxf(ss, 1e-12f));
}

template <int G, int V>
__device__ __forceinline__ void phase_adv(const StepCtx& c, int s, int gid, int ngroups, int lane, unsigned mask) {
  const int d = c.d, B = c.B;
  const int nu = c.ucnt[s];
  const int32_t* useg_user = c.useg_user + int64_t(s) * B;
  const int32_t* useg_off = c.useg_off + int64_t(s) * (B + 1);
  const int32_t* strip = c.strip + int64_t(s) * B;
  const int32_t* ii = c.i + int64_t(s) * B;
  const int32_t* jj = c.j + int64_t(s) * B;
  const int32_t* slot_i = c.slot_i + int64_t(s) * B;
  const int32_t* slot_j = c.slot_j + int64_t(s) * B;
  for (int seg = gid; seg < nu; seg += ngroups) {
    const int user = useg_user[seg];
    const int b0 = useg_off[seg], b1 = useg_off[seg + 1];
    Row<G, V> p, g, pd;
    row_load<G, V>(p, c.P + int64_t(user) * d, lane, d);
    row_load<G, V>(g, c.GP + int64_t(seg) * d, lane, d);
    const float sp = delta_scale<G, V>(g, c.eps, mask);
#pragma unroll
    for (int k = 0; k < V; ++k) pd.v[k] = f4_fma(sp, g.v[k], p.v[k]);
    for (int pos = b0; pos < b1; ++pos) {
      const int b = strip[pos];
      const int it = ii[b], jt = jj[b], si = slot_i[b], sj = slot_j[b];
      const float cf = __ldcg(&c.cbuf[pos]);
      Row<G, V> q, n, gi, gj;
      row_load<G, V>(q, c.Q + int64_t(it) * d, lane, d);
      row_load<G, V>(n, c.Q + int64_t(jt) * d, lane, d);
      row_load<G, V>(gi, c.GQ + int64_t(si) * d, lane, d);
      row_load<G, V>(gj, c.GQ + int64_t(sj) * d, lane, d);
      const float sci = delta_scale<G, V>(gi, c.eps, mask);
      const float scj = delta_scale<G, V>(gj, c.eps, mask);
      Row<G, V> qd, nd;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        qd.v[k] = f4_fma(sci, gi.v[k], q.v[k]);
        nd.v[k] = f4_fma(scj, gj.v[k], n.v[k]);
      }
      const float xa = row_dot<G, V>(pd, qd, mask) - row_dot<G, V>(pd, nd, mask);
      float ra;
      const float ca = c.reg_adv * bpr_coeff(xa, ra);
      Row<G, V> hi, hj;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        g.v[k] = f4_fma(ca, f4_sub(qd.v[k], nd.v[k]), g.v[k]);
        g.v[k] = f4_fma(c.kreg, p.v[k], g.v[k]);
        const float4 t = f4_fma(ca, pd.v[k], f4_scale(p.v[k], cf));  // c p + reg_adv c' (p + dP)
        hi.v[k] = f4_fma(c.kreg, q.v[k], t);
        hj.v[k] = f4_fma(c.kreg, n.v[k], f4_scale(t, -1.f));
      }
      row_red<G, V>(hi, c.HQ + int64_t(si) * d, lane, d);
      row_red<G, V>(hj, c.HQ + int64_t(sj) * d, lane, d);
    }
    adagrad_row<G, V>(c.P + int64_t(user) * d, c.accP + int64_t(user) * d, p, g, lane, d, c.lr);
  }
}

template <int G, int V>
__device__ __forceinline__ void phase_items(const StepCtx& c, int s, int gid, int ngroups, int lane) {
  const int d = c.d, B = c.B;
  const int ni = c.icnt[s];
  const int32_t* iu_item = c.iu_item + int64_t(s) * 2 * B;
  Row<G, V> z;
#pragma unroll
  for (int k = 0; k < V; ++k) z.v[k] = f4_zero();
  for (int slot = gid; slot < ni; slot += ngroups) {
    const int item = iu_item[slot];
    Row<G, V> w, g;
    row_load<G, V>(g, c.HQ + int64_t(slot) * d, lane, d);
    row_load<G, V>(w, c.Q + int64_t(item) * d, lane, d);
    adagrad_row<G, V>(c.Q + int64_t(item) * d, c.accQ + int64_t(item) * d, w, g, lane, d, c.lr);
    row_store<G, V>(z, c.HQ + int64_t(slot) * d, lane, d);
    if (c.adver) row_store<G, V>(z, c.GQ + int64_t(slot) * d, lane, d);
  }
}

constexpr int kThreads = 256;

template <int G, int V, int PHASE>
__global__ void __launch_bounds__(kThreads) step_phase_kernel(StepCtx c, int s) {
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  const unsigned wl = threadIdx.x & 31;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
  if (PHASE == 0) phase_plain<G, V, true>(c, s, gid, ngroups, lane, mask);
  if (PHASE == 1) phase_plain<G, V, false>(c, s, gid, ngroups, lane, mask);
  if (PHASE == 2) phase_adv<G, V>(c, s, gid, ngroups, lane, mask);
  if (PHASE == 3) phase_items<G, V>(c, s, gid, ngroups, lane);
}

// all steps in one cooperative launch
template <int G, int V>
__global__ void __launch_bounds__(kThreads) step_persistent_kernel(StepCtx c) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  const unsigned wl = threadIdx.x & 31;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
  for (int s = 0; s < c.S; ++s) {
    if (c.adver) {
      phase_plain<G, V, false>(c, s, gid, ngroups, lane, mask);
      grid.sync();
      phase_adv<G, V>(c, s, gid, ngroups, lane, mask);
    } else {
      phase_plain<G, V, true>(c, s, gid, ngroups, lane, mask);
    }
    grid.sync();
    phase_items<G, V>(c, s, gid, ngroups, lane);
    grid.sync();
  }
}

template <int G, int V>
static int run_steps(const StepCtx& c, int mode, cudaStream_t st) {
  const int gpb = kThreads / G;
  const int sms = sm_count();
  if (mode == 0) {
    const int cap = sms * 8;
    const int grid_u = max(1, min((c.B + gpb - 1) / gpb, cap));
    const int grid_i = max(1, min((2 * c.B + gpb - 1) / gpb, cap));
    for (int s = 0; s < c.S; ++s) {
      if (c.adver) {
        step_phase_kernel<G, V, 1><<<grid_u, kThreads, 0, st>>>(c, s);
        step_phase_kernel<G, V, 2><<<grid_u, kThreads, 0, st>>>(c, s);
      } else {
        step_phase_kernel<G, V, 0><<<grid_u, kThreads, 0, st>>>(c, s);
      }
      step_phase_kernel<G, V, 3><<<grid_i, kThreads, 0, st>>>(c, s);
    }
    APR_LAUNCH_CHECK();
    return APR_OK;
  }
  int occ = 0;
  APR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_persistent_kernel<G, V>, kThreads, 0));
  if (occ < 1) return APR_E_CUDA;
  int grid = sms * occ;
  grid = max(1, min(grid, (2 * c.B + gpb - 1) / gpb));
  StepCtx cc = c;
  void* args[] = {&cc};
  APR_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)step_persistent_kernel<G, V>, dim3(grid), dim3(kThreads), args, 0, st));
  return APR_OK;
}

static int dispatch_steps(const StepCtx& c, int mode, cudaStream_t st) {
  const int q = c.d / 4;
  if (q <= 4) return run_steps<4, 1>(c, mode, st);
  if (q <= 8) return run_steps<8, 1>(c, mode, st);
  if (q <= 16) return run_steps<16, 1>(c, mode, st);
  if (q <= 32) return run_steps<32, 1>(c, mode, st);
  if (q <= 64) return run_steps<32, 2>(c, mode, st);
  if (q <= 96) return run_steps<32, 3>(c, mode, st);
  return run_steps<32, 4>(c, mode, st);
}

// ------------------------------------------------------------------------------------------------
// loss / accuracy (utils.py:159-175): one CTA per (step, 256-triple tile)
// ------------------------------------------------------------------------------------------------
template <int G, int V>
__global__ void __launch_bounds__(kThreads) loss_acc_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d,
                                                            const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                                            const int32_t* __restrict__ j, int B, int tiles_per_step,
                                                            double* out) {
  constexpr int kTile = 256;
  const int s = blockIdx.x / tiles_per_step;
  const int tile = blockIdx.x - s * tiles_per_step;
  const int lane = threadIdx.x % G, g = threadIdx.x / G;
  const unsigned wl = threadIdx.x & 31;
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
  const int b_end = min(B, (tile + 1) * kTile);
  float loss = 0.f;
  int correct = 0;
  for (int b = tile * kTile + g; b < b_end; b += kThreads / G) {
    const int64_t t = int64_t(s) * B + b;
    Row<G, V> p, q, n;
    row_load<G, V>(p, P + int64_t(u[t]) * d, lane, d);
    row_load<G, V>(q, Q + int64_t(i[t]) * d, lane, d);
    row_load<G, V>(n, Q + int64_t(j[t]) * d, lane, d);
    const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
    const float r = fminf(fmaxf(x, -80.0f), 1e8f);
    loss += softplus_neg(r);
    correct += x > 0.f;
  }
  __shared__ double sh_loss;
  __shared__ int sh_cor;
  if (threadIdx.x == 0) { sh_loss = 0.0; sh_cor = 0; }
  __syncthreads();
  if (lane == 0) { atomicAdd(&sh_loss, double(loss)); atomicAdd(&sh_cor, correct); }
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(&out[2 * s], sh_loss); atomicAdd(&out[2 * s + 1], double(sh_cor)); }
}

template <int G, int V>
static int run_loss_acc(const float* P, const float* Q, int d, const int32_t* u, const int32_t* i, const int32_t* j, int S,
                        int B, double* out, cudaStream_t st) {
  const int tiles = (B + 255) / 256;
  loss_acc_kernel<G, V><<<S * tiles, kThreads, 0, st>>>(P, Q, d, u, i, j, B, tiles, out);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

}  // namespace apr

using namespace apr;

extern "C" {

int64_t apr_train_workspace_bytes(int32_t n_steps, int32_t batch, int32_t d) {
  if (n_steps < 1 || batch < 1 || !valid_dim(d)) return -1;
  return make_layout(n_steps, batch, d).total;
}

int apr_train_workspace_init(void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!ws || ws_bytes <= 0) return APR_E_ARG;
  APR_CUDA_CHECK(cudaMemsetAsync(ws, 0, size_t(ws_bytes), static_cast<cudaStream_t>(stream)));
  return APR_OK;
}

static int prepare_impl(const int32_t* u, const int32_t* i, const int32_t* j, int S, int B, int d, int64_t rows_p,
                        int64_t rows_q, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const TrainLayout L = make_layout(S, B, d);
  if (ws_bytes < L.total) return APR_E_WORKSPACE;
  // tables: keys = -1, values = 0; counts = 0
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_ucnt), 0, size_t(L.off_useg_user - L.off_ucnt), st));
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_tkey_u), 0xFF, size_t(int64_t(S) * L.Tu * 4), st));
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_tval_u), 0, size_t(int64_t(S) * L.Tu * 4), st));
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_tkey_i), 0xFF, size_t(int64_t(S) * L.Ti * 4), st));
  const int64_t n3 = int64_t(S) * 3 * B;
  const int threads = 256;
  const int cap = sm_count() * 16;
  const int grid_a = int(std::min<int64_t>((n3 + threads - 1) / threads, int64_t(cap)));
  prep_insert_kernel<<<grid_a, threads, 0, st>>>(u, i, j, S, B, rows_p, rows_q, L.Tu, L.Ti, at<int32_t>(ws, L.off_tkey_u),
                                                 at<int32_t>(ws, L.off_tval_u), at<int32_t>(ws, L.off_tkey_i),
                                                 at<int32_t>(ws, L.off_uentry), at<int32_t>(ws, L.off_slot_i),
                                                 at<int32_t>(ws, L.off_slot_j), at<int32_t>(ws, L.off_hdr));
  const int64_t nt = int64_t(S) * (L.Tu + L.Ti);
  const int grid_b = int(std::min<int64_t>((nt + threads - 1) / threads, int64_t(cap)));
  prep_compact_kernel<<<grid_b, threads, 0, st>>>(S, B, L.Tu, L.Ti, at<int32_t>(ws, L.off_tkey_u),
                                                  at<int32_t>(ws, L.off_tval_u), at<int32_t>(ws, L.off_tkey_i),
                                                  at<int32_t>(ws, L.off_tval_i), at<int32_t>(ws, L.off_ucnt),
                                                  at<int32_t>(ws, L.off_icnt), at<int32_t>(ws, L.off_useg_user),
                                                  at<int32_t>(ws, L.off_ucursor), at<int32_t>(ws, L.off_iu_item));
  prep_scan_kernel<<<S, 256, 0, st>>>(B, at<int32_t>(ws, L.off_ucnt), at<int32_t>(ws, L.off_ucursor),
                                      at<int32_t>(ws, L.off_useg_off));
  prep_scatter_kernel<<<grid_a, threads, 0, st>>>(S, B, L.Tu, L.Ti, at<int32_t>(ws, L.off_tval_u),
                                                  at<int32_t>(ws, L.off_tval_i), at<int32_t>(ws, L.off_uentry),
                                                  at<int32_t>(ws, L.off_useg_off), at<int32_t>(ws, L.off_ucursor),
                                                  at<int32_t>(ws, L.off_strip), at<int32_t>(ws, L.off_slot_i),
                                                  at<int32_t>(ws, L.off_slot_j));
  APR_LAUNCH_CHECK();
  return APR_OK;
}

int apr_train_prepare(const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, int32_t d, int64_t rows_p,
                      int64_t rows_q, void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!u || !i || !j || !ws || S < 1 || B < 1 || rows_p < 1 || rows_q < 1 || !valid_dim(d)) return APR_E_ARG;
  return prepare_impl(u, i, j, S, B, d, rows_p, rows_q, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

static int check_train_args(const float* P, const float* Q, const float* accP, const float* accQ, int64_t rows_p,
                            int64_t rows_q, int d, const int32_t* u, const int32_t* i, const int32_t* j, int S, int B,
                            const void* ws) {
  if (!P || !Q || !accP || !accQ || !u || !i || !j || !ws) return APR_E_ARG;
  if (S < 1 || B < 1 || rows_p < 1 || rows_q < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q) || !aligned16(accP) || !aligned16(accQ) || !aligned16(ws)) return APR_E_ALIGN;
  return APR_OK;
}

int apr_train_run(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                  const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, float lr, float reg,
                  float reg_adv, float eps, int32_t adver, int32_t mode, void* ws, int64_t ws_bytes, float* stats,
                  apr_stream_t stream) {
  int rc = check_train_args(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, ws);
  if (rc) return rc;
  if (mode != 0 && mode != 1) return APR_E_ARG;
  const TrainLayout L = make_layout(S, B, d);
  if (ws_bytes < L.total) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepCtx c;
  c.P = P; c.Q = Q; c.accP = accP; c.accQ = accQ; c.u = u; c.i = i; c.j = j;
  c.d = d; c.B = B; c.S = S;
  c.lr = lr;
  // k = 2 reg (1 + [adver]) / (B d): the mean-regulariser is added once, or twice when adver (APR.py:153-154,163-165)
  c.kreg = float(2.0 * double(reg) * (adver ? 2.0 : 1.0) / (double(B) * double(d)));
  c.reg_adv = reg_adv; c.eps = eps; c.adver = adver ? 1 : 0;
  c.ucnt = at<int32_t>(ws, L.off_ucnt); c.icnt = at<int32_t>(ws, L.off_icnt);
  c.useg_user = at<int32_t>(ws, L.off_useg_user); c.useg_off = at<int32_t>(ws, L.off_useg_off);
  c.strip = at<int32_t>(ws, L.off_strip); c.slot_i = at<int32_t>(ws, L.off_slot_i); c.slot_j = at<int32_t>(ws, L.off_slot_j);
  c.iu_item = at<int32_t>(ws, L.off_iu_item);
  c.GQ = at<float>(ws, L.off_GQ); c.HQ = at<float>(ws, L.off_HQ); c.GP = at<float>(ws, L.off_GP);
  c.cbuf = at<float>(ws, L.off_cbuf);
  c.stats = stats;
  if (stats) APR_CUDA_CHECK(cudaMemsetAsync(stats, 0, size_t(S) * 2 * sizeof(float), st));
  return dispatch_steps(c, mode, st);
}

int apr_train_steps(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                    const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, float lr, float reg,
                    float reg_adv, float eps, int32_t adver, int32_t mode, void* ws, int64_t ws_bytes, float* stats,
                    apr_stream_t stream) {
  int rc = check_train_args(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, ws);
  if (rc) return rc;
  rc = prepare_impl(u, i, j, S, B, d, rows_p, rows_q, ws, ws_bytes, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return apr_train_run(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, lr, reg, reg_adv, eps, adver, mode, ws, ws_bytes,
                       stats, stream);
}

int apr_train_unique_counts(const void* ws, int32_t S, int32_t B, int32_t d, int32_t* counts_host, apr_stream_t stream) {
  if (!ws || !counts_host || S < 1 || B < 1 || !valid_dim(d)) return APR_E_ARG;
  const TrainLayout L = make_layout(S, B, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* tmp = static_cast<int32_t*>(malloc(size_t(S) * 8));
  if (!tmp) return APR_E_ARG;
  cudaError_t e = cudaMemcpyAsync(tmp, static_cast<const char*>(ws) + L.off_ucnt, size_t(S) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(tmp + S, static_cast<const char*>(ws) + L.off_icnt, size_t(S) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { free(tmp); set_cuda_error(e, "apr_train_unique_counts"); return APR_E_CUDA; }
  for (int s = 0; s < S; ++s) { counts_host[2 * s] = tmp[s]; counts_host[2 * s + 1] = tmp[S + s]; }
  // id-range flag
  int32_t flag = 0;
  e = cudaMemcpyAsync(&flag, static_cast<const char*>(ws) + L.off_hdr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  free(tmp);
  if (e != cudaSuccess) { set_cuda_error(e, "apr_train_unique_counts"); return APR_E_CUDA; }
  return flag ? APR_E_ARG : APR_OK;
}

int apr_loss_acc(const float* P, const float* Q, int32_t d, const int32_t* u, const int32_t* i, const int32_t* j, int32_t S,
                 int32_t B, double* out, apr_stream_t stream) {
  if (!P || !Q || !u || !i || !j || !out || S < 1 || B < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  APR_CUDA_CHECK(cudaMemsetAsync(out, 0, size_t(S) * 2 * sizeof(double), st));
  const int q = d / 4;
  if (q <= 4) return run_loss_acc<4, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 8) return run_loss_acc<8, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 16) return run_loss_acc<16, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 32) return run_loss_acc<32, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 64) return run_loss_acc<32, 2>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 96) return run_loss_acc<32, 3>(P, Q, d, u, i, j, S, B, out, st);
  return run_loss_acc<32, 4>(P, Q, d, u, i, j, S, B, out, st);
}

}  // extern "C"
