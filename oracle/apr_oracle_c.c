/*
 * CPU oracle in plain C (OpenMP) -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Same arithmetic as oracle/apr_oracle.py (the NumPy restatement of SURVEY Appendix A), written as loops so that it can
 * serve as the multi-threaded CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).  Only tests/ and those two
 * bench legs may load it.  Parity status: as oracle/apr_oracle.py (TensorFlow boundary unpinned); this file is
 * additionally checked against the NumPy oracle in tests/test_oracle_c.py.
 *
 *   apr_oracle_step     one batch of training_batch (utils.py:113-119): APR.py:143-165 (loss), 180-191 (Delta), 193-195
 *                       (Adagrad with duplicate rows summed), on the rows the batch touches.  Pull-based: per-row sums
 *                       are formed by walking each row's occurrence list, so the result is deterministic.
 *   apr_oracle_positions  full-rank leave-one-out positions (utils.py:244-254) with the pinned score order
 *                       acc = fmaf(p[k], q[k], acc), k ascending.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { int32_t key; int32_t pos; } kv_t;
static int kv_cmp(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

static float coeff(float x) {              /* d softplus(-clip(x,-80,1e8)) / dx, APR.py:148-150 */
  float r = x < -80.0f ? -80.0f : (x > 1e8f ? 1e8f : x);
  float m = (x >= -80.0f && x <= 1e8f) ? 1.0f : 0.0f;
  return -m / (1.0f + expf(r));
}

int apr_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* explicit thread count (torchrun exports OMP_NUM_THREADS=1 to its children: the bench overrides it) */
void apr_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* u,i,j: [B] row ids into P / Q.  Returns 0, or -1 on allocation failure. */
int apr_oracle_step(float* P, float* Q, float* accP, float* accQ, int d, const int32_t* u, const int32_t* i,
                    const int32_t* j, int B, float lr, float reg, float reg_adv, float eps, int adver) {
  kv_t* ku = (kv_t*)malloc(sizeof(kv_t) * (size_t)B);
  kv_t* ki = (kv_t*)malloc(sizeof(kv_t) * (size_t)B * 2);
  int32_t* ustart = (int32_t*)malloc(sizeof(int32_t) * ((size_t)B + 1));
  int32_t* istart = (int32_t*)malloc(sizeof(int32_t) * ((size_t)2 * B + 1));
  int32_t* uslot = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);       /* triple -> unique-user index */
  int32_t* islot = (int32_t*)malloc(sizeof(int32_t) * (size_t)B * 2);   /* occurrence (b: pos, B+b: neg) -> unique-item index */
  float* c = (float*)malloc(sizeof(float) * (size_t)B);
  float* ca = (float*)malloc(sizeof(float) * (size_t)B);
  if (!ku || !ki || !ustart || !istart || !uslot || !islot || !c || !ca) return -1;
  for (int b = 0; b < B; ++b) { ku[b].key = u[b]; ku[b].pos = b; ki[b].key = i[b]; ki[b].pos = b; ki[B + b].key = j[b]; ki[B + b].pos = B + b; }
  qsort(ku, (size_t)B, sizeof(kv_t), kv_cmp);
  qsort(ki, (size_t)B * 2, sizeof(kv_t), kv_cmp);
  int nu = 0, ni = 0;
  for (int k = 0; k < B; ++k) { if (k == 0 || ku[k].key != ku[k - 1].key) ustart[nu++] = k; uslot[ku[k].pos] = nu - 1; }
  ustart[nu] = B;
  for (int k = 0; k < 2 * B; ++k) { if (k == 0 || ki[k].key != ki[k - 1].key) istart[ni++] = k; islot[ki[k].pos] = ni - 1; }
  istart[ni] = 2 * B;
  float* dP = (float*)calloc((size_t)nu * d, sizeof(float));   /* Delta rows of the touched users / items */
  float* dQ = (float*)calloc((size_t)ni * d, sizeof(float));
  float* gP = (float*)malloc(sizeof(float) * (size_t)nu * d);
  float* gQ = (float*)malloc(sizeof(float) * (size_t)ni * d);
  if (!dP || !dQ || !gP || !gQ) return -1;
  const float kreg = (float)(2.0 * (double)reg * (adver ? 2.0 : 1.0) / ((double)B * (double)d));

  /* plain forward: x_b, c_b */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    const float* p = P + (size_t)u[b] * d; const float* q = Q + (size_t)i[b] * d; const float* n = Q + (size_t)j[b] * d;
    float yp = 0.f, yn = 0.f;
    for (int k = 0; k < d; ++k) { yp += p[k] * q[k]; yn += p[k] * n[k]; }
    c[b] = coeff(yp - yn);
  }
  /* plain per-row gradients G (duplicates summed), then Delta = eps G / sqrt(max(|G|^2, 1e-12)) (APR.py:183-191) */
#pragma omp parallel for schedule(dynamic, 64)
  for (int r = 0; r < nu; ++r) {
    float* g = gP + (size_t)r * d;
    memset(g, 0, sizeof(float) * (size_t)d);
    for (int k = ustart[r]; k < ustart[r + 1]; ++k) {
      const int b = ku[k].pos;
      const float* q = Q + (size_t)i[b] * d; const float* n = Q + (size_t)j[b] * d;
      for (int t = 0; t < d; ++t) g[t] += c[b] * (q[t] - n[t]);
    }
    if (adver) {
      float ss = 0.f;
      for (int t = 0; t < d; ++t) ss += g[t] * g[t];
      const float sc = eps / sqrtf(ss > 1e-12f ? ss : 1e-12f);
      for (int t = 0; t < d; ++t) dP[(size_t)r * d + t] = g[t] * sc;
    }
  }
#pragma omp parallel for schedule(dynamic, 64)
  for (int r = 0; r < ni; ++r) {
    float* g = gQ + (size_t)r * d;
    memset(g, 0, sizeof(float) * (size_t)d);
    for (int k = istart[r]; k < istart[r + 1]; ++k) {
      const int occ = ki[k].pos, b = occ < B ? occ : occ - B;
      const float s = occ < B ? c[b] : -c[b];
      const float* p = P + (size_t)u[b] * d;
      for (int t = 0; t < d; ++t) g[t] += s * p[t];
    }
    if (adver) {
      float ss = 0.f;
      for (int t = 0; t < d; ++t) ss += g[t] * g[t];
      const float sc = eps / sqrtf(ss > 1e-12f ? ss : 1e-12f);
      for (int t = 0; t < d; ++t) dQ[(size_t)r * d + t] = g[t] * sc;
    }
  }
  /* adversarial forward: x'_b, c'_b (APR.py:158-162) */
  if (adver) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
      const float* p = P + (size_t)u[b] * d; const float* q = Q + (size_t)i[b] * d; const float* n = Q + (size_t)j[b] * d;
      const float* ap = dP + (size_t)uslot[b] * d; const float* aq = dQ + (size_t)islot[b] * d; const float* an = dQ + (size_t)islot[B + b] * d;
      float yp = 0.f, yn = 0.f;
      for (int k = 0; k < d; ++k) { const float pd = p[k] + ap[k]; yp += pd * (q[k] + aq[k]); yn += pd * (n[k] + an[k]); }
      ca[b] = reg_adv * coeff(yp - yn);
    }
  } else {
    for (int b = 0; b < B; ++b) ca[b] = 0.f;
  }
  /* total per-row gradients and Adagrad on each unique row (APR.py:195) */
#pragma omp parallel for schedule(dynamic, 64)
  for (int r = 0; r < nu; ++r) {
    float* g = gP + (size_t)r * d;   /* holds the plain part already */
    const int row = ku[ustart[r]].key;
    const float* w = P + (size_t)row * d;
    for (int k = ustart[r]; k < ustart[r + 1]; ++k) {
      const int b = ku[k].pos;
      const float* q = Q + (size_t)i[b] * d; const float* n = Q + (size_t)j[b] * d;
      const float* aq = dQ + (size_t)islot[b] * d; const float* an = dQ + (size_t)islot[B + b] * d;
      for (int t = 0; t < d; ++t) g[t] += ca[b] * ((q[t] + aq[t]) - (n[t] + an[t])) + kreg * w[t];
    }
  }
#pragma omp parallel for schedule(dynamic, 64)
  for (int r = 0; r < ni; ++r) {
    float* g = gQ + (size_t)r * d;
    const int row = ki[istart[r]].key;
    const float* wq = Q + (size_t)row * d;
    for (int k = istart[r]; k < istart[r + 1]; ++k) {
      const int occ = ki[k].pos, b = occ < B ? occ : occ - B;
      const float s = occ < B ? ca[b] : -ca[b];
      const float* p = P + (size_t)u[b] * d; const float* ap = dP + (size_t)uslot[b] * d;
      for (int t = 0; t < d; ++t) g[t] += s * (p[t] + ap[t]) + kreg * wq[t];
    }
  }
  /* the updates are applied only now: every gradient above was taken at the parameters of the start of the step */
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nu; ++r) {
    const int row = ku[ustart[r]].key;
    float* w = P + (size_t)row * d; float* a = accP + (size_t)row * d; const float* g = gP + (size_t)r * d;
    for (int t = 0; t < d; ++t) { a[t] += g[t] * g[t]; w[t] -= lr * g[t] / sqrtf(a[t]); }
  }
#pragma omp parallel for schedule(static)
  for (int r = 0; r < ni; ++r) {
    const int row = ki[istart[r]].key;
    float* w = Q + (size_t)row * d; float* a = accQ + (size_t)row * d; const float* g = gQ + (size_t)r * d;
    for (int t = 0; t < d; ++t) { a[t] += g[t] * g[t]; w[t] -= lr * g[t] / sqrtf(a[t]); }
  }
  free(ku); free(ki); free(ustart); free(istart); free(uslot); free(islot); free(c); free(ca);
  free(dP); free(dQ); free(gP); free(gQ);
  return 0;
}

static float score_chain(const float* p, const float* q, int d) {
  float acc = 0.f;
  for (int k = 0; k < d; ++k) acc = fmaf(p[k], q[k], acc);
  return acc;
}

/* position[k] = #(c in [0,num_items) minus excl(k) : score(users[k], c) >= score(users[k], test_item[k]))
 * excl = sorted CSR of trainList[u] united with the held-out item (utils.py:210-215,253-254). */
void apr_oracle_positions(const float* P, const float* Q, int d, const int32_t* users, const int32_t* test_item, int n_users,
                          int num_items, const int64_t* excl_ptr, const int32_t* excl_idx, int32_t* position) {
#pragma omp parallel for schedule(dynamic, 8)
  for (int k = 0; k < n_users; ++k) {
    const float* p = P + (size_t)users[k] * d;
    const float sp = score_chain(p, Q + (size_t)test_item[k] * d, d);
    int64_t e = excl_ptr[k];
    const int64_t e1 = excl_ptr[k + 1];
    int cnt = 0;
    for (int c = 0; c < num_items; ++c) {
      while (e < e1 && excl_idx[e] < c) ++e;
      if (e < e1 && excl_idx[e] == c) continue;
      cnt += score_chain(p, Q + (size_t)c * d, d) >= sp;
    }
    position[k] = cnt;
  }
}
