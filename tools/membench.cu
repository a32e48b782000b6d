// Ceiling probe for the APR step's memory pattern on B200: per "segment" one lane group reads 6 random rows of D floats
// (P[u], accP[u], Q[i], accQ[i], Q[j], accQ[j]) and writes all 6 back.  Variants isolate what limits throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 [-DROWD=64] -o tools/membench tools/membench.cu && tools/membench
// ROWD = floats per row (default 128 = 512-byte rows, one warp per segment; 64 = 256-byte rows, two segments per warp)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

#ifndef ROWD
#define ROWD 128
#endif
constexpr int D = ROWD;
constexpr int G = D / 4;   // lanes per row (one float4 each)

template <int MODE>  // 0: read only; 1: read+write; 2: read+write with ~reduction chain; 3: read+write, 2 segments in flight
__global__ void __launch_bounds__(256) probe(float* P, float* A, float* Q, float* AQ, const int4* __restrict__ ids, int n,
                                             float* sink) {
  const int lane = threadIdx.x % G;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / G;     // lane-group index
  const int nwarps = (gridDim.x * blockDim.x) / G;
  float acc = 0.f;
  for (int s = warp; s < n; s += nwarps) {
    const int4 t = __ldg(&ids[s]);
    float4* rp = reinterpret_cast<float4*>(P + int64_t(t.x) * D) + lane;
    float4* ra = reinterpret_cast<float4*>(A + int64_t(t.x) * D) + lane;
    float4* rq = reinterpret_cast<float4*>(Q + int64_t(t.y) * D) + lane;
    float4* rb = reinterpret_cast<float4*>(AQ + int64_t(t.y) * D) + lane;
    float4* rn = reinterpret_cast<float4*>(Q + int64_t(t.z) * D) + lane;
    float4* rc = reinterpret_cast<float4*>(AQ + int64_t(t.z) * D) + lane;
    float4 p = __ldcg(rp), a = __ldcg(ra), q = __ldcg(rq), b = __ldcg(rb), nn = __ldcg(rn), c = __ldcg(rc);
    float x = p.x * q.x + p.y * q.y + a.x * b.x + nn.x * c.x;
    if (MODE == 2) {
#pragma unroll
      for (int rep = 0; rep < 6; ++rep) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        x = x * 0.5f + p.z;
      }
    }
    acc += x;
    if (MODE >= 1) {
      p.x += 1e-9f * x; a.x += 1e-9f; q.x += 1e-9f; b.x += 1e-9f; nn.x += 1e-9f; c.x += 1e-9f;
      __stcg(rp, p); __stcg(ra, a); __stcg(rq, q); __stcg(rb, b); __stcg(rn, nn); __stcg(rc, c);
    }
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const int64_t U = 10000000, I = 2000000;
  const int n = 1 << 20;  // segments per launch
  float *P, *A, *Q, *AQ, *sink;
  CK(cudaMalloc(&P, U * D * 4)); CK(cudaMalloc(&A, U * D * 4)); CK(cudaMalloc(&Q, I * D * 4)); CK(cudaMalloc(&AQ, I * D * 4));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(P, 0, U * D * 4)); CK(cudaMemset(A, 0, U * D * 4)); CK(cudaMemset(Q, 0, I * D * 4)); CK(cudaMemset(AQ, 0, I * D * 4));
  std::vector<int4> h(n);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  for (int k = 0; k < n; ++k) h[k] = make_int4(int(rnd() % U), int(rnd() % I), int(rnd() % I), 0);
  int4* ids;
  CK(cudaMalloc(&ids, n * sizeof(int4)));
  CK(cudaMemcpy(ids, h.data(), n * sizeof(int4), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks_list[] = {148 * 2, 148 * 3, 148 * 4, 148 * 8};
  for (int mode = 0; mode < 3; ++mode)
    for (int bl : blocks_list) {
      auto launch = [&]() {
        if (mode == 0) probe<0><<<bl, 256>>>(P, A, Q, AQ, ids, n, sink);
        if (mode == 1) probe<1><<<bl, 256>>>(P, A, Q, AQ, ids, n, sink);
        if (mode == 2) probe<2><<<bl, 256>>>(P, A, Q, AQ, ids, n, sink);
      };
      launch(); launch();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) launch();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      const double bytes = double(n) * 6 * (D * 4) * (mode == 0 ? 1 : 2);
      printf("row %d B  mode %d  warps/SM %2d  %.3f ms  %.0f GB/s  (%.1f ns/segment-chip)\n", D * 4, mode, bl / 148 * 8, ms, bytes / ms / 1e6,
             ms * 1e6 / n);
    }
  return 0;
}
