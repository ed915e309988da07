// Does the digit-major layout of the key-switch intermediate ([ct][digit][limb][N]: 14 reads per thread, 3.9 MB apart)
// cost HBM read bandwidth against a limb-major one ([ct][limb][digit][N]: the 14 rows adjacent)?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int LAYOUT>
__global__ void k(const u64* __restrict__ t, u64* out, int cts, int D, int Lk, int logn) {
  const unsigned N = 1u << logn;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = ((size_t)cts * Lk) << logn;
  if (idx >= total) return;
  unsigned c = idx & (N - 1);
  size_t row = idx >> logn;
  unsigned ct = row % cts, j = row / cts;
  u64 acc = 0;
  const u64* p;
  size_t stride;
  if (LAYOUT == 0) { p = t + ((((size_t)ct * D) * Lk + j) << logn) + c; stride = (size_t)Lk << logn; }
  else { p = t + ((((size_t)ct * Lk + j) * D) << logn) + c; stride = (size_t)1 << logn; }
#pragma unroll 2
  for (int i = 0; i < D; i++) acc += p[i * stride];
  out[idx] = acc;
}
int main() {
  const int cts = 32, D = 14, Lk = 15, logn = 15;
  size_t words = ((size_t)cts * D * Lk) << logn, outw = ((size_t)cts * Lk) << logn;
  u64 *t, *out;
  cudaMalloc(&t, words * 8); cudaMalloc(&out, outw * 8);
  cudaMemset(t, 1, words * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++)
    for (int layout = 0; layout < 2; layout++) {
      unsigned blocks = (unsigned)((outw + 255) / 256);
      cudaEventRecord(e0);
      if (layout == 0) k<0><<<blocks, 256>>>(t, out, cts, D, Lk, logn); else k<1><<<blocks, 256>>>(t, out, cts, D, Lk, logn);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("layout %d: %.3f ms  %.2f TB/s read\n", layout, ms, words * 8 / (ms * 1e-3) / 1e12);
    }
  return 0;
}
