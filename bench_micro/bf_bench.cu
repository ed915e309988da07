// Butterfly throughput micro-benchmark: Harvey/Shoup vs Solinas-split forward butterflies.
#include <cstdio>
#include <cuda_runtime.h>
#include "../fhe_rs_b200/csrc/ntt.cuh"
using namespace fhe_b200;

template <bool SOL>
__global__ void k(u64* out, const u64* tw, int iters) {
  const u64 p = 4611686018427322369ull, p2 = 2 * p;
  const u32 c = (u32)((1ull << 62) - p);
  u64 x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = (threadIdx.x * 8 + j + blockIdx.x) * 0x9e3779b97f4a7c15ull % p;
  u64 w = tw[threadIdx.x & 15], ws = tw[16 + (threadIdx.x & 15)];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      bf_fwd<SOL>(x[0], x[4], w, ws, p, p2, c);
      bf_fwd<SOL>(x[1], x[5], w, ws, p, p2, c);
      bf_fwd<SOL>(x[2], x[6], w, ws, p, p2, c);
      bf_fwd<SOL>(x[3], x[7], w, ws, p, p2, c);
      bf_fwd<SOL>(x[0], x[2], ws ^ w, w, p, p2, c);
      bf_fwd<SOL>(x[1], x[3], ws ^ w, w, p, p2, c);
      bf_fwd<SOL>(x[4], x[6], w + 1, ws, p, p2, c);
      bf_fwd<SOL>(x[5], x[7], w + 1, ws, p, p2, c);
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <bool SOL>
void run(const char* name) {
  u64 *out, *tw;
  int blocks = 148 * 8, threads = 256, iters = 1000;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaMalloc(&tw, 32 * 8);
  u64 h[32];
  for (int i = 0; i < 32; i++) h[i] = 0x123456789abcdefull * (i + 3) % 4611686018427322369ull;
  cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SOL><<<blocks, threads>>>(out, tw, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<SOL><<<blocks, threads>>>(out, tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double rate = (double)blocks * threads * iters * 32 / (ms * 1e-3);
  printf("%-24s %8.3f ms  %6.3f T butterflies/s  (%.2f per SM per clk @1.9GHz; N=2^15 NTT = %.3f us)\n", name, ms,
         rate / 1e12, rate / 148 / 1.9e9, 245760.0 / rate * 1e6);
}
int main() {
  run<false>("shoup butterfly");
  run<true>("solinas-split butterfly");
  return 0;
}
