// Butterfly throughput micro-benchmark: instruction-selection variants of the forward / inverse butterflies.
#include <cstdio>
#include <cuda_runtime.h>
#include "../fhe_rs_b200/csrc/ntt.cuh"
using namespace fhe_b200;

// V: 0 shoup+csub2p | 1 solinas v1 + csub2p | 2 solinas chain + fold63 | 3 solinas v1 + fold63 | 4 solinas chain + csub2p
template <int V>
__device__ __forceinline__ void fwd(u64& x, u64& y, u64 w, u64 ws, u64 p, u64 p2, u32 c) {
  u64 X = (V == 2 || V == 3) ? fold63_solinas(x, 2 * c) : csub2p(x, p2);
  u64 T = V == 0 ? mul_shoup_lazy(y, w, ws, p) : (V == 1 || V == 3) ? mul_solinas_lazy_v1(y, w, ws, c) : mul_solinas_lazy(y, w, ws, c);
  x = X + T;
  y = X + p2 - T;
}
// inverse: 0 shoup+csub2p | 1 solinas v1 + csub2p | 2 solinas chain + addback | 3 solinas v1 + addback
template <int V>
__device__ __forceinline__ void inv(u64& x, u64& y, u64 w, u64 ws, u64 p, u64 p2, u32 c) {
  u64 t = x;
  x = (V == 2 || V == 3) ? addback2p(t + y - p2, p2) : csub2p(t + y, p2);
  u64 d = p2 + t - y;
  y = V == 0 ? mul_shoup_lazy(d, w, ws, p) : (V == 1 || V == 3) ? mul_solinas_lazy_v1(d, w, ws, c) : mul_solinas_lazy(d, w, ws, c);
}

template <int V, bool INV>
__global__ void k(u64* out, const u64* tw, int iters) {
  const u64 p = 4611686018427322369ull, p2 = 2 * p;
  const u32 c = (u32)((1ull << 62) - p);
  u64 x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = (threadIdx.x * 8 + j + blockIdx.x) * 0x9e3779b97f4a7c15ull >> 3;
  u64 w = tw[threadIdx.x & 15], ws = tw[16 + (threadIdx.x & 15)], w2 = tw[(threadIdx.x + 1) & 15], w3 = tw[16 + ((threadIdx.x + 5) & 15)];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (!INV) {
        fwd<V>(x[0], x[4], w, ws, p, p2, c); fwd<V>(x[1], x[5], w, ws, p, p2, c);
        fwd<V>(x[2], x[6], w, ws, p, p2, c); fwd<V>(x[3], x[7], w, ws, p, p2, c);
        fwd<V>(x[0], x[2], w2, ws, p, p2, c); fwd<V>(x[1], x[3], w2, ws, p, p2, c);
        fwd<V>(x[4], x[6], w3, w, p, p2, c); fwd<V>(x[5], x[7], w3, w, p, p2, c);
      } else {
        inv<V>(x[0], x[4], w, ws, p, p2, c); inv<V>(x[1], x[5], w, ws, p, p2, c);
        inv<V>(x[2], x[6], w, ws, p, p2, c); inv<V>(x[3], x[7], w, ws, p, p2, c);
        inv<V>(x[0], x[2], w2, ws, p, p2, c); inv<V>(x[1], x[3], w2, ws, p, p2, c);
        inv<V>(x[4], x[6], w3, w, p, p2, c); inv<V>(x[5], x[7], w3, w, p, p2, c);
      }
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int V, bool INV>
void run(const char* name, int threads, int bpsm) {
  u64 *out, *tw;
  int blocks = 148 * bpsm, iters = 1000;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaMalloc(&tw, 32 * 8);
  u64 h[32];
  for (int i = 0; i < 32; i++) h[i] = 0x123456789abcdefull * (i + 3) % 4611686018427322369ull;
  cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V, INV><<<blocks, threads>>>(out, tw, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<V, INV><<<blocks, threads>>>(out, tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double rate = (double)blocks * threads * iters * 32 / (ms * 1e-3);
  printf("%-34s warps/SMSP=%2d %8.3f ms  %6.3f T bf/s  (%.2f per SM per clk @1.9GHz; N=2^15 NTT = %.3f us)\n", name,
         threads * bpsm / 128, ms, rate / 1e12, rate / 148 / 1.9e9, 245760.0 / rate * 1e6);
  cudaFree(out); cudaFree(tw);
}
int main() {
  for (int cfg = 0; cfg < 2; cfg++) {
    int threads = 256, bpsm = cfg == 0 ? 8 : 4;  // 16 or 8 warps per SMSP
    run<0, false>("fwd shoup + csub2p", threads, bpsm);
    run<1, false>("fwd solinas v1 + csub2p", threads, bpsm);
    run<4, false>("fwd solinas chain + csub2p", threads, bpsm);
    run<3, false>("fwd solinas v1 + fold63", threads, bpsm);
    run<2, false>("fwd solinas chain + fold63", threads, bpsm);
    run<0, true>("inv shoup + csub2p", threads, bpsm);
    run<1, true>("inv solinas v1 + csub2p", threads, bpsm);
    run<3, true>("inv solinas v1 + addback", threads, bpsm);
    run<2, true>("inv solinas chain + addback", threads, bpsm);
  }
  return 0;
}
