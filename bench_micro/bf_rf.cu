// What bounds the radix-8 NTT round below the bare-butterfly rate (bf_bench.cu: 3.42 butterflies/clk/SM)?
// Same 62-bit Harvey/Shoup butterflies, ingredients of the real tile kernels added one at a time:
//   A  two or three distinct twiddles for all butterflies (the bf_bench pattern: operand-reuse friendly)
//   B  a radix-8 group with its 7 distinct twiddle pairs held in registers
//   C  B + the 7 pairs re-read from shared memory every round (LDS.128, as ntt_tma.cuh does)
//   D  C + the 8 data words read from / written to shared memory every round (LDS.64 / STS.64, in place)
//   E  D + one CTA barrier per round
// Prints butterflies per clock per SM for 4 and 8 resident warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bf_rf bf_rf.cu && ./bf_rf
#include <cstdio>
#include <cuda_runtime.h>
#include "../fhe_rs_b200/csrc/ntt.cuh"
using namespace fhe_b200;

__device__ __forceinline__ void stages3(u64* v, const ulonglong2* tw, u64 p, u64 p2) {
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int half = 4 >> u;
#pragma unroll
    for (int m = 0; m < (1 << u); m++) {
      const ulonglong2 w = tw[(1 << u) - 1 + m];
#pragma unroll
      for (int e = 0; e < half; e++) {
        const int jj = m * 2 * half + e;
        bf_fwd<false>(v[jj], v[jj + half], w.x, w.y, p, p2, 0);
      }
    }
  }
}

template <int MODE>
__global__ void k(u64* out, const ulonglong2* tw_g, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  const u64 p = 4611686018427322369ull, p2 = 2 * p;
  ulonglong2* s_tw = reinterpret_cast<ulonglong2*>(sm);                  // [7][blockDim]
  u64* s_x = reinterpret_cast<u64*>(sm + 7 * 16 * blockDim.x);           // [8][blockDim]
  u64 x[8];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = (threadIdx.x * 8 + j + blockIdx.x) * 0x9e3779b97f4a7c15ull >> 3;
  ulonglong2 tw[7];
#pragma unroll
  for (int j = 0; j < 7; j++) {
    tw[j] = tw_g[(threadIdx.x * 7 + j) & 63];
    if (MODE == 0) tw[j] = tw_g[j & 1];   // A: two distinct twiddles (warp-uniform values, maximal operand reuse)
    s_tw[j * blockDim.x + threadIdx.x] = tw[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) s_x[j * blockDim.x + threadIdx.x] = x[j];
  __syncthreads();
  for (int i = 0; i < iters; i++) {
    if (MODE >= 2) {
#pragma unroll
      for (int j = 0; j < 7; j++) tw[j] = s_tw[j * blockDim.x + threadIdx.x];
    }
    if (MODE >= 3) {
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = s_x[j * blockDim.x + threadIdx.x];
    }
    stages3(x, tw, p, p2);
    if (MODE >= 3) {
#pragma unroll
      for (int j = 0; j < 8; j++) s_x[((j + 1) & 7) * blockDim.x + threadIdx.x] = x[j];
    }
    if (MODE >= 4) __syncthreads();
  }
  u64 acc = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) acc ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int threads, int bpsm) {
  u64* out;
  ulonglong2* tw;
  const int blocks = 148 * bpsm, iters = 2000;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaMalloc(&tw, 64 * 16);
  ulonglong2 h[64];
  for (int i = 0; i < 64; i++) {
    h[i].x = 0x123456789abcdefull * (i + 3) % 4611686018427322369ull;
    h[i].y = 0xfedcba987654321ull * (i + 7);
  }
  cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
  const size_t smem = (size_t)threads * (7 * 16 + 8 * 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads, smem>>>(out, tw, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads, smem>>>(out, tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double rate = (double)blocks * threads * iters * 12 / (ms * 1e-3);
  printf("%-46s warps/SMSP=%2d %8.3f ms  %6.3f T bf/s  (%.2f per SM per clk @1.965GHz)  %s\n", name,
         threads * bpsm / 128, ms, rate / 1e12, rate / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(tw);
}

int main() {
  for (int cfg = 0; cfg < 2; cfg++) {
    const int threads = 128, bpsm = cfg == 0 ? 4 : 8;
    run<0>("A two twiddles in registers", threads, bpsm);
    run<1>("B seven twiddle pairs in registers", threads, bpsm);
    run<2>("C + twiddles from shared memory each round", threads, bpsm);
    run<3>("D + data through shared memory each round", threads, bpsm);
    run<4>("E + CTA barrier each round", threads, bpsm);
  }
  return 0;
}
