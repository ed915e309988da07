// Mixed-pipe micro-benchmark: how do IMAD.WIDE / IMAD / IADD3 streams overlap on one SMSP?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int NW, int NL, int NA>
__global__ void k(u64* out, u32 kk, int iters) {
  u64 a[8];
  u32 l[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 77 + i + kk; l[i] = threadIdx.x + i * 3 + kk; b[i] = threadIdx.x ^ (i + kk); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (i < NW) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[i]) : "r"((u32)a[i]), "r"(kk));
        if (i < NL) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(l[i]) : "r"(kk));
#pragma unroll
        for (int j = 0; j < NA / 8; j++) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(kk));
      }
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= a[i] ^ l[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NW, int NL, int NA>
void run() {
  u64* out;
  int blocks = 148 * 8, threads = 256, iters = 1000;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NW, NL, NA><<<blocks, threads>>>(out, 3, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NW, NL, NA><<<blocks, threads>>>(out, 3, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // cycles per warp per "group" (NW wide + NL lo + NA adds) on one SMSP: 16 warps per SMSP
  double groups_per_smsp = 16.0 * iters * 4;
  double clk = ms * 1e-3 * 1.9e9 / groups_per_smsp;
  printf("W=%d L=%d A=%2d : %7.3f ms  -> %6.1f clk per warp-group per SMSP (sum-of-parts model W*3+L*2+A*1 = %d)\n", NW, NL, NA, ms, clk,
         NW * 3 + NL * 2 + NA);
  cudaFree(out);
}
int main() {
  run<8, 0, 0>(); run<0, 8, 0>(); run<0, 0, 32>(); run<0, 0, 64>();
  run<8, 0, 8>(); run<8, 0, 16>(); run<8, 0, 24>(); run<8, 0, 32>(); run<8, 0, 48>();
  run<0, 8, 8>(); run<0, 8, 16>(); run<0, 8, 32>();
  run<8, 8, 0>(); run<8, 8, 16>(); run<8, 8, 32>();
  run<5, 3, 16>(); run<5, 0, 16>(); run<4, 0, 16>(); run<6, 4, 16>();
  return 0;
}
