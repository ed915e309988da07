// Mixed-pipe micro-benchmark: do IMAD.WIDE / IMAD / IADD3 / DFMA streams overlap on one SM sub-partition?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int NW, int NL, int NA, int ND>
__global__ void k(u64* out, u32 kk, int iters) {
  u64 a[8];
  u32 l[8], b[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 77 + i + kk; l[i] = threadIdx.x + i * 3 + kk; b[i] = threadIdx.x ^ (i + kk); d[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
  const double m = 1.0000001;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (i < NW) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[i]) : "r"((u32)a[i]), "r"((u32)(a[(i + 1) & 7] >> 32)));
        if (i < NL) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(l[i]) : "r"(l[(i + 1) & 7]), "r"(l[(i + 3) & 7]));
        if (i < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(d[(i + 1) & 7]));
#pragma unroll
        for (int j = 0; j < NA / 8; j++) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(b[(i + j + 1) & 7]));
      }
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= a[i] ^ l[i] ^ b[i] ^ (u64)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NW, int NL, int NA, int ND>
void run() {
  u64* out;
  int blocks = 148 * 8, threads = 256, iters = 1000;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NW, NL, NA, ND><<<blocks, threads>>>(out, 3, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NW, NL, NA, ND><<<blocks, threads>>>(out, 3, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double groups_per_smsp = 16.0 * iters * 4;   // 16 warps per SMSP
  double clk = ms * 1e-3 * 1.9e9 / groups_per_smsp;
  printf("W=%d L=%d A=%2d D=%d : %7.3f ms -> %6.1f clk per warp-group per SMSP\n", NW, NL, NA, ND, ms, clk);
  cudaFree(out);
}
int main() {
  run<8, 0, 0, 0>(); run<0, 8, 0, 0>(); run<0, 0, 16, 0>(); run<0, 0, 32, 0>(); run<0, 0, 0, 8>();
  run<8, 0, 16, 0>(); run<8, 0, 32, 0>(); run<0, 8, 16, 0>(); run<0, 8, 32, 0>(); run<8, 8, 0, 0>(); run<8, 8, 16, 0>();
  run<8, 0, 0, 8>(); run<0, 8, 0, 8>(); run<0, 0, 16, 8>(); run<8, 0, 16, 8>(); run<4, 0, 8, 8>();
  return 0;
}
