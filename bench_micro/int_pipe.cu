// Micro-benchmark of the B200 integer pipes relevant to 64-bit modular arithmetic:
// IMAD.WIDE.U32, IMAD (lo), IADD3, and the full Harvey/Shoup butterfly.  Prints lane-ops/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int MODE>
__global__ void k(u64* out, u64 seed, int iters) {
  u64 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, c = seed ^ 0x9e3779b97f4a7c15ull, d = a ^ b;
  u64 e = a + 11, f = b + 13, g = c + 17, h = d + 19;
  const u64 p = 4611686018427322369ull, p2 = 2 * p;
  u64 w = (seed | 1) % p, ws = (u64)((((unsigned __int128)w) << 64) / p);
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {  // mad.wide.u32 chains (8 independent)
#pragma unroll
      for (int r = 0; r < 8; r++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a) : "r"((u32)b), "r"((u32)c));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b) : "r"((u32)c), "r"((u32)d));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"((u32)d), "r"((u32)e));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(d) : "r"((u32)e), "r"((u32)f));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(e) : "r"((u32)f), "r"((u32)g));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(f) : "r"((u32)g), "r"((u32)h));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(g) : "r"((u32)h), "r"((u32)a));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(h) : "r"((u32)a), "r"((u32)b));
      }
    } else if (MODE == 1) {  // 32-bit mad.lo
      u32 x0 = a, x1 = b, x2 = c, x3 = d, x4 = e, x5 = f, x6 = g, x7 = h;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(x1), "r"(x2));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(x2), "r"(x3));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(x3), "r"(x4));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(x4), "r"(x5));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x4) : "r"(x5), "r"(x6));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x5) : "r"(x6), "r"(x7));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x6) : "r"(x7), "r"(x0));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x7) : "r"(x0), "r"(x1));
      }
      a = x0; b = x1; c = x2; d = x3; e = x4; f = x5; g = x6; h = x7;
    } else if (MODE == 2) {  // 32-bit add3 (IADD3)
      u32 x0 = a, x1 = b, x2 = c, x3 = d, x4 = e, x5 = f, x6 = g, x7 = h;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x0) : "r"(x1), "r"(x2));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x1) : "r"(x2), "r"(x3));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x2) : "r"(x3), "r"(x4));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x3) : "r"(x4), "r"(x5));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x4) : "r"(x5), "r"(x6));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x5) : "r"(x6), "r"(x7));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x6) : "r"(x7), "r"(x0));
        asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x7) : "r"(x0), "r"(x1));
      }
      a = x0; b = x1; c = x2; d = x3; e = x4; f = x5; g = x6; h = x7;
    } else if (MODE == 3) {  // Harvey/Shoup forward butterflies, 4 independent pairs
#pragma unroll
      for (int r = 0; r < 4; r++) {
        u64* xs[4] = {&a, &c, &e, &g};
        u64* ys[4] = {&b, &d, &f, &h};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          u64 x = *xs[q], y = *ys[q];
          u64 X = x >= p2 ? x - p2 : x;
          u64 qq = __umul64hi(y, ws);
          u64 T = y * w - qq * p;
          *xs[q] = X + T;
          *ys[q] = X + p2 - T;
        }
      }
    } else if (MODE == 4) {  // 64x64->128 lazy MAC (Acc192 style), 2 accumulators
#pragma unroll
      for (int r = 0; r < 8; r++) {
        u64 pl = a * b, ph = __umul64hi(a, b);
        c += pl; u64 cy = c < pl; d += cy; e += d < cy; d += ph; e += d < ph;
        a += 0x1234567; b ^= c;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int MODE>
double run(const char* name, double ops_per_iter, int iters) {
  u64* out;
  int blocks = 148 * 8, threads = 256;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 12345, 10);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 12345, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = (double)blocks * threads * iters * ops_per_iter;
  double rate = total / (ms * 1e-3);
  printf("%-28s %8.3f ms  %8.2f T lane-ops/s  (%.1f per SM per clk @1.9GHz)\n", name, ms, rate / 1e12,
         rate / 148 / 1.9e9);
  cudaFree(out);
  return rate;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s  SMs=%d  clock=%d kHz  L2=%d MB\n", pr.name, pr.multiProcessorCount, pr.clockRate, pr.l2CacheSize >> 20);
  run<0>("mad.wide.u32 (IMAD.WIDE)", 64, 2000);
  run<1>("mad.lo.u32 (IMAD)", 64, 2000);
  run<2>("add.u32 x2 (IADD3)", 128, 2000);
  run<3>("shoup butterfly", 16, 2000);
  run<4>("64x64 lazy MAC", 8, 2000);
  return 0;
}
