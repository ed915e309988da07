// Second integer-pipe micro-benchmark: variants of 32x32 multiplies and 64-bit adds.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define REP8(X) X X X X X X X X
template <int MODE>
__global__ void k(u64* out, u64 seed, int iters) {
  u32 x0 = seed + threadIdx.x, x1 = seed * 3 + blockIdx.x, x2 = seed ^ 0x9e3779b9u, x3 = x0 ^ x1;
  u32 x4 = x0 + 11, x5 = x1 + 13, x6 = x2 + 17, x7 = x3 + 19;
  u64 a = x0, b = x1, c = x2, d = x3, e = x4, f = x5, g = x6, h = x7;
  double fa = x0, fb = 1.0000001, fc = x2, fd = x3;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {  // mul.wide.u32 (no accumulate) -> xor-fold to keep deps
      REP8(
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a) : "r"(x0), "r"(x1));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(b) : "r"(x1), "r"(x2));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(c) : "r"(x2), "r"(x3));
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(x3), "r"(x0));
        x0 = (u32)a; x1 = (u32)(b >> 32); x2 = (u32)c; x3 = (u32)(d >> 32);
      )
    } else if (MODE == 1) {  // mul.hi.u32
      REP8(
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(x4) : "r"(x0), "r"(x1));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(x5) : "r"(x1), "r"(x2));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(x6) : "r"(x2), "r"(x3));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(x7) : "r"(x3), "r"(x0));
        x0 ^= x4; x1 ^= x5; x2 ^= x6; x3 ^= x7;
      )
    } else if (MODE == 2) {  // mad.hi.u32 with accumulate
      REP8(
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x4) : "r"(x0), "r"(x1));
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x5) : "r"(x1), "r"(x2));
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x6) : "r"(x2), "r"(x3));
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x7) : "r"(x3), "r"(x0));
        x0 += 3; x1 += 5; x2 += 7; x3 += 9;
      )
    } else if (MODE == 3) {  // 64-bit add (2 x IADD3 with carry)
      REP8(
        a += b; b += c; c += d; d += e; e += f; f += g; g += h; h += a;
      )
    } else if (MODE == 4) {  // DFMA
      REP8(
        fa = fma(fa, fb, fc); fc = fma(fc, fb, fd); fd = fma(fd, fb, fa); fb = fma(fb, fb, fa);
      )
    } else if (MODE == 5) {  // mad.wide.u32 with separate (non in-place) accumulator registers
      REP8(
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(x0), "r"(x1), "l"(e));
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(b) : "r"(x1), "r"(x2), "l"(f));
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(c) : "r"(x2), "r"(x3), "l"(g));
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(x3), "r"(x0), "l"(h));
        x0 = (u32)a; x1 = (u32)(b >> 32); x2 = (u32)c; x3 = (u32)(d >> 32);
      )
    } else if (MODE == 6) {  // 64-bit mul.lo (3 IMAD)
      REP8(
        a = a * b + 1; b = b * c + 3; c = c * d + 5; d = d * a + 7;
      )
    } else if (MODE == 7) {  // __umul64hi
      REP8(
        a = __umul64hi(a | 1, b) + 1; b = __umul64hi(b | 1, c) + 3; c = __umul64hi(c | 1, d) + 5; d = __umul64hi(d | 1, a) + 7;
      )
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h ^ x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (u64)(fa + fb + fc + fd);
}
template <int MODE>
void run(const char* name, double ops_per_iter, int iters) {
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
  double rate = (double)blocks * threads * iters * ops_per_iter / (ms * 1e-3);
  printf("%-40s %8.3f ms  %8.2f T lane-ops/s  (%.1f per SM per clk @1.9GHz)\n", name, ms, rate / 1e12, rate / 148 / 1.9e9);
  cudaFree(out);
}
int main() {
  run<0>("mul.wide.u32", 32, 2000);
  run<1>("mul.hi.u32", 32, 2000);
  run<2>("mad.hi.u32", 32, 2000);
  run<3>("add.u64 (64-bit adds)", 64, 2000);
  run<4>("DFMA", 32, 2000);
  run<5>("mad.wide.u32 non-inplace acc", 32, 2000);
  run<6>("mul.lo.u64 + add", 32, 2000);
  run<7>("umul64hi + add", 32, 2000);
  return 0;
}
