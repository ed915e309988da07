// Go / no-go micro-benchmark: the scaler's multiply loop with part of the output limbs accumulated on the FP64 pipe.
//   IMAD route: one 62x62-bit term = 4 IMAD.WIDE + 3 carry adds into a 192-bit lazy accumulator (Acc192 of zq.cuh).
//   DFMA route: both operands as three 21-bit limbs in doubles, 9 DFMA into 5 column sums (every partial sum is an
//               integer < 2^49, so the arithmetic is exact); one recombination per output after the loop.
// The two routes use different pipes (profiles/microbench_r1.txt: IMAD.WIDE and DFMA overlap completely), a DFMA costs
// two issue slots.  Variants <GI, GD>: GI outputs on the integer pipe and GD on the FP64 pipe per pass over the sources.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mac_dfma mac_dfma.cu && ./mac_dfma
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

struct Acc192 {
  u32 e0, e1, e2, e3, e4, o1, o2, o3;
  __device__ __forceinline__ void clear() { e0 = e1 = e2 = e3 = e4 = o1 = o2 = o3 = 0; }
  __device__ __forceinline__ void mac(u64 a, u64 b) {
    asm("{\n\t.reg .u32 a0, a1, b0, b1;\n\tmov.b64 {a0, a1}, %8;\n\tmov.b64 {b0, b1}, %9;\n\t"
        "mad.lo.cc.u32 %0, a0, b0, %0;\n\tmadc.hi.cc.u32 %1, a0, b0, %1;\n\tmadc.lo.cc.u32 %2, a1, b1, %2;\n\t"
        "madc.hi.cc.u32 %3, a1, b1, %3;\n\taddc.u32 %4, %4, 0;\n\tmad.lo.cc.u32 %5, a0, b1, %5;\n\t"
        "madc.hi.cc.u32 %6, a0, b1, %6;\n\taddc.u32 %7, %7, 0;\n\tmad.lo.cc.u32 %5, a1, b0, %5;\n\t"
        "madc.hi.cc.u32 %6, a1, b0, %6;\n\taddc.u32 %7, %7, 0;\n\t}"
        : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o1), "+r"(o2), "+r"(o3) : "l"(a), "l"(b));
  }
  __device__ __forceinline__ u64 mix() const { return ((u64)(e0 ^ e2 ^ e4 ^ o2) << 32) | (e1 ^ e3 ^ o1 ^ o3); }
};
struct AccD5 {
  double c0, c1, c2, c3, c4;
  __device__ __forceinline__ void clear() { c0 = c1 = c2 = c3 = c4 = 0.0; }
  __device__ __forceinline__ void mac(double a0, double a1, double a2, double b0, double b1, double b2) {
    c0 = fma(a0, b0, c0);
    c1 = fma(a0, b1, c1); c1 = fma(a1, b0, c1);
    c2 = fma(a0, b2, c2); c2 = fma(a1, b1, c2); c2 = fma(a2, b0, c2);
    c3 = fma(a1, b2, c3); c3 = fma(a2, b1, c3);
    c4 = fma(a2, b2, c4);
  }
  __device__ __forceinline__ u64 mix() const {
    return (u64)__double_as_longlong(c0) ^ (u64)__double_as_longlong(c1) ^ (u64)__double_as_longlong(c2) ^
           (u64)__double_as_longlong(c3) ^ (u64)__double_as_longlong(c4);
  }
};

constexpr int TC = 128, NF = 29, GMAX = 8;

template <int CONV>
__device__ __forceinline__ void split3(u64 x, double& a0, double& a1, double& a2) {
  const u32 lo = (u32)x, hi = (u32)(x >> 32);
  const u32 l0 = lo & 0x1FFFFF, l1 = __funnelshift_r(lo, hi, 21) & 0x1FFFFF, l2 = hi >> 10;
  if (CONV == 0) {   // cvt.rn.f64.u32
    a0 = (double)l0; a1 = (double)l1; a2 = (double)l2;
  } else {           // 2^52 | limb, minus 2^52: one FP64 add each
    a0 = __hiloint2double(0x43300000, l0) - 4503599627370496.0;
    a1 = __hiloint2double(0x43300000, l1) - 4503599627370496.0;
    a2 = __hiloint2double(0x43300000, l2) - 4503599627370496.0;
  }
}

template <int GI, int GD, int CONV, int UNR>
__global__ void __launch_bounds__(TC) k(u64* out, const u64* tile_src, int reps) {
  __shared__ u64 s_r[NF * TC];
  __shared__ __align__(16) u64 s_om[NF * GMAX];
  __shared__ __align__(16) double s_omd[NF * GMAX * 3];
  for (int i = threadIdx.x; i < NF * TC; i += TC) s_r[i] = tile_src[i] & ((1ull << 62) - 1);
  for (int i = threadIdx.x; i < NF * GMAX; i += TC) {
    const u64 w = (tile_src[i] * 0x9E3779B97F4A7C15ull) & ((1ull << 62) - 1);
    s_om[i] = w;
    s_omd[3 * i] = (double)(w & 0x1FFFFF); s_omd[3 * i + 1] = (double)((w >> 21) & 0x1FFFFF); s_omd[3 * i + 2] = (double)(w >> 42);
  }
  __syncthreads();
  u64 sink = 0;
  for (int rep = 0; rep < reps; rep++) {
    Acc192 ai[GI > 0 ? GI : 1];
    AccD5 ad[GD > 0 ? GD : 1];
#pragma unroll
    for (int g = 0; g < GI; g++) ai[g].clear();
#pragma unroll
    for (int g = 0; g < GD; g++) ad[g].clear();
#pragma unroll UNR
    for (int i = 0; i < NF; i++) {
      const u64 r = s_r[i * TC + threadIdx.x] + rep;
#pragma unroll
      for (int g = 0; g < GI; g++) ai[g].mac(r, s_om[i * GMAX + g]);
      if (GD > 0) {
        double a0, a1, a2;
        split3<CONV>(r, a0, a1, a2);
#pragma unroll
        for (int g = 0; g < GD; g++) {
          const double* b = s_omd + (i * GMAX + g) * 3;
          ad[g].mac(a0, a1, a2, b[0], b[1], b[2]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < GI; g++) sink ^= ai[g].mix();
#pragma unroll
    for (int g = 0; g < GD; g++) sink ^= ad[g].mix();
  }
  out[blockIdx.x * TC + threadIdx.x] = sink;
}

template <int GI, int GD, int CONV, int UNR>
void run(const u64* src, u64* out) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<GI, GD, CONV, UNR>, TC, 0);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k<GI, GD, CONV, UNR>);
  if (per_sm > 8) per_sm = 8;
  const int blocks = 148 * per_sm, reps = 400;
  k<GI, GD, CONV, UNR><<<blocks, TC>>>(out, src, 4);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<GI, GD, CONV, UNR><<<blocks, TC>>>(out, src, reps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double terms = (double)blocks * TC * reps * NF * (GI + GD);
  // clk per warp per term on one SM sub-partition: warps per SMSP * elapsed clk / (terms per warp)
  const double warps_per_smsp = per_sm * (TC / 32) / 4.0;
  const double clk = ms * 1e-3 * 1.965e9 / ((double)reps * NF * (GI + GD) * warps_per_smsp);
  printf("GI=%d GD=%d conv=%d unr=%d  regs=%3d  CTAs/SM=%d  %7.3f ms  %6.3f T terms/s  %5.2f clk per warp-term per SMSP\n", GI, GD, CONV, UNR,
         fa.numRegs, per_sm, ms, terms / ms * 1e-9, clk);
}

int main() {
  u64 *src, *out;
  cudaMalloc(&src, sizeof(u64) * NF * TC);
  cudaMalloc(&out, sizeof(u64) * 148 * 8 * TC);
  u64 h[NF * TC];
  u64 s = 88172645463325252ull;
  for (int i = 0; i < NF * TC; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s; }
  cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
  run<4, 0, 0, 2>(src, out);   // what the scaler does today
  run<0, 3, 0, 2>(src, out);
  run<0, 3, 1, 2>(src, out);
  run<2, 2, 1, 2>(src, out);
  run<3, 2, 1, 2>(src, out);
  run<4, 2, 1, 2>(src, out);
  run<4, 3, 1, 2>(src, out);
  run<4, 3, 0, 2>(src, out);
  run<5, 3, 1, 2>(src, out);
  run<5, 2, 1, 2>(src, out);
  run<6, 2, 1, 2>(src, out);
  run<4, 3, 1, 1>(src, out);
  run<5, 3, 1, 1>(src, out);
  run<4, 4, 1, 1>(src, out);
  return 0;
}
