// EXPLORATORY micro-benchmark (not part of the library).
//
// The exact RNS scaler's multiply part -- y_j = sum_i r_i * omega_ji mod q_j for every coefficient (rns/scaler.rs:
// 316-351) -- is the one stage of the path that IS a dense contraction: [coefficients x L] residues against a
// constant [L x E] table.  The library runs it on the integer multiplier pipe (4 IMAD.WIDE per 62x62-bit term, lazy
// 192-bit accumulators), which is what bounds it.  This program measures the alternative the north-star excludes
// ("no tensor cores"): byte-slice both operands and let the legacy integer tensor-core path do the 8x8 byte products,
//   r_i = sum_a r_i[a] 2^(8a),  omega_ji = sum_b w_ji[b] 2^(8b)
//   sum_i r_i * omega_ji = sum_d 2^(8d) * C[j][d],   C[j][d] = sum_i sum_{a+b=d} r_i[a] * w_ji[b]   (d = 0..14, < 2^24)
// i.e. one u8 x u8 -> s32 GEMM  [coefficients x 8L] x [8L x 16E]  (B = the Toeplitz expansion of omega, 16 diagonals per
// output limb, the 16th always zero) followed by a per-output recombination of the 15 diagonals and three Solinas
// folds modulo q_j = 2^62 - c_j.  Integer MMA is exact, so the canonical outputs are bit-identical (checked below).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scaler_imma scaler_imma.cu && ./scaler_imma
//
// Prints, for the two scaler shapes of BASELINE set C (29 -> 14 and 14 -> 15 limbs, N = 2^15): time of the
// IMAD.WIDE form (the library's multiply loop + reduction, residues staged in shared memory), time of the MMA form,
// and whether every output word agrees.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../fhe_rs_b200/csrc/zq.cuh"
using namespace fhe_b200;

constexpr int kMaxOut = 16, kMaxIn = 32;
struct Consts {
  u64 q[kMaxOut];
  u32 c[kMaxOut];
};

// ------------------------------------------------------------------------------------------------ IMAD.WIDE form
// one thread per coefficient, residues in shared memory, four outputs at a time (as scale_mac_group in kernels.cu)
template <int NF, int NO>
__global__ void __launch_bounds__(128) mac_imad(const u64* __restrict__ r, const u64* __restrict__ omega, Consts K,
                                                u64* __restrict__ out, size_t ncoef) {
  __shared__ u64 s_r[NF][128];
  __shared__ u64 s_om[NF][kMaxOut];
  const size_t col = (size_t)blockIdx.x * 128 + threadIdx.x;
  for (int i = 0; i < NF; i++) s_r[i][threadIdx.x] = r[(size_t)i * ncoef + col];
  for (int idx = threadIdx.x; idx < NF * kMaxOut; idx += 128) {
    const int i = idx / kMaxOut, j = idx % kMaxOut;
    s_om[i][j] = j < NO ? omega[(size_t)j * NF + i] : 0;
  }
  __syncthreads();
  for (int j0 = 0; j0 < NO; j0 += 4) {
    Acc192 acc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k].clear();
#pragma unroll 2
    for (int i = 0; i < NF; i++) {
      const u64 x = s_r[i][threadIdx.x];
      const ulonglong2 o0 = *reinterpret_cast<const ulonglong2*>(&s_om[i][j0]);
      const ulonglong2 o1 = *reinterpret_cast<const ulonglong2*>(&s_om[i][j0 + 2]);
      acc[0].mac(x, o0.x);
      acc[1].mac(x, o0.y);
      acc[2].mac(x, o1.x);
      acc[3].mac(x, o1.y);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = j0 + k;
      if (j >= NO) break;
      u64 lo, mid;
      u32 hi;
      acc[k].merged(lo, mid, hi);
      out[(size_t)j * ncoef + col] = csub(fold192_solinas(lo, mid, hi, K.c[j]), K.q[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core form
__device__ __forceinline__ void mma_u8(int (&d)[4], const u32 (&a)[4], const u32 (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// CTA = 4 warps x 32 coefficients.  Per warp and group of 4 outputs: 2 m-tiles x 8 n-tiles of s32 accumulators
// (64 registers), KS k-steps of 4 limbs each; then the 64 diagonals of its 32 coefficients go through shared memory
// so that lane l recombines and reduces the 4 outputs of coefficient l.
constexpr int kCsStride = 72;   // ints per coefficient row of the exchange buffer (64 + padding)
template <int NF, int NO>
__global__ void __launch_bounds__(128) mac_imma(const u64* __restrict__ r, const uint2* __restrict__ bfrag, Consts K,
                                                u64* __restrict__ out, size_t ncoef) {
  constexpr int KS = (NF + 3) / 4, NG = (NO + 3) / 4;
  extern __shared__ __align__(16) unsigned char dyn[];
  u64 (*s_r)[128] = reinterpret_cast<u64 (*)[128]>(dyn);                       // [KS*4][128] residues, rows >= NF zero
  int (*s_c)[32][kCsStride] = reinterpret_cast<int (*)[32][kCsStride]>(dyn + (size_t)KS * 4 * 128 * 8);   // [4 warps][coefficient][n = 16*jo + d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const size_t col0 = (size_t)blockIdx.x * 128;
  for (int i = 0; i < KS * 4; i++) s_r[i][threadIdx.x] = i < NF ? r[(size_t)i * ncoef + col0 + threadIdx.x] : 0;
  __syncthreads();
  const int cbase = warp * 32;
  for (int grp = 0; grp < NG; grp++) {
    int acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0;
#pragma unroll 1
    for (int ks = 0; ks < KS; ks++) {
      u32 a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        const u64 x0 = s_r[4 * ks + tig][cbase + 16 * mt + g], x1 = s_r[4 * ks + tig][cbase + 16 * mt + g + 8];
        a[mt][0] = (u32)x0;
        a[mt][1] = (u32)x1;
        a[mt][2] = (u32)(x0 >> 32);
        a[mt][3] = (u32)(x1 >> 32);
      }
      const uint2* bp = bfrag + ((size_t)(grp * KS + ks) * 8) * 32 + lane;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const uint2 bv = __ldg(bp + nt * 32);
        const u32 b[2] = {bv.x, bv.y};
        mma_u8(acc[0][nt], a[0], b);
        mma_u8(acc[1][nt], a[1], b);
      }
    }
    // exchange: C[row][col] of n-tile nt -> s_c[coefficient][8*nt + col]
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        *reinterpret_cast<int2*>(&s_c[warp][16 * mt + g][8 * nt + 2 * tig]) = make_int2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<int2*>(&s_c[warp][16 * mt + g + 8][8 * nt + 2 * tig]) = make_int2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    __syncwarp();
    // lane l: the four outputs of coefficient l of this warp
    const size_t col = col0 + cbase + lane;
#pragma unroll
    for (int jo = 0; jo < 4; jo++) {
      const int j = 4 * grp + jo;
      if (j >= NO) break;
      const int4* cp = reinterpret_cast<const int4*>(&s_c[warp][lane][16 * jo]);
      u64 t[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {   // t_k = sum of the four diagonals that start inside 32-bit word k
        const int4 v = cp[k];
        t[k] = (u64)(u32)v.x + ((u64)(u32)v.y << 8) + ((u64)(u32)v.z << 16) + ((u64)(u32)v.w << 24);
      }
      // V = t0 + t1 2^32 + t2 2^64 + t3 2^96  (t_k < 2^49)
      const u64 lo = t[0] + (t[1] << 32);
      const u64 c0 = lo < t[0];
      const u64 m0 = (t[1] >> 32) + t[2] + c0;
      const u64 mid = m0 + (t[3] << 32);
      const u64 c1 = mid < m0;
      const u64 hi = (t[3] >> 32) + c1;
      out[(size_t)j * ncoef + col] = csub(fold192_solinas(lo, mid, hi, K.c[j]), K.q[j]);
    }
    __syncwarp();
  }
}

// bare issue rate of the legacy integer MMA: 8 independent accumulator tiles per warp, no memory traffic
__global__ void imma_rate(int* out, int iters) {
  int acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int e = 0; e < 4; e++) acc[i][e] = 0;
  u32 a[4] = {threadIdx.x * 2654435761u, threadIdx.x * 40503u + 7, blockIdx.x + 1u, 0x01020304u};
  u32 b[2] = {threadIdx.x * 97u + 3, 0x05060708u};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) mma_u8(acc[i], a, b);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int e = 0; e < 4; e++) s ^= acc[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------ host
static u64 rnd64(u64& s) {
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return s;
}

template <int NF, int NO>
void run(const char* name, int polys) {
  constexpr int KS = (NF + 3) / 4, NG = (NO + 3) / 4;
  const size_t ncoef = (size_t)polys << 15;
  u64 seed = 0x9e3779b97f4a7c15ull + NF;
  Consts K;
  std::vector<u64> omega((size_t)NO * NF);
  for (int j = 0; j < NO; j++) {
    K.c[j] = (u32)((rnd64(seed) & ((1u << 27) - 1)) | 1);
    K.q[j] = (1ull << 62) - K.c[j];
    for (int i = 0; i < NF; i++) omega[(size_t)j * NF + i] = rnd64(seed) % K.q[j];
  }
  // B fragments of mma.m16n8k32 (col-major B): lane (g, tig) holds, for k-step ks and n-tile nt of group grp,
  //   b0 byte e = B[k = 4 tig + e][n = g], b1 byte e = B[k = 16 + 4 tig + e][n = g]
  // with k <-> (limb 4 ks + tig, byte e / 4 + e) and n = 8 nt + g <-> (output 4 grp + nt / 2, diagonal d = 8 (nt & 1) + g)
  std::vector<uint2> bf((size_t)NG * KS * 8 * 32);
  for (int grp = 0; grp < NG; grp++)
    for (int ks = 0; ks < KS; ks++)
      for (int nt = 0; nt < 8; nt++)
        for (int lane = 0; lane < 32; lane++) {
          const int g = lane >> 2, tig = lane & 3, i = 4 * ks + tig, j = 4 * grp + (nt >> 1), d = 8 * (nt & 1) + g;
          u32 w[2] = {0, 0};
          for (int h = 0; h < 2; h++)
            for (int e = 0; e < 4; e++) {
              const int a = 4 * h + e, b = d - a;
              if (i < NF && j < NO && b >= 0 && b < 8) w[h] |= (u32)((omega[(size_t)j * NF + i] >> (8 * b)) & 0xff) << (8 * e);
            }
          bf[(((size_t)(grp * KS + ks) * 8) + nt) * 32 + lane] = make_uint2(w[0], w[1]);
        }
  std::vector<u64> hr((size_t)NF * ncoef);
  for (auto& v : hr) v = rnd64(seed) >> 2;   // < 2^62
  u64 *d_r, *d_om, *d_o1, *d_o2;
  uint2* d_bf;
  cudaMalloc(&d_r, hr.size() * 8);
  cudaMalloc(&d_om, omega.size() * 8);
  cudaMalloc(&d_bf, bf.size() * 8);
  cudaMalloc(&d_o1, (size_t)NO * ncoef * 8);
  cudaMalloc(&d_o2, (size_t)NO * ncoef * 8);
  cudaMemcpy(d_r, hr.data(), hr.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_om, omega.data(), omega.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_bf, bf.data(), bf.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(d_o1, 0, (size_t)NO * ncoef * 8);
  cudaMemset(d_o2, 0xff, (size_t)NO * ncoef * 8);
  const unsigned blocks = (unsigned)(ncoef / 128);
  const size_t smem_b = (size_t)KS * 4 * 128 * 8 + 4 * 32 * kCsStride * 4;
  cudaFuncSetAttribute(mac_imma<NF, NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int rep = 0; rep < 4; rep++) {
      if (rep == 1) cudaEventRecord(e0);
      if (which == 0) mac_imad<NF, NO><<<blocks, 128>>>(d_r, d_om, K, d_o1, ncoef);
      else mac_imma<NF, NO><<<blocks, 128, smem_b>>>(d_r, d_bf, K, d_o2, ncoef);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[which], e0, e1);
    ms[which] /= 3;
  }
  std::vector<u64> o1((size_t)NO * ncoef), o2((size_t)NO * ncoef);
  cudaMemcpy(o1.data(), d_o1, o1.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(o2.data(), d_o2, o2.size() * 8, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t k = 0; k < o1.size(); k++) bad += o1[k] != o2[k];
  // spot check of the IMAD form itself against 128-bit host arithmetic
  size_t bad_host = 0;
  for (int t = 0; t < 64; t++) {
    const size_t col = (rnd64(seed) % ncoef);
    for (int j = 0; j < NO; j++) {
      unsigned __int128 acc = 0;
      for (int i = 0; i < NF; i++)
        acc = (acc + (unsigned __int128)hr[(size_t)i * ncoef + col] % K.q[j] * omega[(size_t)j * NF + i]) % K.q[j];
      bad_host += (u64)acc != o1[(size_t)j * ncoef + col];
    }
  }
  const double terms = (double)ncoef * NF * NO;
  printf("%-22s %d polys of 2^15: IMAD.WIDE %8.3f ms (%.2f T terms/s) | u8 MMA %8.3f ms (%.2f T terms/s) | speed-up %.2fx | "
         "mismatches %zu, host spot-check mismatches %zu, %s\n",
         name, polys, ms[0], terms / ms[0] * 1e-9, ms[1], terms / ms[1] * 1e-9, ms[0] / ms[1], bad, bad_host,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_r); cudaFree(d_om); cudaFree(d_bf); cudaFree(d_o1); cudaFree(d_o2);
}

void rate() {
  int* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int blocks = 148 * warps / 2, iters = 4000;   // 256-thread CTAs, `warps` warps per scheduler
    imma_rate<<<blocks, 256>>>(out, 10);
    cudaEventRecord(e0);
    imma_rate<<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mma = (double)blocks * 8 * iters * 8;
    printf("IMMA.16832.U8.U8 issue rate, %2d warps/scheduler: %.1f G mma/s = %.3f per SM per clk = %.2f POPS (u8 MAC x2)\n", warps,
           mma / ms * 1e-6, mma / (ms * 1e-3) / 148 / 1.965e9, mma * 4096 * 2 / (ms * 1e-3) * 1e-15);
  }
  cudaFree(out);
}

int main() {
  rate();
  run<29, 14>("scale-down 29 -> 14", 48);
  run<14, 15>("extension  14 -> 15", 64);
  return 0;
}
