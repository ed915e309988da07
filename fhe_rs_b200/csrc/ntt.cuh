// Batched negacyclic NTT over Z_p[x]/(x^N+1), one launch over many
// (ciphertext x poly x limb) rows.  Replaces NttOperator::{forward,backward}[_vt]
// (crates/fhe-math/src/ntt/native.rs:77-233) row by row:
//   forward : natural order in -> bit-reversed order out, canonical residues
//   backward: bit-reversed in  -> natural order out, scaled by N^-1, canonical
// using the same twiddle tables (omegas / zetas_inv, native.rs:44-56) and the same
// Harvey lazy butterflies (native.rs:272-316), so every output word is bit-identical.
//
// Decomposition (N = N1 * N2, row viewed as a [N1][N2] matrix):
//   "cols" tile kernel: the log2(N1) large-stride stages on a tile of C adjacent
//          columns (C*8-byte contiguous segments, stride N2) held in shared memory;
//   "rows" tile kernel: the log2(N2) small-stride stages on R adjacent matrix rows
//          (one contiguous R*N2*8-byte chunk) held in shared memory.
// Forward = cols then rows; backward = rows then cols.  Each thread keeps a radix-8
// group (3 stages) in registers between shared-memory exchanges.
#pragma once
#include "zq.cuh"

namespace fhe_b200 {

constexpr int kMaxPos = 64;  // max limbs of any context (cipher + extension)

struct NttArgs {
  const u64* in;        // source rows
  u64* out;             // destination rows (may alias in)
  const LimbDev* limbs; // per-prime constants/tables
  u32 n_rows;           // rows in `out`
  u32 limbs_per_poly;   // out row r has limb position r % limbs_per_poly
  u32 in_div;           // source row of out row r is r / in_div   (1 = same shape;
                        //  L_ksk = key-switch digit broadcast, rq/mod.rs:563-586)
  u32 reduce_on_load;   // reduce source words modulo the row's prime on load
  u32 logn;             // log2 N
  u32 logn1;            // log2 N1 (0 => single-kernel transform done by the rows kernel)
  u32 lazy_out;         // forward only: leave the outputs in [0,4p) (forward_vt_lazy, ntt/native.rs:142-181)
  unsigned short ids[kMaxPos];  // limb position -> index into `limbs`
};

// x in [0,4p), 2p < 2^63: x - 2p is "negative" exactly when x < 2p
__device__ __forceinline__ u64 csub2p(u64 x, u64 p2) {
  u64 t = x - p2;
  return (long long)t < 0 ? x : t;
}
// Forward butterfly (ntt/native.rs:272-285).
//  generic limb : the reference's ranges -- x, y in [0,4p), X = x - 2p*[x >= 2p], T in [0,2p).
//  Solinas limb : any 64-bit x, y; X = x - 2p*[x >= 2^63] in [0, 2^63+2c), T < 2^62+2^61, so
//                 X + T < 2^64 and X + 2p - T in [0, 2^64): same residues, cheaper range control.
template <bool SOL>
__device__ __forceinline__ void bf_fwd(u64& x, u64& y, u64 w, u64 ws, u64 p, u64 p2, u32 c) {
  u64 X = SOL ? fold63_solinas(x, 2 * c) : csub2p(x, p2);
  u64 T = mul_const_lazy<SOL>(y, w, ws, p, c);
  x = X + T;
  y = X + p2 - T;
}
// canonical residue of a forward-transform output
template <bool SOL>
__device__ __forceinline__ u64 fwd_final(u64 v, u64 p, u64 p2, u32 c) {
  if (SOL) v = fold63_solinas(v, 2 * c);  // < 2^63 + 2c = 2p + 4c
  return csub(csub2p(v, p2), p);           // native.rs:238 reduce3
}
// Inverse butterfly (ntt/native.rs:303-316): x, y in [0,2p) in and out.
template <bool SOL>
__device__ __forceinline__ void bf_inv(u64& x, u64& y, u64 z, u64 zs, u64 p, u64 p2, u32 c) {
  u64 t = x;
  x = SOL ? addback2p(t + y - p2, p2) : csub2p(t + y, p2);
  y = mul_const_lazy<SOL>(p2 + t - y, z, zs, p, c);
}

__device__ __forceinline__ u32 sm_phys(u32 i) { return i + (i >> 5); }

// One shared-memory round: NS (1..3) consecutive stages on groups of 2^NS elements.
// LOGP: log2 points of the in-tile transform, LOGB: log2 batch lanes of the tile,
// COLS: layout (true: phys = a*B + b, false: phys = b*P + a).
// t: first local stage of this round (forward numbering).  s_base: global stage of
// local stage 0.  root0: global block offset of this tile's transform at local stage 0
// (0 for cols, matrix row index for rows; per batch lane for the rows layout).
template <int LOGP, int LOGB, bool COLS, bool INV, int NS, bool SOL>
__device__ __forceinline__ void ntt_round(u64* sm, const LimbDev& L, int t, int s_base, u32 logn,
                                          u32 row0, bool first_global_pass) {
  constexpr int R = 1 << NS;
  constexpr u32 P = 1u << LOGP, B = 1u << LOGB;
  constexpr u32 G = (P * B) >> NS;
  const u64 p = L.p, p2 = L.p2;
  const u32 c = (u32)L.sol_c;
  const int logstride = LOGP - t - NS;      // log2 of the in-group element stride
  for (u32 g = threadIdx.x; g < G; g += blockDim.x) {
    u32 b, a_lo, a_hi;
    if (COLS) {
      b = g & (B - 1);
      u32 rest = g >> LOGB;
      a_lo = rest & ((1u << logstride) - 1);
      a_hi = rest >> logstride;
    } else {
      a_lo = g & ((1u << logstride) - 1);
      u32 rest = g >> logstride;
      a_hi = rest & ((1u << t) - 1);
      b = rest >> t;
    }
    const u32 a0 = (a_hi << (LOGP - t)) + a_lo;
    u64 x[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
      u32 a = a0 + ((u32)j << logstride);
      u32 idx = COLS ? (a << LOGB) + b : (b << LOGP) + a;
      x[j] = sm[sm_phys(idx)];
    }
    const u32 root0 = COLS ? 0u : (row0 + b);
    if (!INV) {
#pragma unroll
      for (int u = 0; u < NS; u++) {
        const int half = R >> (u + 1);
        const int tl = t + u;            // local stage
        const int s = s_base + tl;       // global stage: l = N >> (s+1), m = 2^s blocks
#pragma unroll
        for (int jj = 0; jj < R; jj++) {
          if (jj & half) continue;
          u32 i_loc = (a_hi << u) + (jj >> (NS - u));
          u32 k = (1u << s) + (root0 << tl) + i_loc;   // omegas index m + i
          const ulonglong2 w = __ldg(L.om + k);
          bf_fwd<SOL>(x[jj], x[jj + half], w.x, w.y, p, p2, c);
        }
      }
      if (s_base + t + NS == (int)logn) {  // last global stage: reduce3 (native.rs:238)
#pragma unroll
        for (int j = 0; j < R; j++) x[j] = fwd_final<SOL>(x[j], p, p2, c);
      }
    } else {
#pragma unroll
      for (int u = NS - 1; u >= 0; u--) {
        const int half = R >> (u + 1);
        const int tl = t + u;
        const int s = s_base + tl;
        if (s == 0 && first_global_pass) {
          // last inverse stage fused with the N^-1 scaling (native.rs:230-232)
#pragma unroll
          for (int jj = 0; jj < R; jj++) {
            if (jj & half) continue;
            u64 a = x[jj], b2 = x[jj + half];
            x[jj] = csub(mul_const_lazy<SOL>(a + b2, L.ninv, L.ninv_s, p, c), p);
            x[jj + half] = csub(mul_const_lazy<SOL>(p2 + a - b2, L.zn, L.zn_s, p, c), p);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < R; jj++) {
            if (jj & half) continue;
            u32 i_loc = (a_hi << u) + (jj >> (NS - u));
            u32 k = (1u << logn) - (2u << s) + (root0 << tl) + i_loc;  // zetas_inv index N-2m+i
            const ulonglong2 z = __ldg(L.zi + k);
            bf_inv<SOL>(x[jj], x[jj + half], z.x, z.y, p, p2, c);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < R; j++) {
      u32 a = a0 + ((u32)j << logstride);
      u32 idx = COLS ? (a << LOGB) + b : (b << LOGP) + a;
      sm[sm_phys(idx)] = x[j];
    }
  }
}

template <int LOGP, int LOGB, bool COLS, bool INV, bool SOL>
__device__ __forceinline__ void ntt_tile_transform_mode(u64* sm, const LimbDev& L, int s_base, u32 logn,
                                                        u32 row0, bool first_pass) {
  constexpr int NR = (LOGP + 2) / 3;        // rounds
  constexpr int REM = LOGP - 3 * (NR - 1);  // stages in the short round (1..3)
  if (!INV) {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      if (r < NR - 1) ntt_round<LOGP, LOGB, COLS, INV, 3, SOL>(sm, L, 3 * r, s_base, logn, row0, first_pass);
      else ntt_round<LOGP, LOGB, COLS, INV, REM, SOL>(sm, L, 3 * r, s_base, logn, row0, first_pass);
      __syncthreads();
    }
  } else {
#pragma unroll
    for (int r = NR - 1; r >= 0; r--) {
      if (r < NR - 1) ntt_round<LOGP, LOGB, COLS, INV, 3, SOL>(sm, L, 3 * r, s_base, logn, row0, first_pass);
      else ntt_round<LOGP, LOGB, COLS, INV, REM, SOL>(sm, L, 3 * r, s_base, logn, row0, first_pass);
      __syncthreads();
    }
  }
}

template <int LOGP, int LOGB, bool COLS, bool INV>
__device__ __forceinline__ void ntt_tile_transform(u64* sm, const LimbDev& L, int s_base, u32 logn,
                                                   u32 row0, bool first_pass) {
  if (L.sol_ntt) ntt_tile_transform_mode<LOGP, LOGB, COLS, INV, true>(sm, L, s_base, logn, row0, first_pass);
  else ntt_tile_transform_mode<LOGP, LOGB, COLS, INV, false>(sm, L, s_base, logn, row0, first_pass);
}

// cols kernel: tile = P=N1 points x C=2^LOGB adjacent columns.  grid.x = rows * (N2/C).
template <int LOGP, int LOGB, bool INV>
__global__ void ntt_cols_kernel(NttArgs A) {
  extern __shared__ u64 sm[];
  constexpr u32 P = 1u << LOGP, C = 1u << LOGB;
  const u32 logn2 = A.logn - LOGP;
  const u32 tiles = (1u << logn2) >> LOGB;
  const u32 row = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const LimbDev& L = A.limbs[A.ids[row % A.limbs_per_poly]];
  const u64* src = A.in + ((size_t)(row / A.in_div) << A.logn) + (tile << LOGB);
  u64* dst = A.out + ((size_t)row << A.logn) + (tile << LOGB);
  for (u32 e = threadIdx.x; e < P * C; e += blockDim.x) {
    u32 a = e >> LOGB, c = e & (C - 1);
    u64 v = src[((size_t)a << logn2) + c];
    if (A.reduce_on_load) v = barrett64(v, L.p, L.bhi, L.blo);
    sm[sm_phys(e)] = v;
  }
  __syncthreads();
  ntt_tile_transform<LOGP, LOGB, true, INV>(sm, L, 0, A.logn, 0, true);
  for (u32 e = threadIdx.x; e < P * C; e += blockDim.x) {
    u32 a = e >> LOGB, c = e & (C - 1);
    dst[((size_t)a << logn2) + c] = sm[sm_phys(e)];
  }
}

// rows kernel: tile = R=2^LOGB matrix rows x P=N2 contiguous points.  grid.x = rows * (N1/R).
// With logn1 == 0 (N1 = 1, LOGB = 0) it performs the whole transform of one row.
template <int LOGP, int LOGB, bool INV>
__global__ void ntt_rows_kernel(NttArgs A) {
  extern __shared__ u64 sm[];
  constexpr u32 P = 1u << LOGP, R = 1u << LOGB;
  const u32 tiles = (1u << A.logn1) >> LOGB;
  const u32 row = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const LimbDev& L = A.limbs[A.ids[row % A.limbs_per_poly]];
  const u64* src = A.in + ((size_t)(row / A.in_div) << A.logn) + ((size_t)tile << (LOGB + LOGP));
  u64* dst = A.out + ((size_t)row << A.logn) + ((size_t)tile << (LOGB + LOGP));
  for (u32 e = threadIdx.x; e < P * R; e += blockDim.x) {
    u64 v = src[e];
    if (A.reduce_on_load) v = barrett64(v, L.p, L.bhi, L.blo);
    sm[sm_phys(e)] = v;
  }
  __syncthreads();
  // the transforms compare `stage == logn` to find the last stage, where the outputs are fully reduced; a lazy
  // forward transform never meets that condition
  ntt_tile_transform<LOGP, LOGB, false, INV>(sm, L, (int)A.logn1, (!INV && A.lazy_out) ? 0xffu : A.logn, tile << LOGB,
                                             A.logn1 == 0);
  for (u32 e = threadIdx.x; e < P * R; e += blockDim.x) dst[e] = sm[sm_phys(e)];
}

}  // namespace fhe_b200
