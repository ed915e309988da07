// Element-wise ring ops, tensor product, exact RNS scaler, key-switch inner product,
// Galois gather and modulus switch-down kernels (sm_100a).  64-bit integer modular
// arithmetic, HBM / integer-pipe bound: no tensor cores.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "engine.hpp"
#include "ntt_tma.cuh"

namespace fhe_b200 {

typedef unsigned __int128 u128;

namespace {

// ------------------------------------------------------------------ element-wise
struct EwArgs {
  u64* a;
  const u64* b;
  size_t n_words;
  u32 logn, limbs_per_poly;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};

// Modulus::{add,sub,neg}_vec (zq/mod.rs:240-326, :534-550) over every row of a batch
template <int OP>
__global__ void ew_kernel(EwArgs A) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= A.n_words) return;
  const u64 p = A.limbs[A.ids[(i >> A.logn) % A.limbs_per_poly]].p;
  ulonglong2 x = *reinterpret_cast<ulonglong2*>(A.a + i);
  if (OP == EW_NEG) {
    x.x = csub(p - x.x, p);
    x.y = csub(p - x.y, p);
  } else {
    ulonglong2 y = *reinterpret_cast<const ulonglong2*>(A.b + i);
    if (OP == EW_ADD) {
      x.x = csub(x.x + y.x, p);
      x.y = csub(x.y + y.y, p);
    } else {
      x.x = csub(x.x + p - y.x, p);
      x.y = csub(x.y + p - y.y, p);
    }
  }
  *reinterpret_cast<ulonglong2*>(A.a + i) = x;
}

// ------------------------------------------------------------------ ciphertext x plaintext polynomial
struct MulPlainArgs {
  u64* a;
  const u64* pt;
  u32 cts, parts, n_pt, logn, limbs_per_poly;
  u32 op;   // 0: every part *= pt (ops/mod.rs:229); 1: part 0 += pt (:88-97); 2: part 0 -= pt (:188-197)
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
__global__ void mul_plain_kernel(MulPlainArgs A) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = ((size_t)A.cts * A.parts * A.limbs_per_poly) << A.logn;
  if (idx >= total) return;
  const u32 c = idx & ((1u << A.logn) - 1);
  const size_t row = idx >> A.logn;
  const u32 limb = row % A.limbs_per_poly;
  const u32 ct = (u32)(row / ((size_t)A.limbs_per_poly * A.parts));
  const LimbDev& M = A.limbs[A.ids[limb]];
  const u64 w = A.pt[((((size_t)(ct % A.n_pt)) * A.limbs_per_poly + limb) << A.logn) + c];
  if (A.op == 0) {
    A.a[idx] = mulmod_limb(A.a[idx], w, M);
  } else if ((row / A.limbs_per_poly) % A.parts == 0) {
    const u64 x = A.a[idx];
    A.a[idx] = A.op == 1 ? csub(x + w, M.p) : csub(x + M.p - w, M.p);   // Modulus::add / sub, zq/mod.rs:103-128
  }
}

// ------------------------------------------------------------------ dot_product_scalar
struct DotArgs {
  const u64 *ct, *pt;
  u64* out;
  u32 groups, n_terms, parts, ct_count, pt_count, limbs_per_poly, logn;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
// out[g][part][limb][:] = sum_i ct[g*n + i][part][limb][:] * pt[g*n + i][limb][:]   (an operand with only n entries
// is shared by all groups)
// (bfv/ops/dot_product.rs:55-184: u128 fused multiply-adds per coefficient, one reduction at the end; the lazy
// register here is 160 bits wide, so no term-count threshold / fallback path is needed).  HBM-bound: two words
// read per multiply; four terms in flight per trip.
__global__ void dot_kernel(DotArgs A) {
  const u32 N = 1u << A.logn;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over groups*parts*limbs*N
  size_t total = ((size_t)A.groups * A.parts * A.limbs_per_poly) << A.logn;
  if (idx >= total) return;
  const u32 c = idx & (N - 1);
  size_t row = idx >> A.logn;
  const u32 limb = row % A.limbs_per_poly;
  row /= A.limbs_per_poly;
  const u32 part = row % A.parts, g = (u32)(row / A.parts);
  const LimbDev& M = A.limbs[A.ids[limb]];
  const size_t ct_stride = ((size_t)A.parts * A.limbs_per_poly) << A.logn, pt_stride = (size_t)A.limbs_per_poly << A.logn;
  const u64* cp = A.ct + (((size_t)part * A.limbs_per_poly + limb) << A.logn) + c;
  const u64* pp = A.pt + ((size_t)limb << A.logn) + c;
  Acc192 acc;
  acc.clear();
  // an operand holds either n_terms entries (shared by every group) or groups * n_terms (checked by the caller)
  cp += (A.ct_count == A.n_terms ? 0 : (size_t)g * A.n_terms) * ct_stride;
  pp += (A.pt_count == A.n_terms ? 0 : (size_t)g * A.n_terms) * pt_stride;
  u32 i = 0;
  for (; i + 4 <= A.n_terms; i += 4) {   // eight independent loads in flight
    u64 x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      x[k] = cp[(size_t)(i + k) * ct_stride];
      y[k] = pp[(size_t)(i + k) * pt_stride];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) acc.mac(x[k], y[k]);
  }
  for (; i < A.n_terms; i++) acc.mac(cp[(size_t)i * ct_stride], pp[(size_t)i * pt_stride]);
  A.out[idx] = acc.reduce(M);
}

// ------------------------------------------------------------------ tensor
struct TensorArgs {
  const u64 *a, *b, *xa, *xb;
  u64* out;
  u32 cts, L, nca, ncb, K, logn;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
// c0 = a0*b0, c1 = a0*b1 + a1*b0, c2 = a1*b1 (bfv/ops/mul.rs:198-201; Modulus::mul_vec zq/mod.rs:332).
// Operand x (x = a, b) supplies its first nc_x mul-basis limbs from the ciphertext itself ([ct][2][L][N], the
// common prefix a factor-one extender keeps, rq/scaler.rs:61-65) and the other K - nc_x from the scaled rows
// ([ct][2][K - nc_x][N]).
__global__ void tensor_kernel(TensorArgs A) {
  const u32 N = 1u << A.logn, K = A.K;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over cts*K*N
  size_t total = (size_t)A.cts * K << A.logn;
  if (idx >= total) return;
  u32 c = idx & (N - 1);
  size_t row = idx >> A.logn;
  u32 pos = row % K, ct = row / K;
  const LimbDev& M = A.limbs[A.ids[pos]];
  u64 a0, a1, b0, b1;
  if (pos < A.nca) {
    size_t o = (((size_t)ct * 2) * A.L + pos) << A.logn;
    a0 = A.a[o + c]; a1 = A.a[o + ((size_t)A.L << A.logn) + c];
  } else {
    const u32 E = K - A.nca;
    size_t o = (((size_t)ct * 2) * E + (pos - A.nca)) << A.logn;
    a0 = A.xa[o + c]; a1 = A.xa[o + ((size_t)E << A.logn) + c];
  }
  if (pos < A.ncb) {
    size_t o = (((size_t)ct * 2) * A.L + pos) << A.logn;
    b0 = A.b[o + c]; b1 = A.b[o + ((size_t)A.L << A.logn) + c];
  } else {
    const u32 E = K - A.ncb;
    size_t o = (((size_t)ct * 2) * E + (pos - A.ncb)) << A.logn;
    b0 = A.xb[o + c]; b1 = A.xb[o + ((size_t)E << A.logn) + c];
  }
  u64 c0 = mulmod_limb(a0, b0, M);
  u64 c2 = mulmod_limb(a1, b1, M);
  Acc192 s;                                 // a0*b1 + a1*b0 < 2^125, one reduction
  s.clear();
  s.mac(a0, b1);
  s.mac(a1, b0);
  u64 c1 = s.reduce(M);
  size_t o = (((size_t)ct * 3) * K + pos) << A.logn;
  A.out[o + c] = c0;
  A.out[o + ((size_t)K << A.logn) + c] = c1;
  A.out[o + ((size_t)2 * K << A.logn) + c] = c2;
}

// General part counts of &ct * &ct (bfv/ops/mod.rs:259-358): c[k] = sum_{i+j=k} a_i * b_j over the multiplication
// basis, one thread per (ciphertext, output part k, limb, coefficient).  a: [ct][na][L][N], xa: [ct][na][E][N] etc.
struct TensorNmArgs {
  const u64 *a, *b, *xa, *xb;
  u64* out;
  u32 cts, L, E, na, nb, logn;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
__global__ void tensor_nm_kernel(TensorNmArgs A) {
  const u32 N = 1u << A.logn, K = A.L + A.E, nc = A.na + A.nb - 1;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over cts*nc*K*N
  size_t total = ((size_t)A.cts * nc * K) << A.logn;
  if (idx >= total) return;
  const u32 c = idx & (N - 1);
  size_t row = idx >> A.logn;
  const u32 pos = row % K;
  row /= K;
  const u32 k = row % nc, ct = (u32)(row / nc);
  const LimbDev& M = A.limbs[A.ids[pos]];
  const bool ext = pos >= A.L;
  const u32 rows = ext ? A.E : A.L, r = ext ? pos - A.L : pos;
  const u64* pa = (ext ? A.xa : A.a) + ((((size_t)ct * A.na) * rows + r) << A.logn) + c;
  const u64* pb = (ext ? A.xb : A.b) + ((((size_t)ct * A.nb) * rows + r) << A.logn) + c;
  const size_t ps = (size_t)rows << A.logn;
  Acc192 acc;
  acc.clear();
  const u32 lo = k + 1 > A.nb ? k + 1 - A.nb : 0, hi = k < A.na - 1 ? k : A.na - 1;
  for (u32 i = lo; i <= hi; i++) acc.mac(pa[i * ps], pb[(k - i) * ps]);
  A.out[idx] = acc.reduce(M);
}

// ------------------------------------------------------------------ exact RNS scaler
struct ScaleArgs {
  ScalerDev S;
  const LimbDev* limbs;
  const u64* in;
  u64 *out0, *out1;
  u32 polys, out_rows_per_poly, start, n_out, split3, logn;
};

// 256-bit two's-complement helpers on 4 x u64 (stand-in for ethnum::U256 wrapping arithmetic)
struct U256 {
  u64 w0, w1, w2, w3;
};
__device__ __forceinline__ U256 u256_from_acc(const u32 (&a)[7]) {
  U256 r;
  r.w0 = ((u64)a[1] << 32) | a[0];
  r.w1 = ((u64)a[3] << 32) | a[2];
  r.w2 = ((u64)a[5] << 32) | a[4];
  r.w3 = a[6];
  return r;
}
__device__ __forceinline__ U256 u256_add(U256 a, U256 b) {
  U256 r;
  asm("add.cc.u64 %0, %4, %8;\n\t"
      "addc.cc.u64 %1, %5, %9;\n\t"
      "addc.cc.u64 %2, %6, %10;\n\t"
      "addc.u64 %3, %7, %11;"
      : "=l"(r.w0), "=l"(r.w1), "=l"(r.w2), "=l"(r.w3)
      : "l"(a.w0), "l"(a.w1), "l"(a.w2), "l"(a.w3), "l"(b.w0), "l"(b.w1), "l"(b.w2), "l"(b.w3));
  return r;
}
__device__ __forceinline__ U256 u256_sub(U256 a, U256 b) {
  U256 r;
  asm("sub.cc.u64 %0, %4, %8;\n\t"
      "subc.cc.u64 %1, %5, %9;\n\t"
      "subc.cc.u64 %2, %6, %10;\n\t"
      "subc.u64 %3, %7, %11;"
      : "=l"(r.w0), "=l"(r.w1), "=l"(r.w2), "=l"(r.w3)
      : "l"(a.w0), "l"(a.w1), "l"(a.w2), "l"(a.w3), "l"(b.w0), "l"(b.w1), "l"(b.w2), "l"(b.w3));
  return r;
}
// (128-bit v) * (128-bit theta) mod 2^256
__device__ __forceinline__ U256 u256_mul_128(u128 v, u64 tlo, u64 thi) {
  u32 a[7] = {0, 0, 0, 0, 0, 0, 0}, b[7] = {0, 0, 0, 0, 0, 0, 0};
  mac_theta(a, (u64)v, tlo, thi);
  mac_theta(b, (u64)(v >> 64), tlo, thi);
  U256 lo = u256_from_acc(a), hi = u256_from_acc(b);
  U256 hs = {0, hi.w0, hi.w1, hi.w2};
  return u256_add(lo, hs);
}

// RnsScaler::scale (rns/scaler.rs:249-352): one thread per coefficient column, 128 columns per CTA.
// The n_from source residues of the tile are staged in shared memory (coalesced load, conflict-free
// reads) so the thread keeps only accumulators in registers; the fixed-point sums (v, w) are computed
// exactly as coded in the reference; the output limbs are produced four at a time (four independent
// lazy accumulators per thread give the instruction-level parallelism the single dependent carry
// chain lacks -- ncu: the register-resident one-limb-at-a-time form ran at 0.3 IPC).
constexpr int kScaleTC = 128;

// sum_i r_i * omega_{j0+k, i} for the G (<= 4) output limbs of one group
template <int G, int UNR = 2>
__device__ __forceinline__ void scale_mac_group(Acc192 (&acc)[4], const u64* r_col, const ulonglong2* om, u32 nf,
                                                u32 om_stride) {
#pragma unroll UNR
  for (u32 i = 0; i < nf; i++) {
    const u64 r = r_col[i * kScaleTC];
    const ulonglong2 o0 = om[(size_t)i * om_stride];
    acc[0].mac(r, o0.x);
    if (G > 1) acc[1].mac(r, o0.y);
    if (G > 2) {
      const ulonglong2 o1 = om[(size_t)i * om_stride + 1];
      acc[2].mac(r, o1.x);
      if (G > 3) acc[3].mac(r, o1.y);
    }
  }
}

__global__ void __launch_bounds__(kScaleTC) scale_kernel(ScaleArgs A) {
  extern __shared__ __align__(16) u64 smem[];
  const ScalerDev& S = A.S;
  const u32 nf = S.n_from, n_out = A.n_out;
  const u32 n_out4 = (n_out + 3) & ~3u;
  constexpr u32 TC = kScaleTC;
  u64* s_r = smem;                                // [n_from][TC]
  u64* s_omega = s_r + (size_t)nf * TC;           // [n_from][n_out4]  (transposed)
  u64* s_gamma = s_omega + (size_t)nf * n_out4;   // [n_out4]
  u64* s_tgl = s_gamma + n_out4;                  // theta tables [n_from]
  u64* s_tgh = s_tgl + nf;
  u64* s_tol = s_tgh + nf;
  u64* s_toh = s_tol + nf;
  u32* s_ord = reinterpret_cast<u32*>(s_toh + nf);   // [n_from] term order of the w sum

  const u32 N = 1u << A.logn;
  const u32 per_poly = N / TC;
  const u32 poly = blockIdx.x / per_poly;
  const u32 c0 = (blockIdx.x % per_poly) * TC;
  const u64* src = A.in + (((size_t)poly * nf) << A.logn) + c0;
  const u32 cc = threadIdx.x;

  // Every thread fetches its own column of the source residues with asynchronous 8-byte copies (all n_from
  // requests in flight at once, no registers held, and no CTA barrier is needed for them: a thread only ever reads
  // back what it copied itself).  ncu r1: the former load-then-store loop left 30% of the warp samples waiting at
  // the barrier behind eight dependent HBM round trips.
  {
    const u32 dst0 = (u32)__cvta_generic_to_shared(s_r + cc);
    for (u32 i = 0; i < nf; i++)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst0 + i * TC * 8u),
                   "l"(src + ((size_t)i << A.logn) + cc)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // the (L2-resident) tables, spread over the whole CTA
  for (u32 idx = cc; idx < nf * n_out4; idx += TC) {
    const u32 ii = idx / n_out4, jj = idx - ii * n_out4;
    s_omega[idx] = jj < n_out ? S.omega[(size_t)(A.start + jj) * nf + ii] : 0;
  }
  for (u32 i = cc; i < n_out4; i += TC) s_gamma[i] = i < n_out ? S.gamma[A.start + i] : 0;
  for (u32 i = cc; i < nf; i += TC) {
    s_tgl[i] = S.tgar_lo[i];
    s_tgh[i] = S.tgar_hi[i];
    s_tol[i] = S.to_lo[i];
    s_toh[i] = S.to_hi[i];
    s_ord[i] = i < S.n_terms ? S.to_order[i] : 0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // v = round(sum_i r_i * theta_garner_i / 2^shift)   (:260-272)
  u128 v;
  {
    AccTheta at;
    at.clear();
#pragma unroll 2
    for (u32 i = 0; i < nf; i++) at.mac(s_r[i * TC + cc], s_tgl[i], s_tgh[i]);
    u32 acc[7];
    at.words(acc);
    U256 sg = u256_from_acc(acc);
    // theta_garner_shift is in [123,127] for moduli < 2^62 and <= 64 limbs (:130-142): shift-1 = 64 + bs, 58 <= bs <= 62
    const u32 bs = S.shift - 1 - 64;
    u64 lo = (sg.w1 >> bs) | (sg.w2 << (64 - bs));
    u64 hi = (sg.w2 >> bs) | (sg.w3 << (64 - bs));
    u128 x = ((u128)hi << 64) | lo;
    v = (x >> 1) + (x & 1);
  }
  // w = round((sum_i +/- r_i * theta_omega_i -/+ v * theta_gamma) / 2^127)   (:276-314)
  bool w_sign = false;
  u128 w = 0;
  if (!S.is_one) {
    // one pass per sign keeps a single accumulator set live; the table lists the positive terms first
    U256 s_pos = {0, 0, 0, 0}, s_neg = {0, 0, 0, 0};
#pragma unroll 1
    for (u32 sg = 0; sg < 2; sg++) {
      AccTheta at;
      at.clear();
      const u32 k0 = sg ? S.n_pos : 0, k1 = sg ? S.n_terms : S.n_pos;
#pragma unroll 2
      for (u32 k = k0; k < k1; k++) {
        const u32 i = s_ord[k];
        at.mac(s_r[i * TC + cc], s_tol[i], s_toh[i]);
      }
      u32 wds[7];
      at.words(wds);
      if (sg == 0) s_pos = u256_from_acc(wds);
      else s_neg = u256_from_acc(wds);
    }
    U256 so = u256_sub(s_pos, s_neg);
    U256 vt = u256_mul_128(v, S.tg_lo, S.tg_hi);
    so = S.tg_sign ? u256_add(so, vt) : u256_sub(so, vt);
    w_sign = (so.w3 != 0) || (so.w2 >> 63);
    if (w_sign) {
      u64 n1 = ~so.w1, n2 = ~so.w2, n3 = ~so.w3;
      u128 y = ((u128)((n2 >> 62) | (n3 << 2)) << 64) | ((n1 >> 62) | (n2 << 2));
      w = (y + 1) >> 1;
    } else {
      u128 y = ((u128)((so.w2 >> 62) | (so.w3 << 2)) << 64) | ((so.w1 >> 62) | (so.w2 << 2));
      w = (y >> 1) + (y & 1);
    }
  }

  // outputs (:316-351): y_j = (-(v mod q_j) * gamma_j +/- w + sum_i r_i * omega_ji) mod q_j, four limbs at a time
  for (u32 j0 = 0; j0 < n_out; j0 += 4) {
    Acc192 acc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k].clear();
    const ulonglong2* om = reinterpret_cast<const ulonglong2*>(s_omega + j0);
    // the multiplier pipe bounds this loop (bench_micro/mac_bench.cu), so the last group only multiplies for the
    // limbs it really has (14 outputs = 4+4+4+2, not 16)
    switch (min(4u, n_out - j0)) {
      case 4: scale_mac_group<4>(acc, s_r + cc, om, nf, n_out4 / 2); break;
      case 3: scale_mac_group<3>(acc, s_r + cc, om, nf, n_out4 / 2); break;
      case 2: scale_mac_group<2>(acc, s_r + cc, om, nf, n_out4 / 2); break;
      default: scale_mac_group<1>(acc, s_r + cc, om, nf, n_out4 / 2); break;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 jj = j0 + k;
      if (jj >= n_out) break;
      const LimbDev& M = A.limbs[S.to_ids[A.start + jj]];
      u64 vr = reduce94_limb((u64)v, (u64)(v >> 64), M);   // v < n_from * 2^63
      acc[k].mac(vr ? M.p - vr : 0, s_gamma[jj]);
      if (!S.is_one) {
        u64 wr = reduce94_limb((u64)w, (u64)(w >> 64), M); // w < 2^70
        acc[k].add64(w_sign ? (wr ? M.p - wr : 0) : wr);
      }
      u64 y = acc[k].reduce(M);
      u64* dst;
      if (A.split3) {
        u32 ct = poly / 3, part = poly % 3;
        dst = part < 2 ? A.out0 + ((((size_t)ct * 2 + part) * n_out + jj) << A.logn)
                       : A.out1 + (((size_t)ct * n_out + jj) << A.logn);
      } else {
        dst = A.out0 + (((size_t)poly * A.out_rows_per_poly + jj) << A.logn);
      }
      dst[c0 + cc] = y;
    }
  }
}


// ------------------------------------------------------------------ exact RNS scaler, persistent TMA-fed form
// Same arithmetic as scale_kernel (RnsScaler::scale, rns/scaler.rs:249-352: v and w "as coded", one lazy accumulator
// and one reduction per output limb); what changes is everything around the multiply loop, which ran at 0.28 of the
// multiplier-pipe bound while the loop itself ran at 0.87 (profiles/microbench_r1.txt):
//   * persistent CTAs: the scaler tables (omega, gamma, theta_*, per-limb constants) are staged in shared memory once
//     per CTA instead of once per 128-column tile;
//   * the n_from x 128 source residues of a tile arrive with ONE TMA box copy (tensor map over [rows][N], box
//     {128 columns, n_from rows}) tracked by an mbarrier; the next tile's copy is issued as soon as the last multiply
//     group has read the current one;
//   * the per-limb epilogue works on Solinas limbs only (q_j = 2^62 - c_j; the caller checks): v and w are folded with
//     2^62 == c_j (one IMAD.WIDE each, no conditional subtraction, no canonical intermediate), -(v mod q_j)*gamma_j
//     enters the accumulator as (2q_j - v')*gamma_j, +/-w as one 64-bit addend, and the per-limb constants come from
//     shared memory (the old epilogue fetched the LimbDev record from global memory through two dependent loads).
struct ScaleTmaArgs {
  ScalerDev S;
  const LimbDev* limbs;
  u64 *out0, *out1;
  u32 polys, out_rows_per_poly, start, n_out, split3, logn;
  u32 tiles_total;   // polys * N / 128
};

template <bool IS_ONE, int UNR>
__global__ void __launch_bounds__(kScaleTC) scale_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const ScaleTmaArgs A) {
  using namespace tma;
  extern __shared__ __align__(128) u64 smem[];
  const ScalerDev& S = A.S;
  const u32 nf = S.n_from, n_out = A.n_out;
  const u32 n_out4 = (n_out + 3) & ~3u;
  constexpr u32 TC = kScaleTC;
  u64* s_r = smem;                                // [n_from][TC]   (TMA destination, 128-byte aligned)
  u64* s_omega = s_r + (size_t)nf * TC;           // [n_from][n_out4]  (transposed)
  u64* s_gamma = s_omega + (size_t)nf * n_out4;   // [n_out4]
  u64* s_p2 = s_gamma + n_out4;                   // [n_out4]  2 q_j
  u64* s_c = s_p2 + n_out4;                       // [n_out4]  c_j = 2^62 - q_j
  u64* s_tgl = s_c + n_out4;                      // theta_garner [n_from]
  u64* s_tgh = s_tgl + nf;
  u64* s_tol = s_tgh + nf;                        // theta_omega of the non-zero terms, positive sign first [n_terms]
  u64* s_toh = s_tol + nf;
  u32* s_ord = reinterpret_cast<u32*>(s_toh + nf);   // their source rows, as byte offsets into s_r
  u64* s_bar = reinterpret_cast<u64*>(s_ord + ((nf + 1) & ~1u));
  const u32 bar = smem_u32(s_bar);
  const u32 cc = threadIdx.x;

  for (u32 idx = cc; idx < nf * n_out4; idx += TC) {
    const u32 ii = idx / n_out4, jj = idx - ii * n_out4;
    s_omega[idx] = jj < n_out ? S.omega[(size_t)(A.start + jj) * nf + ii] : 0;
  }
  for (u32 i = cc; i < n_out4; i += TC) {
    const bool live = i < n_out;
    const LimbDev& M = A.limbs[S.to_ids[A.start + (live ? i : 0)]];
    s_gamma[i] = live ? S.gamma[A.start + i] : 0;
    s_p2[i] = M.p2;
    s_c[i] = M.sol_c;
  }
  for (u32 i = cc; i < nf; i += TC) {
    s_tgl[i] = S.tgar_lo[i];
    s_tgh[i] = S.tgar_hi[i];
    const u32 src = (!IS_ONE && i < S.n_terms) ? S.to_order[i] : 0;
    s_tol[i] = IS_ONE ? 0 : S.to_lo[src];
    s_toh[i] = IS_ONE ? 0 : S.to_hi[src];
    s_ord[i] = src * TC * 8;
  }
  if (cc == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const u32 N = 1u << A.logn;
  const u32 per_poly = N / TC;
  const u32 dst_r = smem_u32(s_r);
  const u32 tile_bytes = nf * TC * 8;
  u32 tile = blockIdx.x;
  if (cc == 0 && tile < A.tiles_total) {
    mbar_expect_tx(bar, tile_bytes);
    load_2d(dst_r, &tm_in, (tile % per_poly) * TC, (tile / per_poly) * nf, bar);
  }
  const u32 my_r = dst_r + cc * 8;   // this thread's column of the tile
  for (u32 it = 0; tile < A.tiles_total; tile += gridDim.x, it++) {
    const u32 poly = tile / per_poly;
    const u32 c0 = (tile - poly * per_poly) * TC;
    mbar_wait(bar, it & 1);

    // v = round(sum_i r_i * theta_garner_i / 2^shift)   (:260-272)
    u128 v;
    {
      AccTheta at;
      at.clear();
#pragma unroll 2
      for (u32 i = 0; i < nf; i++) at.mac(lds64(my_r + i * TC * 8), s_tgl[i], s_tgh[i]);
      u32 acc[7];
      at.words(acc);
      U256 sg = u256_from_acc(acc);
      const u32 bs = S.shift - 1 - 64;
      u64 lo = (sg.w1 >> bs) | (sg.w2 << (64 - bs));
      u64 hi = (sg.w2 >> bs) | (sg.w3 << (64 - bs));
      u128 x = ((u128)hi << 64) | lo;
      v = (x >> 1) + (x & 1);
    }
    // w = round((sum_i +/- r_i * theta_omega_i -/+ v * theta_gamma) / 2^127)   (:276-314)
    bool w_sign = false;
    u128 w = 0;
    if (!IS_ONE) {
      U256 s_pos = {0, 0, 0, 0}, s_neg = {0, 0, 0, 0};
#pragma unroll 1
      for (u32 sg = 0; sg < 2; sg++) {
        AccTheta at;
        at.clear();
        const u32 k0 = sg ? S.n_pos : 0, k1 = sg ? S.n_terms : S.n_pos;
#pragma unroll 2
        for (u32 k = k0; k < k1; k++) at.mac(lds64(my_r + s_ord[k]), s_tol[k], s_toh[k]);
        u32 wds[7];
        at.words(wds);
        if (sg == 0) s_pos = u256_from_acc(wds);
        else s_neg = u256_from_acc(wds);
      }
      U256 so = u256_sub(s_pos, s_neg);
      U256 vt = u256_mul_128(v, S.tg_lo, S.tg_hi);
      so = S.tg_sign ? u256_add(so, vt) : u256_sub(so, vt);
      w_sign = (so.w3 != 0) || (so.w2 >> 63);
      if (w_sign) {
        u64 n1 = ~so.w1, n2 = ~so.w2, n3 = ~so.w3;
        u128 y = ((u128)((n2 >> 62) | (n3 << 2)) << 64) | ((n1 >> 62) | (n2 << 2));
        w = (y + 1) >> 1;
      } else {
        u128 y = ((u128)((so.w2 >> 62) | (so.w3 << 2)) << 64) | ((so.w1 >> 62) | (so.w2 << 2));
        w = (y >> 1) + (y & 1);
      }
    }
    // v = vh * 2^62 + vl, w likewise: modulo q_j = 2^62 - c_j they are vh * c_j + vl < 2 q_j
    const u64 mask62 = (1ull << 62) - 1;
    const u64 vl = (u64)v & mask62, wl = (u64)w & mask62;
    const u32 vh = (u32)(v >> 62), wh = (u32)(w >> 62);

    // destination of output limb 0 of this column
    u64* dst;
    size_t dstride = (size_t)1 << A.logn;
    if (A.split3) {
      const u32 ct = poly / 3, part = poly - ct * 3;
      dst = part < 2 ? A.out0 + ((((size_t)ct * 2 + part) * n_out) << A.logn) : A.out1 + (((size_t)ct * n_out) << A.logn);
    } else {
      dst = A.out0 + (((size_t)poly * A.out_rows_per_poly) << A.logn);
    }
    dst += c0 + cc;

    // outputs (:316-351): y_j = (-(v mod q_j) * gamma_j +/- w + sum_i r_i * omega_ji) mod q_j, four limbs at a time
    for (u32 j0 = 0; j0 < n_out; j0 += 4) {
      Acc192 acc[4];
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k].clear();
      const ulonglong2* om = reinterpret_cast<const ulonglong2*>(s_omega + j0);
      const u32 g = min(4u, n_out - j0);
      switch (g) {
        case 4: scale_mac_group<4, UNR>(acc, s_r + cc, om, nf, n_out4 / 2); break;
        case 3: scale_mac_group<3, UNR>(acc, s_r + cc, om, nf, n_out4 / 2); break;
        case 2: scale_mac_group<2, UNR>(acc, s_r + cc, om, nf, n_out4 / 2); break;
        default: scale_mac_group<1, UNR>(acc, s_r + cc, om, nf, n_out4 / 2); break;
      }
      if (j0 + 4 >= n_out) {
        // the tile has been read for the last time: fetch the next one while the last epilogue runs
        __syncthreads();
        const u32 nxt = tile + gridDim.x;
        if (cc == 0 && nxt < A.tiles_total) {
          mbar_expect_tx(bar, tile_bytes);
          load_2d(dst_r, &tm_in, (nxt % per_poly) * TC, (nxt / per_poly) * nf, bar);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const u32 jj = j0 + k;
        if (jj >= n_out) break;
        const u64 p2 = s_p2[jj];
        const u32 c = (u32)s_c[jj];
        acc[k].mac(p2 - ((u64)vh * c + vl), s_gamma[jj]);      // -(v mod q) * gamma, as a positive multiple
        if (!IS_ONE) {
          const u64 wr = (u64)wh * c + wl;                     // w mod q in [0, 2q)
          acc[k].add64(w_sign ? p2 - wr : wr);
        }
        u64 lo, mid;
        u32 hi32;
        acc[k].merged(lo, mid, hi32);
        dst[(size_t)jj * dstride] = csub(fold192_solinas(lo, mid, hi32, c), p2 >> 1);
      }
    }
  }
}

// fallback for rings smaller than one tile (N < 64): one thread per coefficient, same arithmetic
__global__ void scale_small_kernel(ScaleArgs A) {
  const ScalerDev& S = A.S;
  const u32 nf = S.n_from, n_out = A.n_out;
  const u32 N = 1u << A.logn;
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= A.polys * N) return;
  const u32 poly = idx >> A.logn, c = idx & (N - 1);
  const u64* src = A.in + (((size_t)poly * nf) << A.logn) + c;
  u32 av[7] = {0, 0, 0, 0, 0, 0, 0}, ap[7] = {0, 0, 0, 0, 0, 0, 0}, an[7] = {0, 0, 0, 0, 0, 0, 0};
  for (u32 i = 0; i < nf; i++) {
    const u64 r = src[(size_t)i << A.logn];
    mac_theta(av, r, S.tgar_lo[i], S.tgar_hi[i]);
    if (!S.is_one) {
      if (S.to_sign[i]) mac_theta(an, r, S.to_lo[i], S.to_hi[i]);
      else mac_theta(ap, r, S.to_lo[i], S.to_hi[i]);
    }
  }
  U256 sg = u256_from_acc(av);
  const u32 bs = S.shift - 1 - 64;
  u64 lo = (sg.w1 >> bs) | (sg.w2 << (64 - bs));
  u64 hi = (sg.w2 >> bs) | (sg.w3 << (64 - bs));
  u128 x = ((u128)hi << 64) | lo;
  u128 v = (x >> 1) + (x & 1);
  bool w_sign = false;
  u128 w = 0;
  if (!S.is_one) {
    U256 so = u256_sub(u256_from_acc(ap), u256_from_acc(an));
    U256 vt = u256_mul_128(v, S.tg_lo, S.tg_hi);
    so = S.tg_sign ? u256_add(so, vt) : u256_sub(so, vt);
    w_sign = (so.w3 != 0) || (so.w2 >> 63);
    if (w_sign) {
      u64 n1 = ~so.w1, n2 = ~so.w2, n3 = ~so.w3;
      u128 y = ((u128)((n2 >> 62) | (n3 << 2)) << 64) | ((n1 >> 62) | (n2 << 2));
      w = (y + 1) >> 1;
    } else {
      u128 y = ((u128)((so.w2 >> 62) | (so.w3 << 2)) << 64) | ((so.w1 >> 62) | (so.w2 << 2));
      w = (y >> 1) + (y & 1);
    }
  }
  for (u32 j = 0; j < n_out; j++) {
    const LimbDev& M = A.limbs[S.to_ids[A.start + j]];
    Acc192 acc;
    acc.clear();
    for (u32 i = 0; i < nf; i++) acc.mac(src[(size_t)i << A.logn], S.omega[(size_t)(A.start + j) * nf + i]);
    u64 vr = reduce128_limb((u64)v, (u64)(v >> 64), M);
    acc.mac(vr ? M.p - vr : 0, S.gamma[A.start + j]);
    if (!S.is_one) {
      u64 wr = reduce128_limb((u64)w, (u64)(w >> 64), M);
      acc.add64(w_sign ? (wr ? M.p - wr : 0) : wr);
    }
    u64 y = acc.reduce(M);
    u64* dst;
    if (A.split3) {
      u32 ct = poly / 3, part = poly % 3;
      dst = part < 2 ? A.out0 + ((((size_t)ct * 2 + part) * n_out + j) << A.logn)
                     : A.out1 + (((size_t)ct * n_out + j) << A.logn);
    } else {
      dst = A.out0 + (((size_t)poly * A.out_rows_per_poly + j) << A.logn);
    }
    dst[c] = y;
  }
}

// ------------------------------------------------------------------ key-switch MAC
struct KsMacArgs {
  const u64 *inter, *k0, *k1, *base0, *base1;
  u64 *out0, *out1;
  u32 cts, n_dig, Lk, out_ct_rows, logn;
  u32 adjacent;   // digit rows of one (ciphertext, limb) adjacent: inter is [ct][limb][digit][N], else [ct][digit][limb][N]
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
// out0 = base0 + sum_i t_i * k0_i ; out1 = base1 + sum_i t_i * k1_i   (key_switching_key.rs:256-268)
// One thread per (limb j, ciphertext, coefficient), limb-major: consecutive CTAs work on the same key limb for
// every ciphertext of the chunk, so the 2 x n_dig key rows of that limb (7 MB at set C, stored limb-major
// [limb][digit][N]) stay in L2 while the digit rows stream through -- each key word leaves HBM once per chunk instead
// of once per ciphertext.
__global__ void ksmac_kernel(KsMacArgs A) {
  const u32 N = 1u << A.logn;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over Lk*cts*N
  size_t total = ((size_t)A.cts * A.Lk) << A.logn;
  if (idx >= total) return;
  u32 c = idx & (N - 1);
  size_t row = idx >> A.logn;
  u32 ct = row % A.cts, j = row / A.cts;
  const LimbDev& M = A.limbs[A.ids[j]];
  Acc192 a0, a1;
  a0.clear();
  a1.clear();
  const u64* t_ptr = A.adjacent ? A.inter + ((((size_t)ct * A.Lk + j) * A.n_dig) << A.logn) + c
                                : A.inter + ((((size_t)ct * A.n_dig) * A.Lk + j) << A.logn) + c;
  const size_t dstride = A.adjacent ? (size_t)1 << A.logn : (size_t)A.Lk << A.logn;
  const size_t kstride = (size_t)1 << A.logn;
  const u64* k0_ptr = A.k0 + (((size_t)j * A.n_dig) << A.logn) + c;
  const u64* k1_ptr = A.k1 + (((size_t)j * A.n_dig) << A.logn) + c;
  // two digits per trip, the six words of the next trip requested before the multiplies of this one (the kernel
  // is bound by HBM latency, not by the multiplier: 2 x n_dig x 8 IMAD.WIDE per 48 bytes read).  Two
  // coefficients per thread with 16-byte accesses, and four ciphertexts per thread sharing each key word (a third
  // of the L2 -> SM bytes), both measured the same or slower (profiles/microbench_r1.txt)
  u32 i = 0;
  u64 t0 = 0, t1 = 0, x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  if (A.n_dig >= 2) {
    t0 = t_ptr[0], t1 = t_ptr[dstride];
    x0 = __ldg(k0_ptr), x1 = __ldg(k0_ptr + kstride);
    y0 = __ldg(k1_ptr), y1 = __ldg(k1_ptr + kstride);
  }
  for (; i + 2 <= A.n_dig; i += 2) {
    const u64 ct0 = t0, ct1 = t1, cx0 = x0, cx1 = x1, cy0 = y0, cy1 = y1;
    if (i + 4 <= A.n_dig) {
      t0 = t_ptr[(size_t)(i + 2) * dstride], t1 = t_ptr[(size_t)(i + 3) * dstride];
      x0 = __ldg(k0_ptr + (size_t)(i + 2) * kstride), x1 = __ldg(k0_ptr + (size_t)(i + 3) * kstride);
      y0 = __ldg(k1_ptr + (size_t)(i + 2) * kstride), y1 = __ldg(k1_ptr + (size_t)(i + 3) * kstride);
    }
    a0.mac(ct0, cx0);
    a1.mac(ct0, cy0);
    a0.mac(ct1, cx1);
    a1.mac(ct1, cy1);
  }
  if (i < A.n_dig) {
    const u64 tl = t_ptr[(size_t)i * dstride];
    a0.mac(tl, __ldg(k0_ptr + (size_t)i * kstride));
    a1.mac(tl, __ldg(k1_ptr + (size_t)i * kstride));
  }
  const size_t o = (((size_t)ct * A.out_ct_rows + j) << A.logn) + c;
  if (A.base0) a0.add64(A.base0[o]);
  if (A.base1) a1.add64(A.base1[o]);
  A.out0[o] = a0.reduce(M);
  A.out1[o] = a1.reduce(M);
}


// ------------------------------------------------------------------ key-switch MAC, persistent TMA-fed form
// The same sums as ksmac_kernel for the digit-adjacent layout (inter [ct][limb][digit][N], keys [limb][digit][N]).
// The inner product is HBM-bound (14 digit words streamed per pair of outputs) but the per-thread form was bound by
// load latency (~3.1 TB/s of the 6.5 TB/s copy rate): here one elected thread streams every operand with TMA box
// copies and the compute threads only read shared memory.
//   work item = (limb j, 128-coefficient tile tau, ciphertext ct), ct innermost: the two key tiles of (j, tau)
//   ({128, n_dig} boxes, 2 x n_dig KiB) are fetched once and stay in shared memory for every ciphertext of the CTA's
//   range; the digit tile of each ciphertext ({128, n_dig} box: its n_dig rows are adjacent) arrives through a ring of
//   KS_STAGES buffers, refilled as soon as the CTA has consumed it.
constexpr int kKsTC = 128;
struct KsTmaArgs {
  const u64 *base0, *base1;
  u64 *out0, *out1;
  u32 cts, n_dig, Lk, out_ct_rows, logn;
  u32 items_total;   // Lk * (N / 128) * cts, item = (j * tiles + tau) * cts + ct
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};

template <int KS_STAGES>
__global__ void __launch_bounds__(kKsTC) ksmac_tma_kernel(const __grid_constant__ CUtensorMap tm_t,
                                                          const __grid_constant__ CUtensorMap tm_k0,
                                                          const __grid_constant__ CUtensorMap tm_k1, const KsTmaArgs A) {
  using namespace tma;
  extern __shared__ __align__(128) u64 smem[];
  constexpr u32 TC = kKsTC, S = KS_STAGES;
  const u32 nd = A.n_dig;
  const u32 box_bytes = nd * TC * 8;
  u64* s_k0 = smem;                       // [n_dig][TC]
  u64* s_k1 = s_k0 + (size_t)nd * TC;
  u64* s_t = s_k1 + (size_t)nd * TC;      // [S][n_dig][TC]
  u64* s_bar = s_t + (size_t)S * nd * TC; // S full barriers + 1 key barrier
  const u32 bar_full = smem_u32(s_bar), bar_key = bar_full + 8 * S;
  const u32 cc = threadIdx.x;
  if (cc == 0) {
    for (u32 s = 0; s < S; s++) mbar_init(bar_full + 8 * s, 1);
    mbar_init(bar_key, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const u32 tiles = (1u << A.logn) / TC;
  const u32 lo = (u32)(((u64)A.items_total * blockIdx.x) / gridDim.x);
  const u32 hi = (u32)(((u64)A.items_total * (blockIdx.x + 1)) / gridDim.x);
  const u32 n = hi - lo;
  // item -> (jt, ct) walkers: `wl` for the loads thread 0 issues ahead, `w` for the item being computed
  TileWalk wl, w;
  wl.init(lo, A.cts);
  w.init(lo, A.cts);
  u32 loaded = 0;
  auto load_next = [&]() {   // thread 0 only
    const u32 s = loaded % S;
    const u32 j = wl.jt / tiles, tau = wl.jt - j * tiles;
    mbar_expect_tx(bar_full + 8 * s, box_bytes);
    load_2d(smem_u32(s_t + (size_t)s * nd * TC), &tm_t, tau * TC, (wl.p * A.Lk + j) * nd, bar_full + 8 * s);
    wl.next();
    loaded++;
  };
  if (cc == 0)
    while (loaded < n && loaded < S) load_next();

  u32 cur_jt = 0xffffffffu, key_phase = 0;
  const LimbDev* Mp = A.limbs;
  for (u32 i = 0; i < n; i++) {
    if (w.jt != cur_jt) {
      // new (limb, tile): its two key tiles replace the previous ones (every thread has finished with those: the
      // item loop ends with a CTA barrier)
      cur_jt = w.jt;
      const u32 j = cur_jt / tiles, tau = cur_jt - j * tiles;
      Mp = A.limbs + A.ids[j];
      if (cc == 0) {
        mbar_expect_tx(bar_key, 2 * box_bytes);
        load_2d(smem_u32(s_k0), &tm_k0, tau * TC, j * nd, bar_key);
        load_2d(smem_u32(s_k1), &tm_k1, tau * TC, j * nd, bar_key);
      }
      mbar_wait(bar_key, key_phase);
      key_phase ^= 1;
    }
    const u32 j = cur_jt / tiles, tau = cur_jt - j * tiles;
    const u32 s = i % S;
    const size_t o = (((size_t)w.p * A.out_ct_rows + j) << A.logn) + tau * TC + cc;
    u64 b0 = 0, b1 = 0;
    if (A.base0) b0 = A.base0[o];
    if (A.base1) b1 = A.base1[o];
    mbar_wait(bar_full + 8 * s, (i / S) & 1);
    const u64* t = s_t + (size_t)s * nd * TC + cc;
    Acc192 a0, a1;
    a0.clear();
    a1.clear();
#pragma unroll 2
    for (u32 d = 0; d < nd; d++) {
      const u64 td = t[d * TC];
      a0.mac(td, s_k0[d * TC + cc]);
      a1.mac(td, s_k1[d * TC + cc]);
    }
    a0.add64(b0);
    a1.add64(b1);
    A.out0[o] = a0.reduce(*Mp);
    A.out1[o] = a1.reduce(*Mp);
    __syncthreads();                       // stage s (and, before a key change, the key tiles) are free again
    if (cc == 0 && loaded < n) load_next();
    w.next();
  }
}

// base-2^log_base digits of a single-limb power-basis polynomial (key_switching_key.rs:339-345):
// out[poly][d][:] = (in[poly][:] >> (d * log_base)) & (2^log_base - 1)
__global__ void decompose_kernel(const u64* in, u64* out, size_t n_words, u32 n_dig, u32 log_base, u32 logn) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const size_t poly = i >> logn;
  const u32 c = i & ((1u << logn) - 1);
  u64 v = in[i];
  const u64 mask = (1ull << log_base) - 1;
  for (u32 d = 0; d < n_dig; d++) {
    out[((poly * n_dig + d) << logn) + c] = v & mask;
    v >>= log_base;
  }
}

// ------------------------------------------------------------------ gather / switch_down
__global__ void gather_kernel(const u64* in, u64* out, size_t n_words, const int* perm, u32 logn) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  size_t row = i >> logn;
  u32 t = i & ((1u << logn) - 1);
  out[i] = in[(row << logn) + perm[t]];
}

// Poly::substitute, PowerBasis branch (rq/mod.rs:390-408): coefficient j of x^j moves to x^(j*e mod 2N), i.e. to slot
// (j*e) & (N-1), negated when bit N of j*e is set (x^N = -1).  e is odd, so the map is a bijection: a scatter in which
// every output word is written exactly once (reads coalesced, writes strided by e).
struct SubstPowerArgs {
  const u64* in;
  u64* out;
  size_t n_words;
  u32 logn, limbs_per_poly, exponent;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
__global__ void substitute_power_kernel(SubstPowerArgs A) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_words) return;
  const u32 N = 1u << A.logn;
  const size_t row = i >> A.logn;
  const u32 j = (u32)i & (N - 1);
  const u64 p = A.limbs[A.ids[row % A.limbs_per_poly]].p;
  const u32 power = j * A.exponent;          // mod 2^32 keeps the low logn+1 bits exact
  const u64 v = A.in[i];
  A.out[(row << A.logn) + (power & (N - 1))] = (power & N) ? csub(p - v, p) : v;   // Modulus::sub(0, v) / add(0, v)
}

struct SwitchDownArgs {
  SwitchDownDev S;
  const u64* in;
  u64* out;
  u32 polys, L, logn;
  const LimbDev* limbs;
  unsigned short ids[kMaxPos];
};
__global__ void switch_down_kernel(SwitchDownArgs A) {
  const u32 N = 1u << A.logn;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over polys*N
  if (idx >= ((size_t)A.polys << A.logn)) return;
  u32 c = idx & (N - 1);
  size_t poly = idx >> A.logn;
  const u64* src = A.in + ((poly * A.L) << A.logn) + c;
  u64* dst = A.out + ((poly * (A.L - 1)) << A.logn) + c;
  u64 xl = csub(src[(size_t)(A.L - 1) << A.logn] + A.S.q_last_half, A.S.q_last);  // rq/mod.rs:456-458
  for (u32 i = 0; i + 1 < A.L; i++) {
    const LimbDev& M = A.limbs[A.ids[i]];
    u64 tmp = barrett64(xl, M.p, M.bhi, M.blo) + A.S.half_mod[i];   // :469
    u64 v = src[(size_t)i << A.logn] + 3 * M.p - tmp;              // :473
    dst[(size_t)i << A.logn] = mul_shoup(v, A.S.inv[i], A.S.inv_s[i], M.p);  // :476 (always a Shoup pair)
  }
}

// ------------------------------------------------------------------ wire-format bit packing
// transcode_to_bytes / transcode_from_bytes (fhe-util/src/lib.rs:71-146): eight coefficients of nbits bits are
// exactly nbits bytes, so one thread converts one 8-coefficient group.
__global__ void pack_kernel(PackDev P, const u64* words, unsigned char* bytes, size_t n_groups, u32 logn) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const u32 gpr = (1u << logn) >> 3;            // groups per row
  const size_t row = g / gpr;
  const u32 k = (u32)(g % gpr), limb = (u32)(row % P.limbs);
  const u32 nb = P.nbits[limb];
  const u64* src = words + (row << logn) + ((size_t)k << 3);
  unsigned char* dst = bytes + (row / P.limbs) * P.poly_bytes + P.offs[limb] + (size_t)k * nb;
  const u64 mask = nb == 64 ? ~0ull : ((1ull << nb) - 1);
  unsigned __int128 cur = 0;
  u32 have = 0, o = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    cur |= (unsigned __int128)(src[e] & mask) << have;
    have += nb;
    while (have >= 8) {
      dst[o++] = (unsigned char)cur;
      cur >>= 8;
      have -= 8;
    }
  }
}
__global__ void unpack_kernel(PackDev P, const unsigned char* bytes, u64* words, size_t n_groups, u32 logn) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const u32 gpr = (1u << logn) >> 3;
  const size_t row = g / gpr;
  const u32 k = (u32)(g % gpr), limb = (u32)(row % P.limbs);
  const u32 nb = P.nbits[limb];
  const unsigned char* src = bytes + (row / P.limbs) * P.poly_bytes + P.offs[limb] + (size_t)k * nb;
  u64* dst = words + (row << logn) + ((size_t)k << 3);
  const u64 mask = nb == 64 ? ~0ull : ((1ull << nb) - 1);
  unsigned __int128 cur = 0;
  u32 have = 0, o = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    while (have < nb) {
      cur |= (unsigned __int128)src[o++] << have;
      have += 8;
    }
    dst[e] = (u64)cur & mask;
    cur >>= nb;
    have -= nb;
  }
}

void copy_ids(unsigned short* dst, const RowIds& ids) {
  for (int i = 0; i < kMaxPos; i++) dst[i] = ids.ids[i];
}

}  // namespace

void launch_ew(EwOp op, u64* a, const u64* b, size_t n_rows, const RowIds& ids, const LimbDev* limbs, u32 logn,
               cudaStream_t st) {
  EwArgs A;
  A.a = a;
  A.b = b;
  A.n_words = n_rows << logn;
  A.logn = logn;
  A.limbs_per_poly = ids.limbs_per_poly;
  A.limbs = limbs;
  copy_ids(A.ids, ids);
  if (A.n_words == 0) return;
  const u32 threads = 256;
  const size_t blocks = (A.n_words / 2 + threads - 1) / threads;
  if (op == EW_ADD) ew_kernel<EW_ADD><<<(unsigned)blocks, threads, 0, st>>>(A);
  else if (op == EW_SUB) ew_kernel<EW_SUB><<<(unsigned)blocks, threads, 0, st>>>(A);
  else ew_kernel<EW_NEG><<<(unsigned)blocks, threads, 0, st>>>(A);
  g_launches++;
}

void launch_mul_plain(u64* a, const u64* pt, u32 cts, u32 parts, u32 n_pt, const RowIds& ids, const LimbDev* limbs,
                      u32 logn, cudaStream_t st, u32 op) {
  MulPlainArgs A;
  A.a = a; A.pt = pt; A.cts = cts; A.parts = parts; A.n_pt = n_pt; A.logn = logn; A.op = op;
  A.limbs_per_poly = ids.limbs_per_poly; A.limbs = limbs;
  copy_ids(A.ids, ids);
  size_t total = ((size_t)cts * parts * ids.limbs_per_poly) << logn;
  if (!total) return;
  mul_plain_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

void launch_dot(const u64* ct, const u64* pt, u64* out, u32 groups, u32 n_terms, u32 parts, u32 ct_count,
                u32 pt_count, const RowIds& ids, const LimbDev* limbs, u32 logn, cudaStream_t st) {
  DotArgs A;
  A.ct = ct; A.pt = pt; A.out = out; A.groups = groups; A.n_terms = n_terms; A.parts = parts;
  A.ct_count = ct_count; A.pt_count = pt_count; A.logn = logn;
  A.limbs_per_poly = ids.limbs_per_poly; A.limbs = limbs;
  copy_ids(A.ids, ids);
  size_t total = ((size_t)groups * parts * ids.limbs_per_poly) << logn;
  if (!total) return;
  dot_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

void launch_tensor(const u64* a, const u64* b, const u64* xa, const u64* xb, u64* out, u32 cts, u32 L, u32 nca,
                   u32 ncb, u32 K, const RowIds& mul_ids, const LimbDev* limbs, u32 logn, cudaStream_t st) {
  TensorArgs A;
  A.a = a; A.b = b; A.xa = xa; A.xb = xb; A.out = out;
  A.cts = cts; A.L = L; A.nca = nca; A.ncb = ncb; A.K = K; A.logn = logn;
  A.limbs = limbs;
  copy_ids(A.ids, mul_ids);
  size_t total = ((size_t)cts * K) << logn;
  if (!total) return;
  tensor_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

void launch_tensor_nm(const u64* a, const u64* b, const u64* xa, const u64* xb, u64* out, u32 cts, u32 L, u32 E, u32 na,
                      u32 nb, const RowIds& mul_ids, const LimbDev* limbs, u32 logn, cudaStream_t st) {
  TensorNmArgs A;
  A.a = a; A.b = b; A.xa = xa; A.xb = xb; A.out = out;
  A.cts = cts; A.L = L; A.E = E; A.na = na; A.nb = nb; A.logn = logn;
  A.limbs = limbs;
  copy_ids(A.ids, mul_ids);
  size_t total = ((size_t)cts * (na + nb - 1) * (L + E)) << logn;
  if (!total) return;
  tensor_nm_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

namespace {
typedef CUresult (*ScaleEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
ScaleEncodeFn scale_encoder() {
  static ScaleEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (ScaleEncodeFn)p;
  }();
  return fn;
}
int scale_sm_count() {
  static std::mutex mu;
  static std::map<int, int> cache;
  int dev = 0;
  FHE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 0;
  FHE_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return cache[dev] = n;
}
}  // namespace

// The persistent TMA-fed kernel serves N >= 128 when every output limb is a Solinas prime (all 62-bit primes the
// reference's parameter builder generates are); FHE_B200_SCALER=classic keeps the per-tile kernel.
static bool launch_scale_tma(const ScalerDev& S, const LimbDev* limbs, const std::vector<u64>* sol_c_of_out, const u64* in,
                             u64* out0, u64* out1, u32 polys, u32 out_rows_per_poly, u32 start, u32 n_out, int split3,
                             u32 logn, cudaStream_t st) {
  (void)sol_c_of_out;
  static const bool classic = [] { const char* e = getenv("FHE_B200_SCALER"); return e && !strcmp(e, "classic"); }();
  if (classic || !scale_encoder() || logn < 7 || (reinterpret_cast<uintptr_t>(in) & 127)) return false;
  const u32 N = 1u << logn, nf = S.n_from;
  if (nf > 64 || (u64)polys * nf >= (1ull << 31)) return false;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {N, (cuuint64_t)polys * nf};
  const cuuint64_t gstride[1] = {(cuuint64_t)8 << logn};
  const cuuint32_t box[2] = {(cuuint32_t)kScaleTC, nf};
  const cuuint32_t es[2] = {1, 1};
  if (scale_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)in, gdim, gstride, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  ScaleTmaArgs A;
  A.S = S; A.limbs = limbs; A.out0 = out0; A.out1 = out1;
  A.polys = polys; A.out_rows_per_poly = out_rows_per_poly; A.start = start; A.n_out = n_out;
  A.split3 = split3; A.logn = logn;
  A.tiles_total = polys * (N / kScaleTC);
  const size_t n_out4 = (n_out + 3) & ~(size_t)3;
  const size_t smem = (nf * kScaleTC + nf * n_out4 + 3 * n_out4 + 4 * nf) * sizeof(u64) + ((nf + 1) & ~(size_t)1) * 4 + 16;
  // persistent grid = exactly the CTAs that are resident at once (registers and shared memory both limit it)
  auto resident = [&](const void* k) {
    ensure_dynamic_smem(k, smem);
    int per_sm = 0;
    FHE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kScaleTC, smem));
    return (u32)std::min<u64>(A.tiles_total, (u64)scale_sm_count() * std::max(per_sm, 1));
  };
  static const int unr = [] { const char* e = getenv("FHE_B200_SCALE_UNROLL"); return e ? atoi(e) : 2; }();
  if (unr == 4) {
    if (S.is_one) scale_tma_kernel<true, 4><<<resident((const void*)scale_tma_kernel<true, 4>), kScaleTC, smem, st>>>(tm, A);
    else scale_tma_kernel<false, 4><<<resident((const void*)scale_tma_kernel<false, 4>), kScaleTC, smem, st>>>(tm, A);
  } else {
    if (S.is_one) scale_tma_kernel<true, 2><<<resident((const void*)scale_tma_kernel<true, 2>), kScaleTC, smem, st>>>(tm, A);
    else scale_tma_kernel<false, 2><<<resident((const void*)scale_tma_kernel<false, 2>), kScaleTC, smem, st>>>(tm, A);
  }
  g_launches++;
  return true;
}

void launch_scale(const ScalerDev& S, const LimbDev* limbs, const u64* in, u64* out0, u64* out1, u32 polys,
                  u32 out_rows_per_poly, u32 start, u32 n_out, int split3, u32 logn, cudaStream_t st) {
  if (!polys || !n_out) return;
  if (S.all_solinas && launch_scale_tma(S, limbs, nullptr, in, out0, out1, polys, out_rows_per_poly, start, n_out, split3,
                                        logn, st))
    return;
  ScaleArgs A;
  A.S = S; A.limbs = limbs; A.in = in; A.out0 = out0; A.out1 = out1;
  A.polys = polys; A.out_rows_per_poly = out_rows_per_poly; A.start = start; A.n_out = n_out;
  A.split3 = split3; A.logn = logn;
  const u32 N = 1u << logn;
  if (N < (u32)kScaleTC) {
    const u32 total = polys * N;
    scale_small_kernel<<<(total + 63) / 64, 64, 0, st>>>(A);
    g_launches++;
    return;
  }
  const size_t n_out4 = (n_out + 3) & ~(size_t)3, nf = S.n_from;
  const size_t smem = (nf * kScaleTC + nf * n_out4 + n_out4 + 5 * nf) * sizeof(u64);
  ensure_dynamic_smem((const void*)scale_kernel, smem);
  scale_kernel<<<polys * (N / kScaleTC), kScaleTC, smem, st>>>(A);
  g_launches++;
}

void launch_ksmac(const u64* inter, const u64* k0, const u64* k1, const u64* base0, const u64* base1, u64* out0,
                  u64* out1, u32 cts, u32 n_dig, u32 Lk, u32 out_ct_rows, const RowIds& ids, const LimbDev* limbs,
                  u32 logn, cudaStream_t st, bool adjacent) {
  size_t total = ((size_t)cts * Lk) << logn;
  if (!total) return;
  static const bool classic = [] { const char* e = getenv("FHE_B200_KSMAC"); return e && !strcmp(e, "classic"); }();
  // ring depth: 2 buffers -> 4 resident CTAs per SM at set C measured best (4333 products/s; 3 buffers / 3 CTAs 4314,
  // 4 buffers / 2 CTAs 4279): the kernel needs warps more than prefetch depth.  FHE_B200_KS_STAGES = 2 | 3 | 4.
  static const int stages = [] { const char* e = getenv("FHE_B200_KS_STAGES"); int v = e ? atoi(e) : 2; return v < 2 ? 2 : v > 4 ? 4 : v; }();
  const size_t smem_tma = ((size_t)(2 + stages) * n_dig * kKsTC + stages + 1) * sizeof(u64);
  if (adjacent && !classic && scale_encoder() && logn >= 7 && n_dig <= 256 && smem_tma <= 200 * 1024 &&
      !((reinterpret_cast<uintptr_t>(inter) | reinterpret_cast<uintptr_t>(k0) | reinterpret_cast<uintptr_t>(k1)) & 127) &&
      (u64)cts * Lk * n_dig < (1ull << 31)) {
    auto map = [&](const u64* base, u64 rows, CUtensorMap* m) {
      const cuuint64_t gdim[2] = {(cuuint64_t)1 << logn, rows};
      const cuuint64_t gstride[1] = {(cuuint64_t)8 << logn};
      const cuuint32_t box[2] = {(cuuint32_t)kKsTC, n_dig};
      const cuuint32_t es[2] = {1, 1};
      return scale_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)base, gdim, gstride, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    CUtensorMap mt, m0, m1;
    if (map(inter, (u64)cts * Lk * n_dig, &mt) && map(k0, (u64)Lk * n_dig, &m0) && map(k1, (u64)Lk * n_dig, &m1)) {
      KsTmaArgs T;
      T.base0 = base0; T.base1 = base1; T.out0 = out0; T.out1 = out1;
      T.cts = cts; T.n_dig = n_dig; T.Lk = Lk; T.out_ct_rows = out_ct_rows; T.logn = logn;
      T.items_total = Lk * ((1u << logn) / kKsTC) * cts;
      T.limbs = limbs;
      copy_ids(T.ids, ids);
      auto go = [&](auto kern) {
        ensure_dynamic_smem((const void*)kern, smem_tma);
        int per_sm = 0;
        FHE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, kKsTC, smem_tma));
        const u32 grid = (u32)std::min<u64>(T.items_total, (u64)scale_sm_count() * std::max(per_sm, 1));
        kern<<<grid, kKsTC, smem_tma, st>>>(mt, m0, m1, T);
      };
      if (stages == 2) go(ksmac_tma_kernel<2>);
      else if (stages == 3) go(ksmac_tma_kernel<3>);
      else go(ksmac_tma_kernel<4>);
      g_launches++;
      return;
    }
  }
  KsMacArgs A;
  A.adjacent = adjacent ? 1 : 0;
  A.inter = inter; A.k0 = k0; A.k1 = k1; A.base0 = base0; A.base1 = base1; A.out0 = out0; A.out1 = out1;
  A.cts = cts; A.n_dig = n_dig; A.Lk = Lk; A.out_ct_rows = out_ct_rows; A.logn = logn;
  A.limbs = limbs;
  copy_ids(A.ids, ids);
  ksmac_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

void launch_decompose(const u64* in, u64* out, size_t polys, u32 n_dig, u32 log_base, u32 logn, cudaStream_t st) {
  const size_t n_words = polys << logn;
  if (!n_words) return;
  decompose_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(in, out, n_words, n_dig, log_base, logn);
  g_launches++;
}

void launch_gather(const u64* in, u64* out, size_t n_rows, const int* perm, u32 logn, cudaStream_t st) {
  size_t n = n_rows << logn;
  if (!n) return;
  gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n, perm, logn);
  g_launches++;
}

void launch_substitute_power(const u64* in, u64* out, size_t n_rows, u32 exponent, const RowIds& ids,
                             const LimbDev* limbs, u32 logn, cudaStream_t st) {
  SubstPowerArgs A;
  A.in = in; A.out = out; A.n_words = n_rows << logn; A.logn = logn; A.limbs_per_poly = ids.limbs_per_poly;
  A.exponent = exponent; A.limbs = limbs;
  copy_ids(A.ids, ids);
  if (!A.n_words) return;
  substitute_power_kernel<<<(unsigned)((A.n_words + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

void launch_pack(const PackDev& P, const u64* words, unsigned char* bytes, size_t n_rows, u32 logn, cudaStream_t st) {
  const size_t groups = n_rows << (logn - 3);
  if (!groups) return;
  pack_kernel<<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(P, words, bytes, groups, logn);
  g_launches++;
}
void launch_unpack(const PackDev& P, const unsigned char* bytes, u64* words, size_t n_rows, u32 logn, cudaStream_t st) {
  const size_t groups = n_rows << (logn - 3);
  if (!groups) return;
  unpack_kernel<<<(unsigned)((groups + 127) / 128), 128, 0, st>>>(P, bytes, words, groups, logn);
  g_launches++;
}

void launch_switch_down(const SwitchDownDev& S, const u64* in, u64* out, u32 polys, u32 L, const RowIds& ids,
                        const LimbDev* limbs, u32 logn, cudaStream_t st) {
  SwitchDownArgs A;
  A.S = S; A.in = in; A.out = out; A.polys = polys; A.L = L; A.logn = logn; A.limbs = limbs;
  copy_ids(A.ids, ids);
  size_t total = (size_t)polys << logn;
  if (!total) return;
  switch_down_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A);
  g_launches++;
}

}  // namespace fhe_b200
