// Host-side one-off precompute for the B200 BFV engine: prime generation, NTT twiddle
// tables, RNS contexts and exact-scaler tables.  Mirrors (and cites) the reference's
// constructors; runs once per parameter set, results are uploaded as flat device tables.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bigint.hpp"

namespace fhe_b200 {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

struct FheError : std::runtime_error {
  int code;
  FheError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline u64 mulmod_h(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
inline u64 powmod_h(u64 a, u64 e, u64 p) {
  u64 r = 1 % p;
  a %= p;
  while (e) {
    if (e & 1) r = mulmod_h(r, a, p);
    a = mulmod_h(a, a, p);
    e >>= 1;
  }
  return r;
}
// fhe-util is_prime (fhe-util/src/lib.rs:16): deterministic Miller-Rabin for 64-bit inputs.
inline bool is_prime_u64(u64 n) {
  if (n < 2) return false;
  static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (u64 q : small)
    if (n % q == 0) return n == q;
  u64 d = n - 1;
  int s = 0;
  while ((d & 1) == 0) { d >>= 1; s++; }
  for (u64 a : small) {
    u64 x = powmod_h(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < s; i++) {
      x = mulmod_h(x, x, n);
      if (x == n - 1) { comp = false; break; }
    }
    if (comp) return false;
  }
  return true;
}
inline int clz64(u64 v) { return v ? __builtin_clzll(v) : 64; }

// zq/primes.rs:30-59
inline bool generate_prime(int num_bits, u64 modulo, u64 upper_bound, u64* out) {
  if (num_bits < 10 || num_bits > 62) return false;
  const int lz = 64 - num_bits;
  u64 t = upper_bound - 1;
  while (t % modulo != 1 && clz64(t) == lz) t--;
  while (clz64(t) == lz && !is_prime_u64(t) && t >= modulo) t -= modulo;
  if (clz64(t) == lz && is_prime_u64(t)) { *out = t; return true; }
  return false;
}

// modular inverse by extended Euclid (moduli of an RnsContext need not be prime, rns/mod.rs:93-96)
inline bool invmod_h(u64 a, u64 m, u64* out) {
  __int128 t = 0, nt = 1, r = m, nr = a % m;
  while (nr != 0) {
    __int128 q = r / nr;
    __int128 tmp = t - q * nt; t = nt; nt = tmp;
    tmp = r - q * nr; r = nr; nr = tmp;
  }
  if (r != 1) return false;
  if (t < 0) t += m;
  *out = (u64)t;
  return true;
}

// Default 2N-th root when the caller passes none (documented rule; the reference's
// ChaCha8-sampled root, ntt/native.rs:320-336, cannot be reproduced without its RNG crates).
inline u64 default_psi(u64 p, u64 n) {
  const u64 lam = (p - 1) / (2 * n);
  for (u64 g = 2;; g++) {
    u64 psi = powmod_h(g, lam, p);
    if (powmod_h(psi, n, p) == p - 1) return psi;
  }
}

struct ModulusH {  // zq::Modulus, zq/mod.rs:83-98
  u64 p = 0, bhi = 0, blo = 0, c128 = 0;
  explicit ModulusH(u64 q = 2) : p(q) {
    if (q < 2 || (q >> 62) != 0) throw FheError(-2, "InvalidModulus(" + std::to_string(q) + ")");
    BigUint b = (BigUint(1) << 128) / BigUint(q);
    u128 v = b.low_u128();
    bhi = (u64)(v >> 64);
    blo = (u64)v;
    c128 = ((BigUint(1) << 128) % BigUint(q)).to_u64();
  }
  u64 shoup(u64 a) const { return (u64)((((u128)a) << 64) / p); }  // zq/mod.rs:195
};

struct NttTablesH {  // ntt::native::NttOperator::new, ntt/native.rs:35-73
  u64 p, psi, ninv, ninv_s, zn, zn_s;
  std::vector<u64> om, om_s, zi, zi_s;
};
inline NttTablesH make_ntt_tables(u64 p, size_t n, u64 psi) {
  if (!(p % (2 * n) == 1 && is_prime_u64(p)))  // supports_ntt, ntt/mod.rs:17-23
    throw FheError(-4, "NttOperatorUnavailable: modulus " + std::to_string(p));
  if (powmod_h(psi, n, p) != p - 1) throw FheError(-4, "psi is not a primitive 2N-th root");
  ModulusH m(p);
  NttTablesH t;
  t.p = p;
  t.psi = psi;
  t.ninv = powmod_h(n % p, p - 2, p);
  t.ninv_s = m.shoup(t.ninv);
  const u64 psi_inv = powmod_h(psi, p - 2, p);
  std::vector<u64> pw(n), pwi(n);
  u64 v = 1;
  for (size_t i = 0; i < n; i++) { pw[i] = v; v = mulmod_h(v, psi, p); }
  v = psi_inv;
  for (size_t i = 0; i < n; i++) { pwi[i] = v; v = mulmod_h(v, psi_inv, p); }
  int logn = __builtin_ctzll(n);
  t.om.resize(n); t.om_s.resize(n); t.zi.resize(n); t.zi_s.resize(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (int b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);
    t.om[i] = pw[j];
    t.zi[i] = pwi[j];
    t.om_s[i] = m.shoup(t.om[i]);
    t.zi_s[i] = m.shoup(t.zi[i]);
  }
  // the last inverse stage (l = N/2, one block) consumes zetas_inv[N-2] (k runs 0..N-2, native.rs:201-227)
  t.zn = mulmod_h(t.zi[n - 2], t.ninv, p);
  t.zn_s = m.shoup(t.zn);
  return t;
}

struct RnsContextH {  // rns::RnsContext::new, rns/mod.rs:52-116
  std::vector<u64> moduli;
  BigUint product;
  std::vector<BigUint> garner;
  explicit RnsContextH(const std::vector<u64>& q) : moduli(q), product(1) {
    if (q.empty()) throw FheError(-2, "EmptyModuli");
    for (u64 m : q) product = product * BigUint(m);
    for (u64 m : q) {
      BigUint q_star = product / BigUint(m);
      u64 q_tilde;
      if (!invmod_h(q_star.mod_u64(m), m, &q_tilde)) throw FheError(-2, "NonCoprimeModuli");
      garner.push_back(q_star * BigUint(q_tilde));
    }
  }
};

struct ScalerTablesH {  // rns::RnsScaler::new, rns/scaler.rs:79-175
  uint32_t n_from = 0, n_to = 0, is_one = 0, shift = 0;
  std::vector<u64> gamma;             // [n_to]
  u64 theta_gamma_lo = 0, theta_gamma_hi = 0;
  uint32_t theta_gamma_sign = 0;
  std::vector<u64> omega;             // [n_to][n_from]
  std::vector<u64> theta_omega_lo, theta_omega_hi;  // [n_from]
  std::vector<uint8_t> theta_omega_sign;            // [n_from]
  std::vector<u64> theta_garner_lo, theta_garner_hi;  // [n_from]
};

// extract_projection_and_theta, rns/scaler.rs:183-229
inline void extract_projection_and_theta(const std::vector<u64>& to, const BigUint& input, const BigUint& num,
                                         const BigUint& den, bool round_up, std::vector<u64>* projected,
                                         u64* lo, u64* hi, bool* sign) {
  BigUint ni = num * input;
  BigUint gamma = (ni + (den >> 1)) / den;
  projected->clear();
  for (u64 m : to) projected->push_back(gamma.mod_u64(m));
  BigUint theta = ni % den;
  bool s = false;
  if (den > BigUint(1)) {
    if (den.is_odd()) {
      if (theta > (den >> 1)) { s = true; theta = den - theta; }
    } else if (theta >= (den >> 1)) { s = true; theta = den - theta; }
  }
  if (round_up) {
    theta = s ? (theta << 127) / den : ((theta << 127) + den - BigUint(1)) / den;
  } else {
    theta = s ? ((theta << 127) + den - BigUint(1)) / den : (theta << 127) / den;
  }
  u128 v = theta.low_u128();
  *lo = (u64)v;
  *hi = (u64)(v >> 64);
  *sign = s;
}

inline ScalerTablesH make_scaler_tables(const RnsContextH& from, const RnsContextH& to, const BigUint& num,
                                        const BigUint& den) {
  if (den.is_zero()) throw FheError(-1, "scaling factor denominator is zero");
  ScalerTablesH t;
  t.n_from = (uint32_t)from.moduli.size();
  t.n_to = (uint32_t)to.moduli.size();
  t.is_one = num == den;
  bool sg;
  extract_projection_and_theta(to.moduli, from.product, num, den, false, &t.gamma, &t.theta_gamma_lo,
                               &t.theta_gamma_hi, &sg);
  t.theta_gamma_sign = sg;
  t.omega.assign((size_t)t.n_to * t.n_from, 0);
  t.theta_omega_lo.resize(t.n_from);
  t.theta_omega_hi.resize(t.n_from);
  t.theta_omega_sign.resize(t.n_from);
  for (uint32_t i = 0; i < t.n_from; i++) {
    std::vector<u64> proj;
    extract_projection_and_theta(to.moduli, from.garner[i], num, den, true, &proj, &t.theta_omega_lo[i],
                                 &t.theta_omega_hi[i], &sg);
    t.theta_omega_sign[i] = sg;
    for (uint32_t j = 0; j < t.n_to; j++) t.omega[(size_t)j * t.n_from + i] = proj[j] % to.moduli[j];
  }
  // theta_garner_shift, rns/scaler.rs:128-142
  uint32_t shift = 127;
  for (u64 qi : from.moduli) {
    u128 v = (u128)qi * t.n_from;
    uint32_t lg = 0;  // next_power_of_two().ilog2()
    while (((u128)1 << lg) < v) lg++;
    uint32_t s = 192 - 1 - lg;
    if (s < shift) shift = s;
  }
  t.shift = shift;
  t.theta_garner_lo.resize(t.n_from);
  t.theta_garner_hi.resize(t.n_from);
  for (uint32_t i = 0; i < t.n_from; i++) {  // rns/scaler.rs:145-155
    BigUint theta = ((from.garner[i] << shift) + (from.product >> 1)) / from.product;
    u128 v = theta.low_u128();
    t.theta_garner_lo[i] = (u64)v;
    t.theta_garner_hi[i] = (u64)(v >> 64);
  }
  return t;
}

}  // namespace fhe_b200
