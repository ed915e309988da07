/*
 * fhe_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, single-threaded C restatement of the fhe-math hot loops of
 * tlepoint/fhe.rs @ e248cd28 (the reference is pure Rust; no Rust toolchain
 * exists in this image, so the reference itself cannot be built -- see
 * DESIGN.md "Oracle").  Every function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libfhe_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status: the arithmetic (zq, ntt, rns scaler, rq ops) is pinned
 * against the reference's own known-answer data and BigUint property oracles
 * (tests/test_oracle_*.py).  The choice of the 2N-th root psi inside
 * NttOperator::new (ntt/native.rs:320-336) depends on rand_chacha internals
 * that are not vendored: psi is therefore an INPUT here ("parity unpinned" for
 * psi only; every NTT-domain value is bit-exact for a given psi).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -shared -fPIC).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

/* zq::Modulus (crates/fhe-math/src/zq/mod.rs:32-40) minus RNG/arch fields. */
typedef struct {
  u64 p;
  u64 barrett_hi;
  u64 barrett_lo;
  u32 leading_zeros;
  u32 supports_opt;
} zq_modulus;

/* ---------------------------------------------------------------- zq ---- */

/* zq/mod.rs:659  reduce1: x in [0,2p) -> x mod p */
static inline u64 reduce1(u64 x, u64 p) { return x >= p ? x - p : x; }

/* zq/mod.rs:693  lazy_reduce_u128 -> [0,2p) */
static inline u64 lazy_reduce_u128(const zq_modulus *m, u128 a) {
  u64 a_lo = (u64)a, a_hi = (u64)(a >> 64);
  u128 p_lo_lo = ((u128)a_lo * m->barrett_lo) >> 64;
  u128 p_hi_lo = (u128)a_hi * m->barrett_lo;
  u128 p_lo_hi = (u128)a_lo * m->barrett_hi;
  u128 q = ((p_lo_hi + p_hi_lo + p_lo_lo) >> 64) + (u128)a_hi * m->barrett_hi;
  return (u64)(a - q * (u128)m->p);
}

/* zq/mod.rs:712  lazy_reduce (u64) -> [0,2p) */
static inline u64 lazy_reduce(const zq_modulus *m, u64 a) {
  u128 p_lo_lo = ((u128)a * m->barrett_lo) >> 64;
  u128 p_lo_hi = (u128)a * m->barrett_hi;
  u128 q = (p_lo_hi + p_lo_lo) >> 64;
  return (u64)((u128)a - q * (u128)m->p);
}

/* zq/mod.rs:730  lazy_reduce_opt_u128 -> [0,2p), requires a < p^2 */
static inline u64 lazy_reduce_opt_u128(const zq_modulus *m, u128 a) {
  u128 q = (((u128)m->barrett_lo * (a >> 64)) + (a << m->leading_zeros)) >> 64;
  return (u64)(a - q * (u128)m->p);
}

/* zq/mod.rs:744  lazy_reduce_opt (u64) -> [0,2p) */
static inline u64 lazy_reduce_opt(const zq_modulus *m, u64 a) {
  u64 q = a >> (64 - m->leading_zeros);
  return (u64)((u128)a - (u128)q * (u128)m->p);
}

/* zq/mod.rs:594 */
static inline u64 reduce_u128(const zq_modulus *m, u128 a) {
  return reduce1(lazy_reduce_u128(m, a), m->p);
}
/* zq/mod.rs:610 */
static inline u64 reduce_u64(const zq_modulus *m, u64 a) {
  return reduce1(lazy_reduce(m, a), m->p);
}
/* zq/mod.rs:131 */
static inline u64 zq_mul(const zq_modulus *m, u64 a, u64 b) {
  return reduce_u128(m, (u128)a * b);
}
/* zq/mod.rs:151 */
static inline u64 zq_mul_opt(const zq_modulus *m, u64 a, u64 b) {
  return reduce1(lazy_reduce_opt_u128(m, (u128)a * b), m->p);
}
/* zq/mod.rs:195 */
static inline u64 zq_shoup(const zq_modulus *m, u64 a) {
  return (u64)((((u128)a) << 64) / (u128)m->p);
}
/* zq/mod.rs:224 -> [0,2p) for any a < 2^64 */
static inline u64 lazy_mul_shoup(const zq_modulus *m, u64 a, u64 b, u64 b_shoup) {
  u128 q = ((u128)a * b_shoup) >> 64;
  return (u64)((u128)a * b - q * (u128)m->p);
}
/* zq/mod.rs:205 */
static inline u64 mul_shoup(const zq_modulus *m, u64 a, u64 b, u64 b_shoup) {
  return reduce1(lazy_mul_shoup(m, a, b, b_shoup), m->p);
}

u64 orc_zq_mul(const zq_modulus *m, u64 a, u64 b) { return zq_mul(m, a, b); }
u64 orc_zq_mul_opt(const zq_modulus *m, u64 a, u64 b) { return zq_mul_opt(m, a, b); }
u64 orc_zq_shoup(const zq_modulus *m, u64 a) { return zq_shoup(m, a); }
u64 orc_zq_mul_shoup(const zq_modulus *m, u64 a, u64 b, u64 bs) { return mul_shoup(m, a, b, bs); }
u64 orc_zq_lazy_mul_shoup(const zq_modulus *m, u64 a, u64 b, u64 bs) { return lazy_mul_shoup(m, a, b, bs); }
u64 orc_zq_reduce(const zq_modulus *m, u64 a) { return reduce_u64(m, a); }
u64 orc_zq_reduce_u128(const zq_modulus *m, u64 lo, u64 hi) { return reduce_u128(m, ((u128)hi << 64) | lo); }
u64 orc_zq_lazy_reduce(const zq_modulus *m, u64 a) { return lazy_reduce(m, a); }
u64 orc_zq_lazy_reduce_opt(const zq_modulus *m, u64 a) { return lazy_reduce_opt(m, a); }

/* zq/mod.rs:556  pow (square-and-multiply, MSB first) */
u64 orc_zq_pow(const zq_modulus *m, u64 a, u64 n) {
  if (n == 0) return 1;
  if (n == 1) return a;
  u64 r = a;
  int i = 62 - __builtin_clzll(n);
  while (i >= 0) {
    r = zq_mul(m, r, r);
    if ((n >> i) & 1) r = zq_mul(m, r, a);
    i--;
  }
  return r;
}

/* zq/mod.rs:240,254 add_vec[_vt] */
void orc_zq_add_vec(const zq_modulus *m, u64 *a, const u64 *b, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = reduce1(a[i] + b[i], m->p);
}
/* zq/mod.rs:286,300 sub_vec[_vt] */
void orc_zq_sub_vec(const zq_modulus *m, u64 *a, const u64 *b, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = reduce1(a[i] + m->p - b[i], m->p);
}
/* zq/mod.rs:534,545 neg_vec[_vt] */
void orc_zq_neg_vec(const zq_modulus *m, u64 *a, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = reduce1(m->p - a[i], m->p);
}
/* zq/mod.rs:332,378 mul_vec[_vt] */
void orc_zq_mul_vec(const zq_modulus *m, u64 *a, const u64 *b, size_t n) {
  if (m->supports_opt)
    for (size_t i = 0; i < n; i++) a[i] = zq_mul_opt(m, a[i], b[i]);
  else
    for (size_t i = 0; i < n; i++) a[i] = zq_mul(m, a[i], b[i]);
}
/* zq/mod.rs:398 shoup_vec */
void orc_zq_shoup_vec(const zq_modulus *m, const u64 *a, u64 *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = zq_shoup(m, a[i]);
}
/* zq/mod.rs:407,425 mul_shoup_vec[_vt]; accepts lazy a (any u64) */
void orc_zq_mul_shoup_vec(const zq_modulus *m, u64 *a, const u64 *b, const u64 *bs, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = mul_shoup(m, a[i], b[i], bs[i]);
}
/* zq/mod.rs:349 scalar_mul_vec */
void orc_zq_scalar_mul_vec(const zq_modulus *m, u64 *a, u64 b, size_t n) {
  u64 bs = zq_shoup(m, b);
  for (size_t i = 0; i < n; i++) a[i] = mul_shoup(m, a[i], b, bs);
}
/* zq/mod.rs:438 reduce_vec */
void orc_zq_reduce_vec(const zq_modulus *m, u64 *a, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = reduce_u64(m, a[i]);
}
/* zq/mod.rs:756 lazy_reduce_vec -> [0,2p) */
void orc_zq_lazy_reduce_vec(const zq_modulus *m, u64 *a, size_t n) {
  if (m->supports_opt)
    for (size_t i = 0; i < n; i++) a[i] = lazy_reduce_opt(m, a[i]);
  else
    for (size_t i = 0; i < n; i++) a[i] = lazy_reduce(m, a[i]);
}
/* zq/mod.rs:479,494 reduce_vec_i64 */
void orc_zq_reduce_vec_i64(const zq_modulus *m, const int64_t *a, u64 *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    __int128 v = (((__int128)m->p) << 64) + (__int128)a[i];
    out[i] = reduce_u128(m, (u128)v);
  }
}

/* --------------------------------------------------------------- ntt ---- */

/* NttOperator::new table construction, ntt/native.rs:35-73, for a GIVEN psi
 * (= `omega` there).  omegas[i] = psi^{bitrev(i)}, zetas_inv[i] =
 * psi^{-(bitrev(i)+1)}; *_shoup companions; size_inv and its Shoup value.
 * scratch must hold 2*n u64. */
void orc_ntt_tables(const zq_modulus *m, size_t n, u64 psi, u64 psi_inv, u64 n_inv,
                    u64 *omegas, u64 *omegas_shoup, u64 *zetas_inv, u64 *zetas_inv_shoup,
                    u64 *size_inv_shoup, u64 *scratch) {
  u64 *powers = scratch, *powers_inv = scratch + n;
  u64 v = 1;
  for (size_t i = 0; i < n; i++) { powers[i] = v; v = zq_mul(m, v, psi); }
  v = psi_inv;
  for (size_t i = 0; i < n; i++) { powers_inv[i] = v; v = zq_mul(m, v, psi_inv); }
  int logn = __builtin_ctzll(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (int b = 0; b < logn; b++) j |= ((i >> b) & 1) << (logn - 1 - b);
    omegas[i] = powers[j];
    zetas_inv[i] = powers_inv[j];
    omegas_shoup[i] = zq_shoup(m, omegas[i]);
    zetas_inv_shoup[i] = zq_shoup(m, zetas_inv[i]);
  }
  *size_inv_shoup = zq_shoup(m, n_inv);
}

/* ntt/native.rs:142-179  forward_vt_lazy: natural order in ([0,4p) allowed),
 * bit-reversed order out, values in [0,4p).  Butterfly: native.rs:272-285. */
void orc_ntt_forward_lazy(const zq_modulus *m, u64 *a, size_t n,
                          const u64 *omegas, const u64 *omegas_shoup) {
  const u64 p_twice = 2 * m->p;
  size_t l = n >> 1, mm = 1, k = 1;
  while (l > 0) {
    for (size_t i = 0; i < mm; i++) {
      u64 w = omegas[k], ws = omegas_shoup[k];
      k++;
      size_t s = 2 * i * l;
      for (size_t j = s; j < s + l; j++) {
        u64 x = reduce1(a[j], p_twice);
        u64 t = lazy_mul_shoup(m, a[j + l], w, ws);
        a[j + l] = x + p_twice - t;
        a[j] = x + t;
      }
    }
    l >>= 1;
    mm <<= 1;
  }
}

/* ntt/native.rs:183-189 forward_vt (== forward, :77-102): lazy + reduce3 */
void orc_ntt_forward(const zq_modulus *m, u64 *a, size_t n,
                     const u64 *omegas, const u64 *omegas_shoup) {
  orc_ntt_forward_lazy(m, a, n, omegas, omegas_shoup);
  const u64 p_twice = 2 * m->p;
  for (size_t i = 0; i < n; i++) a[i] = reduce1(reduce1(a[i], p_twice), m->p);
}

/* ntt/native.rs:197-233 backward_vt (== backward, :106-132).  Inverse
 * butterfly: native.rs:303-316; final scaling by size_inv with mul_shoup. */
void orc_ntt_backward(const zq_modulus *m, u64 *a, size_t n,
                      const u64 *zetas_inv, const u64 *zetas_inv_shoup,
                      u64 size_inv, u64 size_inv_shoup) {
  const u64 p_twice = 2 * m->p;
  size_t k = 0, mm = n >> 1, l = 1;
  while (mm > 0) {
    for (size_t i = 0; i < mm; i++) {
      size_t s = 2 * i * l;
      u64 z = zetas_inv[k], zs = zetas_inv_shoup[k];
      k++;
      for (size_t j = s; j < s + l; j++) {
        u64 t = a[j];
        u64 y = a[j + l];
        a[j] = reduce1(y + t, p_twice);
        a[j + l] = lazy_mul_shoup(m, p_twice + t - y, z, zs);
      }
    }
    l <<= 1;
    mm >>= 1;
  }
  for (size_t i = 0; i < n; i++) a[i] = mul_shoup(m, a[i], size_inv, size_inv_shoup);
}

/* --------------------------------------------------------- rns scaler ---- */

/* 256-bit wrapping arithmetic standing in for ethnum::U256 1.5.3 as used at
 * rns/scaler.rs:260-313 (wrapping_add / wrapping_sub / mul / shr / not). */
typedef struct { u64 w[4]; } u256;

static inline u256 u256_zero(void) { u256 r = {{0, 0, 0, 0}}; return r; }
static inline u256 u256_add(u256 a, u256 b) {
  u256 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.w[i] + b.w[i]; r.w[i] = (u64)c; c >>= 64; }
  return r;
}
static inline u256 u256_not(u256 a) {
  u256 r; for (int i = 0; i < 4; i++) r.w[i] = ~a.w[i]; return r;
}
static inline u256 u256_sub(u256 a, u256 b) {
  u256 one = {{1, 0, 0, 0}};
  return u256_add(a, u256_add(u256_not(b), one));
}
/* (u64) * (u128) -> 192-bit product, exact inside 256 bits */
static inline u256 u256_mul_64_128(u64 a, u64 b_lo, u64 b_hi) {
  u256 r; u128 lo = (u128)a * b_lo, hi = (u128)a * b_hi;
  r.w[0] = (u64)lo;
  u128 mid = (lo >> 64) + (u64)hi;
  r.w[1] = (u64)mid;
  u128 top = (mid >> 64) + (hi >> 64);
  r.w[2] = (u64)top;
  r.w[3] = (u64)(top >> 64);
  return r;
}
/* (u128) * (u128) -> 256-bit product */
static inline u256 u256_mul_128_128(u128 a, u64 b_lo, u64 b_hi) {
  u256 lo = u256_mul_64_128((u64)a, b_lo, b_hi);
  u256 hi = u256_mul_64_128((u64)(a >> 64), b_lo, b_hi);
  u256 hs = {{0, hi.w[0], hi.w[1], hi.w[2]}};
  return u256_add(lo, hs);
}
static inline u256 u256_shr(u256 a, unsigned s) {
  u256 r = u256_zero();
  unsigned ws = s / 64, bs = s % 64;
  for (unsigned i = 0; i + ws < 4; i++) {
    u64 v = a.w[i + ws] >> bs;
    if (bs && i + ws + 1 < 4) v |= a.w[i + ws + 1] << (64 - bs);
    r.w[i] = v;
  }
  return r;
}
static inline u128 u256_as_u128(u256 a) { return ((u128)a.w[1] << 64) | a.w[0]; }
static inline int u256_nonzero(u256 a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) != 0; }

/* Flat view of rns::RnsScaler (rns/scaler.rs:52-73). omega/omega_shoup are
 * [n_to][n_from] row-major. */
typedef struct {
  u32 n_from, n_to;
  u32 is_one;
  u32 theta_garner_shift;
  const zq_modulus *to_moduli;      /* [n_to] */
  const u64 *gamma, *gamma_shoup;   /* [n_to] */
  u64 theta_gamma_lo, theta_gamma_hi;
  u32 theta_gamma_sign, _pad;
  const u64 *omega, *omega_shoup;   /* [n_to][n_from] */
  const u64 *theta_omega_lo, *theta_omega_hi; /* [n_from] */
  const uint8_t *theta_omega_sign;            /* [n_from] */
  const u64 *theta_garner_lo, *theta_garner_hi; /* [n_from] */
} rns_scaler;

/* RnsScaler::scale, rns/scaler.rs:249-352, for one coefficient column.
 * rests has stride rstride (u64 units), out has stride ostride. */
static void rns_scale_column(const rns_scaler *s, const u64 *rests, size_t rstride,
                             u64 *out, size_t ostride, size_t n_out, size_t starting_index) {
  /* :260-272 */
  u256 sum_theta_garner = u256_zero();
  for (u32 i = 0; i < s->n_from; i++)
    sum_theta_garner = u256_add(sum_theta_garner,
        u256_mul_64_128(rests[i * rstride], s->theta_garner_lo[i], s->theta_garner_hi[i]));
  sum_theta_garner = u256_shr(sum_theta_garner, s->theta_garner_shift - 1);
  u128 v128 = u256_as_u128(sum_theta_garner);
  u128 v = (v128 >> 1) + (v128 & 1); /* div_ceil(2) */

  /* :276-314 */
  int w_sign = 0;
  u128 w = 0;
  if (!s->is_one) {
    u256 sum_theta_omega = u256_zero();
    for (u32 i = 0; i < s->n_from; i++) {
      u256 prod = u256_mul_64_128(rests[i * rstride], s->theta_omega_lo[i], s->theta_omega_hi[i]);
      sum_theta_omega = s->theta_omega_sign[i] ? u256_sub(sum_theta_omega, prod)
                                               : u256_add(sum_theta_omega, prod);
    }
    u256 v_theta_gamma = u256_mul_128_128(v, s->theta_gamma_lo, s->theta_gamma_hi);
    sum_theta_omega = s->theta_gamma_sign ? u256_add(sum_theta_omega, v_theta_gamma)
                                          : u256_sub(sum_theta_omega, v_theta_gamma);
    w_sign = u256_nonzero(u256_shr(sum_theta_omega, 63 + 128));
    if (w_sign) {
      w = u256_as_u128(u256_shr(u256_not(sum_theta_omega), 126)) + 1;
      w /= 2;
    } else {
      w = u256_as_u128(u256_shr(sum_theta_omega, 126));
      w = (w >> 1) + (w & 1);
    }
  }

  /* :316-351 */
  for (size_t i = 0; i < n_out; i++) {
    size_t t = starting_index + i;
    const zq_modulus *qi = &s->to_moduli[t];
    const u64 *omega_i = s->omega + t * s->n_from;
    const u64 *omega_shoup_i = s->omega_shoup + t * s->n_from;
    u128 yi = (u128)(qi->p * 2 - lazy_mul_shoup(qi, reduce_u128(qi, v), s->gamma[t], s->gamma_shoup[t]));
    if (!s->is_one) {
      u64 wi = lazy_reduce_u128(qi, w);
      yi += (u128)(w_sign ? qi->p * 2 - wi : wi);
    }
    for (u32 j = 0; j < s->n_from; j++)
      yi += (u128)lazy_mul_shoup(qi, rests[j * rstride], omega_i[j], omega_shoup_i[j]);
    out[i * ostride] = reduce_u128(qi, yi);
  }
}

/* The per-column loop of rq::scaler::Scaler::scale (rq/scaler.rs:85-94):
 * in  = [n_from][n] power-basis rows, out = [n_out][n] rows written for the
 * `to` limbs starting_index .. starting_index+n_out. */
void orc_rns_scale_columns(const rns_scaler *s, const u64 *in, u64 *out, size_t n,
                           size_t n_out, size_t starting_index) {
  for (size_t c = 0; c < n; c++)
    rns_scale_column(s, in + c, n, out + c, n, n_out, starting_index);
}

/* RnsScaler::scale on one explicit residue vector (rns/scaler.rs:249). */
void orc_rns_scale_one(const rns_scaler *s, const u64 *rests, u64 *out, size_t n_out,
                       size_t starting_index) {
  rns_scale_column(s, rests, 1, out, 1, n_out, starting_index);
}

/* ----------------------------------------------------------------- rq ---- */

/* Poly<PowerBasis>::switch_down inner loops, rq/mod.rs:456-478.
 * rows = [n_limbs][n] in place; after the call rows 0..n_limbs-2 hold the
 * switched-down polynomial (the caller drops the last row). */
void orc_switch_down(const zq_modulus *q, size_t n_limbs, u64 *rows, size_t n,
                     const u64 *inv_last, const u64 *inv_last_shoup) {
  const zq_modulus *q_last = &q[n_limbs - 1];
  u64 q_last_div_2 = q_last->p / 2;
  u64 *last = rows + (n_limbs - 1) * n;
  for (size_t c = 0; c < n; c++) last[c] = reduce1(last[c] + q_last_div_2, q_last->p);
  for (size_t i = 0; i + 1 < n_limbs; i++) {
    const zq_modulus *qi = &q[i];
    u64 q_last_div_2_mod_qi = qi->p - reduce_u64(qi, q_last_div_2);
    u64 *row = rows + i * n;
    for (size_t c = 0; c < n; c++) {
      u64 tmp = lazy_reduce(qi, last[c]) + q_last_div_2_mod_qi;
      u64 v = row[c] + 3 * qi->p - tmp;
      row[c] = mul_shoup(qi, v, inv_last[i], inv_last_shoup[i]);
    }
  }
}

/* Poly::substitute, PowerBasis branch, rq/mod.rs:390-408 (one limb row). */
void orc_substitute_pb_row(const zq_modulus *m, const u64 *in, u64 *out, size_t n, size_t exponent) {
  memset(out, 0, n * sizeof(u64));
  size_t power = 0, mask = n - 1;
  for (size_t j = 0; j < n; j++) {
    size_t d = power & mask;
    if (power & n) out[d] = reduce1(out[d] + m->p - in[j], m->p);
    else out[d] = reduce1(out[d] + in[j], m->p);
    power += exponent;
  }
}

/* One digit of KeySwitchingKey::key_switch (key_switching_key.rs:256-268)
 * for one ksk limb j:  t = lazy NTT_j(lazy_reduce_j(digit)) (rq/mod.rs:563-586);
 * acc0 += t (*) k0 ; acc1 += t (*) k1  (rq/ops.rs:208 mul_shoup_vec_vt + :92 add_vec_vt).
 * scratch holds n u64. */
void orc_key_switch_digit_limb(const zq_modulus *m, const u64 *digit, size_t n,
                               const u64 *omegas, const u64 *omegas_shoup,
                               const u64 *k0, const u64 *k0s, const u64 *k1, const u64 *k1s,
                               u64 *acc0, u64 *acc1, u64 *scratch) {
  memcpy(scratch, digit, n * sizeof(u64));
  orc_zq_lazy_reduce_vec(m, scratch, n);
  orc_ntt_forward_lazy(m, scratch, n, omegas, omegas_shoup);
  for (size_t c = 0; c < n; c++) {
    u64 t = scratch[c];
    acc0[c] = reduce1(acc0[c] + mul_shoup(m, t, k0[c], k0s[c]), m->p);
    acc1[c] = reduce1(acc1[c] + mul_shoup(m, t, k1[c], k1s[c]), m->p);
  }
}
