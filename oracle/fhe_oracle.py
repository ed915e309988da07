"""CPU ORACLE for the fhe.rs BFV hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python (exact big-int precompute + orchestration) over ``fhe_oracle.c`` (the
scalar hot loops).  Restates tlepoint/fhe.rs @ e248cd28; every class / method
cites the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product never does.

Parity status
-------------
* arithmetic pinned: prime lists (zq/primes.rs:68-101, parameters.rs:846-856),
  RNS project/lift KATs (rns/mod.rs:217-238), scaler == BigUint centered
  rounding (rns/scaler.rs:397-414), NTT == negacyclic evaluation at
  psi^(2*bitrev(i)+1) and backward(forward(x)) == x (ntt/mod.rs:50-82),
  decrypt(multiply) == product mod t (ops/mul.rs:263-294).
* PARITY UNPINNED for the 2N-th root psi only: ``NttOperator::primitive_root``
  (ntt/native.rs:320-336) samples from ChaCha8Rng(seed 0) through rand 0.10.2 /
  rand_chacha 0.10.0, which are not vendored under /root/reference and cannot
  be run here (no Rust toolchain).  psi is an explicit input everywhere; the
  default rule below is documented and shared with the product's C ABI
  (NULL psi); a Rust host passes the reference's own psi = omegas[N/2].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfhe_oracle.so")


def build(force: bool = False) -> str:
    """Compile fhe_oracle.c -> libfhe_oracle.so (gcc -O3 -march=native)."""
    src = os.path.join(_HERE, "fhe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-std=gnu11", "-o", _SO, src])
    return _SO


class ZqModulusC(C.Structure):
    _fields_ = [("p", C.c_uint64), ("barrett_hi", C.c_uint64), ("barrett_lo", C.c_uint64),
                ("leading_zeros", C.c_uint32), ("supports_opt", C.c_uint32)]


class RnsScalerC(C.Structure):
    _fields_ = [("n_from", C.c_uint32), ("n_to", C.c_uint32), ("is_one", C.c_uint32),
                ("theta_garner_shift", C.c_uint32),
                ("to_moduli", C.POINTER(ZqModulusC)),
                ("gamma", C.c_void_p), ("gamma_shoup", C.c_void_p),
                ("theta_gamma_lo", C.c_uint64), ("theta_gamma_hi", C.c_uint64),
                ("theta_gamma_sign", C.c_uint32), ("_pad", C.c_uint32),
                ("omega", C.c_void_p), ("omega_shoup", C.c_void_p),
                ("theta_omega_lo", C.c_void_p), ("theta_omega_hi", C.c_void_p),
                ("theta_omega_sign", C.c_void_p),
                ("theta_garner_lo", C.c_void_p), ("theta_garner_hi", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        u64, vp, sz = C.c_uint64, C.c_void_p, C.c_size_t
        mp = C.POINTER(ZqModulusC)
        for name, res, args in [
            ("orc_zq_mul", u64, [mp, u64, u64]), ("orc_zq_mul_opt", u64, [mp, u64, u64]),
            ("orc_zq_shoup", u64, [mp, u64]), ("orc_zq_mul_shoup", u64, [mp, u64, u64, u64]),
            ("orc_zq_lazy_mul_shoup", u64, [mp, u64, u64, u64]),
            ("orc_zq_reduce", u64, [mp, u64]), ("orc_zq_reduce_u128", u64, [mp, u64, u64]),
            ("orc_zq_lazy_reduce", u64, [mp, u64]), ("orc_zq_lazy_reduce_opt", u64, [mp, u64]),
            ("orc_zq_pow", u64, [mp, u64, u64]),
            ("orc_zq_add_vec", None, [mp, vp, vp, sz]), ("orc_zq_sub_vec", None, [mp, vp, vp, sz]),
            ("orc_zq_neg_vec", None, [mp, vp, sz]), ("orc_zq_mul_vec", None, [mp, vp, vp, sz]),
            ("orc_zq_shoup_vec", None, [mp, vp, vp, sz]),
            ("orc_zq_mul_shoup_vec", None, [mp, vp, vp, vp, sz]),
            ("orc_zq_scalar_mul_vec", None, [mp, vp, u64, sz]),
            ("orc_zq_reduce_vec", None, [mp, vp, sz]), ("orc_zq_lazy_reduce_vec", None, [mp, vp, sz]),
            ("orc_zq_reduce_vec_i64", None, [mp, vp, vp, sz]),
            ("orc_ntt_tables", None, [mp, sz, u64, u64, u64, vp, vp, vp, vp, C.POINTER(u64), vp]),
            ("orc_ntt_forward_lazy", None, [mp, vp, sz, vp, vp]),
            ("orc_ntt_forward", None, [mp, vp, sz, vp, vp]),
            ("orc_ntt_backward", None, [mp, vp, sz, vp, vp, u64, u64]),
            ("orc_rns_scale_columns", None, [C.POINTER(RnsScalerC), vp, vp, sz, sz, sz]),
            ("orc_rns_scale_one", None, [C.POINTER(RnsScalerC), vp, vp, sz, sz]),
            ("orc_switch_down", None, [mp, sz, vp, sz, vp, vp]),
            ("orc_substitute_pb_row", None, [mp, vp, vp, sz, sz]),
            ("orc_key_switch_digit_limb", None, [mp, vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        ]:
            f = getattr(_lib, name)
            f.restype, f.argtypes = res, args
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 or a.dtype == np.int64 or a.dtype == np.uint8
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def u64arr(x) -> np.ndarray:
    return np.ascontiguousarray(np.array(x, dtype=np.uint64))


# --------------------------------------------------------------------------- util

def is_prime(n: int) -> bool:
    """fhe-util/src/lib.rs:16 (probably_prime); deterministic Miller-Rabin for u64."""
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for q in small:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def leading_zeros64(p: int) -> int:
    return 64 - p.bit_length()


def supports_opt(p: int) -> bool:
    """zq/primes.rs:10-24 (NFLlib eq. (1))."""
    lz = leading_zeros64(p)
    if lz < 1:
        return False
    middle = 1 << (3 * lz)
    left_side = (middle + 1) << 64
    middle *= (1 << lz) + 1
    middle *= p
    return left_side < middle


def generate_prime(num_bits: int, modulo: int, upper_bound: int) -> Optional[int]:
    """zq/primes.rs:30-59."""
    if not (10 <= num_bits <= 62):
        return None
    lz = 64 - num_bits
    t = upper_bound - 1
    while t % modulo != 1 and leading_zeros64(t) == lz:
        t -= 1
    while leading_zeros64(t) == lz and not is_prime(t) and t >= modulo:
        t -= modulo
    if leading_zeros64(t) == lz and is_prime(t):
        return t
    return None


def bitrev(i: int, logn: int) -> int:
    return int(format(i, "0%db" % logn)[::-1], 2) if logn else 0


def bitrev_table(n: int) -> np.ndarray:
    logn = n.bit_length() - 1
    idx = np.arange(n, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for b in range(logn):
        out |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(logn - 1 - b)
    return out.astype(np.int64)


# ----------------------------------------------------------------------------- zq

class Modulus:
    """zq::Modulus, zq/mod.rs:32-98."""

    def __init__(self, p: int):
        if p < 2 or (p >> 62) != 0:
            raise ValueError("InvalidModulus(%d)" % p)
        self.p = p
        barrett = (1 << 128) // p
        self.c = ZqModulusC(p, barrett >> 64, barrett & ((1 << 64) - 1),
                            leading_zeros64(p), 1 if supports_opt(p) else 0)
        self.supports_opt = bool(self.c.supports_opt)

    def ref(self):
        return C.byref(self.c)

    def shoup(self, a: int) -> int:  # zq/mod.rs:195
        return (a << 64) // self.p

    def inv(self, a: int) -> Optional[int]:  # zq/mod.rs:582
        if not is_prime(self.p) or a == 0:
            return None
        return pow(a, self.p - 2, self.p)


def default_psi(p: int, n: int) -> int:
    """DEFAULT 2n-th primitive root (documented rule shared with the product's
    C ABI when psi == NULL): psi = g^((p-1)/2n) for the smallest g >= 2 that
    makes psi primitive (psi^n == p-1).  NOT the reference's ChaCha8-sampled
    root (ntt/native.rs:320-336) -- that one is unpinned, see module docstring."""
    lam = (p - 1) // (2 * n)
    g = 2
    while True:
        psi = pow(g, lam, p)
        if pow(psi, n, p) == p - 1:
            return psi
        g += 1


class NttOperator:
    """ntt::native::NttOperator, ntt/native.rs:16-73 (tables for a given psi)."""

    def __init__(self, m: Modulus, size: int, psi: Optional[int] = None):
        assert size >= 8 and size & (size - 1) == 0
        if not (m.p % (2 * size) == 1 and is_prime(m.p)):  # ntt/mod.rs:20
            raise ValueError("NttOperatorUnavailable")
        self.m, self.size = m, size
        self.psi = default_psi(m.p, size) if psi is None else psi
        assert pow(self.psi, size, m.p) == m.p - 1, "psi is not a primitive 2n-th root"
        self.size_inv = m.inv(size)
        psi_inv = m.inv(self.psi)
        self.omegas = np.zeros(size, np.uint64)
        self.omegas_shoup = np.zeros(size, np.uint64)
        self.zetas_inv = np.zeros(size, np.uint64)
        self.zetas_inv_shoup = np.zeros(size, np.uint64)
        sis = C.c_uint64(0)
        scratch = np.zeros(2 * size, np.uint64)
        lib().orc_ntt_tables(m.ref(), size, self.psi, psi_inv, self.size_inv,
                             _p(self.omegas), _p(self.omegas_shoup), _p(self.zetas_inv),
                             _p(self.zetas_inv_shoup), C.byref(sis), _p(scratch))
        self.size_inv_shoup = sis.value

    def forward(self, a: np.ndarray):  # native.rs:77 / :183
        lib().orc_ntt_forward(self.m.ref(), _p(a), self.size, _p(self.omegas), _p(self.omegas_shoup))

    def forward_lazy(self, a: np.ndarray):  # native.rs:142
        lib().orc_ntt_forward_lazy(self.m.ref(), _p(a), self.size, _p(self.omegas), _p(self.omegas_shoup))

    def backward(self, a: np.ndarray):  # native.rs:106 / :197
        lib().orc_ntt_backward(self.m.ref(), _p(a), self.size, _p(self.zetas_inv),
                               _p(self.zetas_inv_shoup), self.size_inv, self.size_inv_shoup)


# ---------------------------------------------------------------------------- rns

class RnsContext:
    """rns::RnsContext, rns/mod.rs:24-116."""

    def __init__(self, moduli: Sequence[int]):
        if len(moduli) == 0:
            raise ValueError("EmptyModuli")
        from math import gcd
        for i, a in enumerate(moduli):
            for j, b in enumerate(moduli):
                if i != j and gcd(a, b) != 1:
                    raise ValueError("NonCoprimeModuli")
        self.moduli_u64 = list(moduli)
        self.moduli = [Modulus(q) for q in moduli]
        self.product = 1
        for q in moduli:
            self.product *= q
        self.q_star = [self.product // q for q in moduli]
        self.q_tilde = [pow(qs % q, -1, q) for qs, q in zip(self.q_star, moduli)]
        self.garner = [qs * qt for qs, qt in zip(self.q_star, self.q_tilde)]
        self.moduli_c = (ZqModulusC * len(moduli))(*[m.c for m in self.moduli])

    def project(self, a: int) -> List[int]:  # :126
        return [a % q for q in self.moduli_u64]

    def lift(self, rests: Sequence[int]) -> int:  # :138
        return sum(int(r) * g for r, g in zip(rests, self.garner)) % self.product


class ScalingFactor:
    """rns/scaler.rs:20-48."""

    def __init__(self, numerator: int, denominator: int):
        assert denominator != 0
        self.numerator, self.denominator = numerator, denominator
        self.is_one = numerator == denominator

    @staticmethod
    def one():
        return ScalingFactor(1, 1)


class RnsScaler:
    """rns::RnsScaler, rns/scaler.rs:52-352 (tables by exact integers; `scale` in C)."""

    def __init__(self, frm: RnsContext, to: RnsContext, factor: ScalingFactor):
        self.frm, self.to, self.factor = frm, to, factor
        num, den = factor.numerator, factor.denominator
        # :81-91
        gamma, self.theta_gamma_lo, self.theta_gamma_hi, self.theta_gamma_sign = \
            self._extract(to, frm.product, num, den, False)
        self.gamma = u64arr(gamma)
        self.gamma_shoup = u64arr([q.shoup(g) for g, q in zip(gamma, to.moduli)])
        # :93-124
        nf, nt = len(frm.moduli), len(to.moduli)
        omega = np.zeros((nt, nf), np.uint64)
        omega_shoup = np.zeros((nt, nf), np.uint64)
        tol, toh, tos = [], [], []
        for i, garner_i in enumerate(frm.garner):
            om_i, lo, hi, sg = self._extract(to, garner_i, num, den, True)
            tol.append(lo); toh.append(hi); tos.append(1 if sg else 0)
            for j in range(nt):
                qj = to.moduli[j]
                omega[j, i] = om_i[j] % qj.p
                omega_shoup[j, i] = qj.shoup(int(omega[j, i]))
        self.omega, self.omega_shoup = omega, omega_shoup
        self.theta_omega_lo, self.theta_omega_hi = u64arr(tol), u64arr(toh)
        self.theta_omega_sign = np.ascontiguousarray(np.array(tos, dtype=np.uint8))
        # :128-142
        def npot_ilog2(x):  # next_power_of_two().ilog2()
            return (x - 1).bit_length() if x > 1 else 0
        self.theta_garner_shift = min(
            min(192 - 1 - npot_ilog2(qi * nf) for qi in frm.moduli_u64), 127)
        # :145-155
        tgl, tgh = [], []
        for garner_i in frm.garner:
            theta = ((garner_i << self.theta_garner_shift) + (frm.product >> 1)) // frm.product
            tgh.append(theta >> 64); tgl.append(theta & ((1 << 64) - 1))
        self.theta_garner_lo, self.theta_garner_hi = u64arr(tgl), u64arr(tgh)
        self.c = RnsScalerC(
            nf, nt, 1 if factor.is_one else 0, self.theta_garner_shift,
            C.cast(to.moduli_c, C.POINTER(ZqModulusC)), _p(self.gamma), _p(self.gamma_shoup),
            self.theta_gamma_lo, self.theta_gamma_hi, 1 if self.theta_gamma_sign else 0, 0,
            _p(self.omega), _p(self.omega_shoup), _p(self.theta_omega_lo), _p(self.theta_omega_hi),
            _p(self.theta_omega_sign), _p(self.theta_garner_lo), _p(self.theta_garner_hi))

    @staticmethod
    def _extract(ctx: RnsContext, inp: int, num: int, den: int, round_up: bool):
        """extract_projection_and_theta, rns/scaler.rs:183-229."""
        gamma = (num * inp + (den >> 1)) // den
        projected = ctx.project(gamma)
        theta = (num * inp) % den
        sign = False
        if den > 1:
            if den & 1:
                if theta > (den >> 1):
                    sign, theta = True, den - theta
            elif theta >= (den >> 1):
                sign, theta = True, den - theta
        if round_up:
            theta = (theta << 127) // den if sign else ((theta << 127) + den - 1) // den
        elif sign:
            theta = ((theta << 127) + den - 1) // den
        else:
            theta = (theta << 127) // den
        return projected, theta & ((1 << 64) - 1), theta >> 64, sign

    def scale_one(self, rests: Sequence[int], size: int, starting_index: int = 0) -> List[int]:
        """RnsScaler::scale on one residue vector, rns/scaler.rs:249."""
        r = u64arr(rests)
        out = np.zeros(size, np.uint64)
        lib().orc_rns_scale_one(C.byref(self.c), _p(r), _p(out), size, starting_index)
        return [int(x) for x in out]

    def scale_columns(self, rows: np.ndarray, n_out: int, starting_index: int) -> np.ndarray:
        """per-column loop of rq/scaler.rs:85-94."""
        n = rows.shape[1]
        out = np.zeros((n_out, n), np.uint64)
        lib().orc_rns_scale_columns(C.byref(self.c), _p(rows), _p(out), n, n_out, starting_index)
        return out


# ----------------------------------------------------------------------------- rq

POWER_BASIS, NTT, NTT_SHOUP = 0, 1, 2


class Context:
    """rq::Context, rq/context.rs:9-92.  `psi` maps modulus -> 2N-th root."""

    def __init__(self, moduli: Sequence[int], degree: int, psi: Optional[dict] = None):
        if degree < 8 or degree & (degree - 1):
            raise ValueError("InvalidPolynomialDegree")
        self.moduli = list(moduli)
        self.degree = degree
        self.psi = psi or {}
        self.rns = RnsContext(moduli)
        self.q = [Modulus(q) for q in moduli]
        self.ops = [_ntt_op(q, degree, self.psi.get(q)) for q in moduli]
        self.bitrev = bitrev_table(degree)
        q_last = moduli[-1]
        inv = [pow(q_last % q, -1, q) for q in moduli[:-1]]  # context.rs:65-71
        self.inv_last_qi_mod_qj = np.array(inv, dtype=np.uint64)
        self.inv_last_qi_mod_qj_shoup = np.array(
            [m.shoup(v) for m, v in zip(self.q[:-1], inv)], dtype=np.uint64)
        self._next = None
        self.q_c = (ZqModulusC * len(moduli))(*[m.c for m in self.q])

    @property
    def next_context(self) -> Optional["Context"]:
        if self._next is None and len(self.moduli) >= 2:
            self._next = Context(self.moduli[:-1], self.degree, self.psi)
        return self._next

    def modulus(self) -> int:
        return self.rns.product

    def __eq__(self, o):
        return isinstance(o, Context) and self.moduli == o.moduli and self.degree == o.degree

    def niterations_to(self, other: "Context") -> int:  # context.rs:116
        n, cur = 0, self
        while cur != other:
            cur = cur.next_context
            if cur is None:
                raise ValueError("ContextNotReachable")
            n += 1
        return n


_NTT_CACHE: dict = {}


def _ntt_op(q: int, degree: int, psi: Optional[int]) -> NttOperator:
    key = (q, degree, psi)
    if key not in _NTT_CACHE:
        _NTT_CACHE[key] = NttOperator(Modulus(q), degree, psi)
    return _NTT_CACHE[key]


class Poly:
    """rq::Poly<R>, rq/mod.rs:126-133: coefficients [limb][coeff] row-major u64."""

    def __init__(self, ctx: Context, rep: int, coeffs: Optional[np.ndarray] = None):
        self.ctx, self.rep = ctx, rep
        self.c = np.zeros((len(ctx.moduli), ctx.degree), np.uint64) if coeffs is None \
            else np.ascontiguousarray(coeffs, dtype=np.uint64)
        assert self.c.shape == (len(ctx.moduli), ctx.degree)
        self.shoup = None
        if rep == NTT_SHOUP:
            self.compute_shoup()

    def copy(self) -> "Poly":
        p = Poly(self.ctx, self.rep if self.rep != NTT_SHOUP else NTT, self.c.copy())
        if self.rep == NTT_SHOUP:
            p.rep, p.shoup = NTT_SHOUP, self.shoup.copy()
        return p

    def compute_shoup(self):  # rq/mod.rs:256
        self.shoup = np.zeros_like(self.c)
        for i, m in enumerate(self.ctx.q):
            lib().orc_zq_shoup_vec(m.ref(), _p(self.c[i]), _p(self.shoup[i]), self.ctx.degree)

    @staticmethod
    def from_i64(ctx: Context, v: np.ndarray, rep: int = POWER_BASIS) -> "Poly":
        """try_convert_from(&[i64]) rq/convert.rs:329-354 (PowerBasis) then optional NTT."""
        v = np.ascontiguousarray(v, dtype=np.int64)
        p = Poly(ctx, POWER_BASIS)
        for i, m in enumerate(ctx.q):
            lib().orc_zq_reduce_vec_i64(m.ref(), _p(v), _p(p.c[i]), ctx.degree)
        return p if rep == POWER_BASIS else p.into_ntt()

    @staticmethod
    def from_u64(ctx: Context, v: np.ndarray, rep: int = POWER_BASIS) -> "Poly":
        """try_convert_from(&[u64]): reduce each coefficient modulo every q_i."""
        v = np.ascontiguousarray(v, dtype=np.uint64)
        p = Poly(ctx, POWER_BASIS)
        for i, m in enumerate(ctx.q):
            p.c[i, : len(v)] = v
            lib().orc_zq_reduce_vec(m.ref(), _p(p.c[i]), ctx.degree)
        return p if rep == POWER_BASIS else p.into_ntt()

    @staticmethod
    def random(ctx: Context, rep: int, rng: np.random.Generator) -> "Poly":
        """Poly::random (rq/mod.rs:262): uniform residues (numpy RNG, not ChaCha8)."""
        p = Poly(ctx, NTT if rep == NTT_SHOUP else rep)
        for i, q in enumerate(ctx.moduli):
            p.c[i] = rng.integers(0, q, size=ctx.degree, dtype=np.uint64)
        if rep == NTT_SHOUP:
            p.rep = NTT_SHOUP
            p.compute_shoup()
        return p

    # -- representation changes (rq/mod.rs:335-350, :535, :590)
    def into_ntt(self) -> "Poly":
        assert self.rep == POWER_BASIS
        for i, op in enumerate(self.ctx.ops):
            op.forward(self.c[i])
        self.rep = NTT
        return self

    def into_power_basis(self) -> "Poly":
        assert self.rep in (NTT, NTT_SHOUP)
        for i, op in enumerate(self.ctx.ops):
            op.backward(self.c[i])
        self.rep, self.shoup = POWER_BASIS, None
        return self

    def into_ntt_shoup(self) -> "Poly":
        if self.rep == POWER_BASIS:
            self.into_ntt()
        self.rep = NTT_SHOUP
        self.compute_shoup()
        return self

    # -- arithmetic (rq/ops.rs)
    def iadd(self, o: "Poly") -> "Poly":  # ops.rs:92
        assert self.ctx == o.ctx
        for i, m in enumerate(self.ctx.q):
            lib().orc_zq_add_vec(m.ref(), _p(self.c[i]), _p(o.c[i]), self.ctx.degree)
        return self

    def isub(self, o: "Poly") -> "Poly":  # ops.rs:133
        assert self.ctx == o.ctx
        for i, m in enumerate(self.ctx.q):
            lib().orc_zq_sub_vec(m.ref(), _p(self.c[i]), _p(o.c[i]), self.ctx.degree)
        return self

    def neg(self) -> "Poly":  # ops.rs:354
        r = self.copy()
        for i, m in enumerate(self.ctx.q):
            lib().orc_zq_neg_vec(m.ref(), _p(r.c[i]), self.ctx.degree)
        return r

    def imul(self, o: "Poly") -> "Poly":  # ops.rs:174 (Ntt x Ntt) / :208 (Ntt x NttShoup)
        assert self.ctx == o.ctx and self.rep == NTT
        for i, m in enumerate(self.ctx.q):
            if o.rep == NTT_SHOUP:
                lib().orc_zq_mul_shoup_vec(m.ref(), _p(self.c[i]), _p(o.c[i]), _p(o.shoup[i]), self.ctx.degree)
            else:
                assert o.rep == NTT
                lib().orc_zq_mul_vec(m.ref(), _p(self.c[i]), _p(o.c[i]), self.ctx.degree)
        return self

    def mul(self, o: "Poly") -> "Poly":  # ops.rs:247
        return self.copy().imul(o)

    def mul_scalar_big(self, k: int) -> "Poly":
        """&BigUint * &Poly (rq/ops.rs:296-330): multiply by k mod each q_i."""
        r = self.copy()
        for i, m in enumerate(self.ctx.q):
            lib().orc_zq_scalar_mul_vec(m.ref(), _p(r.c[i]), k % m.p, self.ctx.degree)
        return r

    # -- substitute (rq/mod.rs:99-121, :360-412)
    def substitute(self, exponent: int) -> "Poly":
        n = self.ctx.degree
        exponent %= 2 * n
        if exponent & 1 == 0:
            raise ValueError("InvalidSubstitutionExponent")
        q = Poly(self.ctx, self.rep if self.rep != NTT_SHOUP else NTT)
        if self.rep in (NTT, NTT_SHOUP):
            power = (exponent - 1) // 2 + exponent * np.arange(n, dtype=np.int64)
            power_bitrev = self.ctx.bitrev[power & (n - 1)]
            q.c[:, self.ctx.bitrev] = self.c[:, power_bitrev]
            if self.rep == NTT_SHOUP:
                q.rep = NTT_SHOUP
                q.shoup = np.zeros_like(self.shoup)
                q.shoup[:, self.ctx.bitrev] = self.shoup[:, power_bitrev]
        else:
            for i, m in enumerate(self.ctx.q):
                lib().orc_substitute_pb_row(m.ref(), _p(self.c[i]), _p(q.c[i]), n, exponent)
        return q

    # -- switch_down (rq/mod.rs:433-492)
    def switch_down(self) -> "Poly":
        assert self.rep == POWER_BASIS
        nxt = self.ctx.next_context
        if nxt is None:
            raise ValueError("NoMoreContext")
        L = len(self.ctx.moduli)
        lib().orc_switch_down(C.cast(self.ctx.q_c, C.POINTER(ZqModulusC)), L, _p(self.c), self.ctx.degree,
                              _p(self.ctx.inv_last_qi_mod_qj), _p(self.ctx.inv_last_qi_mod_qj_shoup))
        self.c = np.ascontiguousarray(self.c[: L - 1])
        self.ctx = nxt
        return self

    def switch_down_to(self, ctx: Context) -> "Poly":  # rq/mod.rs:498
        for _ in range(self.ctx.niterations_to(ctx)):
            self.switch_down()
        return self

    def to_bigints(self) -> List[int]:
        """Vec<BigUint>::from(&Poly) (rq/convert.rs): CRT-lift each coefficient."""
        assert self.rep == POWER_BASIS
        return [self.ctx.rns.lift([int(x) for x in self.c[:, j]]) for j in range(self.ctx.degree)]


def transcode_to_bytes(a: Sequence[int], nbits: int) -> bytes:
    """fhe_util::transcode_to_bytes, fhe-util/src/lib.rs:71-108 (LSB-first bit stream)."""
    mask = (1 << nbits) - 1
    cur, have, out = 0, 0, bytearray()
    for v in a:
        cur |= (int(v) & mask) << have
        have += nbits
        while have >= 8:
            out.append(cur & 0xFF)
            cur >>= 8
            have -= 8
    if have > 0:
        out.append(cur & 0xFF)
    return bytes(out)


def transcode_from_bytes(b: bytes, nbits: int) -> List[int]:
    """fhe_util::transcode_from_bytes, fhe-util/src/lib.rs:112-146."""
    mask = (1 << nbits) - 1
    cur, have, out = 0, 0, []
    for byte in b:
        cur |= byte << have
        have += 8
        while have >= nbits:
            out.append(cur & mask)
            cur >>= nbits
            have -= nbits
    if have > 0:
        out.append(cur)
    return out


def poly_to_rq_coefficients(p: Poly) -> bytes:
    """`Rq::from(&Poly).coefficients`, rq/convert.rs:17-44: always power basis, per-limb serialize_vec (zq/mod.rs:783)."""
    q = p.copy()
    if q.rep != POWER_BASIS:
        q.into_power_basis()
    out = b""
    for i, m in enumerate(q.ctx.moduli):
        out += transcode_to_bytes(q.c[i], (m - 1).bit_length())
    return out


def poly_from_rq_coefficients(ctx: Context, blob: bytes, rep: int) -> Poly:
    """TryConvertFrom<&Rq> for Poly<PowerBasis|Ntt>, rq/convert.rs:46-131."""
    n, idx, rows = ctx.degree, 0, []
    for m in ctx.moduli:
        nb = (m - 1).bit_length()
        size = nb * n // 8
        rows.append(transcode_from_bytes(blob[idx: idx + size], nb)[:n])
        idx += size
    if idx != len(blob):
        raise ValueError("InvalidCoefficientCount")
    p = Poly(ctx, POWER_BASIS, np.array(rows, dtype=np.uint64))
    return p if rep == POWER_BASIS else p.into_ntt()


def lazy_constant_ntt(row: np.ndarray, ctx: Context) -> Poly:
    """create_constant_ntt_polynomial_with_lazy_coefficients_and_variable_time,
    rq/mod.rs:563-586: values in [0,4q_j)."""
    p = Poly(ctx, NTT)
    for j, (m, op) in enumerate(zip(ctx.q, ctx.ops)):
        p.c[j] = row
        lib().orc_zq_lazy_reduce_vec(m.ref(), _p(p.c[j]), ctx.degree)
        op.forward_lazy(p.c[j])
    return p


class Scaler:
    """rq::scaler::Scaler, rq/scaler.rs:18-127."""

    def __init__(self, frm: Context, to: Context, factor: ScalingFactor):
        assert frm.degree == to.degree
        self.frm, self.to = frm, to
        if factor.is_one:  # :35-43
            n = 0
            for a, b in zip(frm.moduli, to.moduli):
                if a != b:
                    break
                n += 1
            self.number_common_moduli = n
        else:
            self.number_common_moduli = 0
        self.scaler = RnsScaler(frm.rns, to.rns, factor)

    def scale(self, p: Poly) -> Poly:  # :55-127
        assert p.ctx == self.frm and p.rep in (POWER_BASIS, NTT)
        nc, nt = self.number_common_moduli, len(self.to.moduli)
        new = np.zeros((nt, self.to.degree), np.uint64)
        if nc > 0:
            new[:nc] = p.c[:nc]
        if nc < nt:
            if p.rep != POWER_BASIS:
                pb = p.c.copy()
                for i, op in enumerate(p.ctx.ops):
                    op.backward(pb[i])
            else:
                pb = p.c
            new[nc:] = self.scaler.scale_columns(pb, nt - nc, nc)
            if p.rep != POWER_BASIS:
                for j in range(nc, nt):
                    self.to.ops[j].forward(new[j])
        return Poly(self.to, p.rep, new)


class Switcher(Scaler):
    """rq/switcher.rs:11-26."""

    def __init__(self, frm: Context, to: Context):
        super().__init__(frm, to, ScalingFactor(to.modulus(), frm.modulus()))

    def switch(self, p: Poly) -> Poly:
        return self.scale(p)


# ---------------------------------------------------------------------------- bfv

class BfvParameters:
    """fhe::bfv::BfvParameters (+Builder::build), bfv/parameters.rs:560-738.
    Per-level tables are built lazily."""

    def __init__(self, degree: int, plaintext: int, moduli: Optional[Sequence[int]] = None,
                 moduli_sizes: Optional[Sequence[int]] = None, variance: int = 10,
                 psi: Optional[dict] = None):
        if degree < 8 or degree & (degree - 1) or degree > 65536:
            raise ValueError("invalid degree")
        if (moduli is None) == (moduli_sizes is None):
            raise ValueError("exactly one of moduli / moduli_sizes")
        self.degree, self.plaintext, self.variance = degree, plaintext, variance
        self.psi = dict(psi or {})
        self.moduli = list(moduli) if moduli is not None else self.generate_moduli(moduli_sizes, degree)
        for q in self.moduli:
            Modulus(q)
            if q % (2 * degree) != 1 or not is_prime(q):
                raise ValueError("CiphertextModulusNotNttFriendly")
        if len(set(self.moduli)) != len(self.moduli):
            raise ValueError("DuplicateModuli")
        self.moduli_sizes = [q.bit_length() for q in self.moduli]
        # plaintext context (:579-595)
        acc = cnt = 0
        for s in self.moduli_sizes:
            acc += s
            cnt += 1
            if acc >= plaintext.bit_length() + 60:
                break
        cnt = min(max(cnt, 1), len(self.moduli))
        self.plaintext_context = Context(self.moduli[:cnt], degree, self.psi)
        # extended basis (:660-676)
        self.extended_basis: List[int] = []
        ub = 1 << 62
        while len(self.extended_basis) != len(self.moduli) + 1:
            ub = generate_prime(62, 2 * degree, ub)
            if ub is None:
                raise ValueError("NotEnoughPrimes")
            if ub not in self.extended_basis and ub not in self.moduli:
                self.extended_basis.append(ub)
        self._levels: dict = {}
        # SIMD index map (:713-726)
        row, m = degree >> 1, degree << 1
        logn = degree.bit_length() - 1
        self.matrix_reps_index_map = [0] * degree
        pos = 1
        for i in range(row):
            self.matrix_reps_index_map[i] = bitrev((pos - 1) >> 1, logn)
            self.matrix_reps_index_map[row | i] = bitrev((m - pos - 1) >> 1, logn)
            pos = (pos * 3) & (m - 1)

    @staticmethod
    def generate_moduli(sizes: Sequence[int], degree: int) -> List[int]:
        """parameters.rs:391-431."""
        moduli: List[int] = []
        for size in sizes:
            if size > 62 or size < 10:
                raise ValueError("InvalidModulusSize")
            ub = 1 << size
            while True:
                prime = generate_prime(size, 2 * degree, ub)
                if prime is None:
                    raise ValueError("NotEnoughPrimes")
                if prime not in moduli:
                    moduli.append(prime)
                    break
                ub = prime
        return moduli

    def max_level(self) -> int:
        return len(self.moduli) - 1

    def level(self, level: int) -> "ContextLevel":
        if level not in self._levels:
            if not (0 <= level < len(self.moduli)):
                raise ValueError("InvalidLevel")
            self._levels[level] = ContextLevel(self, level)
        return self._levels[level]

    def context_at_level(self, level: int) -> Context:
        return self.level(level).poly_context


class ContextLevel:
    """context/level.rs + cipher_plain_context.rs + MultiplicationParameters
    (parameters.rs:600-700, :793-813) for one level."""

    def __init__(self, par: BfvParameters, level: int):
        L = len(par.moduli) - level
        mods = par.moduli[:L]
        self.level = level
        self.poly_context = Context(mods, par.degree, par.psi)
        t = par.plaintext
        # delta (:608-630): residues (-t)^-1 mod q_i; NTT of a constant is the constant row.
        self.delta_rests = [pow((-t) % q, -1, q) for q in mods]
        self.delta = Poly(self.poly_context, NTT,
                          np.repeat(u64arr(self.delta_rests)[:, None], par.degree, axis=1))
        self.delta.rep = NTT_SHOUP
        self.delta.compute_shoup()
        self.q_mod_t = self.poly_context.modulus() % t
        self.plain_threshold = (t + 1) >> 1
        self._scaler = None
        self._mul = None
        self.par = par

    @property
    def scaler(self) -> Scaler:  # :639-643 ciphertext -> plaintext context, factor t/Q
        if self._scaler is None:
            self._scaler = Scaler(self.poly_context, self.par.plaintext_context,
                                  ScalingFactor(self.par.plaintext, self.poly_context.modulus()))
        return self._scaler

    @property
    def mul_params(self):  # :686-700
        if self._mul is None:
            par = self.par
            L = len(self.poly_context.moduli)
            modulus_size = sum(par.moduli_sizes[:L])
            n_moduli = -(-(modulus_size + 60) // 62)
            mul_moduli = par.moduli[:L] + par.extended_basis[:n_moduli]
            to = Context(mul_moduli, par.degree, par.psi)
            self._mul = MultiplicationParameters(
                self.poly_context, to, ScalingFactor.one(),
                ScalingFactor(par.plaintext, self.poly_context.modulus()))
        return self._mul


class MultiplicationParameters:
    """parameters.rs:793-813."""

    def __init__(self, frm: Context, to: Context, up: ScalingFactor, down: ScalingFactor):
        self.extender = Scaler(frm, to, up)
        self.down_scaler = Scaler(to, frm, down)
        self.frm, self.to = frm, to


def sample_vec_cbd(n: int, variance: int, rng: np.random.Generator) -> np.ndarray:
    """fhe-util/src/lib.rs:22-67 (centered binomial; numpy bit source)."""
    assert 1 <= variance <= 32
    bits = rng.integers(0, 2, size=(n, 4 * variance), dtype=np.int64)
    return bits[:, : 2 * variance].sum(axis=1) - bits[:, 2 * variance:].sum(axis=1)


class Ciphertext:
    """bfv::Ciphertext, ciphertext.rs:18: parts are Poly<Ntt> at `level`."""

    def __init__(self, par: BfvParameters, c: List[Poly], level: int):
        self.par, self.c, self.level = par, c, level

    def copy(self):
        return Ciphertext(self.par, [p.copy() for p in self.c], self.level)

    def to_array(self) -> np.ndarray:
        return np.stack([p.c for p in self.c])

    @staticmethod
    def from_array(par: BfvParameters, arr: np.ndarray, level: int) -> "Ciphertext":
        ctx = par.context_at_level(level)
        return Ciphertext(par, [Poly(ctx, NTT, a) for a in arr], level)

    def add(self, o):  # ops/mod.rs:15-45
        assert self.level == o.level and len(self.c) == len(o.c)
        return Ciphertext(self.par, [a.copy().iadd(b) for a, b in zip(self.c, o.c)], self.level)

    def sub(self, o):  # ops/mod.rs:109-140
        assert self.level == o.level and len(self.c) == len(o.c)
        return Ciphertext(self.par, [a.copy().isub(b) for a, b in zip(self.c, o.c)], self.level)

    def neg(self):  # ops/mod.rs:205-227
        return Ciphertext(self.par, [a.neg() for a in self.c], self.level)

    def mul(self, o):
        """operator &ct * &ct, ops/mod.rs:259-358 (both branches compute this)."""
        assert self.level == o.level
        mp = self.par.level(self.level).mul_params
        sc = [mp.extender.scale(p) for p in self.c]
        oc = [mp.extender.scale(p) for p in o.c]
        c = [Poly(mp.to, NTT) for _ in range(len(sc) + len(oc) - 1)]
        for i, a in enumerate(sc):
            for j, b in enumerate(oc):
                c[i + j].iadd(a.mul(b))
        c = [mp.down_scaler.scale(p) for p in c]
        return Ciphertext(self.par, c, self.level)

    def switch_down(self):  # ciphertext.rs:148-161
        if self.level >= self.par.max_level():
            raise ValueError("NoMoreContext")
        self.c = [p.copy().into_power_basis().switch_down().into_ntt() for p in self.c]
        self.level += 1
        return self

    def switch_to_level(self, target_level: int):  # ciphertext.rs:164-184
        if target_level < self.level or target_level > self.par.max_level():
            raise ValueError("InvalidLevel")
        while self.level < target_level:
            self.switch_down()
        return self


class SecretKey:
    """keys/secret_key.rs (client side; oracle-only test plumbing)."""

    def __init__(self, par: BfvParameters, rng: np.random.Generator):
        self.par = par
        self.coeffs = sample_vec_cbd(par.degree, par.variance, rng)

    def s_ntt(self, ctx: Context) -> Poly:
        return Poly.from_i64(ctx, self.coeffs, NTT)

    def encrypt_poly(self, m_scaled: Poly, level: int, rng) -> Ciphertext:  # :97-136
        ctx = self.par.context_at_level(level)
        a = Poly.random(ctx, NTT, rng)
        b = Poly.from_i64(ctx, sample_vec_cbd(self.par.degree, self.par.variance, rng), NTT)
        b.isub(a.mul(self.s_ntt(ctx)))
        b.iadd(m_scaled)
        return Ciphertext(self.par, [b, a], level)

    def encrypt(self, values: Sequence[int], level: int, rng) -> Ciphertext:
        """Encrypt a Poly-encoded plaintext (coefficients = values mod t)."""
        return self.encrypt_poly(plaintext_to_poly(self.par, values, level), level, rng)

    def phase(self, ct: Ciphertext) -> Poly:  # :199-222
        ctx = ct.c[0].ctx
        s = self.s_ntt(ctx)
        si = s.copy()
        c = ct.c[0].copy()
        for i in range(1, len(ct.c)):
            c.iadd(ct.c[i].mul(si))
            if i + 1 < len(ct.c):
                si.imul(s)
        return c

    def decrypt(self, ct: Ciphertext) -> np.ndarray:  # :199-259 (small-t path)
        par = self.par
        c_pb = self.phase(ct).into_power_basis()
        d = par.level(ct.level).scaler.scale(c_pb)
        w = d.c[0].copy() + np.uint64(par.plaintext)
        q0 = Modulus(par.moduli[0])
        lib().orc_zq_reduce_vec(q0.ref(), _p(w), par.degree)
        return w % np.uint64(par.plaintext)

    def measure_noise(self, ct: Ciphertext) -> int:  # :61-95
        m = plaintext_to_poly(self.par, self.decrypt(ct), ct.level)
        c = self.phase(ct).isub(m).into_power_basis()
        Q = c.ctx.modulus()
        return max(min(v.bit_length(), (Q - v).bit_length()) for v in c.to_bigints())


def plaintext_to_poly(par: BfvParameters, values: Sequence[int], level: int) -> Poly:
    """Plaintext::to_poly, plaintext.rs:172-197 (small plaintext modulus)."""
    lvl = par.level(level)
    t = par.plaintext
    v = np.zeros(par.degree, np.uint64)
    vals = np.array([int(x) % t for x in values], dtype=np.uint64)
    v[: len(vals)] = vals
    tm = Modulus(t)
    lib().orc_zq_scalar_mul_vec(tm.ref(), _p(v), lvl.q_mod_t, par.degree)
    m = Poly.from_u64(lvl.poly_context, v, NTT)
    m.imul(lvl.delta)
    return m


def simd_encode(par: BfvParameters, values: Sequence[int]) -> np.ndarray:
    """Encoding::simd, plaintext_vec.rs:81-95: scatter then inverse NTT mod t."""
    t = par.plaintext
    op = _ntt_op(t, par.degree, par.psi.get(t))
    w = np.zeros(par.degree, np.uint64)
    for i, v in enumerate(values):
        w[par.matrix_reps_index_map[i]] = int(v) % t
    op.backward(w)
    return w


def simd_decode(par: BfvParameters, coeffs: np.ndarray) -> np.ndarray:
    """plaintext.rs:155-170: forward NTT mod t then gather."""
    t = par.plaintext
    op = _ntt_op(t, par.degree, par.psi.get(t))
    w = np.ascontiguousarray(coeffs, dtype=np.uint64).copy()
    op.forward(w)
    return w[np.array(par.matrix_reps_index_map)]


def _ksk_log_base(ctx_ksk: "Context"):
    """key_switching_key.rs:92-97: a single-modulus key level decomposes in base 2^(log_modulus / 2);
    returns (log_base, n_digits), (0, 0) for the RNS-digit variant."""
    if len(ctx_ksk.moduli) != 1:
        return 0, 0
    log_modulus = (ctx_ksk.moduli[0] - 1).bit_length()      # next_power_of_two().ilog2()
    log_base = log_modulus // 2
    return log_base, -(-log_modulus // log_base)


class KeySwitchingKey:
    """keys/key_switching_key.rs:22-362: RNS-digit variant (log_base == 0) and, when the key level has a single
    modulus, the base-2^log_base decomposition variant (:92-110, :196-236, :323-362)."""

    def __init__(self, sk: SecretKey, frm: Poly, ciphertext_level: int, ksk_level: int, rng):
        par = sk.par
        self.par = par
        self.ctx_ksk = par.context_at_level(ksk_level)
        self.ctx_ciphertext = par.context_at_level(ciphertext_level)
        self.ciphertext_level, self.ksk_level = ciphertext_level, ksk_level
        assert frm.ctx == self.ctx_ksk and frm.rep == POWER_BASIS
        self.log_base, n_dec = _ksk_log_base(self.ctx_ksk)
        size = n_dec if self.log_base else len(self.ctx_ciphertext.moduli)
        # generate_c1 (:130-146): uniform NttShoup polys (numpy RNG instead of seeded ChaCha8)
        self.c1 = [Poly.random(self.ctx_ksk, NTT_SHOUP, rng) for _ in range(size)]
        # generate_c0 (:149-194) / generate_c0_decomposition (:196-236)
        s = sk.s_ntt(self.ctx_ksk)
        rns = None if self.log_base else RnsContext(par.moduli[:size])
        self.c0 = []
        for i, c1i in enumerate(self.c1):
            a_s = Poly(self.ctx_ksk, NTT, c1i.c.copy()).imul(s).into_power_basis()
            b = Poly.from_i64(self.ctx_ksk, sample_vec_cbd(par.degree, par.variance, rng))
            b.isub(a_s)
            b.iadd(frm.mul_scalar_big(1 << (i * self.log_base) if self.log_base else rns.garner[i]))
            self.c0.append(b.into_ntt_shoup())

    @staticmethod
    def from_arrays(par: BfvParameters, c0: np.ndarray, c1: np.ndarray, ciphertext_level: int = 0,
                    ksk_level: int = 0) -> "KeySwitchingKey":
        """Rebuild a key from its NTT-domain words (as TryConvertFrom<&KeySwitchingKeyProto>,
        key_switching_key.rs:418-482, minus the seed expansion)."""
        k = KeySwitchingKey.__new__(KeySwitchingKey)
        k.par = par
        k.ctx_ksk = par.context_at_level(ksk_level)
        k.ctx_ciphertext = par.context_at_level(ciphertext_level)
        k.ciphertext_level, k.ksk_level = ciphertext_level, ksk_level
        k.log_base, n_dec = _ksk_log_base(k.ctx_ksk)
        assert len(c0) == len(c1) == (n_dec if k.log_base else len(k.ctx_ciphertext.moduli))
        k.c0 = [Poly(k.ctx_ksk, NTT_SHOUP, a) for a in c0]
        k.c1 = [Poly(k.ctx_ksk, NTT_SHOUP, a) for a in c1]
        return k

    def key_switch(self, p: Poly):  # :241-270, :323-362
        assert p.ctx == self.ctx_ciphertext and p.rep == POWER_BASIS
        c0 = Poly(self.ctx_ksk, NTT)
        c1 = Poly(self.ctx_ksk, NTT)
        if self.log_base:
            coeffs = p.c[0].copy()
            mask = np.uint64((1 << self.log_base) - 1)
            digits = []
            for _ in range(len(self.c0)):
                digits.append(coeffs & mask)
                coeffs = coeffs >> np.uint64(self.log_base)
        else:
            digits = [p.c[i] for i in range(len(self.c0))]
        for i, d in enumerate(digits):
            c2_i = lazy_constant_ntt(d, self.ctx_ksk)
            c0.iadd(c2_i.mul(self.c0[i]))
            c2_i.imul(self.c1[i])
            c1.iadd(c2_i)
        return c0, c1

    def arrays(self):
        """Flat key material: (c0, c1) each [n_digits][n_ksk_limbs][N] (NTT values)."""
        return np.stack([p.c for p in self.c0]), np.stack([p.c for p in self.c1])


class RGSWCiphertext:
    """bfv/rgsw_ciphertext.rs:20-155: encryption of a plaintext polynomial as two key-switching keys (m, m*s)."""

    def __init__(self, sk: SecretKey, m_scaled_ntt: Poly, level: int, rng):
        # FheEncrypter<Plaintext, RGSWCiphertext> for SecretKey (:93-120): m = pt.poly_ntt (unscaled message, NTT)
        ctx = sk.par.context_at_level(level)
        m = m_scaled_ntt.copy().into_power_basis()
        m_s = sk.s_ntt(ctx).imul(m_scaled_ntt).into_power_basis()
        self.ksk0 = KeySwitchingKey(sk, m, level, level, rng)
        self.ksk1 = KeySwitchingKey(sk, m_s, level, level, rng)
        self.level = level

    def external_product(self, ct: Ciphertext) -> Ciphertext:  # :122-155
        assert ct.level == self.level and len(ct.c) == 2
        ct0 = ct.c[0].copy().into_power_basis()
        ct1 = ct.c[1].copy().into_power_basis()
        c0, c1 = self.ksk0.key_switch(ct0)
        c0p, c1p = self.ksk1.key_switch(ct1)
        return Ciphertext(ct.par, [c0.iadd(c0p), c1.iadd(c1p)], ct.level)


def _post_key_switch(c0: Poly, c1: Poly, target: Context):
    """relinearization_key.rs:88-95 / galois_key.rs:69-76."""
    if c0.ctx != target:
        c0 = c0.into_power_basis().switch_down_to(target).into_ntt()
        c1 = c1.into_power_basis().switch_down_to(target).into_ntt()
    return c0, c1


class RelinearizationKey:
    """keys/relinearization_key.rs:23-111."""

    def __init__(self, sk: SecretKey, rng, ciphertext_level: int = 0, key_level: int = 0):
        par = sk.par
        ctx_rk = par.context_at_level(key_level)
        ctx_ct = par.context_at_level(ciphertext_level)
        if len(ctx_rk.moduli) == 1:
            raise ValueError("KeySwitchingNotSupported")
        s = sk.s_ntt(ctx_ct)
        s2 = s.mul(s).into_power_basis()
        s2_up = Switcher(ctx_ct, ctx_rk).switch(s2)
        self.ksk = KeySwitchingKey(sk, s2_up, ciphertext_level, key_level, rng)

    @staticmethod
    def from_ksk(ksk: "KeySwitchingKey") -> "RelinearizationKey":
        rk = RelinearizationKey.__new__(RelinearizationKey)
        rk.ksk = ksk
        return rk

    def relinearizes(self, ct: Ciphertext) -> Ciphertext:  # :70-103
        assert len(ct.c) == 3 and ct.level == self.ksk.ciphertext_level
        c2 = ct.c[2].copy().into_power_basis()
        c0, c1 = self.ksk.key_switch(c2)
        c0, c1 = _post_key_switch(c0, c1, ct.c[0].ctx)
        return Ciphertext(ct.par, [ct.c[0].copy().iadd(c0), ct.c[1].copy().iadd(c1)], ct.level)


class GaloisKey:
    """keys/galois_key.rs:18-124."""

    def __init__(self, sk: SecretKey, exponent: int, rng, ciphertext_level: int = 0, key_level: int = 0):
        par = sk.par
        ctx_gk = par.context_at_level(key_level)
        ctx_ct = par.context_at_level(ciphertext_level)
        self.exponent = exponent % (2 * par.degree)
        s = Poly.from_i64(ctx_ct, sk.coeffs)
        s_sub = s.substitute(self.exponent)
        s_sub_up = Switcher(ctx_ct, ctx_gk).switch(s_sub)
        self.ksk = KeySwitchingKey(sk, s_sub_up, ciphertext_level, key_level, rng)

    def relinearize(self, ct: Ciphertext) -> Ciphertext:  # :63-86
        assert len(ct.c) == 2 and ct.level == self.ksk.ciphertext_level
        c2 = ct.c[1].substitute(self.exponent).into_power_basis()
        c0, c1 = self.ksk.key_switch(c2)
        c0, c1 = _post_key_switch(c0, c1, ct.c[0].ctx)
        c0.iadd(ct.c[0].substitute(self.exponent))
        return Ciphertext(ct.par, [c0, c1], self.ksk.ciphertext_level)


def dot_product_scalar(cts: Sequence["Ciphertext"], pts: Sequence["Poly"]) -> "Ciphertext":
    """bfv/ops/dot_product.rs:55-184: sum_i ct_i (.) pt_i with wide accumulators and one reduction per coefficient
    (both branches of the reference -- u128 fma below the 2^(2*lz) term threshold, rq::dot_product above it -- give the
    canonical residue of the exact sum)."""
    cts, pts = list(cts), list(pts)
    if not cts or not pts:
        raise ValueError("EmptyInput")
    if len(cts) != len(pts):
        raise ValueError("OperandCountMismatch")
    first = cts[0]
    ctx = first.par.context_at_level(first.level)
    for ct, pt in zip(cts, pts):
        if ct.level != first.level or pt.ctx != ctx or pt.rep != NTT:
            raise ValueError("InvalidLevel")
        if len(ct.c) != len(first.c):
            raise ValueError("CiphertextPolynomialCountMismatch")
    out = []
    for part in range(len(first.c)):
        acc = np.zeros((len(ctx.moduli), first.par.degree), dtype=object)
        for ct, pt in zip(cts, pts):
            acc += ct.c[part].c.astype(object) * pt.c.astype(object)
        rows = np.zeros(acc.shape, np.uint64)
        for i, q in enumerate(ctx.moduli):
            rows[i] = np.array([int(v) % q for v in acc[i]], dtype=np.uint64)
        out.append(Poly(ctx, NTT, rows))
    return Ciphertext(first.par, out, first.level)


def computes_inner_sum(par: BfvParameters, gks: dict, ct: Ciphertext) -> Ciphertext:
    """EvaluationKey::computes_inner_sum, evaluation_key.rs:56-100 (gks: exponent -> GaloisKey)."""
    out = ct.copy()
    i = 1
    while i < par.degree // 2:
        out = out.add(gks[pow(3, i, 2 * par.degree)].relinearize(out))
        i *= 2
    return out.add(gks[2 * par.degree - 1].relinearize(out))


def expansion_monomial(par: BfvParameters, l: int, level: int = 0) -> Poly:
    """evaluation_key.rs:465-474: -x^(N - 2^l) as Poly<NttShoup>."""
    v = np.zeros(par.degree, np.int64)
    v[par.degree - (1 << l)] = -1
    return Poly.from_i64(par.context_at_level(level), v).into_ntt_shoup()


def expands(par: BfvParameters, gks: dict, ct: Ciphertext, size: int) -> List[Ciphertext]:
    """EvaluationKey::expands, evaluation_key.rs:192-256."""
    level = (size - 1).bit_length()
    out: List[Optional[Ciphertext]] = [None] * (1 << level)
    out[0] = ct.copy()
    for l in range(level):
        mono = expansion_monomial(par, l, ct.level)
        gk = gks[(par.degree >> l) + 1]
        step = 1 << l
        for i in range(step):
            sub = gk.relinearize(out[i])
            j = step | i
            if j < size:
                tgt = out[i].sub(sub)
                tgt.c = [p.imul(mono) for p in tgt.c]
                out[j] = tgt
            out[i] = out[i].add(sub)
    return out[:size]


def rotation_exponent(par: BfvParameters, i: int) -> int:
    """column rotation by i <-> 3^i mod 2N (evaluation_key.rs:278-286); row swap <-> 2N-1 (:118)."""
    return pow(3, i, 2 * par.degree)


class Multiplicator:
    """bfv::Multiplicator, ops/mul.rs:22-243."""

    def __init__(self, par: BfvParameters, lhs: ScalingFactor, rhs: ScalingFactor,
                 extended_basis: Sequence[int], post: ScalingFactor, level: int = 0):
        self.par, self.level = par, level
        self.base_ctx = par.context_at_level(level)
        self.mul_ctx = Context(list(extended_basis), par.degree, par.psi)
        self.extender_lhs = Scaler(self.base_ctx, self.mul_ctx, lhs)
        self.extender_rhs = Scaler(self.base_ctx, self.mul_ctx, rhs)
        self.down_scaler = Scaler(self.mul_ctx, self.base_ctx, post)
        self.rk: Optional[RelinearizationKey] = None
        self.mod_switch = False

    @staticmethod
    def default(rk: RelinearizationKey) -> "Multiplicator":  # mul.rs:101-138
        par = rk.ksk.par
        ctx = par.context_at_level(rk.ksk.ciphertext_level)
        L = len(ctx.moduli)
        n_moduli = -(-(sum(par.moduli_sizes[:L]) + 60) // 62)
        ext = list(ctx.moduli)
        ub = 1 << 62
        while len(ext) != L + n_moduli:
            ub = generate_prime(62, 2 * par.degree, ub)
            if ub is None:
                raise ValueError("NotEnoughPrimes")
            if ub not in ext and ub not in ctx.moduli:
                ext.append(ub)
        m = Multiplicator(par, ScalingFactor.one(), ScalingFactor.one(), ext,
                          ScalingFactor(par.plaintext, ctx.modulus()), rk.ksk.ciphertext_level)
        m.enable_relinearization(rk)
        return m

    def enable_relinearization(self, rk: RelinearizationKey):  # :141-151
        if self.par.context_at_level(rk.ksk.ciphertext_level) != self.base_ctx:
            raise ValueError("ParameterMismatch")
        self.rk = rk

    def enable_mod_switching(self):  # :155-162
        if self.par.context_at_level(self.par.max_level()) == self.base_ctx:
            raise ValueError("NoMoreContext")
        self.mod_switch = True

    def multiply(self, lhs: Ciphertext, rhs: Ciphertext) -> Ciphertext:  # :165-243
        if lhs.level != self.level or rhs.level != self.level:
            raise ValueError("InvalidLevel")
        if len(lhs.c) != 2 or len(rhs.c) != 2:
            raise ValueError("MultiplicationPolynomialCount")
        c00 = self.extender_lhs.scale(lhs.c[0]); c01 = self.extender_lhs.scale(lhs.c[1])
        c10 = self.extender_rhs.scale(rhs.c[0]); c11 = self.extender_rhs.scale(rhs.c[1])
        c0 = c00.mul(c10)
        c1 = c00.mul(c11).iadd(c01.mul(c10))
        c2 = c01.mul(c11)
        c = [self.down_scaler.scale(x) for x in (c0, c1, c2)]
        if self.rk is not None:
            c2_pb = c[2].copy().into_power_basis()
            c0r, c1r = self.rk.ksk.key_switch(c2_pb)
            c0r, c1r = _post_key_switch(c0r, c1r, c[0].ctx)
            c[0].iadd(c0r)
            c[1].iadd(c1r)
            c = c[:2]
        out = Ciphertext(self.par, c, self.level)
        if self.mod_switch:
            out.switch_down()
        return out
