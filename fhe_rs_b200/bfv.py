"""Host-side mirror of the reference's `fhe::bfv` interface for the accelerated path
(BfvParameters, Ciphertext, Multiplicator, RelinearizationKey, GaloisKey / EvaluationKey),
implemented purely on top of the C ABI in include/fhe_b200.h.

Same names, argument meaning and error behaviour as the reference items cited in each
docstring (paths relative to /root/reference/crates).  Differences forced by the device:
a `Ciphertext` here is a *batch* of ciphertexts of one level resident in HBM (the reference's
operators act on one ciphertext; per-ciphertext FFI would be launch/PCIe bound),
and keys are constructed from their NTT-domain words (as after deserialization,
key_switching_key.rs:418-482) -- key generation is client-side code outside this path.
No CPU fallback exists: every operation is a CUDA launch behind the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _capi, wire
from ._capi import NTT, POWER_BASIS, FheError, check
from .wire import WireError

__all__ = ["BfvParameters", "BfvParametersBuilder", "Ciphertext", "KeySwitchingKey", "RelinearizationKey", "RGSWCiphertext",
           "GaloisKey", "EvaluationKey", "Multiplicator", "ScalingFactor", "dot_product_scalar", "FheError", "WireError", "NTT", "POWER_BASIS"]


def _release(free_name: str, handle) -> None:
    """Call a C-ABI destructor from __del__; at interpreter shutdown the module globals may already be gone, in which
    case the process is about to release everything anyway."""
    try:
        getattr(_capi.lib(), free_name)(handle)
    except Exception:  # noqa: BLE001
        pass


def _ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class BfvParameters:
    """fhe::bfv::BfvParameters (bfv/parameters.rs:88-114), built by
    BfvParametersBuilder::build (:560-738).  `device=-1` builds the host tables only."""

    def __init__(self, degree: int, plaintext_modulus: int, moduli: Optional[Sequence[int]] = None,
                 moduli_sizes: Optional[Sequence[int]] = None, psi: Optional[Sequence[int]] = None,
                 device: int = 0):
        L = _capi.lib()
        if (moduli is None) == (moduli_sizes is None):
            # parameters.rs:455-466
            raise FheError(_capi.INVALID_ARGUMENT, "exactly one of moduli / moduli_sizes must be given")
        pt = int(plaintext_modulus)
        if pt <= 0:
            raise FheError(_capi.INVALID_ARGUMENT, "plaintext modulus must be positive")
        pt_bytes = pt.to_bytes(max(1, (pt.bit_length() + 7) // 8), "little")
        buf = (C.c_uint8 * len(pt_bytes)).from_buffer_copy(pt_bytes)
        h = C.c_void_p()
        if moduli is not None:
            m = np.ascontiguousarray(np.array([int(x) for x in moduli], dtype=np.uint64))
            ps = None
            if psi is not None:
                ps = np.ascontiguousarray(np.array([int(x) for x in psi], dtype=np.uint64))
                if len(ps) != 2 * len(m) + 1:
                    raise FheError(_capi.INVALID_ARGUMENT, "psi needs one root per modulus and extension prime")
            check(L.fhe_b200_params_create(device, degree, _ptr(m), len(m), C.addressof(buf), len(pt_bytes),
                                           _ptr(ps) if ps is not None else None, C.byref(h)))
        else:
            if psi is not None:
                raise FheError(_capi.INVALID_ARGUMENT, "psi requires explicit moduli")
            s = np.ascontiguousarray(np.array(list(moduli_sizes), dtype=np.uint32))
            check(L.fhe_b200_params_create_from_sizes(device, degree, s.ctypes.data, len(s), C.addressof(buf),
                                                      len(pt_bytes), C.byref(h)))
        self._h = h
        self.device = device
        self._plaintext = pt
        self._degree = L.fhe_b200_params_degree(h)
        n = L.fhe_b200_params_n_moduli(h)
        out = np.zeros(n, np.uint64)
        check(L.fhe_b200_params_moduli(h, _ptr(out)))
        self._moduli = [int(x) for x in out]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _release("fhe_b200_params_destroy", h)

    def degree(self) -> int:  # parameters.rs:130
        return self._degree

    def moduli(self):  # parameters.rs:136
        return list(self._moduli)

    def plaintext(self) -> int:
        return self._plaintext

    def max_level(self) -> int:  # parameters.rs:173
        return len(self._moduli) - 1

    def mul_basis(self, level: int = 0):
        """level moduli followed by the extension primes (parameters.rs:660-700)."""
        L = _capi.lib()
        n = C.c_uint32()
        check(L.fhe_b200_params_mul_basis(self._h, level, None, C.byref(n)))
        out = np.zeros(n.value, np.uint64)
        check(L.fhe_b200_params_mul_basis(self._h, level, _ptr(out), C.byref(n)))
        return [int(x) for x in out]

    def psi(self, q: int) -> int:
        r = C.c_uint64()
        check(_capi.lib().fhe_b200_params_psi(self._h, q, C.byref(r)))
        return r.value

    def scaler_tables(self, level: int, which: int) -> Dict[str, object]:
        """Host precompute inspection: RnsScaler tables (rns/scaler.rs:52-73)."""
        L = _capi.lib()
        nf, nt, sh = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(L.fhe_b200_debug_scaler_tables(self._h, level, which, C.byref(nf), C.byref(nt), C.byref(sh),
                                             *([None] * 8)))
        f, t = nf.value, nt.value
        gamma, omega, tg = np.zeros(t, np.uint64), np.zeros((t, f), np.uint64), np.zeros(3, np.uint64)
        tol, toh, tos = np.zeros(f, np.uint64), np.zeros(f, np.uint64), np.zeros(f, np.uint8)
        tgl, tgh = np.zeros(f, np.uint64), np.zeros(f, np.uint64)
        check(L.fhe_b200_debug_scaler_tables(self._h, level, which, C.byref(nf), C.byref(nt), C.byref(sh),
                                             _ptr(gamma), _ptr(omega), _ptr(tg), _ptr(tol), _ptr(toh), _ptr(tos),
                                             _ptr(tgl), _ptr(tgh)))
        return dict(n_from=f, n_to=t, shift=sh.value, gamma=gamma, omega=omega, theta_gamma=tg,
                    theta_omega_lo=tol, theta_omega_hi=toh, theta_omega_sign=tos,
                    theta_garner_lo=tgl, theta_garner_hi=tgh)

    def ntt_tables(self, q: int) -> Dict[str, object]:
        n = self._degree
        om, oms, zi, zis = (np.zeros(n, np.uint64) for _ in range(4))
        ninv = C.c_uint64()
        check(_capi.lib().fhe_b200_debug_ntt_tables(self._h, q, _ptr(om), _ptr(oms), _ptr(zi), _ptr(zis),
                                                    C.byref(ninv)))
        return dict(omegas=om, omegas_shoup=oms, zetas_inv=zi, zetas_inv_shoup=zis, size_inv=ninv.value)


class BfvParametersBuilder:
    """fhe::bfv::BfvParametersBuilder (bfv/parameters.rs:319-388)."""

    def __init__(self):
        self._degree = 0
        self._plaintext = 0
        self._moduli = None
        self._sizes = None
        self._psi = None

    def set_degree(self, degree: int):
        self._degree = degree
        return self

    def set_plaintext_modulus(self, t: int):
        self._plaintext = t
        return self

    def set_moduli(self, moduli: Sequence[int]):
        self._moduli = list(moduli)
        return self

    def set_moduli_sizes(self, sizes: Sequence[int]):
        self._sizes = list(sizes)
        return self

    def set_ntt_roots(self, psi: Sequence[int]):
        """2N-th roots per [moduli..., extension primes...] (the reference's own, for interchange)."""
        self._psi = list(psi)
        return self

    def build(self, device: int = 0) -> BfvParameters:
        return BfvParameters(self._degree, self._plaintext, self._moduli, self._sizes, self._psi, device)

    build_arc = build


class Ciphertext:
    """A device-resident batch of fhe::bfv::Ciphertext (bfv/ciphertext.rs:18-32): `count`
    ciphertexts of `parts` polynomials at `level`, words [count][parts][limbs][N]."""

    def __init__(self, par: BfvParameters, count: int, parts: int = 2, level: int = 0, repr: int = NTT,
                 stream: int = 0, mul_basis: bool = False):
        h = C.c_void_p()
        f = _capi.lib().fhe_b200_batch_alloc_mul_basis if mul_basis else _capi.lib().fhe_b200_batch_alloc
        check(f(par._h, count, parts, level, repr, C.byref(h)))
        self._h, self.par, self.stream = h, par, stream

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _release("fhe_b200_batch_free", h)

    # -- shape
    def _info(self):
        c, p, lv, lm, r = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int()
        check(_capi.lib().fhe_b200_batch_info(self._h, C.byref(c), C.byref(p), C.byref(lv), C.byref(lm), C.byref(r)))
        return c.value, p.value, lv.value, lm.value, r.value

    @property
    def count(self):
        return self._info()[0]

    def __len__(self):  # Ciphertext::len -> number of polynomials (ciphertext.rs:118)
        return self._info()[1]

    @property
    def level(self):
        return self._info()[2]

    @property
    def limbs(self):
        return self._info()[3]

    @property
    def representation(self):
        return self._info()[4]

    def shape(self):
        c, p, _, lm, _ = self._info()
        return (c, p, lm, self.par.degree())

    # -- transfer
    @staticmethod
    def from_host(par: BfvParameters, words: np.ndarray, level: int = 0, repr: int = NTT, stream: int = 0,
                  mul_basis: bool = False) -> "Ciphertext":
        """words: u64 [count][parts][limbs][N] as Vec<u64>::from(&Poly) (rq/convert.rs:474-503)."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        if words.ndim != 4 or words.shape[3] != par.degree():
            raise FheError(_capi.INVALID_ARGUMENT, "expected [count][parts][limbs][N] words")
        ct = Ciphertext(par, words.shape[0], words.shape[1], level, repr, stream, mul_basis)
        if ct.limbs != words.shape[2]:
            raise FheError(_capi.CONTEXT_MISMATCH, "limb count does not match the level")
        check(_capi.lib().fhe_b200_batch_upload(ct._h, 0, words.shape[0], _ptr(words), stream))
        check(_capi.lib().fhe_b200_sync(stream))
        return ct

    def _check_words(self, words: np.ndarray, first: int):
        """[n][parts][limbs][N] contiguous u64 with first + n <= count -- anything else would overrun a buffer"""
        shape = self.shape()
        if words.dtype != np.uint64 or not words.flags["C_CONTIGUOUS"]:
            raise FheError(_capi.INVALID_ARGUMENT, "expected a C-contiguous uint64 array")
        if words.ndim != 4 or tuple(words.shape[1:]) != tuple(shape[1:]):
            raise FheError(_capi.INVALID_ARGUMENT, "expected [n][%d][%d][%d] words" % tuple(shape[1:]))
        if first < 0 or first + words.shape[0] > shape[0]:
            raise FheError(_capi.INVALID_ARGUMENT, "range exceeds batch")

    def upload(self, words: np.ndarray, first: int = 0):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        self._check_words(words, first)
        check(_capi.lib().fhe_b200_batch_upload(self._h, first, words.shape[0], _ptr(words), self.stream))
        check(_capi.lib().fhe_b200_sync(self.stream))   # `words` may be a temporary: do not return before it is read

    def to_host(self, out: Optional[np.ndarray] = None, first: int = 0) -> np.ndarray:
        """ciphertexts [first, first + len(out)) (all of them from `first` on when out is None)"""
        if out is None:
            shape = self.shape()
            out = np.empty((shape[0] - first,) + tuple(shape[1:]), np.uint64)
        self._check_words(out, first)
        check(_capi.lib().fhe_b200_batch_download(self._h, first, out.shape[0], _ptr(out), self.stream))
        return out

    # -- wire format (rq/convert.rs:17-131): Rq.coefficients blobs, one per polynomial
    def packed_bytes_per_poly(self) -> int:
        n = C.c_size_t()
        check(_capi.lib().fhe_b200_batch_packed_bytes(self._h, C.byref(n)))   # per batch: a mul-basis batch has L + E limbs
        return n.value

    def to_packed(self) -> np.ndarray:
        """[count][parts][packed_bytes] uint8: what `Rq::from(&poly).coefficients` holds for every polynomial"""
        c, p, _, _, _ = self._info()
        out = np.empty((c, p, self.packed_bytes_per_poly()), np.uint8)
        check(_capi.lib().fhe_b200_batch_pack(self._h, 0, c, _ptr(out), self.stream))
        return out

    @staticmethod
    def from_packed(par: "BfvParameters", blobs: np.ndarray, level: int = 0, repr: int = NTT,
                    stream: int = 0) -> "Ciphertext":
        """inverse of to_packed: `Poly::<R>::try_convert_from(&Rq, ctx, ..)` for every polynomial"""
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        ct = Ciphertext(par, blobs.shape[0], blobs.shape[1], level, repr, stream)
        if blobs.shape[2] != ct.packed_bytes_per_poly():
            raise FheError(_capi.INVALID_ARGUMENT, "InvalidCoefficientCount: blob size does not match the context")
        check(_capi.lib().fhe_b200_batch_unpack(ct._h, 0, blobs.shape[0], _ptr(blobs), stream))
        check(_capi.lib().fhe_b200_sync(stream))
        return ct

    # -- protobuf messages (bfv/ciphertext.rs:230-317; fhe_traits::Serialize / DeserializeParametrized)
    def to_bytes(self) -> list:
        """`ct.to_bytes()` for every ciphertext of the batch: one encoded fhers.bfv.Ciphertext each, every part an
        fhers.rq.Rq with representation NTT (the unseeded branch of ciphertext.rs:240-252).  The coefficient packing
        runs on the device; only the few bytes of framing are host work."""
        c, p, lv, _, r = self._info()
        if r != NTT:
            raise FheError(_capi.INVALID_REPRESENTATION, "Ciphertext polynomials are Poly<Ntt> (ciphertext.rs:18-32)")
        blobs = self.to_packed()
        self.sync()
        deg = self.par.degree()
        return [wire.encode_ciphertext([wire.encode_rq(wire.REP_NTT, deg, memoryview(blobs[i, j])) for j in range(p)],
                                       b"", lv) for i in range(c)]

    @staticmethod
    def from_bytes(par: "BfvParameters", messages: Sequence[bytes], seeded_halves: Optional[np.ndarray] = None,
                   stream: int = 0) -> "Ciphertext":
        """`Ciphertext::from_bytes(bytes, &par)` (ciphertext.rs:259-317) for a batch of messages of one level and part
        count.  A message that carries a seed instead of its last polynomial needs `seeded_halves[i]`: the NTT words
        [limbs][N] of `Poly::random_from_seed(ctx, seed)`, expanded by the Rust host (include/fhe_b200.h explains why
        the device does not)."""
        if len(messages) == 0:
            raise FheError(_capi.INVALID_ARGUMENT, "no messages")
        dec = [wire.decode_ciphertext(m) for m in messages]
        level = dec[0][2]
        if level > par.max_level():
            raise WireError("InvalidLevel", _capi.INVALID_LEVEL, "level %d, max %d" % (level, par.max_level()))
        n_rq = len(dec[0][0])
        seeded = bool(dec[0][1])
        for c, seed, lv in dec:
            if lv != level or len(c) != n_rq or bool(seed) != seeded:
                raise FheError(_capi.INVALID_ARGUMENT, "a batch holds ciphertexts of one level, part count and kind")
            if seed and len(seed) != 32:
                raise WireError("InvalidSeedSize", detail="%d bytes, expected 32" % len(seed))
        body = Ciphertext(par, len(dec), n_rq, level, NTT, stream)
        _unpack_rq(body, [c for c, _, _ in dec], wire.REP_NTT)
        if not seeded:
            return body
        if seeded_halves is None:
            raise WireError("SeedExpansion", _capi.UNSUPPORTED,
                            "the message carries a seed: pass the host-expanded last polynomial (ciphertext.rs:287-300)")
        halves = np.ascontiguousarray(seeded_halves, dtype=np.uint64)
        if halves.shape != (len(dec), body.limbs, par.degree()):
            raise FheError(_capi.INVALID_ARGUMENT, "expected seeded_halves as [count][limbs][N]")
        words = np.concatenate([body.to_host(), halves[:, None]], axis=1)
        return Ciphertext.from_host(par, words, level, NTT, stream)

    def device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_size_t()
        check(_capi.lib().fhe_b200_batch_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value

    def sync(self):
        check(_capi.lib().fhe_b200_sync(self.stream))

    def _like(self, parts: Optional[int] = None, level: Optional[int] = None) -> "Ciphertext":
        c, p, lv, _, r = self._info()
        return Ciphertext(self.par, c, parts or p, lv if level is None else level, r, self.stream)

    def clone(self) -> "Ciphertext":
        out = self._like()
        check(_capi.lib().fhe_b200_batch_copy(out._h, self._h, self.stream))
        return out

    # -- representation (Poly::into_ntt / into_power_basis, rq/mod.rs:535, :590)
    def into_ntt(self) -> "Ciphertext":
        check(_capi.lib().fhe_b200_ntt_forward(self._h, self.stream))
        return self

    def into_power_basis(self) -> "Ciphertext":
        check(_capi.lib().fhe_b200_ntt_backward(self._h, self.stream))
        return self

    # -- operators (bfv/ops/mod.rs:15-358)
    def __iadd__(self, rhs: "Ciphertext"):
        check(_capi.lib().fhe_b200_add(self._h, rhs._h, self.stream))
        return self

    def __isub__(self, rhs: "Ciphertext"):
        check(_capi.lib().fhe_b200_sub(self._h, rhs._h, self.stream))
        return self

    def __add__(self, rhs: "Ciphertext") -> "Ciphertext":
        out = self.clone()
        out += rhs
        return out

    def __sub__(self, rhs: "Ciphertext") -> "Ciphertext":
        out = self.clone()
        out -= rhs
        return out

    def __neg__(self) -> "Ciphertext":
        out = self.clone()
        check(_capi.lib().fhe_b200_neg(out._h, out.stream))
        return out

    def __mul__(self, rhs: "Ciphertext") -> "Ciphertext":
        """&Ciphertext * &Ciphertext, no relinearization (ops/mod.rs:259-358): n x m parts -> n + m - 1 parts."""
        out = self._like(parts=len(self) + len(rhs) - 1)
        check(_capi.lib().fhe_b200_mul(self._h, rhs._h, out._h, self.stream))
        return out

    def mul_plain(self, poly_ntt: np.ndarray) -> "Ciphertext":
        """Ciphertext *= &Plaintext (ops/mod.rs:229-238), in place.  poly_ntt: the plaintext's `poly_ntt` words,
        [limbs][N] (shared by the batch) or [count][limbs][N] (one plaintext per ciphertext)."""
        w = np.ascontiguousarray(poly_ntt, dtype=np.uint64)
        n = 1 if w.ndim == 2 else w.shape[0]
        check(_capi.lib().fhe_b200_mul_plain(self._h, _ptr(w), n, self.stream))
        return self

    def switch_down(self) -> "Ciphertext":
        """Ciphertext::switch_down (ciphertext.rs:148-161), in place."""
        check(_capi.lib().fhe_b200_switch_down(self._h, self.stream))
        return self

    def add_plain(self, poly: np.ndarray, subtract: bool = False) -> "Ciphertext":
        """Ciphertext += &Plaintext / -= &Plaintext (ops/mod.rs:88-97, :188-197), in place: `poly` is the plaintext's
        `to_poly()` words (delta-scaled, NTT), [limbs][N] shared by the batch or [count][limbs][N]."""
        w = np.ascontiguousarray(poly, dtype=np.uint64)
        n = 1 if w.ndim == 2 else w.shape[0]
        check(_capi.lib().fhe_b200_add_plain(self._h, _ptr(w), n, 1 if subtract else 0, self.stream))
        return self

    def sub_plain(self, poly: np.ndarray) -> "Ciphertext":
        return self.add_plain(poly, subtract=True)

    def max_switchable_level(self) -> int:  # ciphertext.rs:187-189
        return self.par.max_level()

    def switch_to_level(self, target_level: int) -> "Ciphertext":
        """Ciphertext::switch_to_level (ciphertext.rs:164-184): only moves down; InvalidLevel otherwise."""
        if target_level < self.level or target_level > self.max_switchable_level():
            raise FheError(_capi.INVALID_LEVEL, "InvalidLevel: level %d, min %d, max %d"
                           % (target_level, self.level, self.max_switchable_level()))
        while self.level < target_level:
            self.switch_down()
        return self

    def substitute(self, exponent: int) -> "Ciphertext":
        """Poly::substitute on every polynomial, in either representation (rq/mod.rs:360-408)."""
        out = self._like()
        check(_capi.lib().fhe_b200_substitute(self._h, exponent, out._h, self.stream))
        return out

    def scale(self, which: int) -> "Ciphertext":
        """Poly::scale with the level's extender (0) / down scaler (1) (rq/mod.rs:669, rq/scaler.rs:55)."""
        c, p, lv, _, r = self._info()
        out = Ciphertext(self.par, c, p, lv, r, self.stream, mul_basis=(which == 0))
        check(_capi.lib().fhe_b200_scale(self._h, which, out._h, self.stream))
        return out


def _unpack_rq(batch: "Ciphertext", rq_messages, want_rep: int) -> None:
    """`Poly::<R>::from_bytes(bytes, ctx)` (rq/serialize.rs:23-31 -> rq/convert.rs:46-161) for every polynomial of
    `batch`: rq_messages[i][j] is the encoded Rq of part j of ciphertext i.  Framing and checks here, unpacking (and
    the forward NTT of an NTT batch) on the device."""
    par, nbytes = batch.par, batch.packed_bytes_per_poly()
    count, parts = batch.count, len(batch)
    limbs, deg = batch.limbs, par.degree()
    blobs = np.zeros((count, parts, nbytes), np.uint8)
    for i, polys in enumerate(rq_messages):
        for j, msg in enumerate(polys):
            rep, degree, coeffs = wire.decode_rq(msg)
            if degree * nbytes != len(coeffs) * deg:          # sum_i serialization_length(degree) (convert.rs:76-88)
                raise WireError("InvalidCoefficientCount", detail="%d bytes for degree %d" % (len(coeffs), degree))
            if rep != want_rep:
                raise WireError("RepresentationMismatch", _capi.INVALID_REPRESENTATION,
                                "found %d, expected %d" % (rep, want_rep))
            if degree != deg and (limbs != 1 or degree > deg):
                # TryConvertFrom<Vec<u64>> for Poly<PowerBasis> (convert.rs:148-192): q.len() * degree words, or -- one
                # modulus only -- a shorter low-order polynomial, zero-extended (the zero bytes already in `blobs`)
                raise WireError("InvalidCoefficientCount", detail="degree %d in a context of degree %d" % (degree, deg))
            blobs[i, j, :len(coeffs)] = np.frombuffer(coeffs, np.uint8)
    check(_capi.lib().fhe_b200_batch_unpack(batch._h, 0, count, _ptr(blobs), batch.stream))
    check(_capi.lib().fhe_b200_sync(batch.stream))


class KeySwitchingKey:
    """fhe::bfv::KeySwitchingKey (keys/key_switching_key.rs:22-45) from its NTT-domain words:
    c0, c1 = [n_digits][ksk_limbs][N] (the `coefficients` of the Poly<NttShoup> elements)."""

    def __init__(self, par: BfvParameters, c0: np.ndarray, c1: np.ndarray, ciphertext_level: int = 0,
                 ksk_level: int = 0):
        c0 = np.ascontiguousarray(c0, dtype=np.uint64)
        c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        if c0.shape != c1.shape or c0.ndim != 3 or c0.shape[2] != par.degree():
            raise FheError(_capi.INVALID_ARGUMENT, "expected c0, c1 as [digits][limbs][N]")
        if c0.shape[1] != len(par.moduli()) - ksk_level:
            raise FheError(_capi.CONTEXT_MISMATCH, "key limb count does not match ksk_level")
        h = C.c_void_p()
        check(_capi.lib().fhe_b200_ksk_upload(par._h, ciphertext_level, ksk_level, _ptr(c0), _ptr(c1),
                                              c0.shape[0], C.byref(h)))
        self._h, self.par = h, par
        self.ciphertext_level, self.ksk_level = ciphertext_level, ksk_level
        self._words = (c0, c1)                   # the caller's key material, for to_bytes (no device read-back entry)
        # key_switching_key.rs:92-97: a key level with one modulus decomposes in base 2^(log_modulus / 2)
        self.log_base = ((int(par.moduli()[0]) - 1).bit_length() // 2) if c0.shape[1] == 1 else 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _release("fhe_b200_ksk_free", h)

    # -- protobuf message (keys/key_switching_key.rs:365-482)
    def to_bytes(self) -> bytes:
        """KeySwitchingKeyProto::from(&ksk).encode_to_vec(), unseeded branch: every c0_i and c1_i as an Rq with
        representation NTTSHOUP (its coefficients are the power-basis words, packed on the device)."""
        c0, c1 = self._words
        par, nd = self.par, c0.shape[0]
        tmp = Ciphertext.from_host(par, np.ascontiguousarray(np.stack([c0, c1], axis=1)), self.ksk_level, NTT)
        blobs = tmp.to_packed()
        tmp.sync()
        deg = par.degree()
        enc = [[wire.encode_rq(wire.REP_NTTSHOUP, deg, memoryview(blobs[i, j])) for i in range(nd)] for j in range(2)]
        return wire.encode_ksk(enc[0], enc[1], b"", self.ciphertext_level, self.ksk_level, self.log_base)

    @staticmethod
    def from_bytes(par: "BfvParameters", data: bytes, seeded_c1: Optional[np.ndarray] = None) -> "KeySwitchingKey":
        """KeySwitchingKey::try_convert_from(&KeySwitchingKeyProto, par) (key_switching_key.rs:388-482).  A key whose
        c1 row travels as a seed needs `seeded_c1`: [digits][limbs][N] NTT words of generate_c1 (:130-146), expanded by
        the Rust host."""
        k = wire.decode_ksk(data)
        ct_level, ksk_level, log_base = k["ciphertext_level"], k["ksk_level"], k["log_base"]
        for lv in (ksk_level, ct_level):
            if lv > par.max_level():
                raise WireError("InvalidLevel", _capi.INVALID_LEVEL, "level %d, max %d" % (lv, par.max_level()))
        if log_base != 0:
            if ksk_level != par.max_level() or ct_level != par.max_level():
                raise WireError("InvalidKeySwitchingDecompositionLevels", _capi.INVALID_LEVEL)
            # as coded (:406-408): the first modulus of the parameter set sizes the decomposition
            log_modulus = (int(par.moduli()[0]) - 1).bit_length()
            c0_size = -(-log_modulus // log_base)
        else:
            c0_size = len(par.moduli()) - ct_level
        if len(k["c0"]) != c0_size:
            raise WireError("WrongPolynomialCount", _capi.BAD_POLY_COUNT,
                            "KeySwitchingKeyC0: expected %d, found %d" % (c0_size, len(k["c0"])))
        seed = k["seed"]
        if not seed:
            if len(k["c1"]) != c0_size:
                raise WireError("WrongPolynomialCount", _capi.BAD_POLY_COUNT,
                                "KeySwitchingKeyC1: expected %d, found %d" % (c0_size, len(k["c1"])))
            tmp = Ciphertext(par, c0_size, 2, ksk_level, NTT)
            _unpack_rq(tmp, list(zip(k["c0"], k["c1"])), wire.REP_NTTSHOUP)
            words = tmp.to_host()
            tmp.sync()
            c0, c1 = np.ascontiguousarray(words[:, 0]), np.ascontiguousarray(words[:, 1])
        else:
            if len(seed) != 32:
                raise WireError("InvalidKeySwitchingSeedLength", detail="%d bytes, expected 32" % len(seed))
            if seeded_c1 is None:
                raise WireError("SeedExpansion", _capi.UNSUPPORTED,
                                "the key carries a seed: pass the host-expanded c1 row (key_switching_key.rs:130-146)")
            tmp = Ciphertext(par, c0_size, 1, ksk_level, NTT)
            _unpack_rq(tmp, [(m,) for m in k["c0"]], wire.REP_NTTSHOUP)
            c0 = np.ascontiguousarray(tmp.to_host()[:, 0])
            tmp.sync()
            c1 = np.ascontiguousarray(seeded_c1, dtype=np.uint64)
            if c1.shape != c0.shape:
                raise FheError(_capi.INVALID_ARGUMENT, "expected seeded_c1 as [digits][limbs][N]")
        key = KeySwitchingKey(par, c0, c1, ct_level, ksk_level)
        if key.log_base != log_base:
            # the reference would compute with whatever base the message names; the device derives the base from the
            # key level (fhe_b200_ksk_upload), so a message that disagrees is refused rather than reinterpreted
            raise WireError("InvalidKeySwitchingDecompositionLevels", _capi.UNSUPPORTED,
                            "log_base %d does not match the key level (expected %d)" % (log_base, key.log_base))
        return key

    @staticmethod
    def from_arrays(par: BfvParameters, c0, c1, ciphertext_level: int = 0, key_level: int = 0) -> "KeySwitchingKey":
        return KeySwitchingKey(par, c0, c1, ciphertext_level, key_level)

    def key_switch(self, p: Ciphertext, part: int = 0) -> Ciphertext:
        """KeySwitchingKey::key_switch (key_switching_key.rs:241-270, :323-362) on polynomial `part` of a
        POWER_BASIS batch; returns the (c0, c1) pair as a 2-part NTT batch at the key level."""
        out = Ciphertext(self.par, p.count, 2, self.ksk_level, NTT, p.stream)
        check(_capi.lib().fhe_b200_key_switch(p._h, part, self._h, out._h, p.stream))
        return out


class RGSWCiphertext:
    """fhe::bfv::RGSWCiphertext (bfv/rgsw_ciphertext.rs:20-24): two key-switching keys (for m and m*s)."""

    def __init__(self, ksk0: KeySwitchingKey, ksk1: KeySwitchingKey):
        if ksk0.ksk_level != ksk0.ciphertext_level or ksk1.ksk_level != ksk1.ciphertext_level \
                or ksk0.ciphertext_level != ksk1.ciphertext_level:
            raise FheError(_capi.INVALID_LEVEL, "RGSW key-switching keys must share one level")  # rgsw_ciphertext.rs:58-70
        self.ksk0, self.ksk1 = ksk0, ksk1

    @staticmethod
    def from_arrays(par: BfvParameters, k0c0, k0c1, k1c0, k1c1, level: int = 0) -> "RGSWCiphertext":
        return RGSWCiphertext(KeySwitchingKey(par, k0c0, k0c1, level, level), KeySwitchingKey(par, k1c0, k1c1, level, level))

    def to_bytes(self) -> bytes:  # rgsw_ciphertext.rs:30-37
        return wire.encode_rgsw(self.ksk0.to_bytes(), self.ksk1.to_bytes())

    @staticmethod
    def from_bytes(par: BfvParameters, data: bytes) -> "RGSWCiphertext":  # rgsw_ciphertext.rs:39-71
        m0, m1 = wire.decode_rgsw(data)
        k0, k1 = KeySwitchingKey.from_bytes(par, m0), KeySwitchingKey.from_bytes(par, m1)
        if k0.ksk_level != k0.ciphertext_level or k0.ciphertext_level != k1.ciphertext_level \
                or k1.ciphertext_level != k1.ksk_level:
            raise WireError("InconsistentKeySwitchingLevels", _capi.INVALID_LEVEL)
        return RGSWCiphertext(k0, k1)

    def external_product(self, ct: Ciphertext) -> Ciphertext:
        """&Ciphertext * &RGSWCiphertext (rgsw_ciphertext.rs:122-155): key-switch both parts, add."""
        if ct.level != self.ksk0.ciphertext_level:
            raise FheError(_capi.INVALID_LEVEL, "Ciphertext and RGSWCiphertext must have the same level")
        if len(ct) != 2:
            raise FheError(_capi.BAD_POLY_COUNT, "Ciphertext must have two parts")
        pb = ct.clone().into_power_basis()
        out = self.ksk0.key_switch(pb, part=0)
        out += self.ksk1.key_switch(pb, part=1)
        return out


class RelinearizationKey:
    """fhe::bfv::RelinearizationKey (keys/relinearization_key.rs:23-26)."""

    def __init__(self, ksk: KeySwitchingKey):
        self.ksk = ksk

    @staticmethod
    def from_arrays(par: BfvParameters, c0, c1, ciphertext_level: int = 0, key_level: int = 0):
        return RelinearizationKey(KeySwitchingKey(par, c0, c1, ciphertext_level, key_level))

    def to_bytes(self) -> bytes:  # relinearization_key.rs:113-119, :137-141
        return wire.encode_relinearization_key(self.ksk.to_bytes())

    @staticmethod
    def from_bytes(par: BfvParameters, data: bytes) -> "RelinearizationKey":  # relinearization_key.rs:121-135
        return RelinearizationKey(KeySwitchingKey.from_bytes(par, wire.decode_relinearization_key(data)))

    def relinearizes(self, ct: Ciphertext) -> Ciphertext:
        """RelinearizationKey::relinearizes (relinearization_key.rs:70-103): (c0,c1,c2) -> (c0,c1).
        (The reference mutates `ct`; a device batch changes shape, so the result is returned.)"""
        out = ct._like(parts=2)
        check(_capi.lib().fhe_b200_relinearize(ct._h, self.ksk._h, out._h, ct.stream))
        return out


class GaloisKey:
    """fhe::bfv::GaloisKey (keys/galois_key.rs:18-22)."""

    def __init__(self, exponent: int, ksk: KeySwitchingKey):
        self.exponent, self.ksk = exponent, ksk

    @staticmethod
    def from_arrays(par: BfvParameters, exponent: int, c0, c1, ciphertext_level: int = 0, key_level: int = 0):
        return GaloisKey(exponent, KeySwitchingKey(par, c0, c1, ciphertext_level, key_level))

    def to_bytes(self) -> bytes:  # galois_key.rs:146-153
        return wire.encode_galois_key(self.ksk.to_bytes(), self.exponent)

    @staticmethod
    def from_bytes(par: BfvParameters, data: bytes) -> "GaloisKey":  # galois_key.rs:155-173
        msg, exponent = wire.decode_galois_key(data)
        ksk = KeySwitchingKey.from_bytes(par, msg)
        exponent %= 2 * par.degree()            # SubstitutionExponent::new (rq/mod.rs:99-106)
        if exponent & 1 == 0:
            raise WireError("InvalidSubstitutionExponent", _capi.INVALID_EXPONENT, str(exponent))
        return GaloisKey(exponent, ksk)

    def relinearize(self, ct: Ciphertext) -> Ciphertext:
        """GaloisKey::relinearize (galois_key.rs:63-86)."""
        out = ct._like()
        check(_capi.lib().fhe_b200_galois(ct._h, self.exponent, self.ksk._h, out._h, ct.stream))
        return out


class EvaluationKey:
    """The rotation subset of fhe::bfv::EvaluationKey (keys/evaluation_key.rs:110-170):
    a map Galois exponent -> GaloisKey."""

    def __init__(self, par: BfvParameters):
        self.par = par
        self.gk: Dict[int, GaloisKey] = {}

    def add_galois_key(self, gk: GaloisKey):
        self.gk[gk.exponent % (2 * self.par.degree())] = gk

    def supports_row_rotation(self) -> bool:
        return (2 * self.par.degree() - 1) in self.gk

    def supports_column_rotation_by(self, i: int) -> bool:
        return pow(3, i, 2 * self.par.degree()) in self.gk

    def supports_inner_sum(self) -> bool:  # evaluation_key.rs:40-53
        n = self.par.degree()
        i, ok = 1, self.supports_row_rotation()
        while i < n // 2:
            ok = ok and self.supports_column_rotation_by(i)
            i *= 2
        return ok

    def computes_inner_sum(self, ct: Ciphertext) -> Ciphertext:
        """EvaluationKey::computes_inner_sum (evaluation_key.rs:56-100)."""
        if not self.supports_inner_sum():
            raise FheError(_capi.INVALID_ARGUMENT, "EvaluationKeyError: inner sum not supported by this key")
        out = ct.clone()
        i = 1
        while i < self.par.degree() // 2:
            out += self.gk[pow(3, i, 2 * self.par.degree())].relinearize(out)
            i *= 2
        out += self.gk[2 * self.par.degree() - 1].relinearize(out)
        return out

    def expands(self, ct: Ciphertext, size: int, monomials: Sequence[np.ndarray]):
        """EvaluationKey::expands (evaluation_key.rs:192-256), oblivious expansion of eprint 2019/1483.
        monomials[l]: NTT words [limbs][N] of -x^(N - 2^l) (evaluation_key.rs:465-474)."""
        n = self.par.degree()
        if size == 0 or size > n:
            raise FheError(_capi.INVALID_ARGUMENT, "EvaluationKeyError: invalid expansion size")
        level = (size - 1).bit_length()
        out = [None] * (1 << level)
        out[0] = ct.clone()
        for l in range(level):
            gk = self.gk.get((n >> l) + 1)
            if gk is None or l >= len(monomials):
                raise FheError(_capi.INVALID_ARGUMENT, "EvaluationKeyError: expansion not supported by this key")
            step = 1 << l
            for i in range(step):
                sub = gk.relinearize(out[i])
                j = step | i
                if j < size:
                    tgt = out[i].clone()
                    tgt -= sub
                    tgt.mul_plain(monomials[l])
                    out[j] = tgt
                out[i] += sub
        return out[:size]

    def rotates_rows(self, ct: Ciphertext) -> Ciphertext:  # evaluation_key.rs:110-126
        e = 2 * self.par.degree() - 1
        if e not in self.gk:
            raise FheError(_capi.INVALID_ARGUMENT, "EvaluationKeyError: row rotation not supported by this key")
        return self.gk[e].relinearize(ct)

    def rotates_columns_by(self, ct: Ciphertext, i: int) -> Ciphertext:  # evaluation_key.rs:145-170
        e = pow(3, i, 2 * self.par.degree())  # :278-286
        if e not in self.gk:
            raise FheError(_capi.INVALID_ARGUMENT, "EvaluationKeyError: column rotation not supported by this key")
        return self.gk[e].relinearize(ct)


def dot_product_scalar(cts: "Ciphertext", pts, n_terms: Optional[int] = None) -> "Ciphertext":
    """fhe::bfv::dot_product_scalar (bfv/ops/dot_product.rs:55-184): sum_i cts[i] * pts[i].

    `cts` is a batch of ciphertexts, `pts` a batch of NTT plaintext polynomials (a one-part `Ciphertext` batch, or
    u64 words [count][limbs][N] = Plaintext::poly_ntt).  With `n_terms` smaller than the batch, the call computes
    count / n_terms independent dot products at once; an operand holding exactly n_terms entries is shared by all of
    them (the expanded PIR query of examples/mulpir.rs:153-181)."""
    if not isinstance(pts, Ciphertext):
        w = np.ascontiguousarray(pts, dtype=np.uint64)
        if w.ndim != 3:
            raise FheError(_capi.INVALID_ARGUMENT, "expected [count][limbs][N] plaintext words")
        if w.shape[0] == 0:
            raise FheError(_capi.INVALID_ARGUMENT, "DotProductError::EmptyInput")
        pts = Ciphertext.from_host(cts.par, w[:, None], cts.level, NTT, cts.stream)
    n = n_terms if n_terms is not None else max(cts.count, pts.count)
    if n <= 0:
        raise FheError(_capi.INVALID_ARGUMENT, "DotProductError::EmptyInput")
    groups = max(cts.count, pts.count) // n
    out = Ciphertext(cts.par, max(groups, 1), len(cts), cts.level, NTT, cts.stream)
    check(_capi.lib().fhe_b200_dot_product_scalar(cts._h, pts._h, n, out._h, cts.stream))
    return out


class ScalingFactor:
    """fhe_math::rns::ScalingFactor (rns/scaler.rs:20-58): numerator / denominator."""

    def __init__(self, numerator: int, denominator: int):
        if denominator == 0:
            raise FheError(_capi.INVALID_ARGUMENT, "The denominator of a scaling factor should be non-zero")
        self.numerator, self.denominator = int(numerator), int(denominator)

    @staticmethod
    def one() -> "ScalingFactor":
        return ScalingFactor(1, 1)

    @property
    def is_one(self) -> bool:
        return self.numerator == self.denominator


def _le(x: int):
    b = int(x).to_bytes(max(1, (int(x).bit_length() + 7) // 8), "little")
    return (C.c_uint8 * len(b)).from_buffer_copy(b), len(b)


class Multiplicator:
    """fhe::bfv::Multiplicator (bfv/ops/mul.rs:22-33).  `default(rk)` is the default strategy (mul.rs:101-138: extend
    by factor 1, scale by t/Q, relinearize) on the fused path; `new` / `new_leveled` (mul.rs:37-75) build a custom
    strategy from scaling factors and an extended basis."""

    def __init__(self, rk: Optional[RelinearizationKey] = None, par: Optional[BfvParameters] = None, level: int = 0):
        self.rk = rk
        self.par = rk.ksk.par if rk is not None else par
        self.level = rk.ksk.ciphertext_level if rk is not None else level
        self.mod_switch = False
        self._h = None   # custom-strategy handle

    @staticmethod
    def default(rk: RelinearizationKey) -> "Multiplicator":
        return Multiplicator(rk)

    @staticmethod
    def new(lhs: ScalingFactor, rhs: ScalingFactor, extended_basis, post: ScalingFactor,
            par: BfvParameters, psi=None) -> "Multiplicator":
        return Multiplicator.new_leveled(lhs, rhs, extended_basis, post, 0, par, psi)

    @staticmethod
    def new_leveled(lhs: ScalingFactor, rhs: ScalingFactor, extended_basis, post: ScalingFactor, level: int,
                    par: BfvParameters, psi=None) -> "Multiplicator":
        m = Multiplicator(None, par, level)
        basis = np.ascontiguousarray(np.array([int(q) for q in extended_basis], dtype=np.uint64))
        ps = None
        if psi is not None:
            ps = np.ascontiguousarray(np.array([int(psi[int(q)]) for q in basis], dtype=np.uint64))
        args = []
        for f in (lhs, rhs):
            for v in (f.numerator, f.denominator):
                args += list(_le(v))
        pn, pd = _le(post.numerator), _le(post.denominator)
        h = C.c_void_p()
        check(_capi.lib().fhe_b200_multiplicator_create(
            par._h, level, *args, basis.ctypes.data, len(basis), ps.ctypes.data if ps is not None else None,
            pn[0], pn[1], pd[0], pd[1], C.byref(h)))
        m._h = h
        m.extended_basis = [int(q) for q in basis]
        return m

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _release("fhe_b200_multiplicator_free", h)

    def enable_relinearization(self, rk: RelinearizationKey):  # mul.rs:141-151
        if rk.ksk.par is not self.par or rk.ksk.ciphertext_level != self.level:
            raise FheError(_capi.CONTEXT_MISMATCH, "ParameterMismatch")
        self.rk = rk
        return self

    def enable_mod_switching(self):  # mul.rs:155-162
        if self.level >= self.par.max_level():
            raise FheError(_capi.NO_MORE_CONTEXT, "NoMoreContext")
        self.mod_switch = True
        return self

    def multiply(self, lhs: Ciphertext, rhs: Ciphertext) -> Ciphertext:
        """Multiplicator::multiply (mul.rs:165-243)."""
        if lhs.level != self.level or rhs.level != self.level:
            raise FheError(_capi.INVALID_LEVEL, "InvalidLevel")  # mul.rs:168-181
        ms = 1 if self.mod_switch else 0
        if self._h is None:
            out = lhs._like(parts=2, level=self.level + ms)
            check(_capi.lib().fhe_b200_mul_relin(lhs._h, rhs._h, self.rk.ksk._h, ms, out._h, lhs.stream))
            return out
        out = lhs._like(parts=2 if self.rk is not None else 3, level=self.level + ms)
        check(_capi.lib().fhe_b200_multiplicator_multiply(
            self._h, lhs._h, rhs._h, self.rk.ksk._h if self.rk is not None else None, ms, out._h, lhs.stream))
        return out
