"""ORACLE (test infrastructure, never on the product path): the reference's protobuf messages around the hot path.

Restates, for the CPU oracle objects of `fhe_oracle.py`:
  * `Serialize for Poly` / `DeserializeWithContext for Poly`   crates/fhe-math/src/rq/serialize.rs:10-31
    `From<&Poly<R>> for Rq`, `parse_proto`                     crates/fhe-math/src/rq/convert.rs:17-98
  * `From<&Ciphertext> for CiphertextProto` and back            crates/fhe/src/bfv/ciphertext.rs:230-317
  * `KeySwitchingKeyProto` both ways                            crates/fhe/src/bfv/keys/key_switching_key.rs:365-482
  * `RelinearizationKeyProto`, `GaloisKeyProto`                 keys/relinearization_key.rs:113-135, keys/galois_key.rs:146-173
  * `RGSWCiphertextProto`                                       bfv/rgsw_ciphertext.rs:30-71

The reference encodes with prost; here the schema of `fhe-math/src/proto/rq.proto:5-17` and `fhe/src/proto/bfv.proto:5-37`
is rebuilt as descriptors of the `google.protobuf` runtime (an implementation independent of both prost and of the
hand-written codec in `fhe_rs_b200/wire.py`), which emits the same canonical proto3 bytes for these messages:
fields in field-number order, zero scalars and empty singular `bytes` omitted, every element of a repeated `bytes`
present.  Parity status: the framing is pinned against this third implementation, not against bytes produced by a Rust
process (no Rust toolchain here, and the reference's tests hold no serialized golden vectors: they are round trips).

Seeded messages (`seed` non-empty): the seed's expansion (Poly::random_from_seed, rq/mod.rs:276-292) is NOT restated --
see include/fhe_b200.h; `*_from_bytes` take the expanded half from the caller.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

import fhe_oracle as orc

_T = descriptor_pb2.FieldDescriptorProto


def _build_pool():
    pool = descriptor_pool.DescriptorPool()
    rq = descriptor_pb2.FileDescriptorProto(name="oracle_rq.proto", package="fhers.rq", syntax="proto3")
    e = rq.enum_type.add(name="Representation")                       # rq.proto:5-10
    for name, number in (("UNKNOWN", 0), ("POWERBASIS", 1), ("NTT", 2), ("NTTSHOUP", 3)):
        e.value.add(name=name, number=number)
    m = rq.message_type.add(name="Rq")                                # rq.proto:12-17
    m.field.add(name="representation", number=1, type=_T.TYPE_ENUM, type_name=".fhers.rq.Representation",
                label=_T.LABEL_OPTIONAL)
    m.field.add(name="degree", number=2, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    m.field.add(name="coefficients", number=3, type=_T.TYPE_BYTES, label=_T.LABEL_OPTIONAL)
    m.field.add(name="allow_variable_time", number=4, type=_T.TYPE_BOOL, label=_T.LABEL_OPTIONAL)
    pool.Add(rq)

    bfv = descriptor_pb2.FileDescriptorProto(name="oracle_bfv.proto", package="fhers.bfv", syntax="proto3")
    m = bfv.message_type.add(name="Ciphertext")                       # bfv.proto:5-9
    m.field.add(name="c", number=1, type=_T.TYPE_BYTES, label=_T.LABEL_REPEATED)
    m.field.add(name="seed", number=2, type=_T.TYPE_BYTES, label=_T.LABEL_OPTIONAL)
    m.field.add(name="level", number=3, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    m = bfv.message_type.add(name="KeySwitchingKey")                  # bfv.proto:16-23
    m.field.add(name="c0", number=1, type=_T.TYPE_BYTES, label=_T.LABEL_REPEATED)
    m.field.add(name="c1", number=2, type=_T.TYPE_BYTES, label=_T.LABEL_REPEATED)
    m.field.add(name="seed", number=3, type=_T.TYPE_BYTES, label=_T.LABEL_OPTIONAL)
    m.field.add(name="ciphertext_level", number=4, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    m.field.add(name="ksk_level", number=5, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    m.field.add(name="log_base", number=6, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    m = bfv.message_type.add(name="RGSWCiphertext")                   # bfv.proto:11-14
    m.field.add(name="ksk0", number=1, type=_T.TYPE_MESSAGE, type_name=".fhers.bfv.KeySwitchingKey",
                label=_T.LABEL_OPTIONAL)
    m.field.add(name="ksk1", number=2, type=_T.TYPE_MESSAGE, type_name=".fhers.bfv.KeySwitchingKey",
                label=_T.LABEL_OPTIONAL)
    m = bfv.message_type.add(name="RelinearizationKey")               # bfv.proto:25-27
    m.field.add(name="ksk", number=1, type=_T.TYPE_MESSAGE, type_name=".fhers.bfv.KeySwitchingKey",
                label=_T.LABEL_OPTIONAL)
    m = bfv.message_type.add(name="GaloisKey")                        # bfv.proto:29-32
    m.field.add(name="ksk", number=1, type=_T.TYPE_MESSAGE, type_name=".fhers.bfv.KeySwitchingKey",
                label=_T.LABEL_OPTIONAL)
    m.field.add(name="exponent", number=2, type=_T.TYPE_UINT32, label=_T.LABEL_OPTIONAL)
    pool.Add(bfv)
    return pool


_POOL = _build_pool()


def _cls(name: str):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(name))


Rq = _cls("fhers.rq.Rq")
CiphertextProto = _cls("fhers.bfv.Ciphertext")
KeySwitchingKeyProto = _cls("fhers.bfv.KeySwitchingKey")
RGSWCiphertextProto = _cls("fhers.bfv.RGSWCiphertext")
RelinearizationKeyProto = _cls("fhers.bfv.RelinearizationKey")
GaloisKeyProto = _cls("fhers.bfv.GaloisKey")

_REP_TO_PROTO = {orc.POWER_BASIS: 1, orc.NTT: 2, orc.NTT_SHOUP: 3}


class WireError(ValueError):
    """PolynomialSerializationError / SerializationError of the reference, by variant name"""


# ------------------------------------------------------------------------------------ polynomials
def poly_to_bytes(p: "orc.Poly") -> bytes:
    """Poly::to_bytes = Rq::from(self).encode_to_vec() (serialize.rs:10-14, convert.rs:17-44)"""
    m = Rq()
    m.representation = _REP_TO_PROTO[p.rep]
    m.degree = p.ctx.degree
    m.coefficients = orc.poly_to_rq_coefficients(p)
    m.allow_variable_time = False          # convert.rs:39-41: never serialized as true
    return m.SerializeToString()


def poly_from_bytes(ctx: "orc.Context", data: bytes, rep: int) -> "orc.Poly":
    """Poly::<R>::from_bytes (serialize.rs:23-31) -> parse_proto (convert.rs:46-98) -> try_convert_from (:100-161)"""
    m = Rq()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:                  # PolynomialSerializationError::Decode
        raise WireError("Decode") from e
    if m.representation not in (0, 1, 2, 3):
        raise WireError("InvalidRepresentation")
    if m.representation == 0:
        raise WireError("UnknownRepresentation")
    degree = m.degree
    if degree % 8 != 0 or degree < 8:
        raise WireError("InvalidDegree")
    expected = sum((q - 1).bit_length() * degree // 8 for q in ctx.moduli)
    if len(m.coefficients) != expected:
        raise WireError("InvalidCoefficientCount")
    if m.representation != _REP_TO_PROTO[rep]:
        raise WireError("RepresentationMismatch")
    coeffs = m.coefficients
    if degree != ctx.degree:
        # Poly::<PowerBasis>::try_convert_from(Vec<u64>) (convert.rs:148-192): a vector of q.len() * degree words is taken
        # as is; a SHORTER one of at most `degree` words (only possible with one modulus) is a low-order polynomial,
        # zero-extended and reduced; anything else is InvalidCoefficientCount.
        if len(ctx.moduli) * degree > ctx.degree or len(ctx.moduli) != 1:
            raise WireError("InvalidCoefficientCount")
        nb = (ctx.moduli[0] - 1).bit_length()
        vals = orc.transcode_from_bytes(coeffs, nb)[:degree] + [0] * (ctx.degree - degree)
        row = np.array([v % ctx.moduli[0] for v in vals], dtype=np.uint64)[None]
        p = orc.Poly(ctx, orc.POWER_BASIS, row)
        if rep != orc.POWER_BASIS:
            p.into_ntt()
    else:
        p = orc.poly_from_rq_coefficients(ctx, coeffs, orc.POWER_BASIS if rep == orc.POWER_BASIS else orc.NTT)
    p.rep = rep
    return p


# ------------------------------------------------------------------------------------ ciphertexts
def ciphertext_to_bytes(ct: "orc.Ciphertext", seed: Optional[bytes] = None) -> bytes:
    """ciphertext.rs:230-257: all polynomials, or all but the last plus the seed the last one was drawn from"""
    m = CiphertextProto()
    if ct.c:
        for p in ct.c[:-1]:
            m.c.append(poly_to_bytes(p))
        if seed is not None:
            m.seed = bytes(seed)
        else:
            m.c.append(poly_to_bytes(ct.c[-1]))
    m.level = ct.level
    return m.SerializeToString()


def ciphertext_from_bytes(par: "orc.BfvParameters", data: bytes,
                          seeded_half: Optional[np.ndarray] = None) -> "orc.Ciphertext":
    """ciphertext.rs:259-317.  `seeded_half` = the NTT words of Poly::random_from_seed(ctx, seed), host-expanded."""
    m = CiphertextProto()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:
        raise WireError("Decode") from e
    if len(m.c) == 0 or (len(m.c) == 1 and len(m.seed) == 0):
        raise WireError("InvalidCiphertextPolynomialCount")
    if m.level > par.max_level():
        raise WireError("InvalidLevel")
    ctx = par.context_at_level(m.level)
    c = [poly_from_bytes(ctx, b, orc.NTT) for b in m.c]
    if len(m.seed):
        if len(m.seed) != 32:
            raise WireError("InvalidSeedSize")
        if seeded_half is None:
            raise WireError("SeedExpansionOutsideTheOracle")
        c.append(orc.Poly(ctx, orc.NTT, np.array(seeded_half, dtype=np.uint64)))
    return orc.Ciphertext(par, c, m.level)


# ------------------------------------------------------------------------------------ key-switching keys
def _ksk_proto(k: "orc.KeySwitchingKey", seed: Optional[bytes] = None):
    m = KeySwitchingKeyProto()                                       # key_switching_key.rs:365-385
    if seed is not None:
        m.seed = bytes(seed)
    else:
        for p in k.c1:
            m.c1.append(poly_to_bytes(p))
    for p in k.c0:
        m.c0.append(poly_to_bytes(p))
    m.ciphertext_level = k.ciphertext_level
    m.ksk_level = k.ksk_level
    m.log_base = k.log_base
    return m


def ksk_to_bytes(k: "orc.KeySwitchingKey", seed: Optional[bytes] = None) -> bytes:
    return _ksk_proto(k, seed).SerializeToString()


def _ksk_from_proto(par: "orc.BfvParameters", m, seeded_c1: Optional[np.ndarray] = None) -> "orc.KeySwitchingKey":
    """key_switching_key.rs:388-482"""
    ct_level, ksk_level = m.ciphertext_level, m.ksk_level
    if ksk_level > par.max_level() or ct_level > par.max_level():
        raise WireError("InvalidLevel")
    ctx_ksk = par.context_at_level(ksk_level)
    ctx_ct = par.context_at_level(ct_level)
    if m.log_base != 0:
        if ksk_level != par.max_level() or ct_level != par.max_level():
            raise WireError("InvalidKeySwitchingDecompositionLevels")
        # as coded (:406-407): the FIRST modulus of the parameter set, not the key level's own last one
        log_modulus = (par.moduli[0] - 1).bit_length()              # next_power_of_two().ilog2()
        c0_size = -(-log_modulus // m.log_base)
    else:
        c0_size = len(ctx_ct.moduli)
    if len(m.c0) != c0_size:
        raise WireError("WrongPolynomialCount:KeySwitchingKeyC0")
    if len(m.seed) == 0:
        if len(m.c1) != c0_size:
            raise WireError("WrongPolynomialCount:KeySwitchingKeyC1")
        c1 = [poly_from_bytes(ctx_ksk, b, orc.NTT_SHOUP) for b in m.c1]
    else:
        if len(m.seed) != 32:
            raise WireError("InvalidKeySwitchingSeedLength")
        if seeded_c1 is None:
            raise WireError("SeedExpansionOutsideTheOracle")
        c1 = [orc.Poly(ctx_ksk, orc.NTT_SHOUP, np.array(a, dtype=np.uint64)) for a in seeded_c1]
    c0 = [poly_from_bytes(ctx_ksk, b, orc.NTT_SHOUP) for b in m.c0]
    k = orc.KeySwitchingKey.__new__(orc.KeySwitchingKey)
    k.par, k.ctx_ksk, k.ctx_ciphertext = par, ctx_ksk, ctx_ct
    k.ciphertext_level, k.ksk_level, k.log_base = ct_level, ksk_level, m.log_base
    k.c0, k.c1 = c0, c1
    return k


def ksk_from_bytes(par, data: bytes, seeded_c1=None) -> "orc.KeySwitchingKey":
    m = KeySwitchingKeyProto()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:
        raise WireError("Decode") from e
    return _ksk_from_proto(par, m, seeded_c1)


def relin_key_to_bytes(rk: "orc.RelinearizationKey") -> bytes:      # relinearization_key.rs:113-119, :137-141
    m = RelinearizationKeyProto()
    m.ksk.CopyFrom(_ksk_proto(rk.ksk))
    return m.SerializeToString()


def relin_key_from_bytes(par, data: bytes) -> "orc.RelinearizationKey":   # relinearization_key.rs:121-135
    m = RelinearizationKeyProto()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:
        raise WireError("Decode") from e
    if not m.HasField("ksk"):
        raise WireError("MissingField:RelinearizationKeySwitchingKey")
    return orc.RelinearizationKey.from_ksk(_ksk_from_proto(par, m.ksk))


def galois_key_to_bytes(gk: "orc.GaloisKey") -> bytes:              # galois_key.rs:146-153
    m = GaloisKeyProto()
    m.ksk.CopyFrom(_ksk_proto(gk.ksk))
    m.exponent = gk.exponent
    return m.SerializeToString()


def galois_key_from_bytes(par, data: bytes) -> "orc.GaloisKey":     # galois_key.rs:155-173
    m = GaloisKeyProto()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:
        raise WireError("Decode") from e
    if not m.HasField("ksk"):
        raise WireError("MissingField:GaloisKeySwitchingKey")
    ksk = _ksk_from_proto(par, m.ksk)
    exponent = m.exponent % (2 * par.degree)    # SubstitutionExponent::new (rq/mod.rs:99-106): odd exponents only
    if exponent & 1 == 0:
        raise WireError("InvalidSubstitutionExponent")
    gk = orc.GaloisKey.__new__(orc.GaloisKey)
    gk.exponent, gk.ksk = exponent, ksk
    return gk


def rgsw_to_bytes(r: "orc.RGSWCiphertext") -> bytes:                # rgsw_ciphertext.rs:30-37
    m = RGSWCiphertextProto()
    m.ksk0.CopyFrom(_ksk_proto(r.ksk0))
    m.ksk1.CopyFrom(_ksk_proto(r.ksk1))
    return m.SerializeToString()


def rgsw_from_bytes(par, data: bytes) -> "orc.RGSWCiphertext":      # rgsw_ciphertext.rs:39-71
    m = RGSWCiphertextProto()
    try:
        m.ParseFromString(bytes(data))
    except Exception as e:
        raise WireError("Decode") from e
    if not m.HasField("ksk0"):
        raise WireError("MissingField:RgswKeySwitchingKey0")
    if not m.HasField("ksk1"):
        raise WireError("MissingField:RgswKeySwitchingKey1")
    k0, k1 = _ksk_from_proto(par, m.ksk0), _ksk_from_proto(par, m.ksk1)
    if k0.ksk_level != k0.ciphertext_level or k0.ciphertext_level != k1.ciphertext_level \
            or k1.ciphertext_level != k1.ksk_level:
        raise WireError("InconsistentKeySwitchingLevels")
    r = orc.RGSWCiphertext.__new__(orc.RGSWCiphertext)
    r.ksk0, r.ksk1, r.level = k0, k1, k0.ciphertext_level
    return r
