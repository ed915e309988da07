// fhe_b200_wire.hpp -- the protobuf messages either side of the accelerated path, for C++ hosts (header-only).
//
// The reference serialises with prost (paths under /root/reference/crates):
//   fhers.rq.Rq                   fhe-math/src/proto/rq.proto:12-17, written by fhe-math/src/rq/convert.rs:17-44
//   fhers.bfv.Ciphertext          fhe/src/proto/bfv.proto:5-9,       fhe/src/bfv/ciphertext.rs:230-317
//   fhers.bfv.KeySwitchingKey     bfv.proto:16-23,                   fhe/src/bfv/keys/key_switching_key.rs:365-482
//   fhers.bfv.RelinearizationKey  bfv.proto:25-27 (keys/relinearization_key.rs:113-135), GaloisKey :29-32
//                                 (keys/galois_key.rs:146-173)
// `Rq.coefficients` -- the bit-packed power-basis words, all but a few bytes of every message -- is produced and consumed
// on the device (fhe_b200_batch_pack / fhe_b200_batch_unpack); this header is the proto3 framing around it, emitting
// what prost emits (fields in field-number order, zero scalars and empty singular `bytes` omitted) and accepting what
// prost accepts (any order, unknown fields skipped, last scalar wins), plus the checks of the reference's decoders
// under the reference's variant names (WireError::variant).  The Python mirror's fhe_rs_b200/wire.py is the same
// codec; both are tested byte for byte against the google.protobuf runtime.
//
// Seeded messages carry a 32-byte ChaCha8 seed instead of their last polynomial (row c1 for keys).  Expanding it is
// the Rust host's job (see fhe_b200.h): the *_from_bytes functions take the expanded words as an argument.
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "fhe_b200.hpp"

namespace fhe_b200 {

// PolynomialSerializationError (fhe-math/src/errors.rs) / SerializationError (fhe/src/errors.rs) by variant name
struct WireError : Error {
  std::string variant;
  WireError(const std::string& v, int c = FHE_B200_INVALID_ARGUMENT, const std::string& detail = "")
      : Error(c, detail.empty() ? v : v + ": " + detail), variant(v) {}
};

namespace wire {

enum : int32_t { REP_UNKNOWN = 0, REP_POWERBASIS = 1, REP_NTT = 2, REP_NTTSHOUP = 3 };  // rq.proto:5-10

struct Span {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

inline void put_varint(std::string& out, uint64_t v) {
  while (v >= 0x80) {
    out.push_back((char)(v | 0x80));
    v >>= 7;
  }
  out.push_back((char)v);
}
inline void put_uint(std::string& out, uint32_t field, uint64_t v) {
  if (!v) return;  // proto3: default values are not written
  put_varint(out, (uint64_t)field << 3);
  put_varint(out, v);
}
inline void put_len(std::string& out, uint32_t field, const void* data, size_t n) {
  put_varint(out, ((uint64_t)field << 3) | 2);
  put_varint(out, n);
  out.append((const char*)data, n);
}
inline void put_len(std::string& out, uint32_t field, const std::string& s) { put_len(out, field, s.data(), s.size()); }

// one message, field by field.  Follows prost's decoder: keys are 32-bit with a field number >= 1, varints are at most
// ten bytes, unknown groups are skipped whole (they never occur in these messages), an unmatched end-group or wire
// types 6 / 7 are errors, a known field with another wire type is an error (expect()).
class Reader {
 public:
  Reader(const void* p, size_t n) : p_((const uint8_t*)p), end_((const uint8_t*)p + n) {}
  // false at the end of the message; otherwise field / wire_type and (varint, fixed) value or (bytes) span
  bool next() {
    if (p_ >= end_) return false;
    key(field, wire_type);
    skip_or_read(field, wire_type, 0, true);
    return true;
  }
  void expect(int wt) const {
    if (wire_type != wt) fail("unexpected wire type for a known field");
  }
  uint32_t field = 0;
  int wire_type = 0;
  uint64_t value = 0;
  Span span;

 private:
  [[noreturn]] static void fail(const char* why) { throw WireError("Decode", FHE_B200_INVALID_ARGUMENT, why); }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0;; shift += 7) {
      if (p_ >= end_) fail("truncated varint");
      uint8_t b = *p_++;
      if (shift == 63 && b > 1) fail("varint overflows 64 bits");
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
  }
  void key(uint32_t& f, int& wt) {
    uint64_t k = varint();
    if (k > 0xFFFFFFFFull) fail("key does not fit 32 bits");
    if ((k >> 3) == 0) fail("field number 0");
    f = (uint32_t)(k >> 3);
    wt = (int)(k & 7);
  }
  void skip_or_read(uint32_t f, int wt, int depth, bool keep) {
    switch (wt) {
      case 0: {
        uint64_t v = varint();
        if (keep) value = v;
        break;
      }
      case 2: {
        uint64_t n = varint();
        if (n > (uint64_t)(end_ - p_)) fail("length overruns the buffer");
        if (keep) { span.p = p_; span.n = (size_t)n; }
        p_ += n;
        break;
      }
      case 1:
      case 5: {
        size_t n = wt == 1 ? 8 : 4;
        if (n > (size_t)(end_ - p_)) fail("truncated fixed-width field");
        if (keep) {
          value = 0;
          for (size_t i = 0; i < n; i++) value |= (uint64_t)p_[i] << (8 * i);
        }
        p_ += n;
        break;
      }
      case 3: {
        if (depth >= 100) fail("recursion limit");
        for (;;) {
          if (p_ >= end_) fail("unterminated group");
          uint32_t gf;
          int gw;
          key(gf, gw);
          if (gw == 4) {
            if (gf != f) fail("mismatched end of group");
            break;
          }
          skip_or_read(gf, gw, depth + 1, false);
        }
        if (keep) value = 0;
        break;
      }
      default: fail("unsupported wire type");
    }
  }
  const uint8_t *p_, *end_;
};

// ---- Rq ---------------------------------------------------------------------------------------------------------
// Rq::from(&poly).encode_to_vec(); allow_variable_time is never true on the wire (rq/convert.rs:39-41)
inline std::string encode_rq(int32_t representation, uint32_t degree, const uint8_t* coeffs, size_t n) {
  std::string out;
  out.reserve(n + 16);
  put_uint(out, 1, (uint32_t)representation);
  put_uint(out, 2, degree);
  if (n) put_len(out, 3, coeffs, n);
  return out;
}
struct Rq {
  int32_t representation = 0;
  uint32_t degree = 0;
  Span coefficients;
};
// the context-free checks of parse_proto (rq/convert.rs:46-75)
inline Rq decode_rq(const void* data, size_t n) {
  Rq m;
  Reader r(data, n);
  while (r.next()) {
    if (r.field == 1) { r.expect(0); m.representation = (int32_t)(uint32_t)r.value; }
    else if (r.field == 2) { r.expect(0); m.degree = (uint32_t)r.value; }
    else if (r.field == 3) { r.expect(2); m.coefficients = r.span; }
    else if (r.field == 4) { r.expect(0); }   // the timing flag on the wire grants nothing
  }
  if (m.representation < 0 || m.representation > 3)
    throw WireError("InvalidRepresentation", FHE_B200_INVALID_REPRESENTATION, std::to_string(m.representation));
  if (m.representation == REP_UNKNOWN) throw WireError("UnknownRepresentation", FHE_B200_INVALID_REPRESENTATION);
  if (m.degree % 8 != 0 || m.degree < 8) throw WireError("InvalidDegree", FHE_B200_INVALID_DEGREE, std::to_string(m.degree));
  return m;
}

// ---- Ciphertext -------------------------------------------------------------------------------------------------
struct CiphertextMsg {
  std::vector<Span> c;
  Span seed;
  uint32_t level = 0;
};
inline std::string encode_ciphertext(const std::vector<std::string>& polys, const std::string& seed, uint32_t level) {
  std::string out;
  size_t total = 16 + seed.size();
  for (auto& p : polys) total += p.size() + 8;
  out.reserve(total);
  for (auto& p : polys) put_len(out, 1, p);
  if (!seed.empty()) put_len(out, 2, seed);
  put_uint(out, 3, level);
  return out;
}
inline CiphertextMsg decode_ciphertext(const void* data, size_t n) {
  CiphertextMsg m;
  Reader r(data, n);
  while (r.next()) {
    if (r.field == 1) { r.expect(2); m.c.push_back(r.span); }
    else if (r.field == 2) { r.expect(2); m.seed = r.span; }
    else if (r.field == 3) { r.expect(0); m.level = (uint32_t)r.value; }
  }
  if (m.c.empty() || (m.c.size() == 1 && m.seed.n == 0))   // ciphertext.rs:261-269
    throw WireError("InvalidCiphertextPolynomialCount", FHE_B200_BAD_POLY_COUNT);
  return m;
}

// ---- KeySwitchingKey and the messages that wrap it -----------------------------------------------------------------
struct KskMsg {
  std::vector<Span> c0, c1;
  Span seed;
  uint32_t ciphertext_level = 0, ksk_level = 0, log_base = 0;
};
inline std::string encode_ksk(const std::vector<std::string>& c0, const std::vector<std::string>& c1, const std::string& seed,
                              uint32_t ciphertext_level, uint32_t ksk_level, uint32_t log_base) {
  std::string out;
  for (auto& p : c0) put_len(out, 1, p);
  for (auto& p : c1) put_len(out, 2, p);
  if (!seed.empty()) put_len(out, 3, seed);
  put_uint(out, 4, ciphertext_level);
  put_uint(out, 5, ksk_level);
  put_uint(out, 6, log_base);
  return out;
}
inline KskMsg decode_ksk(const void* data, size_t n) {
  KskMsg m;
  Reader r(data, n);
  while (r.next()) {
    if (r.field == 1) { r.expect(2); m.c0.push_back(r.span); }
    else if (r.field == 2) { r.expect(2); m.c1.push_back(r.span); }
    else if (r.field == 3) { r.expect(2); m.seed = r.span; }
    else if (r.field == 4) { r.expect(0); m.ciphertext_level = (uint32_t)r.value; }
    else if (r.field == 5) { r.expect(0); m.ksk_level = (uint32_t)r.value; }
    else if (r.field == 6) { r.expect(0); m.log_base = (uint32_t)r.value; }
  }
  return m;
}
inline std::string encode_relinearization_key(const std::string& ksk) {   // relinearization_key.rs:113-119
  std::string out;
  put_len(out, 1, ksk);
  return out;
}
inline std::string encode_galois_key(const std::string& ksk, uint32_t exponent) {   // galois_key.rs:146-153
  std::string out;
  put_len(out, 1, ksk);
  put_uint(out, 2, exponent);
  return out;
}
// sub-message `field` of a wrapper message; *scalar2 = varint field 2 when present
inline Span sub_message(const void* data, size_t n, uint32_t field, const char* missing, uint32_t* scalar2 = nullptr) {
  Span s;
  bool found = false;
  Reader r(data, n);
  while (r.next()) {
    if (r.field == field) { r.expect(2); s = r.span; found = true; }
    else if (scalar2 && r.field == 2 && r.wire_type == 0) *scalar2 = (uint32_t)r.value;
  }
  if (!found) throw WireError("MissingField", FHE_B200_INVALID_ARGUMENT, missing);
  return s;
}

}  // namespace wire

namespace bfv {

// `Poly::<R>::from_bytes` for every polynomial of `batch` (rq/serialize.rs:23-31, rq/convert.rs:46-161): msgs[i * parts + j]
// is the encoded Rq of part j of ciphertext i.  Framing and checks on the host, unpacking (+ forward NTT) on the device.
inline void unpack_rq(Ciphertext& batch, const std::vector<wire::Span>& msgs, int32_t want_rep) {
  const size_t nbytes = batch.packed_bytes(), deg = batch.par()->degree();
  const uint32_t count = batch.count(), parts = batch.len(), limbs = batch.limbs();
  if (msgs.size() != (size_t)count * parts) throw Error(FHE_B200_INVALID_ARGUMENT, "message count does not match the batch");
  std::vector<uint8_t> blobs(msgs.size() * nbytes, 0);
  for (size_t k = 0; k < msgs.size(); k++) {
    wire::Rq m = wire::decode_rq(msgs[k].p, msgs[k].n);
    if ((uint64_t)m.degree * nbytes != (uint64_t)m.coefficients.n * deg)    // convert.rs:76-88
      throw WireError("InvalidCoefficientCount");
    if (m.representation != want_rep) throw WireError("RepresentationMismatch", FHE_B200_INVALID_REPRESENTATION);
    // convert.rs:148-192: q.len() * degree words, or -- one modulus only -- a shorter low-order polynomial, zero-extended
    if (m.degree != deg && (limbs != 1 || m.degree > deg)) throw WireError("InvalidCoefficientCount");
    std::memcpy(blobs.data() + k * nbytes, m.coefficients.p, m.coefficients.n);
  }
  check(fhe_b200_batch_unpack(batch.handle(), 0, count, blobs.data(), batch.stream()));
  batch.sync();
}

// ct.to_bytes() for every ciphertext of the batch (ciphertext.rs:230-257, unseeded branch)
inline std::vector<std::string> to_bytes(const Ciphertext& ct) {
  const uint32_t count = ct.count(), parts = ct.len(), level = ct.level();
  const size_t nbytes = ct.packed_bytes();
  const uint32_t deg = (uint32_t)ct.par()->degree();
  std::vector<uint8_t> blobs = ct.to_packed();
  ct.sync();
  std::vector<std::string> out(count);
  for (uint32_t i = 0; i < count; i++) {
    std::vector<std::string> polys(parts);
    for (uint32_t j = 0; j < parts; j++)
      polys[j] = wire::encode_rq(wire::REP_NTT, deg, blobs.data() + ((size_t)i * parts + j) * nbytes, nbytes);
    out[i] = wire::encode_ciphertext(polys, std::string(), level);
  }
  return out;
}

// Ciphertext::from_bytes (ciphertext.rs:259-317) for a batch of messages of one level, part count and kind.
// seeded_halves: [count][limbs][N] NTT words of Poly::random_from_seed for messages that carry a seed (host-expanded).
inline Ciphertext ciphertext_from_bytes(std::shared_ptr<BfvParameters> par, const std::vector<std::string>& messages,
                                        const uint64_t* seeded_halves = nullptr) {
  if (messages.empty()) throw Error(FHE_B200_INVALID_ARGUMENT, "no messages");
  std::vector<wire::CiphertextMsg> dec;
  for (auto& m : messages) dec.push_back(wire::decode_ciphertext(m.data(), m.size()));
  const uint32_t level = dec[0].level;
  if (level > par->max_level()) throw WireError("InvalidLevel", FHE_B200_INVALID_LEVEL);
  const size_t n_rq = dec[0].c.size();
  const bool seeded = dec[0].seed.n != 0;
  std::vector<wire::Span> rq;
  for (auto& d : dec) {
    if (d.level != level || d.c.size() != n_rq || (d.seed.n != 0) != seeded)
      throw Error(FHE_B200_INVALID_ARGUMENT, "a batch holds ciphertexts of one level, part count and kind");
    if (d.seed.n && d.seed.n != 32) throw WireError("InvalidSeedSize");
    rq.insert(rq.end(), d.c.begin(), d.c.end());
  }
  const uint32_t count = (uint32_t)dec.size();
  Ciphertext body(par, count, (uint32_t)n_rq, level);
  unpack_rq(body, rq, wire::REP_NTT);
  if (!seeded) return body;
  if (!seeded_halves)
    throw WireError("SeedExpansion", FHE_B200_UNSUPPORTED, "pass the host-expanded last polynomial (ciphertext.rs:287-300)");
  const size_t poly = (size_t)body.limbs() * par->degree();
  std::vector<uint64_t> w = body.to_host(), all((size_t)count * (n_rq + 1) * poly);
  for (uint32_t i = 0; i < count; i++) {
    std::memcpy(&all[(size_t)i * (n_rq + 1) * poly], &w[(size_t)i * n_rq * poly], n_rq * poly * 8);
    std::memcpy(&all[((size_t)i * (n_rq + 1) + n_rq) * poly], seeded_halves + (size_t)i * poly, poly * 8);
  }
  return Ciphertext::from_host(par, all, count, (uint32_t)n_rq + 1, level);
}

// KeySwitchingKey::try_convert_from(&KeySwitchingKeyProto, par) (key_switching_key.rs:388-482).
// seeded_c1: [digits][limbs][N] NTT words of generate_c1 (:130-146) for a key that carries a seed.
inline std::shared_ptr<KeySwitchingKey> key_switching_key_from_bytes(std::shared_ptr<BfvParameters> par, const void* data,
                                                                     size_t n, const uint64_t* seeded_c1 = nullptr) {
  wire::KskMsg k = wire::decode_ksk(data, n);
  if (k.ksk_level > par->max_level() || k.ciphertext_level > par->max_level())
    throw WireError("InvalidLevel", FHE_B200_INVALID_LEVEL);
  auto bits = [](uint64_t v) { uint32_t b = 0; while (v) { b++; v >>= 1; } return b; };
  const std::vector<uint64_t> q = par->moduli();
  size_t c0_size;
  if (k.log_base) {
    if (k.ksk_level != par->max_level() || k.ciphertext_level != par->max_level())
      throw WireError("InvalidKeySwitchingDecompositionLevels", FHE_B200_INVALID_LEVEL);
    const uint32_t log_modulus = bits(q[0] - 1);   // as coded (:406-408): the first modulus of the parameter set
    c0_size = (log_modulus + k.log_base - 1) / k.log_base;
  } else {
    c0_size = q.size() - k.ciphertext_level;
  }
  if (k.c0.size() != c0_size) throw WireError("WrongPolynomialCount", FHE_B200_BAD_POLY_COUNT, "KeySwitchingKeyC0");
  const uint32_t ksk_limbs = (uint32_t)(q.size() - k.ksk_level);
  const uint32_t expect_base = ksk_limbs == 1 ? bits(q[0] - 1) / 2 : 0;   // key_switching_key.rs:92-97
  if (k.log_base != expect_base)   // the device derives the base from the key level; a message that disagrees is refused
    throw WireError("InvalidKeySwitchingDecompositionLevels", FHE_B200_UNSUPPORTED, "log_base does not match the key level");
  const size_t poly = (size_t)ksk_limbs * par->degree();
  std::vector<uint64_t> c0(c0_size * poly), c1(c0_size * poly);
  if (k.seed.n == 0) {
    if (k.c1.size() != c0_size) throw WireError("WrongPolynomialCount", FHE_B200_BAD_POLY_COUNT, "KeySwitchingKeyC1");
    Ciphertext tmp(par, (uint32_t)c0_size, 2, k.ksk_level);
    std::vector<wire::Span> rq;
    for (size_t i = 0; i < c0_size; i++) { rq.push_back(k.c0[i]); rq.push_back(k.c1[i]); }
    unpack_rq(tmp, rq, wire::REP_NTTSHOUP);
    std::vector<uint64_t> w = tmp.to_host();
    for (size_t i = 0; i < c0_size; i++) {
      std::memcpy(&c0[i * poly], &w[(2 * i) * poly], poly * 8);
      std::memcpy(&c1[i * poly], &w[(2 * i + 1) * poly], poly * 8);
    }
  } else {
    if (k.seed.n != 32) throw WireError("InvalidKeySwitchingSeedLength");
    if (!seeded_c1)
      throw WireError("SeedExpansion", FHE_B200_UNSUPPORTED, "pass the host-expanded c1 row (key_switching_key.rs:130-146)");
    Ciphertext tmp(par, (uint32_t)c0_size, 1, k.ksk_level);
    unpack_rq(tmp, k.c0, wire::REP_NTTSHOUP);
    c0 = tmp.to_host();
    std::memcpy(c1.data(), seeded_c1, c1.size() * 8);
  }
  return std::make_shared<KeySwitchingKey>(par, c0, c1, (uint32_t)c0_size, k.ciphertext_level, k.ksk_level);
}

// RelinearizationKey::from_bytes (relinearization_key.rs:121-135)
inline RelinearizationKey relinearization_key_from_bytes(std::shared_ptr<BfvParameters> par, const std::string& data) {
  wire::Span s = wire::sub_message(data.data(), data.size(), 1, "RelinearizationKeySwitchingKey");
  return RelinearizationKey(key_switching_key_from_bytes(std::move(par), s.p, s.n));
}
// GaloisKey::from_bytes (galois_key.rs:155-173)
inline GaloisKey galois_key_from_bytes(std::shared_ptr<BfvParameters> par, const std::string& data) {
  uint32_t exponent = 0;
  wire::Span s = wire::sub_message(data.data(), data.size(), 1, "GaloisKeySwitchingKey", &exponent);
  const uint32_t two_n = 2 * (uint32_t)par->degree();
  auto ksk = key_switching_key_from_bytes(std::move(par), s.p, s.n);
  exponent %= two_n;                        // SubstitutionExponent::new (rq/mod.rs:99-106)
  if (!(exponent & 1)) throw WireError("InvalidSubstitutionExponent", FHE_B200_INVALID_EXPONENT);
  return GaloisKey(exponent, std::move(ksk));
}

}  // namespace bfv
}  // namespace fhe_b200
