// fhe_b200.hpp -- C++17 host-side mirror of the reference's `fhe::bfv` interface for the
// accelerated path, header-only, implemented purely on the C ABI of fhe_b200.h.
//
// Names and argument meaning follow tlepoint/fhe.rs (paths under /root/reference/crates/fhe/src):
//   bfv::BfvParameters / BfvParametersBuilder   bfv/parameters.rs:88, :319
//   bfv::Ciphertext                             bfv/ciphertext.rs:18  (here: a device-resident batch)
//   bfv::KeySwitchingKey / RelinearizationKey   bfv/keys/key_switching_key.rs:22, relinearization_key.rs:23
//   bfv::RGSWCiphertext                         bfv/rgsw_ciphertext.rs:20 (external product = two key switches)
//   bfv::GaloisKey / EvaluationKey              bfv/keys/galois_key.rs:18, evaluation_key.rs:110-170
//   bfv::Multiplicator                          bfv/ops/mul.rs:22
// Fallible reference calls return Result<_, fhe::Error>; here they throw fhe_b200::Error carrying
// the fhe_b200_status code (same variants, see fhe_b200.h).
#pragma once
#include <cstdint>
#include <map>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fhe_b200.h"

namespace fhe_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int code) {
  if (code != FHE_B200_OK) throw Error(code, fhe_b200_last_error());
}

namespace bfv {

enum class Representation : int { PowerBasis = FHE_B200_POWER_BASIS, Ntt = FHE_B200_NTT };

class BfvParameters {
 public:
  BfvParameters(const BfvParameters&) = delete;
  BfvParameters& operator=(const BfvParameters&) = delete;
  ~BfvParameters() { fhe_b200_params_destroy(h_); }
  size_t degree() const { return fhe_b200_params_degree(h_); }
  std::vector<uint64_t> moduli() const {
    std::vector<uint64_t> m(fhe_b200_params_n_moduli(h_));
    check(fhe_b200_params_moduli(h_, m.data()));
    return m;
  }
  size_t max_level() const { return fhe_b200_params_n_moduli(h_) - 1; }
  std::vector<uint64_t> mul_basis(uint32_t level) const {
    uint32_t n = 0;
    check(fhe_b200_params_mul_basis(h_, level, nullptr, &n));
    std::vector<uint64_t> m(n);
    check(fhe_b200_params_mul_basis(h_, level, m.data(), &n));
    return m;
  }
  const fhe_b200_params* handle() const { return h_; }

 private:
  friend class BfvParametersBuilder;
  explicit BfvParameters(fhe_b200_params* h) : h_(h) {}
  fhe_b200_params* h_;
};

class BfvParametersBuilder {
 public:
  BfvParametersBuilder& set_degree(size_t d) { degree_ = (uint32_t)d; return *this; }
  BfvParametersBuilder& set_plaintext_modulus(uint64_t t) { plaintext_ = t; return *this; }
  BfvParametersBuilder& set_moduli(const std::vector<uint64_t>& m) { moduli_ = m; return *this; }
  BfvParametersBuilder& set_moduli_sizes(const std::vector<uint32_t>& s) { sizes_ = s; return *this; }
  BfvParametersBuilder& set_ntt_roots(const std::vector<uint64_t>& psi) { psi_ = psi; return *this; }
  BfvParametersBuilder& set_device(int device) { device_ = device; return *this; }
  // BfvParametersBuilder::build_arc (bfv/parameters.rs:555)
  std::shared_ptr<BfvParameters> build_arc() const {
    uint8_t pt[8];
    for (int i = 0; i < 8; i++) pt[i] = (uint8_t)(plaintext_ >> (8 * i));
    fhe_b200_params* h = nullptr;
    if (!moduli_.empty() && !sizes_.empty())
      throw Error(FHE_B200_INVALID_ARGUMENT, "ConflictingCiphertextModulusSpecifications");
    if (!moduli_.empty())
      check(fhe_b200_params_create(device_, degree_, moduli_.data(), (uint32_t)moduli_.size(), pt, 8,
                                   psi_.empty() ? nullptr : psi_.data(), &h));
    else
      check(fhe_b200_params_create_from_sizes(device_, degree_, sizes_.data(), (uint32_t)sizes_.size(), pt, 8, &h));
    return std::shared_ptr<BfvParameters>(new BfvParameters(h));
  }

 private:
  uint32_t degree_ = 0;
  uint64_t plaintext_ = 0;
  std::vector<uint64_t> moduli_, psi_;
  std::vector<uint32_t> sizes_;
  int device_ = 0;
};

// A batch of `count` ciphertexts with `parts` polynomials each, at one level, resident in HBM.
// Page-locked host staging memory (fhe_b200_host_alloc) for the asynchronous transfers below; write_combined for
// upload-only buffers the host fills front to back.
class PinnedWords {
 public:
  explicit PinnedWords(size_t n_words, bool write_combined = false) : n_(n_words) {
    void* p = nullptr;
    check(fhe_b200_host_alloc(n_words * sizeof(uint64_t), write_combined ? 1 : 0, &p));
    p_ = static_cast<uint64_t*>(p);
  }
  PinnedWords(const PinnedWords&) = delete;
  PinnedWords& operator=(const PinnedWords&) = delete;
  ~PinnedWords() { fhe_b200_host_free(p_); }
  uint64_t* data() { return p_; }
  const uint64_t* data() const { return p_; }
  size_t size() const { return n_; }

 private:
  uint64_t* p_ = nullptr;
  size_t n_ = 0;
};

class Ciphertext {
 public:
  Ciphertext(std::shared_ptr<BfvParameters> par, uint32_t count, uint32_t parts = 2, uint32_t level = 0,
             Representation r = Representation::Ntt, void* stream = nullptr)
      : par_(std::move(par)), stream_(stream) {
    check(fhe_b200_batch_alloc(par_->handle(), count, parts, level, (int)r, &h_));
  }
  Ciphertext(Ciphertext&& o) noexcept : par_(std::move(o.par_)), h_(o.h_), stream_(o.stream_) { o.h_ = nullptr; }
  Ciphertext(const Ciphertext&) = delete;
  ~Ciphertext() { fhe_b200_batch_free(h_); }

  // host words [count][parts][limbs][N] == Vec<u64>::from(&Poly) per part (rq/convert.rs:474)
  static Ciphertext from_host(std::shared_ptr<BfvParameters> par, const std::vector<uint64_t>& words, uint32_t count,
                              uint32_t parts = 2, uint32_t level = 0, Representation r = Representation::Ntt) {
    Ciphertext ct(std::move(par), count, parts, level, r);
    if (words.size() != ct.words()) throw Error(FHE_B200_INVALID_ARGUMENT, "word count does not match the batch shape");
    check(fhe_b200_batch_upload(ct.h_, 0, count, words.data(), ct.stream_));
    check(fhe_b200_sync(ct.stream_));
    return ct;
  }
  std::vector<uint64_t> to_host() const {
    std::vector<uint64_t> w(words());
    check(fhe_b200_batch_download(h_, 0, count(), w.data(), stream_));
    return w;
  }
  // enqueue-only transfers of ciphertexts [first, first + n) on the batch's stream; `host` must be page-locked
  // (PinnedWords) for them to be asynchronous and must stay valid until sync()
  void upload_async(const uint64_t* host, uint32_t first, uint32_t n) { check(fhe_b200_batch_upload(h_, first, n, host, stream_)); }
  void download_async(uint64_t* host, uint32_t first, uint32_t n) const { check(fhe_b200_batch_download_async(h_, first, n, host, stream_)); }
  void sync() const { check(fhe_b200_sync(stream_)); }
  uint32_t count() const { uint32_t c; check(fhe_b200_batch_info(h_, &c, nullptr, nullptr, nullptr, nullptr)); return c; }
  uint32_t len() const { uint32_t p; check(fhe_b200_batch_info(h_, nullptr, &p, nullptr, nullptr, nullptr)); return p; }
  uint32_t level() const { uint32_t l; check(fhe_b200_batch_info(h_, nullptr, nullptr, &l, nullptr, nullptr)); return l; }
  uint32_t limbs() const { uint32_t l; check(fhe_b200_batch_info(h_, nullptr, nullptr, nullptr, &l, nullptr)); return l; }
  size_t words() const { return (size_t)count() * len() * limbs() * par_->degree(); }

  Ciphertext clone() const {
    Ciphertext c(par_, count(), len(), level(), representation(), stream_);
    check(fhe_b200_batch_copy(c.h_, h_, stream_));
    return c;
  }
  // bfv/ops/mod.rs:54, :148, :205
  Ciphertext& operator+=(const Ciphertext& rhs) { check(fhe_b200_add(h_, rhs.h_, stream_)); return *this; }
  Ciphertext& operator-=(const Ciphertext& rhs) { check(fhe_b200_sub(h_, rhs.h_, stream_)); return *this; }
  Ciphertext operator-() const { Ciphertext c = clone(); check(fhe_b200_neg(c.h_, stream_)); return c; }
  // &Ciphertext * &Ciphertext -> 3 parts (bfv/ops/mod.rs:259)
  Ciphertext operator*(const Ciphertext& rhs) const {
    Ciphertext out(par_, count(), len() + rhs.len() - 1, level(), Representation::Ntt, stream_);
    check(fhe_b200_mul(h_, rhs.h_, out.h_, stream_));
    return out;
  }
  // Ciphertext += / -= &Plaintext (bfv/ops/mod.rs:88, :188): poly = Plaintext::to_poly() words, [limbs][N]
  Ciphertext& add_plain(const std::vector<uint64_t>& poly, bool subtract = false) {
    check(fhe_b200_add_plain(h_, poly.data(), 1, subtract ? 1 : 0, stream_));
    return *this;
  }
  // Ciphertext *= &Plaintext (bfv/ops/mod.rs:229): poly_ntt = [limbs][N] words shared by the batch
  Ciphertext& mul_plain(const std::vector<uint64_t>& poly_ntt) {
    check(fhe_b200_mul_plain(h_, poly_ntt.data(), 1, stream_));
    return *this;
  }
  // Poly::into_ntt / into_power_basis on every polynomial (rq/mod.rs:535, :590)
  Ciphertext& into_ntt() { check(fhe_b200_ntt_forward(h_, stream_)); return *this; }
  Ciphertext& into_power_basis() { check(fhe_b200_ntt_backward(h_, stream_)); return *this; }
  Representation representation() const {
    int r;
    check(fhe_b200_batch_info(h_, nullptr, nullptr, nullptr, nullptr, &r));
    return (Representation)r;
  }
  // Poly::substitute on every polynomial, in either representation (rq/mod.rs:360-408)
  Ciphertext substitute(uint32_t exponent) const {
    Ciphertext out(par_, count(), len(), level(), representation(), stream_);
    check(fhe_b200_substitute(h_, exponent, out.h_, stream_));
    return out;
  }
  // Ciphertext::switch_down (bfv/ciphertext.rs:148)
  void switch_down() { check(fhe_b200_switch_down(h_, stream_)); }
  // Ciphertext::switch_to_level (ciphertext.rs:164-184): only moves down
  void switch_to_level(uint32_t target_level) {
    if (target_level < level() || target_level > par_->max_level())
      throw Error(FHE_B200_INVALID_LEVEL, "InvalidLevel");
    while (level() < target_level) switch_down();
  }
  // Rq.coefficients of every polynomial (rq/convert.rs:17-44): count*parts blobs of packed_bytes() each
  size_t packed_bytes() const {
    size_t n = 0;
    check(fhe_b200_batch_packed_bytes(h_, &n));   // per batch: a multiplication-basis batch has L + E limbs
    return n;
  }
  std::vector<uint8_t> to_packed() const {
    std::vector<uint8_t> out((size_t)count() * len() * packed_bytes());
    check(fhe_b200_batch_pack(h_, 0, count(), out.data(), stream_));
    return out;
  }
  // TryConvertFrom<&Rq> for Poly<Ntt> (rq/convert.rs:116-131) for every polynomial of the batch
  static Ciphertext from_packed(std::shared_ptr<BfvParameters> par, const std::vector<uint8_t>& blobs, uint32_t count,
                                uint32_t parts = 2, uint32_t level = 0, Representation r = Representation::Ntt) {
    Ciphertext ct(std::move(par), count, parts, level, r);
    if (blobs.size() != (size_t)count * parts * ct.packed_bytes())
      throw Error(FHE_B200_INVALID_ARGUMENT, "InvalidCoefficientCount");
    check(fhe_b200_batch_unpack(ct.h_, 0, count, blobs.data(), ct.stream_));
    check(fhe_b200_sync(ct.stream_));
    return ct;
  }

  fhe_b200_batch* handle() const { return h_; }
  const std::shared_ptr<BfvParameters>& par() const { return par_; }
  void* stream() const { return stream_; }

 private:
  std::shared_ptr<BfvParameters> par_;
  fhe_b200_batch* h_ = nullptr;
  void* stream_ = nullptr;
};

class KeySwitchingKey {
 public:
  // c0, c1: NTT-domain words [n_digits][ksk_limbs][N] of the key polynomials (key_switching_key.rs:22-45)
  KeySwitchingKey(std::shared_ptr<BfvParameters> par, const std::vector<uint64_t>& c0, const std::vector<uint64_t>& c1,
                  uint32_t n_digits, uint32_t ciphertext_level = 0, uint32_t ksk_level = 0)
      : par_(std::move(par)), ciphertext_level_(ciphertext_level), ksk_level_(ksk_level) {
    check(fhe_b200_ksk_upload(par_->handle(), ciphertext_level, ksk_level, c0.data(), c1.data(), n_digits, &h_));
  }
  KeySwitchingKey(const KeySwitchingKey&) = delete;
  ~KeySwitchingKey() { fhe_b200_ksk_free(h_); }
  const fhe_b200_ksk* handle() const { return h_; }
  uint32_t ciphertext_level() const { return ciphertext_level_; }
  uint32_t ksk_level() const { return ksk_level_; }
  // KeySwitchingKey::key_switch (key_switching_key.rs:241-270, :323-362) on polynomial `part` of a power-basis batch:
  // the (c0, c1) pair as a 2-part NTT batch at the key level
  Ciphertext key_switch(const Ciphertext& p, uint32_t part = 0) const {
    Ciphertext out(par_, p.count(), 2, ksk_level_, Representation::Ntt, p.stream());
    check(fhe_b200_key_switch(p.handle(), part, h_, out.handle(), p.stream()));
    return out;
  }
  const std::shared_ptr<BfvParameters>& par() const { return par_; }

 private:
  std::shared_ptr<BfvParameters> par_;
  fhe_b200_ksk* h_ = nullptr;
  uint32_t ciphertext_level_, ksk_level_;
};

class RelinearizationKey {
 public:
  explicit RelinearizationKey(std::shared_ptr<KeySwitchingKey> ksk) : ksk(std::move(ksk)) {}
  // RelinearizationKey::relinearizes (relinearization_key.rs:70): (c0,c1,c2) -> (c0,c1)
  Ciphertext relinearizes(const Ciphertext& ct) const {
    Ciphertext out(ct.par(), ct.count(), 2, ct.level(), Representation::Ntt, ct.stream());
    check(fhe_b200_relinearize(ct.handle(), ksk->handle(), out.handle(), ct.stream()));
    return out;
  }
  std::shared_ptr<KeySwitchingKey> ksk;
};

class GaloisKey {
 public:
  GaloisKey(uint32_t exponent, std::shared_ptr<KeySwitchingKey> ksk) : exponent(exponent), ksk(std::move(ksk)) {}
  // GaloisKey::relinearize (galois_key.rs:63)
  Ciphertext relinearize(const Ciphertext& ct) const {
    Ciphertext out(ct.par(), ct.count(), 2, ct.level(), Representation::Ntt, ct.stream());
    check(fhe_b200_galois(ct.handle(), exponent, ksk->handle(), out.handle(), ct.stream()));
    return out;
  }
  uint32_t exponent;
  std::shared_ptr<KeySwitchingKey> ksk;
};

// rotation subset of EvaluationKey (evaluation_key.rs:110-170)
// fhe::bfv::RGSWCiphertext (bfv/rgsw_ciphertext.rs:20-24): two key-switching keys (for m and m*s) of one level
class RGSWCiphertext {
 public:
  RGSWCiphertext(std::shared_ptr<KeySwitchingKey> k0, std::shared_ptr<KeySwitchingKey> k1) : ksk0(std::move(k0)), ksk1(std::move(k1)) {
    if (ksk0->ksk_level() != ksk0->ciphertext_level() || ksk1->ksk_level() != ksk1->ciphertext_level() ||
        ksk0->ciphertext_level() != ksk1->ciphertext_level())
      throw Error(FHE_B200_INVALID_LEVEL, "InconsistentKeySwitchingLevels");   // rgsw_ciphertext.rs:58-70
  }
  // &Ciphertext * &RGSWCiphertext (rgsw_ciphertext.rs:122-155): key-switch both parts, add
  Ciphertext external_product(const Ciphertext& ct) const {
    if (ct.level() != ksk0->ciphertext_level()) throw Error(FHE_B200_INVALID_LEVEL, "Ciphertext and RGSWCiphertext must have the same level");
    if (ct.len() != 2) throw Error(FHE_B200_BAD_POLY_COUNT, "Ciphertext must have two parts");
    Ciphertext pb = ct.clone();
    pb.into_power_basis();
    Ciphertext out = ksk0->key_switch(pb, 0);
    out += ksk1->key_switch(pb, 1);
    return out;
  }
  std::shared_ptr<KeySwitchingKey> ksk0, ksk1;
};

class EvaluationKey {
 public:
  explicit EvaluationKey(std::shared_ptr<BfvParameters> par) : par_(std::move(par)) {}
  void add_galois_key(std::shared_ptr<GaloisKey> gk) { gk_[gk->exponent % (2 * (uint32_t)par_->degree())] = std::move(gk); }
  Ciphertext rotates_rows(const Ciphertext& ct) const { return at(2 * (uint32_t)par_->degree() - 1).relinearize(ct); }
  Ciphertext rotates_columns_by(const Ciphertext& ct, uint32_t i) const {
    uint64_t e = 1, m = 2 * par_->degree();
    for (uint32_t k = 0; k < i; k++) e = e * 3 % m;  // evaluation_key.rs:278-286
    return at((uint32_t)e).relinearize(ct);
  }

 private:
  const GaloisKey& at(uint32_t e) const {
    auto it = gk_.find(e);
    if (it == gk_.end()) throw Error(FHE_B200_INVALID_ARGUMENT, "EvaluationKeyError: rotation not supported by this key");
    return *it->second;
  }
  std::shared_ptr<BfvParameters> par_;
  std::map<uint32_t, std::shared_ptr<GaloisKey>> gk_;
};

// fhe::bfv::dot_product_scalar (bfv/ops/dot_product.rs:55): out[g] = sum_{i<n_terms} cts[g*n_terms+i] * pts[g*n_terms+i];
// pts is a batch of one-part NTT polynomials (Plaintext::poly_ntt).  An operand with exactly n_terms entries is shared
// by every group.
inline Ciphertext dot_product_scalar(const Ciphertext& cts, const Ciphertext& pts, uint32_t n_terms) {
  if (n_terms == 0) throw Error(FHE_B200_INVALID_ARGUMENT, "DotProductError::EmptyInput");
  const uint32_t groups = std::max(cts.count(), pts.count()) / n_terms;
  Ciphertext out(cts.par(), groups ? groups : 1, cts.len(), cts.level(), Representation::Ntt, cts.stream());
  check(fhe_b200_dot_product_scalar(cts.handle(), pts.handle(), n_terms, out.handle(), cts.stream()));
  return out;
}

// fhe_math::rns::ScalingFactor (rns/scaler.rs:20-58): numerator / denominator as little-endian byte strings
// (BigUint::to_bytes_le)
struct ScalingFactor {
  std::vector<uint8_t> numerator, denominator;
  static ScalingFactor one() { return ScalingFactor{{1}, {1}}; }
  static ScalingFactor from_u64(uint64_t num, uint64_t den) {
    auto le = [](uint64_t v) {
      std::vector<uint8_t> b;
      do { b.push_back((uint8_t)v); v >>= 8; } while (v);
      return b;
    };
    return ScalingFactor{le(num), le(den)};
  }
};

class Multiplicator {
 public:
  // Multiplicator::default (ops/mul.rs:101)
  static Multiplicator default_(const RelinearizationKey& rk) {
    Multiplicator m(rk.ksk->par(), rk.ksk->ciphertext_level());
    m.rk_ = std::make_shared<RelinearizationKey>(rk);
    return m;
  }
  // Multiplicator::new (ops/mul.rs:37-53)
  static Multiplicator new_(const ScalingFactor& lhs, const ScalingFactor& rhs, const std::vector<uint64_t>& extended_basis,
                            const ScalingFactor& post, std::shared_ptr<BfvParameters> par) {
    return new_leveled(lhs, rhs, extended_basis, post, 0, std::move(par));
  }
  // Multiplicator::new_leveled (ops/mul.rs:56-75)
  static Multiplicator new_leveled(const ScalingFactor& lhs, const ScalingFactor& rhs,
                                   const std::vector<uint64_t>& extended_basis, const ScalingFactor& post,
                                   uint32_t level, std::shared_ptr<BfvParameters> par) {
    Multiplicator m(par, level);
    fhe_b200_multiplicator* h = nullptr;
    check(fhe_b200_multiplicator_create(par->handle(), level, lhs.numerator.data(), (uint32_t)lhs.numerator.size(),
                                        lhs.denominator.data(), (uint32_t)lhs.denominator.size(), rhs.numerator.data(),
                                        (uint32_t)rhs.numerator.size(), rhs.denominator.data(),
                                        (uint32_t)rhs.denominator.size(), extended_basis.data(),
                                        (uint32_t)extended_basis.size(), nullptr, post.numerator.data(),
                                        (uint32_t)post.numerator.size(), post.denominator.data(),
                                        (uint32_t)post.denominator.size(), &h));
    m.h_ = std::shared_ptr<fhe_b200_multiplicator>(h, [](fhe_b200_multiplicator* x) { fhe_b200_multiplicator_free(x); });
    return m;
  }
  // Multiplicator::enable_relinearization (ops/mul.rs:141-151)
  void enable_relinearization(const RelinearizationKey& rk) {
    if (rk.ksk->par() != par_ || rk.ksk->ciphertext_level() != level_)
      throw Error(FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch");
    rk_ = std::make_shared<RelinearizationKey>(rk);
  }
  // Multiplicator::enable_mod_switching (ops/mul.rs:155)
  void enable_mod_switching() {
    if (level_ >= par_->max_level()) throw Error(FHE_B200_NO_MORE_CONTEXT, "NoMoreContext");
    mod_switch_ = true;
  }
  // Multiplicator::multiply (ops/mul.rs:165)
  Ciphertext multiply(const Ciphertext& lhs, const Ciphertext& rhs) const {
    if (lhs.level() != level_ || rhs.level() != level_) throw Error(FHE_B200_INVALID_LEVEL, "InvalidLevel");
    const uint32_t parts = rk_ ? 2 : 3;
    Ciphertext out(lhs.par(), lhs.count(), parts, level_ + (mod_switch_ ? 1 : 0), Representation::Ntt, lhs.stream());
    if (!h_) {
      check(fhe_b200_mul_relin(lhs.handle(), rhs.handle(), rk_->ksk->handle(), mod_switch_ ? 1 : 0, out.handle(),
                               lhs.stream()));
    } else {
      check(fhe_b200_multiplicator_multiply(h_.get(), lhs.handle(), rhs.handle(), rk_ ? rk_->ksk->handle() : nullptr,
                                            mod_switch_ ? 1 : 0, out.handle(), lhs.stream()));
    }
    return out;
  }

 private:
  Multiplicator(std::shared_ptr<BfvParameters> par, uint32_t level) : par_(std::move(par)), level_(level) {}
  std::shared_ptr<BfvParameters> par_;
  std::shared_ptr<RelinearizationKey> rk_;
  std::shared_ptr<fhe_b200_multiplicator> h_;   // custom strategy; empty = fused default path
  uint32_t level_;
  bool mod_switch_ = false;
};

}  // namespace bfv
}  // namespace fhe_b200
