// Minimal arbitrary-precision unsigned integer for the one-off host precompute
// (RNS products, Garner coefficients, scaler theta/omega tables).  Stands in for
// num-bigint's BigUint at the call sites rns/mod.rs:52-116 and rns/scaler.rs:79-229
// of the reference.  32-bit limbs, little endian, schoolbook mul, Knuth-D division.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace fhe_b200 {

class BigUint {
 public:
  std::vector<uint32_t> w;  // little endian, no trailing zero limbs

  BigUint() = default;
  BigUint(uint64_t v) {  // NOLINT(implicit)
    if (v) w.push_back(static_cast<uint32_t>(v));
    if (v >> 32) w.push_back(static_cast<uint32_t>(v >> 32));
  }
  static BigUint from_u128(unsigned __int128 v) {
    BigUint r;
    for (int i = 0; i < 4; i++) r.w.push_back(static_cast<uint32_t>(v >> (32 * i)));
    r.trim();
    return r;
  }
  static BigUint from_le_bytes(const uint8_t* b, size_t n) {
    BigUint r;
    r.w.assign((n + 3) / 4, 0);
    for (size_t i = 0; i < n; i++) r.w[i / 4] |= static_cast<uint32_t>(b[i]) << (8 * (i % 4));
    r.trim();
    return r;
  }
  void trim() {
    while (!w.empty() && w.back() == 0) w.pop_back();
  }
  bool is_zero() const { return w.empty(); }
  bool is_odd() const { return !w.empty() && (w[0] & 1); }
  size_t bits() const {
    if (w.empty()) return 0;
    return 32 * (w.size() - 1) + (32 - __builtin_clz(w.back()));
  }
  bool fits_u64() const { return w.size() <= 2; }
  uint64_t to_u64() const {
    uint64_t v = 0;
    if (w.size() > 0) v |= w[0];
    if (w.size() > 1) v |= static_cast<uint64_t>(w[1]) << 32;
    return v;
  }
  unsigned __int128 low_u128() const {
    unsigned __int128 v = 0;
    for (size_t i = 0; i < 4 && i < w.size(); i++) v |= static_cast<unsigned __int128>(w[i]) << (32 * i);
    return v;
  }

  static int cmp(const BigUint& a, const BigUint& b) {
    if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
    for (size_t i = a.w.size(); i-- > 0;)
      if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
  }
  bool operator<(const BigUint& o) const { return cmp(*this, o) < 0; }
  bool operator>(const BigUint& o) const { return cmp(*this, o) > 0; }
  bool operator<=(const BigUint& o) const { return cmp(*this, o) <= 0; }
  bool operator>=(const BigUint& o) const { return cmp(*this, o) >= 0; }
  bool operator==(const BigUint& o) const { return cmp(*this, o) == 0; }
  bool operator!=(const BigUint& o) const { return cmp(*this, o) != 0; }

  BigUint operator+(const BigUint& o) const {
    BigUint r;
    size_t n = std::max(w.size(), o.w.size());
    r.w.resize(n + 1);
    uint64_t c = 0;
    for (size_t i = 0; i < n; i++) {
      c += (i < w.size() ? w[i] : 0ull) + (i < o.w.size() ? o.w[i] : 0ull);
      r.w[i] = static_cast<uint32_t>(c);
      c >>= 32;
    }
    r.w[n] = static_cast<uint32_t>(c);
    r.trim();
    return r;
  }
  // requires *this >= o
  BigUint operator-(const BigUint& o) const {
    if (cmp(*this, o) < 0) throw std::runtime_error("BigUint underflow");
    BigUint r;
    r.w.resize(w.size());
    int64_t b = 0;
    for (size_t i = 0; i < w.size(); i++) {
      int64_t d = static_cast<int64_t>(w[i]) - (i < o.w.size() ? o.w[i] : 0) - b;
      b = d < 0;
      r.w[i] = static_cast<uint32_t>(d + (b ? (1ll << 32) : 0));
    }
    r.trim();
    return r;
  }
  BigUint operator*(const BigUint& o) const {
    BigUint r;
    if (is_zero() || o.is_zero()) return r;
    r.w.assign(w.size() + o.w.size(), 0);
    for (size_t i = 0; i < w.size(); i++) {
      uint64_t c = 0;
      for (size_t j = 0; j < o.w.size(); j++) {
        c += static_cast<uint64_t>(w[i]) * o.w[j] + r.w[i + j];
        r.w[i + j] = static_cast<uint32_t>(c);
        c >>= 32;
      }
      r.w[i + o.w.size()] = static_cast<uint32_t>(c);
    }
    r.trim();
    return r;
  }
  BigUint operator<<(size_t s) const {
    if (is_zero()) return *this;
    BigUint r;
    size_t ws = s / 32, bs = s % 32;
    r.w.assign(w.size() + ws + 1, 0);
    for (size_t i = 0; i < w.size(); i++) {
      uint64_t v = static_cast<uint64_t>(w[i]) << bs;
      r.w[i + ws] |= static_cast<uint32_t>(v);
      r.w[i + ws + 1] |= static_cast<uint32_t>(v >> 32);
    }
    r.trim();
    return r;
  }
  BigUint operator>>(size_t s) const {
    BigUint r;
    size_t ws = s / 32, bs = s % 32;
    if (ws >= w.size()) return r;
    r.w.assign(w.size() - ws, 0);
    for (size_t i = ws; i < w.size(); i++) {
      uint64_t v = w[i];
      if (i + 1 < w.size()) v |= static_cast<uint64_t>(w[i + 1]) << 32;
      r.w[i - ws] = static_cast<uint32_t>(v >> bs);
    }
    r.trim();
    return r;
  }

  // Knuth algorithm D.  Returns quotient; remainder via out parameter.
  static BigUint divmod(const BigUint& a, const BigUint& b, BigUint* rem) {
    if (b.is_zero()) throw std::runtime_error("BigUint division by zero");
    if (cmp(a, b) < 0) {
      if (rem) *rem = a;
      return BigUint();
    }
    if (b.w.size() == 1) {
      BigUint q;
      q.w.assign(a.w.size(), 0);
      uint64_t r = 0, d = b.w[0];
      for (size_t i = a.w.size(); i-- > 0;) {
        uint64_t cur = (r << 32) | a.w[i];
        q.w[i] = static_cast<uint32_t>(cur / d);
        r = cur % d;
      }
      q.trim();
      if (rem) *rem = BigUint(r);
      return q;
    }
    int s = __builtin_clz(b.w.back());
    BigUint u = a << s, v = b << s;
    const size_t n = v.w.size();
    u.w.resize(a.w.size() + 1, 0);  // the normalising shift adds at most one limb
    const size_t m = a.w.size() - n;
    BigUint q;
    q.w.assign(m + 1, 0);
    const uint64_t B = 1ull << 32;
    for (size_t j = m + 1; j-- > 0;) {
      uint64_t num = (static_cast<uint64_t>(u.w[j + n]) << 32) | u.w[j + n - 1];
      uint64_t qhat = num / v.w[n - 1], rhat = num % v.w[n - 1];
      while (qhat >= B || qhat * v.w[n - 2] > ((rhat << 32) | u.w[j + n - 2])) {
        qhat--;
        rhat += v.w[n - 1];
        if (rhat >= B) break;
      }
      int64_t borrow = 0;
      uint64_t carry = 0;
      for (size_t i = 0; i < n; i++) {
        uint64_t p = qhat * v.w[i] + carry;
        carry = p >> 32;
        int64_t t = static_cast<int64_t>(u.w[i + j]) - borrow - static_cast<int64_t>(p & 0xffffffffull);
        borrow = t < 0;
        u.w[i + j] = static_cast<uint32_t>(t);
      }
      int64_t t = static_cast<int64_t>(u.w[j + n]) - borrow - static_cast<int64_t>(carry);
      borrow = t < 0;
      u.w[j + n] = static_cast<uint32_t>(t);
      if (borrow) {
        qhat--;
        uint64_t c = 0;
        for (size_t i = 0; i < n; i++) {
          c += static_cast<uint64_t>(u.w[i + j]) + v.w[i];
          u.w[i + j] = static_cast<uint32_t>(c);
          c >>= 32;
        }
        u.w[j + n] += static_cast<uint32_t>(c);
      }
      q.w[j] = static_cast<uint32_t>(qhat);
    }
    q.trim();
    if (rem) {
      u.w.resize(n);
      u.trim();
      *rem = u >> s;
    }
    return q;
  }
  BigUint operator/(const BigUint& o) const { return divmod(*this, o, nullptr); }
  BigUint operator%(const BigUint& o) const {
    BigUint r;
    divmod(*this, o, &r);
    return r;
  }
  uint64_t mod_u64(uint64_t m) const {
    unsigned __int128 r = 0;
    for (size_t i = w.size(); i-- > 0;) r = ((r << 32) | w[i]) % m;
    return static_cast<uint64_t>(r);
  }
};

}  // namespace fhe_b200
