// CPU test driver of include/fhe_b200_wire.hpp's codec: decodes every record of <in> and re-encodes it canonically
// into <out>; tests/test_wire_cpu.py compares the output with the google.protobuf runtime byte for byte.
// Record: kind (1 byte: 'q' Rq, 'c' Ciphertext, 'k' KeySwitchingKey, 'r' RelinearizationKey, 'g' GaloisKey),
// u32 length, bytes.  Output record: u32 length + bytes, or length 0xFFFFFFFF followed by u32 n + the error variant.
#include <cstdio>
#include <fstream>
#include <iterator>

#include "fhe_b200_wire.hpp"

using namespace fhe_b200;

static std::string str(const wire::Span& s) { return std::string((const char*)s.p, s.n); }
static std::vector<std::string> strs(const std::vector<wire::Span>& v) {
  std::vector<std::string> o;
  for (auto& s : v) o.push_back(str(s));
  return o;
}
static std::string ksk_again(const void* p, size_t n) {
  wire::KskMsg k = wire::decode_ksk(p, n);
  return wire::encode_ksk(strs(k.c0), strs(k.c1), str(k.seed), k.ciphertext_level, k.ksk_level, k.log_base);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::ofstream out(argv[2], std::ios::binary);
  size_t pos = 0;
  while (pos + 5 <= data.size()) {
    char kind = data[pos];
    uint32_t n;
    memcpy(&n, &data[pos + 1], 4);
    const char* p = data.data() + pos + 5;
    pos += 5 + n;
    std::string res;
    try {
      if (kind == 'q') {
        wire::Rq m = wire::decode_rq(p, n);
        res = wire::encode_rq(m.representation, m.degree, m.coefficients.p, m.coefficients.n);
      } else if (kind == 'c') {
        wire::CiphertextMsg m = wire::decode_ciphertext(p, n);
        res = wire::encode_ciphertext(strs(m.c), str(m.seed), m.level);
      } else if (kind == 'k') {
        res = ksk_again(p, n);
      } else if (kind == 'r') {
        wire::Span s = wire::sub_message(p, n, 1, "RelinearizationKeySwitchingKey");
        res = wire::encode_relinearization_key(ksk_again(s.p, s.n));
      } else if (kind == 'g') {
        uint32_t e = 0;
        wire::Span s = wire::sub_message(p, n, 1, "GaloisKeySwitchingKey", &e);
        res = wire::encode_galois_key(ksk_again(s.p, s.n), e);
      }
      uint32_t len = (uint32_t)res.size();
      out.write((const char*)&len, 4);
      out.write(res.data(), len);
    } catch (const WireError& e) {
      uint32_t bad = 0xFFFFFFFFu, len = (uint32_t)e.variant.size();
      out.write((const char*)&bad, 4);
      out.write((const char*)&len, 4);
      out.write(e.variant.data(), len);
    }
  }
  return 0;
}
