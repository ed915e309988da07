// C++ host-API smoke test: drives the engine through include/fhe_b200.hpp (i.e. through the C ABI) and checks
// size-independent properties on the GPU: NTT-domain linearity of add/sub/neg, (a*b) relinearized twice gives
// identical results, and the error behaviour of Multiplicator::multiply (ops/mul.rs:168-189).
// With a third argument (a directory prepared by the test from the CPU oracle) it also consumes the reference's
// protobuf messages through include/fhe_b200_wire.hpp -- ciphertexts, a relinearization key, a Galois key -- and writes
// back words and messages for the test to compare with the oracle's.
// Built and run by tests/test_gpu_cpp_host.py.   usage: host_api_test <degree> <n_moduli> [<message dir>]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>

#include "fhe_b200.hpp"
#include "fhe_b200_wire.hpp"

using namespace fhe_b200::bfv;

static std::vector<std::string> read_records(const std::string& path) {   // u32 length + bytes, repeated
  std::ifstream in(path, std::ios::binary);
  std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::vector<std::string> out;
  for (size_t pos = 0; pos + 4 <= data.size();) {
    uint32_t n;
    memcpy(&n, &data[pos], 4);
    out.push_back(data.substr(pos + 4, n));
    pos += 4 + n;
  }
  return out;
}
static void write_records(const std::string& path, const std::vector<std::string>& recs) {
  std::ofstream out(path, std::ios::binary);
  for (auto& r : recs) {
    uint32_t n = (uint32_t)r.size();
    out.write((const char*)&n, 4);
    out.write(r.data(), n);
  }
}

int main(int argc, char** argv) {
  const uint32_t degree = argc > 1 ? (uint32_t)atoi(argv[1]) : 64, nmod = argc > 2 ? (uint32_t)atoi(argv[2]) : 3;
  try {
    auto par = BfvParametersBuilder().set_degree(degree).set_plaintext_modulus(1153).set_moduli_sizes(
        std::vector<uint32_t>(nmod, 62)).build_arc();
    auto q = par->moduli();
    std::mt19937_64 rng(42);
    const uint32_t count = 3;
    auto rnd = [&](uint32_t polys) {
      std::vector<uint64_t> w((size_t)polys * nmod * degree);
      for (uint32_t p = 0; p < polys; p++)
        for (uint32_t i = 0; i < nmod; i++)
          for (uint32_t c = 0; c < degree; c++) w[((size_t)p * nmod + i) * degree + c] = rng() % q[i];
      return w;
    };
    auto wa = rnd(count * 2), wb = rnd(count * 2);
    auto A = Ciphertext::from_host(par, wa, count), B = Ciphertext::from_host(par, wb, count);
    // (A + B) - B == A ; -(-A) == A
    auto S = A.clone();
    S += B;
    S -= B;
    if (S.to_host() != wa) { printf("FAIL add/sub\n"); return 1; }
    if ((-(-A)).to_host() != wa) { printf("FAIL neg\n"); return 1; }
    // keys: random key material is enough for determinism / shape checks
    auto k0 = rnd(nmod), k1 = rnd(nmod);
    auto ksk = std::make_shared<KeySwitchingKey>(par, k0, k1, nmod);
    RelinearizationKey rk(ksk);
    auto m = Multiplicator::default_(rk);
    auto P1 = m.multiply(A, B).to_host();
    auto C3 = A * B;
    if (C3.len() != 3) { printf("FAIL parts\n"); return 1; }
    auto P2 = rk.relinearizes(C3).to_host();
    if (P1 != P2) { printf("FAIL multiply != mul + relinearizes\n"); return 1; }
    // multiplication is commutative bit for bit (canonical residues)
    if (m.multiply(B, A).to_host() != P1) { printf("FAIL commutativity\n"); return 1; }
    // wire format round trip (Rq.coefficients blobs)
    if (Ciphertext::from_packed(par, A.to_packed(), count).to_host() != wa) { printf("FAIL wire round trip\n"); return 1; }
    // custom strategy (Multiplicator::new, ops/mul.rs:37): the default strategy rebuilt from its parts must agree
    // bit for bit with the fused default path
    {
      uint32_t nb = 0;
      fhe_b200::check(fhe_b200_params_mul_basis(par->handle(), 0, nullptr, &nb));
      std::vector<uint64_t> basis(nb);
      fhe_b200::check(fhe_b200_params_mul_basis(par->handle(), 0, basis.data(), &nb));
      // post factor t/Q: Q as little-endian bytes by schoolbook multiplication of the moduli
      std::vector<uint64_t> mods(nmod);
      fhe_b200::check(fhe_b200_params_moduli(par->handle(), mods.data()));
      std::vector<uint8_t> Q{1};
      for (uint64_t q : mods) {
        std::vector<uint8_t> r(Q.size() + 8, 0);
        for (size_t i = 0; i < Q.size(); i++) {
          unsigned __int128 carry = 0;
          for (size_t j = 0; j < 8; j++) {
            unsigned __int128 cur = (unsigned __int128)r[i + j] + (unsigned __int128)Q[i] * ((q >> (8 * j)) & 0xff) + carry;
            r[i + j] = (uint8_t)cur;
            carry = cur >> 8;
          }
          for (size_t k = i + 8; carry && k < r.size(); k++) {
            unsigned __int128 cur = (unsigned __int128)r[k] + carry;
            r[k] = (uint8_t)cur;
            carry = cur >> 8;
          }
        }
        Q = r;
      }
      ScalingFactor post = ScalingFactor::from_u64(1153, 1);
      post.denominator = Q;
      auto mc = Multiplicator::new_(ScalingFactor::one(), ScalingFactor::one(), basis, post, par);
      if (mc.multiply(A, B).len() != 3) { printf("FAIL custom parts\n"); return 1; }
      mc.enable_relinearization(rk);
      if (mc.multiply(A, B).to_host() != P1) { printf("FAIL custom strategy != default\n"); return 1; }
    }
    // pinned staging + enqueue-only transfers (fhe_b200_host_alloc, upload / download_async, sync)
    {
      PinnedWords up(wa.size(), true), down(wa.size());
      for (size_t i = 0; i < wa.size(); i++) up.data()[i] = wa[i];
      Ciphertext X(par, count);
      X.upload_async(up.data(), 0, count);
      X += B;
      X -= B;
      X.download_async(down.data(), 0, count);
      X.sync();
      for (size_t i = 0; i < wa.size(); i++)
        if (down.data()[i] != wa[i]) { printf("FAIL pinned round trip\n"); return 1; }
    }
    // protobuf messages: round trip of the batch, then the oracle's messages if the test supplied them
    {
      auto msgs = to_bytes(A);
      if (ciphertext_from_bytes(par, msgs).to_host() != wa) { printf("FAIL message round trip\n"); return 1; }
      if (to_bytes(ciphertext_from_bytes(par, msgs)) != msgs) { printf("FAIL message bytes not stable\n"); return 1; }
      try {
        ciphertext_from_bytes(par, {msgs[0].substr(0, msgs[0].size() - 3)});
        printf("FAIL expected Decode\n");
        return 1;
      } catch (const fhe_b200::WireError& e) {
        if (e.variant != "Decode") { printf("FAIL wrong variant %s\n", e.variant.c_str()); return 1; }
      }
    }
    if (argc > 3) {
      const std::string dir = argv[3];
      auto X = ciphertext_from_bytes(par, read_records(dir + "/cts.bin"));
      auto xw = X.to_host();
      std::ofstream(dir + "/out_words.bin", std::ios::binary).write((const char*)xw.data(), (std::streamsize)(xw.size() * 8));
      write_records(dir + "/out_msgs.bin", to_bytes(X));
      auto ork = relinearization_key_from_bytes(par, read_records(dir + "/relin.bin")[0]);
      auto ogk = galois_key_from_bytes(par, read_records(dir + "/galois.bin")[0]);
      write_records(dir + "/out_mul.bin", to_bytes(Multiplicator::default_(ork).multiply(X, X)));
      write_records(dir + "/out_rot.bin", to_bytes(ogk.relinearize(X)));
      // seeded form of the same ciphertexts: the host supplies the expanded halves
      auto halves = read_records(dir + "/halves.bin")[0];
      auto S = ciphertext_from_bytes(par, read_records(dir + "/cts_seeded.bin"), (const uint64_t*)halves.data());
      if (S.to_host() != xw) { printf("FAIL seeded messages\n"); return 1; }
      try {
        ciphertext_from_bytes(par, read_records(dir + "/cts_seeded.bin"));
        printf("FAIL expected SeedExpansion\n");
        return 1;
      } catch (const fhe_b200::WireError& e) {
        if (e.variant != "SeedExpansion" || e.code != FHE_B200_UNSUPPORTED) { printf("FAIL wrong seeded error\n"); return 1; }
      }
    }
    // error behaviour
    try {
      m.multiply(C3, B);
      printf("FAIL expected MultiplicationPolynomialCount\n");
      return 1;
    } catch (const fhe_b200::Error& e) {
      if (e.code != FHE_B200_BAD_POLY_COUNT) { printf("FAIL wrong code %d\n", e.code); return 1; }
    }
    printf("OK degree=%u moduli=%u\n", degree, nmod);
    return 0;
  } catch (const fhe_b200::Error& e) {
    printf("ERROR %d: %s\n", e.code, e.what());
    return 2;
  }
}
