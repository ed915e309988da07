// C ABI of the engine (include/fhe_b200.h): parameter precompute + upload, device batches,
// key material, and the batched homomorphic operations built from the kernels of
// ntt.cu / kernels.cu.  No CPU fallback: compute entry points require a CUDA device.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fhe_b200.h"
#include "engine.hpp"
#include "host/precompute.hpp"

using namespace fhe_b200;

namespace {

thread_local std::string g_last_error;

struct ScalerData {
  ScalerTablesH h;
  ScalerDev dev;
};

struct LevelData {
  u32 level = 0, L = 0, E = 0, K = 0;
  RowIds ctx_ids, mul_ids;
  std::vector<u64> mul_moduli;
  ScalerData ext, down;
  bool has_sd = false;
  SwitchDownDev sd;
};

// Device copy of one RnsScaler's tables (rns/scaler.rs:79-175).  `to_dev` uploads a vector and keeps ownership of the
// allocation, `index_of` maps a modulus of the `to` basis to its slot in the limb table the kernels will be given.
template <typename ToDev, typename IndexOf>
void upload_scaler_tables(ScalerData& s, const std::vector<u64>& to_moduli, ToDev&& to_dev, IndexOf&& index_of) {
  ScalerDev& d = s.dev;
  std::memset(&d, 0, sizeof(d));
  d.n_from = s.h.n_from; d.n_to = s.h.n_to; d.is_one = s.h.is_one; d.shift = s.h.shift;
  d.tg_lo = s.h.theta_gamma_lo; d.tg_hi = s.h.theta_gamma_hi; d.tg_sign = s.h.theta_gamma_sign;
  for (size_t j = 0; j < to_moduli.size(); j++) d.to_ids[j] = (unsigned short)index_of(to_moduli[j]);
  d.all_solinas = getenv("FHE_B200_NO_SOLINAS") ? 0 : 1;
  for (u64 q : to_moduli)
    if ((q >> 61) != 1 || ((1ull << 62) - q) >= (1ull << 28)) d.all_solinas = 0;
  d.gamma = to_dev(s.h.gamma);
  d.omega = to_dev(s.h.omega);
  d.to_lo = to_dev(s.h.theta_omega_lo);
  d.to_hi = to_dev(s.h.theta_omega_hi);
  d.to_sign = to_dev(s.h.theta_omega_sign);
  // source indices of the theta_omega terms, positive sign first (the kernel makes one pass per sign).  Terms whose
  // fractional part is exactly zero add nothing (rns/scaler.rs:282-298 multiplies them by zero) and are left out:
  // in the down scaler of the multiplication basis that is every extension limb (garner_i * t / Q is an integer).
  std::vector<unsigned char> order;
  d.n_pos = 0;
  for (int sg = 0; sg < 2; sg++)
    for (size_t i = 0; i < s.h.theta_omega_sign.size(); i++)
      if ((int)s.h.theta_omega_sign[i] == sg && (s.h.theta_omega_lo[i] | s.h.theta_omega_hi[i]) != 0) {
        order.push_back((unsigned char)i);
        d.n_pos += sg == 0;
      }
  d.n_terms = (u32)order.size();
  if (order.empty()) order.push_back(0);   // keep the table non-empty (never read: n_terms == 0)
  d.to_order = to_dev(order);
  d.tgar_lo = to_dev(s.h.theta_garner_lo);
  d.tgar_hi = to_dev(s.h.theta_garner_hi);
}

}  // namespace

struct fhe_b200_params {
  // Arc-like lifetime (the reference shares Arc<BfvParameters>): batches and keys hold a
  // reference, so the tables outlive every handle that points at them regardless of the order
  // in which a garbage-collected host releases its objects.
  std::atomic<int> refs{1};
  int device = -1;
  u32 N = 0, logn = 0, Lmax = 0;
  std::vector<u64> moduli, ext, primes, psi;
  std::vector<u32> moduli_sizes;
  BigUint t;
  std::vector<NttTablesH> tables;  // host copies (kept for inspection / host-only handles)
  std::vector<LimbDev> h_limbs;
  LimbDev* d_limbs = nullptr;
  std::vector<void*> d_allocs;
  // scratch of the batched operations comes from a stream-ordered pool this parameter set owns (the device's default
  // pool, which a host application may be using for its own cudaMallocAsync calls, is left untouched)
  cudaMemPool_t pool = nullptr;
  // side streams over which ChunkRunner deals the chunks of one batched call (created on first use)
  mutable cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
  mutable std::mutex mu;
  mutable std::map<u32, std::unique_ptr<LevelData>> levels;
  mutable std::map<u32, int*> perms;

  template <typename T>
  T* to_dev(const std::vector<T>& v) const {
    if (device < 0 || v.empty()) return nullptr;
    T* d = nullptr;
    // tables are built lazily, possibly from a thread whose current device differs: select ours, put theirs back
    struct Restore {
      int prev = -1;
      ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    if (cudaGetDevice(&restore.prev) != cudaSuccess) { cudaGetLastError(); restore.prev = -1; }
    if (restore.prev == device) restore.prev = -1;
    else FHE_CUDA(cudaSetDevice(device));
    FHE_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
    const_cast<fhe_b200_params*>(this)->d_allocs.push_back(d);   // owned from here on (freed with the set)
    FHE_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  int prime_index(u64 q) const {
    for (size_t i = 0; i < primes.size(); i++)
      if (primes[i] == q) return (int)i;
    return -1;
  }
  void upload_scaler(ScalerData& s, const std::vector<u64>& to_moduli) const {
    upload_scaler_tables(s, to_moduli, [&](const auto& v) { return to_dev(v); }, [&](u64 q) { return prime_index(q); });
  }

  // ContextLevel + MultiplicationParameters of one level (bfv/parameters.rs:600-700, :793-813)
  const LevelData& level(u32 lv) const {
    std::lock_guard<std::mutex> g(mu);
    auto it = levels.find(lv);
    if (it != levels.end()) return *it->second;
    if (lv >= Lmax) throw FheError(FHE_B200_INVALID_LEVEL, "InvalidLevel: level " + std::to_string(lv));
    std::unique_ptr<LevelData> d(new LevelData());
    d->level = lv;
    d->L = Lmax - lv;
    u32 bits = 0;
    for (u32 i = 0; i < d->L; i++) bits += moduli_sizes[i];
    d->E = (bits + 60 + 61) / 62;  // (modulus_size + 60).div_ceil(62), parameters.rs:689
    d->K = d->L + d->E;
    if (d->K > (u32)kMaxPos) throw FheError(FHE_B200_UNSUPPORTED, "too many limbs");
    std::vector<u64> ctx(moduli.begin(), moduli.begin() + d->L);
    d->mul_moduli = ctx;
    d->mul_moduli.insert(d->mul_moduli.end(), ext.begin(), ext.begin() + d->E);
    std::memset(&d->ctx_ids, 0, sizeof(RowIds));
    std::memset(&d->mul_ids, 0, sizeof(RowIds));
    d->ctx_ids.limbs_per_poly = d->L;
    d->mul_ids.limbs_per_poly = d->K;
    for (u32 i = 0; i < d->L; i++) d->ctx_ids.ids[i] = d->mul_ids.ids[i] = (unsigned short)i;
    for (u32 j = 0; j < d->E; j++) d->mul_ids.ids[d->L + j] = (unsigned short)(Lmax + j);
    RnsContextH from(ctx), to(d->mul_moduli);
    d->ext.h = make_scaler_tables(from, to, BigUint(1), BigUint(1));
    d->down.h = make_scaler_tables(to, from, t, from.product);
    upload_scaler(d->ext, d->mul_moduli);
    upload_scaler(d->down, ctx);
    if (d->L >= 2) {  // rq/context.rs:65-71 and rq/mod.rs:444-468
      d->has_sd = true;
      u64 ql = ctx.back();
      std::vector<u64> half_mod, inv, inv_s;
      for (u32 i = 0; i + 1 < d->L; i++) {
        u64 qi = ctx[i], iv;
        if (!invmod_h(ql % qi, qi, &iv)) throw FheError(FHE_B200_INVALID_MODULUS, "NonCoprimeModuli");
        half_mod.push_back(qi - (ql / 2) % qi);
        inv.push_back(iv);
        inv_s.push_back(ModulusH(qi).shoup(iv));
      }
      d->sd.q_last = ql;
      d->sd.q_last_half = ql / 2;
      d->sd.half_mod = to_dev(half_mod);
      d->sd.inv = to_dev(inv);
      d->sd.inv_s = to_dev(inv_s);
    }
    auto* raw = d.get();
    levels[lv] = std::move(d);
    return *raw;
  }

  // SubstitutionExponent::new (rq/mod.rs:99-121) folded with ctx.bitrev into one gather table:
  // out[bitrev(j)] = in[bitrev((j*e + (e-1)/2) mod N)]
  const int* perm(u32 exponent) const {
    std::lock_guard<std::mutex> g(mu);
    auto it = perms.find(exponent);
    if (it != perms.end()) return it->second;
    std::vector<int> p(N);
    auto brev = [&](u32 x) {
      u32 r = 0;
      for (u32 b = 0; b < logn; b++) r |= ((x >> b) & 1) << (logn - 1 - b);
      return r;
    };
    u64 power = (exponent - 1) / 2;
    for (u32 j = 0; j < N; j++) {
      p[brev(j)] = (int)brev((u32)(power & (N - 1)));
      power += exponent;
    }
    int* d = to_dev(p);
    perms[exponent] = d;
    return d;
  }
};

struct fhe_b200_batch {
  const fhe_b200_params* par;
  u32 count, parts, level, limbs;
  int repr;
  bool mul_basis;
  u64* d;
  size_t words_per_ct() const { return ((size_t)parts * limbs) << par->logn; }
};

struct fhe_b200_ksk {
  const fhe_b200_params* par;
  u32 ct_level, ksk_level, n_dig, Lk;
  u32 log_base;   // 0: RNS-digit variant; else base-2^log_base decomposition (single-modulus key level)
  u64 *k0, *k1;
};

// Multiplicator::new / new_leveled (bfv/ops/mul.rs:37-98): custom scaling factors and extended basis.
struct fhe_b200_multiplicator {
  const fhe_b200_params* par;
  u32 level = 0, L = 0, K = 0;
  u32 nc_l = 0, nc_r = 0, nc_d = 0;     // Scaler::number_common_moduli of the two extenders and the down scaler
  std::vector<u64> mul_moduli, plan_primes;
  RowIds mul_ids;
  std::vector<LimbDev> h_limbs;          // the parameter set's limbs followed by the primes only this basis has
  LimbDev* d_limbs = nullptr;
  ScalerData ext_l, ext_r, down;
  std::vector<void*> d_allocs;
  template <typename T>
  T* to_dev(const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    T* d = nullptr;
    FHE_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
    d_allocs.push_back(d);
    FHE_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  int prime_index(u64 q) const {
    for (size_t i = 0; i < plan_primes.size(); i++)
      if (plan_primes[i] == q) return (int)i;
    return -1;
  }
  void upload_scaler(ScalerData& sd, const std::vector<u64>& to_moduli) {
    upload_scaler_tables(sd, to_moduli, [&](const auto& v) { return to_dev(v); }, [&](u64 q) { return prime_index(q); });
  }
};

namespace {

void params_release(const fhe_b200_params* cp) {
  fhe_b200_params* p = const_cast<fhe_b200_params*>(cp);
  if (!p || p->refs.fetch_sub(1) != 1) return;
  if (p->device >= 0) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    for (void* d : p->d_allocs) cudaFree(d);
    for (cudaStream_t ss : p->side)
      if (ss) cudaStreamDestroy(ss);
    if (p->pool) {
      cudaDeviceSynchronize();   // scratch freed with cudaFreeAsync must have retired before its pool goes away
      cudaMemPoolDestroy(p->pool);
    }
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
  }
  delete p;
}
const fhe_b200_params* params_retain(const fhe_b200_params* p) {
  const_cast<fhe_b200_params*>(p)->refs.fetch_add(1);
  return p;
}

// Per-prime device constants and twiddle tables (NttOperator::new, ntt/native.rs:35-73, as (value, companion) pairs).
template <typename ToDev>
LimbDev make_limb_dev(u64 q, const NttTablesH& t, ToDev&& to_dev) {
  ModulusH m(q);
  LimbDev d;
  std::memset(&d, 0, sizeof(d));
  d.p = q; d.p2 = 2 * q; d.bhi = m.bhi; d.blo = m.blo; d.c128 = m.c128;
  d.ninv = t.ninv; d.zn = t.zn;
  // limb mode: p = 2^62 - c with c < 2^28 takes the Solinas constant-multiplication form
  const u64 cc = (1ull << 62) - q;
  const bool sol = (q >> 61) == 1 && cc < (1ull << 28) && !getenv("FHE_B200_NO_SOLINAS");
  // NTT butterflies: Shoup pairs by default (bench_micro/bf_bench: 3.21 vs <= 3.02 butterflies/clk/SM for every
  // Solinas instruction selection tried); FHE_B200_SOLINAS_NTT=1 selects the (w, w*2^32 mod p) pairs instead
  const bool sol_ntt = getenv("FHE_B200_SOLINAS_NTT") != nullptr;
  auto pairs = [&](const std::vector<u64>& v, const std::vector<u64>& shoup) {
    std::vector<ulonglong2> o(v.size());
    for (size_t k = 0; k < v.size(); k++) {
      o[k].x = v[k];
      o[k].y = (sol && sol_ntt) ? (u64)((((u128)v[k]) << 32) % q) : shoup[k];
    }
    return o;
  };
  d.sol_c = sol ? cc : 0;
  d.sol_ntt = (sol && sol_ntt) ? 1 : 0;
  d.ninv_s = (sol && sol_ntt) ? (u64)((((u128)t.ninv) << 32) % q) : t.ninv_s;
  d.zn_s = (sol && sol_ntt) ? (u64)((((u128)t.zn) << 32) % q) : t.zn_s;
  d.om = to_dev(pairs(t.om, t.om_s));
  d.zi = to_dev(pairs(t.zi, t.zi_s));
  return d;
}

// selects the parameter set's device for the duration of an API call and puts the caller's device back afterwards
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const fhe_b200_params* p) {
    if (p->device < 0) throw FheError(FHE_B200_NO_DEVICE, "parameter set was created without a CUDA device");
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != p->device) FHE_CUDA(cudaSetDevice(p->device));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// the same for the release paths: never throws, puts the caller's device back
struct ScopedDevice {
  int prev = -1;
  explicit ScopedDevice(int device) {
    if (device < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev == device) prev = -1;
    else cudaSetDevice(device);
    cudaGetLastError();
  }
  ~ScopedDevice() {
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
  }
};

// owning device pointer for the construction of handles (released on commit)
struct DevPtr {
  void* p = nullptr;
  ~DevPtr() { if (p) { cudaFree(p); cudaGetLastError(); } }
  void* release() { void* r = p; p = nullptr; return r; }
};

// stream-ordered scratch memory from the parameter set's own pool
struct Workspace {
  cudaStream_t st;
  cudaMemPool_t pool;
  std::vector<void*> ptrs;
  Workspace(const fhe_b200_params* par, cudaStream_t s) : st(s), pool(par->pool) {}
  u64* words(size_t n) {
    void* p = nullptr;
    FHE_CUDA(cudaMallocFromPoolAsync(&p, n * sizeof(u64), pool, st));
    ptrs.push_back(p);
    return (u64*)p;
  }
  ~Workspace() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
  }
};

u32 chunk_size() {
  static u32 c = [] {
    const char* e = getenv("FHE_B200_CHUNK");
    int v = e ? atoi(e) : 256;   // ~28 GB of scratch per in-flight chunk at set C (108 MB per ciphertext); products/s at
                                 // chunk 64 / 128 / 256 / 512: 4412 / 4458 / 4480 / 4490
    return (u32)(v < 1 ? 1 : v);
  }();
  return c;
}

// A batched call works through its ciphertexts chunk by chunk.  On ONE stream every kernel boundary costs ~16 us (tail
// of one persistent grid, ring fill and table staging of the next: profiles/microbench_r2.txt), 18 boundaries per chunk.
// When a call has more than one chunk the runner therefore deals the chunks over side streams of the parameter set
// (two by default, FHE_B200_STREAMS=1..4): the kernels of one chunk fill the SMs that the kernels of another leave at
// their boundaries, and kernels bound by different resources (integer pipe, HBM) share an SM.  The caller's stream is the
// only one it ever sees: the side streams start behind an event recorded on it and it waits for all of them before the
// call returns.  Each side stream works on chunks of chunk_size() / streams ciphertexts, so the scratch in flight is what
// one full chunk takes.  FHE_B200_STREAMS=1 keeps everything on the caller's stream.
struct ChunkRunner {
  const fhe_b200_params* par;
  cudaStream_t user;
  u32 count, chunk, ns;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  static u32 streams() {
    static const u32 n = [] {
      const char* e = getenv("FHE_B200_STREAMS");
      const int v = e ? atoi(e) : 2;
      return (u32)(v < 1 ? 1 : v > 4 ? 4 : v);
    }();
    return n;
  }
  ChunkRunner(const fhe_b200_params* p, u32 n, cudaStream_t st) : par(p), user(st), count(n), chunk(chunk_size()), ns(1) {
    if (streams() < 2 || count <= chunk || chunk < streams()) return;
    ns = streams();
    chunk = (chunk + ns - 1) / ns;
    {
      std::lock_guard<std::mutex> g(par->mu);
      for (u32 i = 0; i < ns; i++)
        if (!par->side[i]) FHE_CUDA(cudaStreamCreateWithFlags(&par->side[i], cudaStreamNonBlocking));
    }
    for (u32 i = 0; i <= ns; i++) FHE_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    FHE_CUDA(cudaEventRecord(ev[ns], user));
    for (u32 i = 0; i < ns; i++) FHE_CUDA(cudaStreamWaitEvent(par->side[i], ev[ns], 0));
  }
  // body(first ciphertext, number of ciphertexts, stream)
  template <class F>
  void run(F&& body) {
    u32 k = 0;
    for (u32 c0 = 0; c0 < count; c0 += chunk, k++) body(c0, std::min(chunk, count - c0), ns > 1 ? par->side[k % ns] : user);
    join();
  }
  void join() {
    if (ns < 2 || joined) return;
    joined = true;
    for (u32 i = 0; i < ns; i++)
      if (cudaEventRecord(ev[i], par->side[i]) == cudaSuccess) cudaStreamWaitEvent(user, ev[i], 0);
  }
  ~ChunkRunner() {
    join();                         // also when a chunk failed half way: the caller's stream still has to wait
    for (cudaEvent_t e : ev)
      if (e) cudaEventDestroy(e);   // released once the recorded work has completed
  }
  bool joined = false;
};

void check_same(const fhe_b200_batch* a, const fhe_b200_batch* b) {
  if (a->par != b->par) throw FheError(FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch: batches use different parameters");
  if (a->level != b->level) throw FheError(FHE_B200_INVALID_LEVEL, "InvalidLevel: operands are at different levels");
  if (a->mul_basis != b->mul_basis || a->limbs != b->limbs)
    throw FheError(FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
}
void need_repr(const fhe_b200_batch* b, int repr) {
  if (b->repr != repr) throw FheError(FHE_B200_INVALID_REPRESENTATION, "IncorrectRepresentation");
}
const RowIds& ids_of(const fhe_b200_batch* b) {
  const LevelData& lv = b->par->level(b->level);
  return b->mul_basis ? lv.mul_ids : lv.ctx_ids;
}

// KeySwitchingKey::key_switch core on a contiguous power-basis buffer c2 [cts][L][N]
// (key_switching_key.rs:241-270): out0/out1 (+ optional bases), rows (ct, j) at (ct*out_ct_rows + j).
void key_switch_core(const fhe_b200_params* par, const fhe_b200_ksk* k, const u64* c2, u32 cts, const u64* base0,
                     const u64* base1, u64* out0, u64* out1, u32 out_ct_rows, Workspace& ws, cudaStream_t st) {
  const LevelData& kl = par->level(k->ksk_level);
  const u32 L = k->n_dig, Lk = k->Lk;
  u64* inter = ws.words(((size_t)cts * L * Lk) << par->logn);
  bool adjacent = false;
  if (k->log_base) {
    // key_switch_decomposition (key_switching_key.rs:323-362): the digits of the single residue, each below
    // 2^log_base < q, transformed lazily like the RNS digits
    u64* dig = ws.words(((size_t)cts * L) << par->logn);
    launch_decompose(c2, dig, cts, L, k->log_base, par->logn, st);
    launch_ntt(dig, inter, cts * L, kl.ctx_ids, par->d_limbs, par->logn, false, 1, false, st, true);
  } else {
  // digit broadcast (rq/mod.rs:563-586), then NTT of every (digit, limb) row.  The reference lazily reduces the
  // digit modulo q_j before its lazy transform; the forward butterflies accept any input below 4*q_j, so the
  // reduction on load is only needed when a digit (< max q_i) can reach 4 * min q_j (mixed modulus sizes).
  u64 qmax = 0, qmin = ~0ull;
  for (u32 i = 0; i < L; i++) qmax = std::max(qmax, par->moduli[i]);
  for (u32 j = 0; j < Lk; j++) qmin = std::min(qmin, par->moduli[j]);
  const bool reduce = qmax > 4 * qmin - 1 || qmin < (1ull << 8);
  // forward_vt_lazy (rq/mod.rs:580): the digits stay in [0,4q_j); the lazy accumulator of the inner product takes
  // any 64-bit operand and reduces once
  // with the TMA kernels the transform deposits the digits of one (ciphertext, limb) in adjacent rows, which the
  // inner product streams fastest (bench_micro/stride_read.cu)
  adjacent = Lk > 1 && ntt_uses_tma(cts * L * Lk, kl.ctx_ids, par->logn, Lk, c2, inter);
  launch_ntt(c2, inter, cts * L * Lk, kl.ctx_ids, par->d_limbs, par->logn, false, Lk, reduce, st, true, adjacent, L);
  }
  launch_ksmac(inter, k->k0, k->k1, base0, base1, out0, out1, cts, L, Lk, out_ct_rows, kl.ctx_ids, par->d_limbs,
               par->logn, st, adjacent);
}

// key switch + the reference's post-processing (relinearization_key.rs:88-95, galois_key.rs:69-76):
// when the key lives at a lower level number than the ciphertext (more moduli), the (c0, c1) pair is
// taken to power basis, switched down to the ciphertext context and transformed back before it is added.
// c2: [cts][L][N] power basis at the ciphertext level; out: [cts][2][L][N]; base (nullable) is added:
// base_mode 0: none, 1: out += result (in place), 2: out = result + (base part 0 only; base is [cts][2][L][N])
void key_switch_apply(const fhe_b200_params* par, const fhe_b200_ksk* k, const u64* c2, u32 cts, u64* out,
                      int base_mode, u64* base, Workspace& ws, cudaStream_t st) {
  const LevelData& cl = par->level(k->ct_level);
  const u32 L = cl.L, Lk = k->Lk, logn = par->logn;
  const size_t row = (size_t)1 << logn;
  if (Lk == L) {
    const u64* b0 = base_mode == 1 ? out : base_mode == 2 ? base : nullptr;
    const u64* b1 = base_mode == 1 ? out + L * row : nullptr;
    key_switch_core(par, k, c2, cts, b0, b1, out, out + L * row, 2 * L, ws, st);
    return;
  }
  u64* cur = ws.words((size_t)cts * 2 * Lk * row);
  key_switch_core(par, k, c2, cts, nullptr, nullptr, cur, cur + Lk * row, 2 * Lk, ws, st);
  const LevelData& kl = par->level(k->ksk_level);
  launch_ntt(cur, cur, cts * 2 * Lk, kl.ctx_ids, par->d_limbs, logn, true, 1, false, st);
  for (u32 lv = k->ksk_level; lv < k->ct_level; lv++) {  // Poly::switch_down_to, rq/mod.rs:498-507
    const LevelData& from = par->level(lv);
    u64* nxt = ws.words((size_t)cts * 2 * (from.L - 1) * row);
    launch_switch_down(from.sd, cur, nxt, cts * 2, from.L, from.ctx_ids, par->d_limbs, logn, st);
    cur = nxt;
  }
  launch_ntt(cur, cur, cts * 2 * L, cl.ctx_ids, par->d_limbs, logn, false, 1, false, st);
  if (base_mode == 1) {
    launch_ew(EW_ADD, out, cur, (size_t)cts * 2 * L, cl.ctx_ids, par->d_limbs, logn, st);
  } else {
    FHE_CUDA(cudaMemcpyAsync(out, cur, (size_t)cts * 2 * L * row * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    if (base_mode == 2) {
      // only part 0 of `base` takes part: clear its part 1 (a scratch buffer of the caller) and add everything
      FHE_CUDA(cudaMemset2DAsync(base + L * row, 2 * L * row * sizeof(u64), 0, L * row * sizeof(u64), cts, st));
      launch_ew(EW_ADD, out, base, (size_t)cts * 2 * L, cl.ctx_ids, par->d_limbs, logn, st);
    }
  }
}

// extend -> tensor -> scale down of bfv/ops/mul.rs:192-206 for `cts` ciphertext pairs.
// a, b: [cts][2][L][N] NTT.  split == 0: out0 = [cts][3][L][N] power basis (all three parts);
// split == 1: out0 = [cts][2][L][N] (c0, c1), out1 = [cts][L][N] (c2), all power basis.
void mul_core(const fhe_b200_params* par, const LevelData& lv, const u64* a, const u64* b, u32 cts, u64* out0,
              u64* out1, int split, Workspace& ws, cudaStream_t st) {
  const u32 L = lv.L, E = lv.E, K = lv.K, logn = par->logn;
  const size_t row = (size_t)1 << logn;
  u64* A_l = ws.words((size_t)cts * 2 * L * row);
  u64* A_r = ws.words((size_t)cts * 2 * L * row);
  u64* X_l = ws.words((size_t)cts * 2 * E * row);
  u64* X_r = ws.words((size_t)cts * 2 * E * row);
  u64* T = ws.words((size_t)cts * 3 * K * row);
  RowIds ext_ids;
  std::memset(&ext_ids, 0, sizeof(ext_ids));
  ext_ids.limbs_per_poly = E;
  for (u32 j = 0; j < E; j++) ext_ids.ids[j] = lv.mul_ids.ids[L + j];
  // rq/scaler.rs:69-79: backward NTT of the source rows
  launch_ntt(a, A_l, cts * 2 * L, lv.ctx_ids, par->d_limbs, logn, true, 1, false, st);
  launch_ntt(b, A_r, cts * 2 * L, lv.ctx_ids, par->d_limbs, logn, true, 1, false, st);
  // rq/scaler.rs:85-94: exact base extension to the E new limbs (common prefix is kept as is, :61-65)
  launch_scale(lv.ext.dev, par->d_limbs, A_l, X_l, nullptr, cts * 2, E, L, E, 0, logn, st);
  launch_scale(lv.ext.dev, par->d_limbs, A_r, X_r, nullptr, cts * 2, E, L, E, 0, logn, st);
  // rq/scaler.rs:97-115: forward NTT of the new rows
  launch_ntt(X_l, X_l, cts * 2 * E, ext_ids, par->d_limbs, logn, false, 1, false, st);
  launch_ntt(X_r, X_r, cts * 2 * E, ext_ids, par->d_limbs, logn, false, 1, false, st);
  // mul.rs:198-201 tensor product, then mul.rs:204-206 scale down by t/Q (backward NTT of the 3K rows, exact scaling
  // K -> L); product and first inverse pass run as one kernel where the TMA kernels serve the shape
  if (!launch_tensor_inverse_ntt(a, b, X_l, X_r, T, cts, L, K, lv.mul_ids, par->d_limbs, logn, st)) {
    launch_tensor(a, b, X_l, X_r, T, cts, L, L, L, K, lv.mul_ids, par->d_limbs, logn, st);
    launch_ntt(T, T, cts * 3 * K, lv.mul_ids, par->d_limbs, logn, true, 1, false, st);
  }
  launch_scale(lv.down.dev, par->d_limbs, T, out0, out1, cts * 3, L, 0, L, split, logn, st);
}

// &ct * &ct for any part counts (ops/mod.rs:259-358): out [cts][na+nb-1][L][N] power basis
void mul_core_parts(const fhe_b200_params* par, const LevelData& lv, const u64* a, u32 na, const u64* b, u32 nb, u32 cts,
                    u64* out, Workspace& ws, cudaStream_t st) {
  const u32 L = lv.L, E = lv.E, K = lv.K, logn = par->logn, nc = na + nb - 1;
  const size_t row = (size_t)1 << logn;
  RowIds ext_ids;
  std::memset(&ext_ids, 0, sizeof(ext_ids));
  ext_ids.limbs_per_poly = E;
  for (u32 j = 0; j < E; j++) ext_ids.ids[j] = lv.mul_ids.ids[L + j];
  const u64* src[2] = {a, b};
  const u32 np[2] = {na, nb};
  u64* X[2];
  for (int s = 0; s < 2; s++) {
    u64* pb = ws.words((size_t)cts * np[s] * L * row);
    X[s] = ws.words((size_t)cts * np[s] * E * row);
    launch_ntt(src[s], pb, cts * np[s] * L, lv.ctx_ids, par->d_limbs, logn, true, 1, false, st);
    launch_scale(lv.ext.dev, par->d_limbs, pb, X[s], nullptr, cts * np[s], E, L, E, 0, logn, st);
    launch_ntt(X[s], X[s], cts * np[s] * E, ext_ids, par->d_limbs, logn, false, 1, false, st);
  }
  u64* T = ws.words((size_t)cts * nc * K * row);
  launch_tensor_nm(a, b, X[0], X[1], T, cts, L, E, na, nb, lv.mul_ids, par->d_limbs, logn, st);
  launch_ntt(T, T, cts * nc * K, lv.mul_ids, par->d_limbs, logn, true, 1, false, st);
  launch_scale(lv.down.dev, par->d_limbs, T, out, nullptr, cts * nc, L, 0, L, 0, logn, st);
}

// The same pipeline for a custom strategy (mul.rs:192-206 with the Scalers of Multiplicator::new): every extender
// keeps its common prefix only when its factor is one (rq/scaler.rs:35-43), so a side with a non-unit factor gets all
// K limbs from the exact scaler.  out: [cts][3][L][N] power basis.
void mul_core_general(const fhe_b200_multiplicator* m, const u64* a, const u64* b, u32 cts, u64* out, Workspace& ws,
                      cudaStream_t st) {
  const fhe_b200_params* par = m->par;
  const LevelData& lv = par->level(m->level);
  const u32 L = m->L, K = m->K, logn = par->logn;
  const size_t row = (size_t)1 << logn;
  const u64* src[2] = {a, b};
  const ScalerData* ext[2] = {&m->ext_l, &m->ext_r};
  const u32 nc[2] = {m->nc_l, m->nc_r};
  u64* X[2] = {nullptr, nullptr};
  for (int s = 0; s < 2; s++) {
    const u32 E = K - nc[s];
    if (!E) continue;
    u64* pb = ws.words((size_t)cts * 2 * L * row);
    X[s] = ws.words((size_t)cts * 2 * E * row);
    launch_ntt(src[s], pb, cts * 2 * L, lv.ctx_ids, par->d_limbs, logn, true, 1, false, st);
    launch_scale(ext[s]->dev, m->d_limbs, pb, X[s], nullptr, cts * 2, E, nc[s], E, 0, logn, st);
    RowIds ids;
    std::memset(&ids, 0, sizeof(ids));
    ids.limbs_per_poly = E;
    for (u32 j = 0; j < E; j++) ids.ids[j] = m->mul_ids.ids[nc[s] + j];
    launch_ntt(X[s], X[s], cts * 2 * E, ids, m->d_limbs, logn, false, 1, false, st);
  }
  u64* T = ws.words((size_t)cts * 3 * K * row);
  launch_tensor(a, b, X[0], X[1], T, cts, L, nc[0], nc[1], K, m->mul_ids, m->d_limbs, logn, st);
  launch_ntt(T, T, cts * 3 * K, m->mul_ids, m->d_limbs, logn, true, 1, false, st);
  if (m->nc_d)  // common prefix of a factor-one down scaler: kept as is (power basis here, transformed by the caller)
    FHE_CUDA(cudaMemcpy2DAsync(out, L * row * 8, T, K * row * 8, m->nc_d * row * 8, (size_t)cts * 3,
                               cudaMemcpyDeviceToDevice, st));
  launch_scale(m->down.dev, m->d_limbs, T, out + m->nc_d * row, nullptr, cts * 3, L, m->nc_d, L - m->nc_d, 0, logn, st);
}

}  // namespace

extern "C" {

#define API_BEGIN try {
#define API_END                                                                   \
  }                                                                               \
  catch (const FheError& e) { g_last_error = e.what(); return e.code; }           \
  catch (const CudaFail& f) {                                                     \
    g_last_error = std::string(f.what) + ": " + cudaGetErrorString(f.err);        \
    cudaGetLastError();                                                           \
    return f.err == cudaErrorMemoryAllocation ? FHE_B200_OUT_OF_MEMORY : FHE_B200_CUDA_ERROR; \
  }                                                                               \
  catch (const std::bad_alloc&) { g_last_error = "host out of memory"; return FHE_B200_OUT_OF_MEMORY; } \
  catch (const std::exception& e) { g_last_error = e.what(); return FHE_B200_INVALID_ARGUMENT; } \
  return FHE_B200_OK;
#define REQUIRE(c, code, msg) \
  do { if (!(c)) throw FheError(code, msg); } while (0)

const char* fhe_b200_version(void) { return "fhe_b200 0.1 (sm_100a)"; }
const char* fhe_b200_last_error(void) { return g_last_error.c_str(); }
uint64_t fhe_b200_launch_count(void) { return g_launches.load(); }

static int params_build(int device, uint32_t degree, const std::vector<u64>& moduli, const uint8_t* pt,
                        uint32_t pt_len, const uint64_t* psi, fhe_b200_params** out) {
  API_BEGIN
  REQUIRE(out && pt && pt_len, FHE_B200_INVALID_ARGUMENT, "null argument");
  // BfvParametersBuilder::validate_configuration (parameters.rs:440-468)
  REQUIRE(degree >= 8 && degree <= 65536 && (degree & (degree - 1)) == 0, FHE_B200_INVALID_DEGREE,
          "InvalidPolynomialDegree: " + std::to_string(degree));
  REQUIRE(!moduli.empty() && moduli.size() < 32, FHE_B200_INVALID_ARGUMENT, "MissingCiphertextModulusSpecification");
  // released through params_release on every exit path (frees the device tables and the pool of a half-built set)
  std::unique_ptr<fhe_b200_params, void (*)(const fhe_b200_params*)> p(new fhe_b200_params(), params_release);
  struct Restore {   // the caller's current device is left as it was
    int prev = -1;
    ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  if (device >= 0 && cudaGetDevice(&restore.prev) != cudaSuccess) { cudaGetLastError(); restore.prev = -1; }
  p->device = device;
  p->N = degree;
  p->logn = (u32)__builtin_ctz(degree);
  p->Lmax = (u32)moduli.size();
  p->moduli = moduli;
  p->t = BigUint::from_le_bytes(pt, pt_len);
  REQUIRE(!p->t.is_zero(), FHE_B200_INVALID_ARGUMENT, "plaintext modulus is zero");
  // validate_moduli (parameters.rs:471-552)
  BigUint Q(1);
  for (size_t i = 0; i < moduli.size(); i++) {
    ModulusH m(moduli[i]);
    for (size_t j = 0; j < i; j++) REQUIRE(moduli[j] != moduli[i], FHE_B200_INVALID_MODULUS, "DuplicateModuli");
    REQUIRE(moduli[i] % (2 * (u64)degree) == 1 && is_prime_u64(moduli[i]), FHE_B200_NTT_UNAVAILABLE,
            "CiphertextModulusNotNttFriendly: " + std::to_string(moduli[i]));
    u64 tm = p->t.mod_u64(moduli[i]), dummy;
    REQUIRE(tm != 0 && invmod_h(tm, moduli[i], &dummy), FHE_B200_INVALID_MODULUS, "PlaintextModulusNotCoprime");
    p->moduli_sizes.push_back(64 - (u32)clz64(moduli[i]));
    Q = Q * BigUint(moduli[i]);
  }
  REQUIRE(p->t < Q, FHE_B200_INVALID_ARGUMENT, "PlaintextModulusExceedsCiphertextModulus");
  // extended basis (parameters.rs:660-676)
  u64 ub = 1ull << 62;
  while (p->ext.size() != moduli.size() + 1) {
    REQUIRE(generate_prime(62, 2 * (u64)degree, ub, &ub), FHE_B200_INVALID_MODULUS, "NotEnoughPrimes");
    bool dup = false;
    for (u64 q : p->ext) dup |= q == ub;
    for (u64 q : moduli) dup |= q == ub;
    if (!dup) p->ext.push_back(ub);
  }
  p->primes = moduli;
  p->primes.insert(p->primes.end(), p->ext.begin(), p->ext.end());
  if (device >= 0) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) {
      cudaGetLastError();
      throw FheError(FHE_B200_NO_DEVICE, "CUDA device " + std::to_string(device) + " not available");
    }
    FHE_CUDA(cudaSetDevice(device));
    cudaMemPoolProps props;
    std::memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    FHE_CUDA(cudaMemPoolCreate(&p->pool, &props));
    unsigned long long thr = ~0ull;   // keep freed scratch for the next chunk instead of returning it to the OS
    FHE_CUDA(cudaMemPoolSetAttribute(p->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // scratch freed on one stream must not be handed to another stream through an inserted dependency: that
    // serialises callers that pipeline chunks over several streams (bench.py e2e); let each stream keep its own
    int off = 0;
    FHE_CUDA(cudaMemPoolSetAttribute(p->pool, cudaMemPoolReuseAllowInternalDependencies, &off));
  }
  for (size_t i = 0; i < p->primes.size(); i++) {
    u64 q = p->primes[i];
    u64 r = psi ? psi[i] : default_psi(q, degree);
    p->psi.push_back(r);
    p->tables.push_back(make_ntt_tables(q, degree, r));
    p->h_limbs.push_back(make_limb_dev(q, p->tables.back(), [&](const std::vector<ulonglong2>& v) { return p->to_dev(v); }));
  }
  p->d_limbs = p->to_dev(p->h_limbs);
  *out = p.release();
  API_END
}

int fhe_b200_params_create(int device, uint32_t degree, const uint64_t* moduli, uint32_t n_moduli,
                           const uint8_t* plaintext_le, uint32_t plaintext_len, const uint64_t* psi,
                           fhe_b200_params** out) {
  if (!moduli || !n_moduli) { g_last_error = "null moduli"; return FHE_B200_INVALID_ARGUMENT; }
  std::vector<u64> m(moduli, moduli + n_moduli);
  return params_build(device, degree, m, plaintext_le, plaintext_len, psi, out);
}

int fhe_b200_params_create_from_sizes(int device, uint32_t degree, const uint32_t* sizes, uint32_t n_moduli,
                                      const uint8_t* plaintext_le, uint32_t plaintext_len, fhe_b200_params** out) {
  if (!sizes || !n_moduli) { g_last_error = "null sizes"; return FHE_B200_INVALID_ARGUMENT; }
  if (degree < 8 || (degree & (degree - 1))) { g_last_error = "InvalidPolynomialDegree"; return FHE_B200_INVALID_DEGREE; }
  // BfvParametersBuilder::generate_moduli (parameters.rs:391-431)
  std::vector<u64> m;
  for (uint32_t i = 0; i < n_moduli; i++) {
    if (sizes[i] > 62 || sizes[i] < 10) { g_last_error = "InvalidModulusSize"; return FHE_B200_INVALID_MODULUS; }
    u64 ub = 1ull << sizes[i];
    for (;;) {
      u64 q;
      if (!generate_prime((int)sizes[i], 2 * (u64)degree, ub, &q)) { g_last_error = "NotEnoughPrimes"; return FHE_B200_INVALID_MODULUS; }
      bool dup = false;
      for (u64 x : m) dup |= x == q;
      if (!dup) { m.push_back(q); break; }
      ub = q;
    }
  }
  return params_build(device, degree, m, plaintext_le, plaintext_len, nullptr, out);
}

int fhe_b200_params_destroy(fhe_b200_params* p) {
  params_release(p);
  return FHE_B200_OK;
}
uint32_t fhe_b200_params_degree(const fhe_b200_params* p) { return p ? p->N : 0; }
uint32_t fhe_b200_params_n_moduli(const fhe_b200_params* p) { return p ? p->Lmax : 0; }
int fhe_b200_params_moduli(const fhe_b200_params* p, uint64_t* out) {
  if (!p || !out) return FHE_B200_INVALID_ARGUMENT;
  for (u32 i = 0; i < p->Lmax; i++) out[i] = p->moduli[i];
  return FHE_B200_OK;
}
int fhe_b200_params_mul_basis(const fhe_b200_params* p, uint32_t level, uint64_t* out, uint32_t* n) {
  API_BEGIN
  REQUIRE(p && n, FHE_B200_INVALID_ARGUMENT, "null argument");
  const LevelData& lv = p->level(level);
  *n = lv.K;
  if (out) for (u32 i = 0; i < lv.K; i++) out[i] = lv.mul_moduli[i];
  API_END
}
int fhe_b200_params_psi(const fhe_b200_params* p, uint64_t q, uint64_t* psi) {
  if (!p || !psi) return FHE_B200_INVALID_ARGUMENT;
  int i = p->prime_index(q);
  if (i < 0) { g_last_error = "prime not in parameter set"; return FHE_B200_INVALID_MODULUS; }
  *psi = p->psi[i];
  return FHE_B200_OK;
}

// ------------------------------------------------------------------------------ batches
static int batch_alloc(const fhe_b200_params* p, uint32_t count, uint32_t parts, uint32_t level, int repr,
                       bool mul_basis, fhe_b200_batch** out) {
  API_BEGIN
  REQUIRE(p && out, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(count > 0 && parts > 0, FHE_B200_INVALID_ARGUMENT, "empty batch");
  REQUIRE(repr == FHE_B200_POWER_BASIS || repr == FHE_B200_NTT, FHE_B200_INVALID_REPRESENTATION, "bad representation");
  DeviceGuard g(p);
  const LevelData& lv = p->level(level);
  std::unique_ptr<fhe_b200_batch> b(new fhe_b200_batch());
  b->par = p; b->count = count; b->parts = parts; b->level = level; b->repr = repr;
  b->mul_basis = mul_basis;
  b->limbs = mul_basis ? lv.K : lv.L;
  b->d = nullptr;
  FHE_CUDA(cudaMalloc(&b->d, b->words_per_ct() * count * sizeof(u64)));
  params_retain(p);
  *out = b.release();
  API_END
}
int fhe_b200_batch_alloc(const fhe_b200_params* p, uint32_t count, uint32_t parts, uint32_t level, int repr,
                         fhe_b200_batch** out) {
  return batch_alloc(p, count, parts, level, repr, false, out);
}
int fhe_b200_batch_alloc_mul_basis(const fhe_b200_params* p, uint32_t count, uint32_t parts, uint32_t level,
                                   int repr, fhe_b200_batch** out) {
  return batch_alloc(p, count, parts, level, repr, true, out);
}
int fhe_b200_batch_free(fhe_b200_batch* b) {
  if (!b) return FHE_B200_OK;
  {
    ScopedDevice g(b->par->device);
    cudaFree(b->d);
    cudaGetLastError();
  }
  params_release(b->par);
  delete b;
  return FHE_B200_OK;
}
int fhe_b200_batch_info(const fhe_b200_batch* b, uint32_t* count, uint32_t* parts, uint32_t* level, uint32_t* limbs,
                        int* repr) {
  if (!b) return FHE_B200_INVALID_ARGUMENT;
  if (count) *count = b->count;
  if (parts) *parts = b->parts;
  if (level) *level = b->level;
  if (limbs) *limbs = b->limbs;
  if (repr) *repr = b->repr;
  return FHE_B200_OK;
}
int fhe_b200_batch_upload(fhe_b200_batch* b, uint32_t first, uint32_t n, const uint64_t* host, void* stream) {
  API_BEGIN
  REQUIRE(b && host, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE((uint64_t)first + n <= b->count, FHE_B200_INVALID_ARGUMENT, "range exceeds batch");
  DeviceGuard g(b->par);
  size_t w = b->words_per_ct();
  FHE_CUDA(cudaMemcpyAsync(b->d + w * first, host, w * n * sizeof(u64), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  API_END
}
int fhe_b200_batch_download(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint64_t* host, void* stream) {
  API_BEGIN
  REQUIRE(b && host, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE((uint64_t)first + n <= b->count, FHE_B200_INVALID_ARGUMENT, "range exceeds batch");
  DeviceGuard g(b->par);
  size_t w = b->words_per_ct();
  FHE_CUDA(cudaMemcpyAsync(host, b->d + w * first, w * n * sizeof(u64), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FHE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  API_END
}
int fhe_b200_batch_download_async(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint64_t* host, void* stream) {
  API_BEGIN
  REQUIRE(b && host, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE((uint64_t)first + n <= b->count, FHE_B200_INVALID_ARGUMENT, "range exceeds batch");
  DeviceGuard g(b->par);
  size_t w = b->words_per_ct();
  FHE_CUDA(cudaMemcpyAsync(host, b->d + w * first, w * n * sizeof(u64), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  API_END
}
int fhe_b200_batch_copy(fhe_b200_batch* dst, const fhe_b200_batch* src, void* stream) {
  API_BEGIN
  REQUIRE(dst && src, FHE_B200_INVALID_ARGUMENT, "null argument");
  check_same(dst, src);
  REQUIRE(dst->parts == src->parts && dst->count == src->count, FHE_B200_BAD_POLY_COUNT, "shapes differ");
  DeviceGuard g(src->par);
  FHE_CUDA(cudaMemcpyAsync(dst->d, src->d, src->words_per_ct() * src->count * sizeof(u64), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  dst->repr = src->repr;
  API_END
}
int fhe_b200_host_alloc(size_t bytes, int write_combined, void** out) {
  API_BEGIN
  REQUIRE(out && bytes, FHE_B200_INVALID_ARGUMENT, "null argument");
  void* p = nullptr;
  FHE_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  *out = p;
  API_END
}
int fhe_b200_host_free(void* p) {
  API_BEGIN
  if (p) FHE_CUDA(cudaFreeHost(p));
  API_END
}
int fhe_b200_batch_device_ptr(const fhe_b200_batch* b, uint64_t** dptr, size_t* n_words) {
  if (!b || !dptr) return FHE_B200_INVALID_ARGUMENT;
  *dptr = (uint64_t*)b->d;
  if (n_words) *n_words = b->words_per_ct() * b->count;
  return FHE_B200_OK;
}

// ------------------------------------------------------------------------------ keys
int fhe_b200_ksk_upload(const fhe_b200_params* p, uint32_t ciphertext_level, uint32_t ksk_level, const uint64_t* c0,
                        const uint64_t* c1, uint32_t n_digits, fhe_b200_ksk** out) {
  API_BEGIN
  REQUIRE(p && c0 && c1 && out, FHE_B200_INVALID_ARGUMENT, "null argument");
  DeviceGuard g(p);
  const LevelData& cl = p->level(ciphertext_level);
  const LevelData& kl = p->level(ksk_level);
  REQUIRE(ksk_level <= ciphertext_level, FHE_B200_INVALID_LEVEL, "key level must not exceed the ciphertext level");
  u32 log_base = 0;
  if (kl.L == 1) {
    // KeySwitchingKey::new (key_switching_key.rs:92-97): base-2^(log_modulus/2) decomposition of the single residue
    REQUIRE(cl.L == 1, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch: a single-modulus key serves the last level only");
    const u64 q = p->moduli[0];
    const u32 log_modulus = 64 - (u32)clz64(q - 1);   // next_power_of_two().ilog2()
    log_base = log_modulus / 2;
    REQUIRE(log_base >= 1, FHE_B200_UNSUPPORTED, "modulus too small for the decomposition");
    REQUIRE(n_digits == (log_modulus + log_base - 1) / log_base, FHE_B200_CONTEXT_MISMATCH,
            "n_digits must be ceil(log_modulus / log_base) for a single-modulus key");
  } else {
    REQUIRE(n_digits == cl.L, FHE_B200_CONTEXT_MISMATCH, "n_digits must equal the ciphertext level's limb count");
  }
  std::unique_ptr<fhe_b200_ksk> k(new fhe_b200_ksk());
  k->par = p; k->ct_level = ciphertext_level; k->ksk_level = ksk_level; k->n_dig = n_digits; k->Lk = kl.L;
  k->log_base = log_base;
  size_t bytes = ((size_t)n_digits * kl.L << p->logn) * sizeof(u64);
  DevPtr g0, g1;   // freed again if anything below fails
  FHE_CUDA(cudaMalloc(&g0.p, bytes));
  FHE_CUDA(cudaMalloc(&g1.p, bytes));
  // host layout [digit][limb][N] -> device layout [limb][digit][N]: the inner product walks the digits of one limb
  const size_t rowb = sizeof(u64) << p->logn;
  for (u32 i = 0; i < n_digits; i++) {
    FHE_CUDA(cudaMemcpy2D((char*)g0.p + i * rowb, n_digits * rowb, (const char*)c0 + (size_t)i * kl.L * rowb, rowb, rowb,
                          kl.L, cudaMemcpyHostToDevice));
    FHE_CUDA(cudaMemcpy2D((char*)g1.p + i * rowb, n_digits * rowb, (const char*)c1 + (size_t)i * kl.L * rowb, rowb, rowb,
                          kl.L, cudaMemcpyHostToDevice));
  }
  k->k0 = (u64*)g0.release();
  k->k1 = (u64*)g1.release();
  params_retain(p);
  *out = k.release();
  API_END
}
int fhe_b200_ksk_free(fhe_b200_ksk* k) {
  if (!k) return FHE_B200_OK;
  {
    ScopedDevice g(k->par->device);
    cudaFree(k->k0);
    cudaFree(k->k1);
    cudaGetLastError();
  }
  params_release(k->par);
  delete k;
  return FHE_B200_OK;
}

// ------------------------------------------------------------------------------ primitives
static int ntt_batch(fhe_b200_batch* b, bool inverse, void* stream) {
  API_BEGIN
  REQUIRE(b, FHE_B200_INVALID_ARGUMENT, "null argument");
  need_repr(b, inverse ? FHE_B200_NTT : FHE_B200_POWER_BASIS);
  DeviceGuard g(b->par);
  launch_ntt(b->d, b->d, b->count * b->parts * b->limbs, ids_of(b), b->par->d_limbs, b->par->logn, inverse, 1, false,
             (cudaStream_t)stream);
  FHE_CUDA(cudaGetLastError());
  b->repr = inverse ? FHE_B200_POWER_BASIS : FHE_B200_NTT;
  API_END
}
int fhe_b200_ntt_forward(fhe_b200_batch* b, void* stream) { return ntt_batch(b, false, stream); }
int fhe_b200_ntt_backward(fhe_b200_batch* b, void* stream) { return ntt_batch(b, true, stream); }

static int ew(EwOp op, fhe_b200_batch* a, const fhe_b200_batch* b, void* stream) {
  API_BEGIN
  REQUIRE(a && (b || op == EW_NEG), FHE_B200_INVALID_ARGUMENT, "null argument");
  if (b) {
    check_same(a, b);
    REQUIRE(a->parts == b->parts && a->count == b->count, FHE_B200_BAD_POLY_COUNT, "operand shapes differ");
    REQUIRE(a->repr == b->repr, FHE_B200_INVALID_REPRESENTATION, "IncorrectRepresentation");
  }
  DeviceGuard g(a->par);
  launch_ew(op, a->d, b ? b->d : nullptr, (size_t)a->count * a->parts * a->limbs, ids_of(a), a->par->d_limbs,
            a->par->logn, (cudaStream_t)stream);
  FHE_CUDA(cudaGetLastError());
  API_END
}
int fhe_b200_add(fhe_b200_batch* a, const fhe_b200_batch* b, void* stream) { return ew(EW_ADD, a, b, stream); }
int fhe_b200_sub(fhe_b200_batch* a, const fhe_b200_batch* b, void* stream) { return ew(EW_SUB, a, b, stream); }
int fhe_b200_neg(fhe_b200_batch* a, void* stream) { return ew(EW_NEG, a, nullptr, stream); }

static int plain_op(fhe_b200_batch* a, const uint64_t* host_polys, uint32_t n_polys, u32 op, void* stream) {
  API_BEGIN
  REQUIRE(a && host_polys, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(n_polys == 1 || n_polys == a->count, FHE_B200_INVALID_ARGUMENT, "n_polys must be 1 or the batch size");
  REQUIRE(a->parts >= 1, FHE_B200_BAD_POLY_COUNT, "empty ciphertext");
  need_repr(a, FHE_B200_NTT);
  const fhe_b200_params* par = a->par;
  DeviceGuard g(par);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(par, st);
  const size_t words = ((size_t)n_polys * a->limbs) << par->logn;
  u64* pt = ws.words(words);
  FHE_CUDA(cudaMemcpyAsync(pt, host_polys, words * sizeof(u64), cudaMemcpyHostToDevice, st));
  launch_mul_plain(a->d, pt, a->count, a->parts, n_polys, ids_of(a), par->d_limbs, par->logn, st, op);
  FHE_CUDA(cudaGetLastError());
  // No synchronisation: a pageable host_polys has been staged by the runtime when cudaMemcpyAsync returns; a pinned
  // one must stay valid until the stream reaches this point -- the same contract as fhe_b200_batch_upload.
  API_END
}
int fhe_b200_mul_plain(fhe_b200_batch* a, const uint64_t* host_polys, uint32_t n_polys, void* stream) {
  return plain_op(a, host_polys, n_polys, 0, stream);
}
int fhe_b200_add_plain(fhe_b200_batch* a, const uint64_t* host_polys, uint32_t n_polys, int subtract, void* stream) {
  return plain_op(a, host_polys, n_polys, subtract ? 2 : 1, stream);
}

int fhe_b200_dot_product_scalar(const fhe_b200_batch* cts, const fhe_b200_batch* pts, uint32_t n_terms,
                                fhe_b200_batch* out, void* stream) {
  API_BEGIN
  REQUIRE(cts && pts && out, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(n_terms > 0 && cts->count > 0 && pts->count > 0, FHE_B200_INVALID_ARGUMENT, "DotProductError::EmptyInput");
  check_same(cts, pts);
  check_same(cts, out);
  REQUIRE(pts->parts == 1, FHE_B200_BAD_POLY_COUNT, "plaintext batch must hold one polynomial per entry");
  REQUIRE(out->parts == cts->parts, FHE_B200_BAD_POLY_COUNT, "DotProductError::CiphertextPolynomialCountMismatch");
  const size_t total = (size_t)out->count * n_terms;
  REQUIRE((cts->count == total || cts->count == n_terms) && (pts->count == total || pts->count == n_terms),
          FHE_B200_INVALID_ARGUMENT, "DotProductError::OperandCountMismatch");
  need_repr(cts, FHE_B200_NTT);
  need_repr(pts, FHE_B200_NTT);
  DeviceGuard g(cts->par);
  launch_dot(cts->d, pts->d, out->d, out->count, n_terms, cts->parts, cts->count, pts->count, ids_of(cts),
             cts->par->d_limbs, cts->par->logn, (cudaStream_t)stream);
  FHE_CUDA(cudaGetLastError());
  out->repr = FHE_B200_NTT;
  API_END
}

int fhe_b200_mul(const fhe_b200_batch* a, const fhe_b200_batch* b, fhe_b200_batch* out3, void* stream) {
  API_BEGIN
  REQUIRE(a && b && out3, FHE_B200_INVALID_ARGUMENT, "null argument");
  check_same(a, b);
  check_same(a, out3);
  REQUIRE(!a->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  REQUIRE(a->parts >= 1 && b->parts >= 1 && out3->parts == a->parts + b->parts - 1, FHE_B200_BAD_POLY_COUNT,
          "MultiplicationPolynomialCount: expected n x m -> n + m - 1 parts");
  REQUIRE(a->count == b->count && a->count == out3->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(a, FHE_B200_NTT);
  need_repr(b, FHE_B200_NTT);
  DeviceGuard g(a->par);
  cudaStream_t st_user = (cudaStream_t)stream;
  const fhe_b200_params* par = a->par;
  const LevelData& lv = par->level(a->level);
  const size_t row = (size_t)1 << par->logn;
  const u32 na = a->parts, nb = b->parts, nc = na + nb - 1;
  ChunkRunner chunks(par, a->count, st_user);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    u64* o = out3->d + (size_t)c0 * nc * lv.L * row;
    const u64* pa = a->d + (size_t)c0 * na * lv.L * row;
    const u64* pb = b->d + (size_t)c0 * nb * lv.L * row;
    if (na == 2 && nb == 2) mul_core(par, lv, pa, pb, n, o, nullptr, 0, ws, st);
    else mul_core_parts(par, lv, pa, na, pb, nb, n, o, ws, st);
    // rq/scaler.rs:97-115 forward NTT of the scaled result
    launch_ntt(o, o, n * nc * lv.L, lv.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
  });
  FHE_CUDA(cudaGetLastError());
  out3->repr = FHE_B200_NTT;
  API_END
}

static void check_ksk(const fhe_b200_ksk* k, const fhe_b200_params* par, u32 level) {
  REQUIRE(k->par == par, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch: key belongs to other parameters");
  REQUIRE(k->ct_level == level, FHE_B200_INVALID_LEVEL, "InvalidLevel: key is for another ciphertext level");
}

int fhe_b200_relinearize(const fhe_b200_batch* ct3, const fhe_b200_ksk* rk, fhe_b200_batch* out2, void* stream) {
  API_BEGIN
  REQUIRE(ct3 && rk && out2, FHE_B200_INVALID_ARGUMENT, "null argument");
  check_same(ct3, out2);
  REQUIRE(!ct3->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  REQUIRE(ct3->parts == 3 && out2->parts == 2, FHE_B200_BAD_POLY_COUNT, "InvalidPolynomialCount: expected 3 -> 2");
  REQUIRE(ct3->count == out2->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(ct3, FHE_B200_NTT);
  check_ksk(rk, ct3->par, ct3->level);
  DeviceGuard g(ct3->par);
  cudaStream_t st_user = (cudaStream_t)stream;
  const fhe_b200_params* par = ct3->par;
  const LevelData& lv = par->level(ct3->level);
  const size_t row = (size_t)1 << par->logn, L = lv.L;
  ChunkRunner chunks(par, ct3->count, st_user);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    const u64* src = ct3->d + (size_t)c0 * 3 * L * row;
    u64* dst = out2->d + (size_t)c0 * 2 * L * row;
    u64* c2 = ws.words((size_t)n * L * row);
    FHE_CUDA(cudaMemcpy2DAsync(dst, 2 * L * row * 8, src, 3 * L * row * 8, 2 * L * row * 8, n, cudaMemcpyDeviceToDevice, st));
    FHE_CUDA(cudaMemcpy2DAsync(c2, L * row * 8, src + 2 * L * row, 3 * L * row * 8, L * row * 8, n, cudaMemcpyDeviceToDevice, st));
    // relinearization_key.rs:85: c2 -> power basis
    launch_ntt(c2, c2, n * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, true, 1, false, st);
    key_switch_apply(par, rk, c2, n, dst, 1, nullptr, ws, st);
  });
  FHE_CUDA(cudaGetLastError());
  out2->repr = FHE_B200_NTT;
  API_END
}

int fhe_b200_mul_relin(const fhe_b200_batch* a, const fhe_b200_batch* b, const fhe_b200_ksk* rk, int mod_switch,
                       fhe_b200_batch* out2, void* stream) {
  API_BEGIN
  REQUIRE(a && b && rk && out2, FHE_B200_INVALID_ARGUMENT, "null argument");
  check_same(a, b);
  REQUIRE(a->par == out2->par, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch");
  REQUIRE(!a->mul_basis && !out2->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  REQUIRE(a->parts == 2 && b->parts == 2 && out2->parts == 2, FHE_B200_BAD_POLY_COUNT,
          "MultiplicationPolynomialCount: expected 2 x 2 -> 2");
  REQUIRE(a->count == b->count && a->count == out2->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(a, FHE_B200_NTT);
  need_repr(b, FHE_B200_NTT);
  check_ksk(rk, a->par, a->level);
  const fhe_b200_params* par = a->par;
  const LevelData& lv = par->level(a->level);
  if (mod_switch) {
    REQUIRE(lv.L >= 2, FHE_B200_NO_MORE_CONTEXT, "NoMoreContext");  // mul.rs:155-162
    REQUIRE(out2->level == a->level + 1, FHE_B200_INVALID_LEVEL, "output batch must be one level down");
  } else {
    REQUIRE(out2->level == a->level, FHE_B200_INVALID_LEVEL, "output batch must be at the operand level");
  }
  DeviceGuard g(par);
  const size_t row = (size_t)1 << par->logn, L = lv.L;
  ChunkRunner chunks(par, a->count, (cudaStream_t)stream);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    u64* o = mod_switch ? ws.words((size_t)n * 2 * L * row) : out2->d + (size_t)c0 * 2 * L * row;
    u64* c2 = ws.words((size_t)n * L * row);
    mul_core(par, lv, a->d + (size_t)c0 * 2 * L * row, b->d + (size_t)c0 * 2 * L * row, n, o, c2, 1, ws, st);
    // c0, c1 back to NTT.  c2 stays in power basis: mul.rs:206 + :212 forward- then inverse-transform it,
    // and backward(forward(x)) == x for reduced x (ntt/mod.rs:73-74), so skipping both is bit-exact.
    launch_ntt(o, o, n * 2 * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
    key_switch_apply(par, rk, c2, n, o, 1, nullptr, ws, st);
    if (mod_switch) {  // Ciphertext::switch_down, ciphertext.rs:148-161
      launch_ntt(o, o, n * 2 * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, true, 1, false, st);
      u64* dst = out2->d + (size_t)c0 * 2 * (L - 1) * row;
      launch_switch_down(lv.sd, o, dst, n * 2, (u32)L, lv.ctx_ids, par->d_limbs, par->logn, st);
      const LevelData& nl = par->level(a->level + 1);
      launch_ntt(dst, dst, n * 2 * (u32)(L - 1), nl.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
    }
  });
  FHE_CUDA(cudaGetLastError());
  out2->repr = FHE_B200_NTT;
  API_END
}

// ---- custom multiplication strategies
int fhe_b200_multiplicator_create(const fhe_b200_params* p, uint32_t level, const uint8_t* lhs_num, uint32_t lhs_num_len,
                                  const uint8_t* lhs_den, uint32_t lhs_den_len, const uint8_t* rhs_num,
                                  uint32_t rhs_num_len, const uint8_t* rhs_den, uint32_t rhs_den_len,
                                  const uint64_t* extended_basis, uint32_t n_basis, const uint64_t* psi,
                                  const uint8_t* post_num, uint32_t post_num_len, const uint8_t* post_den,
                                  uint32_t post_den_len, fhe_b200_multiplicator** out) {
  API_BEGIN
  REQUIRE(p && out && extended_basis && n_basis && lhs_num && lhs_den && rhs_num && rhs_den && post_num && post_den,
          FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(n_basis <= (u32)kMaxPos, FHE_B200_UNSUPPORTED, "too many limbs");
  const LevelData& lv = p->level(level);   // context_at_level: InvalidLevel when out of range
  DeviceGuard g(p);
  std::unique_ptr<fhe_b200_multiplicator> m(new fhe_b200_multiplicator());
  struct Cleanup {   // frees the device tables if construction throws
    fhe_b200_multiplicator* m;
    ~Cleanup() { if (m) { for (void* d : m->d_allocs) cudaFree(d); cudaGetLastError(); } }
  } cleanup{m.get()};
  m->par = p;
  m->level = level;
  m->L = lv.L;
  m->K = n_basis;
  m->mul_moduli.assign(extended_basis, extended_basis + n_basis);
  // Context::new(extended_basis) (rq/context.rs:42-92): distinct NTT-friendly primes
  for (u32 i = 0; i < n_basis; i++) {
    const u64 q = extended_basis[i];
    for (u32 j = 0; j < i; j++) REQUIRE(extended_basis[j] != q, FHE_B200_INVALID_MODULUS, "DuplicateModuli");
    REQUIRE(q >= 2 && (q >> 62) == 0, FHE_B200_INVALID_MODULUS, "InvalidModulus: " + std::to_string(q));
    REQUIRE(q % (2 * (u64)p->N) == 1 && is_prime_u64(q), FHE_B200_NTT_UNAVAILABLE,
            "modulus does not support the NTT: " + std::to_string(q));
  }
  m->plan_primes = p->primes;
  m->h_limbs = p->h_limbs;
  std::memset(&m->mul_ids, 0, sizeof(RowIds));
  m->mul_ids.limbs_per_poly = n_basis;
  for (u32 i = 0; i < n_basis; i++) {
    const u64 q = extended_basis[i];
    int idx = m->prime_index(q);
    if (idx >= 0 && psi && psi[i] != p->psi[(size_t)idx] && (size_t)idx < p->primes.size())
      throw FheError(FHE_B200_INVALID_ARGUMENT, "psi differs from the parameter set's root for " + std::to_string(q));
    if (idx < 0) {
      const u64 r = psi ? psi[i] : default_psi(q, p->N);
      NttTablesH t = make_ntt_tables(q, p->N, r);
      idx = (int)m->plan_primes.size();
      m->plan_primes.push_back(q);
      m->h_limbs.push_back(make_limb_dev(q, t, [&](const std::vector<ulonglong2>& v) { return m->to_dev(v); }));
    }
    m->mul_ids.ids[i] = (unsigned short)idx;
  }
  m->d_limbs = m->to_dev(m->h_limbs);
  std::vector<u64> base(p->moduli.begin(), p->moduli.begin() + lv.L);
  RnsContextH from(base), to(m->mul_moduli);
  auto factor = [](const uint8_t* b, uint32_t n) { return BigUint::from_le_bytes(b, n); };
  const BigUint ln = factor(lhs_num, lhs_num_len), ld = factor(lhs_den, lhs_den_len);
  const BigUint rn = factor(rhs_num, rhs_num_len), rd = factor(rhs_den, rhs_den_len);
  const BigUint pn = factor(post_num, post_num_len), pd = factor(post_den, post_den_len);
  REQUIRE(!ld.is_zero() && !rd.is_zero() && !pd.is_zero(), FHE_B200_INVALID_ARGUMENT, "zero denominator");
  m->ext_l.h = make_scaler_tables(from, to, ln, ld);
  m->ext_r.h = make_scaler_tables(from, to, rn, rd);
  m->down.h = make_scaler_tables(to, from, pn, pd);
  auto common = [&](const ScalerData& sd, const std::vector<u64>& x, const std::vector<u64>& y) {
    u32 n = 0;
    if (sd.h.is_one)
      while (n < x.size() && n < y.size() && x[n] == y[n]) n++;
    return n;
  };
  m->nc_l = common(m->ext_l, base, m->mul_moduli);
  m->nc_r = common(m->ext_r, base, m->mul_moduli);
  m->nc_d = common(m->down, m->mul_moduli, base);
  m->upload_scaler(m->ext_l, m->mul_moduli);
  m->upload_scaler(m->ext_r, m->mul_moduli);
  m->upload_scaler(m->down, base);
  params_retain(p);
  cleanup.m = nullptr;
  *out = m.release();
  API_END
}

int fhe_b200_multiplicator_free(fhe_b200_multiplicator* m) {
  if (!m) return FHE_B200_OK;
  if (m->par->device >= 0) {
    ScopedDevice g(m->par->device);
    for (void* d : m->d_allocs) cudaFree(d);
    cudaGetLastError();
  }
  params_release(m->par);
  delete m;
  return FHE_B200_OK;
}

int fhe_b200_multiplicator_multiply(const fhe_b200_multiplicator* m, const fhe_b200_batch* a, const fhe_b200_batch* b,
                                    const fhe_b200_ksk* rk, int mod_switch, fhe_b200_batch* out, void* stream) {
  API_BEGIN
  REQUIRE(m && a && b && out, FHE_B200_INVALID_ARGUMENT, "null argument");
  const fhe_b200_params* par = m->par;
  REQUIRE(a->par == par && b->par == par && out->par == par, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch");
  REQUIRE(a->level == m->level && b->level == m->level, FHE_B200_INVALID_LEVEL, "InvalidLevel");  // mul.rs:168-181
  REQUIRE(!a->mul_basis && !b->mul_basis && !out->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  const u32 out_parts = rk ? 2 : 3;
  REQUIRE(a->parts == 2 && b->parts == 2 && out->parts == out_parts, FHE_B200_BAD_POLY_COUNT,
          "MultiplicationPolynomialCount");
  REQUIRE(a->count == b->count && a->count == out->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(a, FHE_B200_NTT);
  need_repr(b, FHE_B200_NTT);
  if (rk) {   // enable_relinearization (mul.rs:141-151): the key must live at the multiplicator's context
    REQUIRE(rk->par == par && rk->ct_level == m->level, FHE_B200_CONTEXT_MISMATCH,
            "ParameterMismatch: relinearization key and multiplicator contexts differ");
  }
  const LevelData& lv = par->level(m->level);
  if (mod_switch) {
    REQUIRE(lv.L >= 2, FHE_B200_NO_MORE_CONTEXT, "NoMoreContext");  // mul.rs:155-162
    REQUIRE(out->level == m->level + 1, FHE_B200_INVALID_LEVEL, "output batch must be one level down");
  } else {
    REQUIRE(out->level == m->level, FHE_B200_INVALID_LEVEL, "output batch must be at the operand level");
  }
  DeviceGuard g(par);
  const size_t row = (size_t)1 << par->logn, L = lv.L;
  ChunkRunner chunks(par, a->count, (cudaStream_t)stream);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    const bool direct = !rk && !mod_switch;
    u64* W = direct ? out->d + (size_t)c0 * 3 * L * row : ws.words((size_t)n * 3 * L * row);
    mul_core_general(m, a->d + (size_t)c0 * 2 * L * row, b->d + (size_t)c0 * 2 * L * row, n, W, ws, st);
    u64* o = W;   // [n][out_parts][L][N]
    if (rk) {
      o = mod_switch ? ws.words((size_t)n * 2 * L * row) : out->d + (size_t)c0 * 2 * L * row;
      u64* c2 = ws.words((size_t)n * L * row);
      FHE_CUDA(cudaMemcpy2DAsync(o, 2 * L * row * 8, W, 3 * L * row * 8, 2 * L * row * 8, n, cudaMemcpyDeviceToDevice, st));
      FHE_CUDA(cudaMemcpy2DAsync(c2, L * row * 8, W + 2 * L * row, 3 * L * row * 8, L * row * 8, n,
                                 cudaMemcpyDeviceToDevice, st));
      launch_ntt(o, o, n * 2 * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
      key_switch_apply(par, rk, c2, n, o, 1, nullptr, ws, st);   // mul.rs:210-228
    } else {
      launch_ntt(o, o, n * 3 * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
    }
    if (mod_switch) {  // Ciphertext::switch_down, ciphertext.rs:148-161
      launch_ntt(o, o, n * out_parts * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, true, 1, false, st);
      u64* dst = out->d + (size_t)c0 * out_parts * (L - 1) * row;
      launch_switch_down(lv.sd, o, dst, n * out_parts, (u32)L, lv.ctx_ids, par->d_limbs, par->logn, st);
      const LevelData& nl = par->level(m->level + 1);
      launch_ntt(dst, dst, n * out_parts * (u32)(L - 1), nl.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
    }
  });
  FHE_CUDA(cudaGetLastError());
  out->repr = FHE_B200_NTT;
  API_END
}

int fhe_b200_substitute(const fhe_b200_batch* in, uint32_t exponent, fhe_b200_batch* out, void* stream) {
  API_BEGIN
  REQUIRE(in && out && in != out, FHE_B200_INVALID_ARGUMENT, "null or aliased argument");
  check_same(in, out);
  REQUIRE(in->parts == out->parts && in->count == out->count, FHE_B200_BAD_POLY_COUNT, "shapes differ");
  const fhe_b200_params* par = in->par;
  exponent %= 2 * par->N;
  REQUIRE(exponent & 1, FHE_B200_INVALID_EXPONENT, "InvalidSubstitutionExponent");
  DeviceGuard g(par);
  const size_t rows = (size_t)in->count * in->parts * in->limbs;
  if (in->repr == FHE_B200_NTT)   // rq/mod.rs:360-389: a permutation of the bit-reversed evaluation points
    launch_gather(in->d, out->d, rows, par->perm(exponent), par->logn, (cudaStream_t)stream);
  else                            // rq/mod.rs:390-408: a signed permutation of the coefficients
    launch_substitute_power(in->d, out->d, rows, exponent, ids_of(in), par->d_limbs, par->logn, (cudaStream_t)stream);
  FHE_CUDA(cudaGetLastError());
  out->repr = in->repr;
  API_END
}

int fhe_b200_galois(const fhe_b200_batch* ct, uint32_t exponent, const fhe_b200_ksk* gk, fhe_b200_batch* out,
                    void* stream) {
  API_BEGIN
  REQUIRE(ct && gk && out && ct != out, FHE_B200_INVALID_ARGUMENT, "null or aliased argument");
  check_same(ct, out);
  REQUIRE(!ct->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  REQUIRE(ct->parts == 2 && out->parts == 2, FHE_B200_BAD_POLY_COUNT, "InvalidPolynomialCount: expected 2");
  REQUIRE(ct->count == out->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(ct, FHE_B200_NTT);
  check_ksk(gk, ct->par, ct->level);
  const fhe_b200_params* par = ct->par;
  exponent %= 2 * par->N;
  REQUIRE(exponent & 1, FHE_B200_INVALID_EXPONENT, "InvalidSubstitutionExponent");
  DeviceGuard g(par);
  const LevelData& lv = par->level(ct->level);
  const size_t row = (size_t)1 << par->logn, L = lv.L;
  const int* perm = par->perm(exponent);
  ChunkRunner chunks(par, ct->count, (cudaStream_t)stream);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    const u64* src = ct->d + (size_t)c0 * 2 * L * row;
    u64* dst = out->d + (size_t)c0 * 2 * L * row;
    u64* s = ws.words((size_t)n * 2 * L * row);
    u64* c2 = ws.words((size_t)n * L * row);
    // galois_key.rs:66: substitute both parts; part 1 becomes the key-switch input
    launch_gather(src, s, (size_t)n * 2 * L, perm, par->logn, st);
    FHE_CUDA(cudaMemcpy2DAsync(c2, L * row * 8, s + L * row, 2 * L * row * 8, L * row * 8, n, cudaMemcpyDeviceToDevice, st));
    launch_ntt(c2, c2, n * (u32)L, lv.ctx_ids, par->d_limbs, par->logn, true, 1, false, st);
    // galois_key.rs:67 + :78: out0 = key_switch0 + substitute(ct[0]); out1 = key_switch1
    key_switch_apply(par, gk, c2, n, dst, 2, s, ws, st);
  });
  FHE_CUDA(cudaGetLastError());
  out->repr = FHE_B200_NTT;
  API_END
}

int fhe_b200_key_switch(const fhe_b200_batch* pb, uint32_t part, const fhe_b200_ksk* k, fhe_b200_batch* out2,
                        void* stream) {
  API_BEGIN
  REQUIRE(pb && k && out2, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(pb->par == out2->par && pb->par == k->par, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch");
  REQUIRE(!pb->mul_basis && !out2->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  REQUIRE(part < pb->parts && out2->parts == 2, FHE_B200_BAD_POLY_COUNT, "bad part index / output parts");
  REQUIRE(pb->level == k->ct_level && out2->level == k->ksk_level, FHE_B200_INVALID_LEVEL, "InvalidLevel");
  REQUIRE(pb->count == out2->count, FHE_B200_INVALID_ARGUMENT, "batch sizes differ");
  need_repr(pb, FHE_B200_POWER_BASIS);
  const fhe_b200_params* par = pb->par;
  DeviceGuard g(par);
  cudaStream_t st_user = (cudaStream_t)stream;
  const size_t row = (size_t)1 << par->logn, L = pb->limbs, Lk = k->Lk;
  ChunkRunner chunks(par, pb->count, st_user);
  chunks.run([&](u32 c0, u32 n, cudaStream_t st) {
    Workspace ws(par, st);
    u64* c2 = ws.words((size_t)n * L * row);
    FHE_CUDA(cudaMemcpy2DAsync(c2, L * row * 8, pb->d + ((size_t)c0 * pb->parts + part) * L * row,
                               pb->parts * L * row * 8, L * row * 8, n, cudaMemcpyDeviceToDevice, st));
    u64* dst = out2->d + (size_t)c0 * 2 * Lk * row;
    key_switch_core(par, k, c2, n, nullptr, nullptr, dst, dst + Lk * row, 2 * (u32)Lk, ws, st);
  });
  FHE_CUDA(cudaGetLastError());
  out2->repr = FHE_B200_NTT;
  API_END
}

int fhe_b200_switch_down(fhe_b200_batch* b, void* stream) {
  API_BEGIN
  REQUIRE(b, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE(!b->mul_basis, FHE_B200_CONTEXT_MISMATCH, "PolynomialContextMismatch");
  need_repr(b, FHE_B200_NTT);
  const fhe_b200_params* par = b->par;
  const LevelData& lv = par->level(b->level);
  REQUIRE(lv.L >= 2, FHE_B200_NO_MORE_CONTEXT, "NoMoreContext");
  DeviceGuard g(par);
  cudaStream_t st = (cudaStream_t)stream;
  const LevelData& nl = par->level(b->level + 1);
  const u32 polys = b->count * b->parts;
  // Stream-ordered and in place: the L-1 surviving rows of every polynomial go through scratch memory and the forward
  // transform writes them back, compacted, at the start of the batch's own allocation (which keeps its size: the
  // pointer handed out by fhe_b200_batch_device_ptr stays valid, nothing is allocated, freed or synchronised here).
  Workspace ws(par, st);
  u64* tmp = ws.words((size_t)polys * nl.L << par->logn);
  launch_ntt(b->d, b->d, polys * lv.L, lv.ctx_ids, par->d_limbs, par->logn, true, 1, false, st);
  launch_switch_down(lv.sd, b->d, tmp, polys, lv.L, lv.ctx_ids, par->d_limbs, par->logn, st);
  launch_ntt(tmp, b->d, polys * nl.L, nl.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
  FHE_CUDA(cudaGetLastError());
  b->level += 1;
  b->limbs = nl.L;
  API_END
}

int fhe_b200_scale(const fhe_b200_batch* in, int which, fhe_b200_batch* out, void* stream) {
  API_BEGIN
  REQUIRE(in && out && in != out, FHE_B200_INVALID_ARGUMENT, "null or aliased argument");
  REQUIRE(in->par == out->par, FHE_B200_CONTEXT_MISMATCH, "ParameterMismatch");
  REQUIRE(in->level == out->level, FHE_B200_INVALID_LEVEL, "InvalidLevel");
  REQUIRE(in->count == out->count && in->parts == out->parts, FHE_B200_BAD_POLY_COUNT, "shapes differ");
  REQUIRE(which == 0 || which == 1, FHE_B200_INVALID_ARGUMENT, "which must be 0 or 1");
  REQUIRE(in->mul_basis == (which == 1) && out->mul_basis == (which == 0), FHE_B200_CONTEXT_MISMATCH,
          "PolynomialContextMismatch");
  need_repr(in, FHE_B200_NTT);
  const fhe_b200_params* par = in->par;
  DeviceGuard g(par);
  cudaStream_t st = (cudaStream_t)stream;
  const LevelData& lv = par->level(in->level);
  const size_t row = (size_t)1 << par->logn;
  const u32 polys = in->count * in->parts;
  Workspace ws(par, st);
  u64* pb = ws.words((size_t)polys * in->limbs * row);
  launch_ntt(in->d, pb, polys * in->limbs, ids_of(in), par->d_limbs, par->logn, true, 1, false, st);
  if (which == 0) {  // extender: common prefix copied, E new rows computed (rq/scaler.rs:61-65, :85-115)
    FHE_CUDA(cudaMemcpy2DAsync(out->d, lv.K * row * 8, in->d, lv.L * row * 8, lv.L * row * 8, polys,
                               cudaMemcpyDeviceToDevice, st));
    u64* x = ws.words((size_t)polys * lv.E * row);
    launch_scale(lv.ext.dev, par->d_limbs, pb, x, nullptr, polys, lv.E, lv.L, lv.E, 0, par->logn, st);
    RowIds ext_ids;
    std::memset(&ext_ids, 0, sizeof(ext_ids));
    ext_ids.limbs_per_poly = lv.E;
    for (u32 j = 0; j < lv.E; j++) ext_ids.ids[j] = lv.mul_ids.ids[lv.L + j];
    launch_ntt(x, x, polys * lv.E, ext_ids, par->d_limbs, par->logn, false, 1, false, st);
    FHE_CUDA(cudaMemcpy2DAsync(out->d + lv.L * row, lv.K * row * 8, x, lv.E * row * 8, lv.E * row * 8, polys,
                               cudaMemcpyDeviceToDevice, st));
  } else {
    launch_scale(lv.down.dev, par->d_limbs, pb, out->d, nullptr, polys, lv.L, 0, lv.L, 0, par->logn, st);
    launch_ntt(out->d, out->d, polys * lv.L, lv.ctx_ids, par->d_limbs, par->logn, false, 1, false, st);
  }
  FHE_CUDA(cudaGetLastError());
  out->repr = FHE_B200_NTT;
  API_END
}

// ------------------------------------------------------------------------------ wire format
static PackDev pack_desc(const fhe_b200_batch* b) {
  const fhe_b200_params* par = b->par;
  const LevelData& lv = par->level(b->level);
  PackDev P;
  std::memset(&P, 0, sizeof(P));
  P.limbs = b->limbs;
  u32 off = 0;
  for (u32 i = 0; i < b->limbs; i++) {
    const u64 q = b->mul_basis ? lv.mul_moduli[i] : par->moduli[i];
    const u32 nb = 64 - (u32)clz64(q - 1);          // Modulus::serialize_vec, zq/mod.rs:784
    P.nbits[i] = (unsigned char)nb;
    P.offs[i] = off;
    off += nb * (par->N / 8);                        // serialization_length, zq/mod.rs:773-777
  }
  P.poly_bytes = off;
  return P;
}
int fhe_b200_poly_packed_bytes(const fhe_b200_params* p, uint32_t level, size_t* nbytes) {
  API_BEGIN
  REQUIRE(p && nbytes, FHE_B200_INVALID_ARGUMENT, "null argument");
  const LevelData& lv = p->level(level);
  size_t n = 0;
  for (u32 i = 0; i < lv.L; i++) n += (size_t)(64 - clz64(p->moduli[i] - 1)) * (p->N / 8);
  *nbytes = n;
  API_END
}
int fhe_b200_batch_packed_bytes(const fhe_b200_batch* b, size_t* nbytes) {
  API_BEGIN
  REQUIRE(b && nbytes, FHE_B200_INVALID_ARGUMENT, "null argument");
  *nbytes = pack_desc(b).poly_bytes;
  API_END
}
int fhe_b200_batch_pack(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint8_t* host_out, void* stream) {
  API_BEGIN
  REQUIRE(b && host_out, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE((uint64_t)first + n <= b->count, FHE_B200_INVALID_ARGUMENT, "range exceeds batch");
  const fhe_b200_params* par = b->par;
  DeviceGuard g(par);
  cudaStream_t st = (cudaStream_t)stream;
  const PackDev P = pack_desc(b);
  const size_t rows = (size_t)n * b->parts * b->limbs, row = (size_t)1 << par->logn;
  Workspace ws(par, st);
  const u64* src = b->d + b->words_per_ct() * first;
  if (b->repr == FHE_B200_NTT) {   // rq/convert.rs:20-24: serialization is always in power basis
    u64* pb = ws.words(rows * row);
    launch_ntt(src, pb, (u32)rows, ids_of(b), par->d_limbs, par->logn, true, 1, false, st);
    src = pb;
  }
  const size_t nbytes = (size_t)n * b->parts * P.poly_bytes;
  unsigned char* dbytes = (unsigned char*)ws.words((nbytes + 7) / 8);
  launch_pack(P, src, dbytes, rows, par->logn, st);
  FHE_CUDA(cudaGetLastError());
  FHE_CUDA(cudaMemcpyAsync(host_out, dbytes, nbytes, cudaMemcpyDeviceToHost, st));
  FHE_CUDA(cudaStreamSynchronize(st));
  API_END
}
int fhe_b200_batch_unpack(fhe_b200_batch* b, uint32_t first, uint32_t n, const uint8_t* host_in, void* stream) {
  API_BEGIN
  REQUIRE(b && host_in, FHE_B200_INVALID_ARGUMENT, "null argument");
  REQUIRE((uint64_t)first + n <= b->count, FHE_B200_INVALID_ARGUMENT, "range exceeds batch");
  const fhe_b200_params* par = b->par;
  DeviceGuard g(par);
  cudaStream_t st = (cudaStream_t)stream;
  const PackDev P = pack_desc(b);
  const size_t rows = (size_t)n * b->parts * b->limbs;
  Workspace ws(par, st);
  const size_t nbytes = (size_t)n * b->parts * P.poly_bytes;
  unsigned char* dbytes = (unsigned char*)ws.words((nbytes + 7) / 8);
  FHE_CUDA(cudaMemcpyAsync(dbytes, host_in, nbytes, cudaMemcpyHostToDevice, st));
  u64* dst = b->d + b->words_per_ct() * first;
  launch_unpack(P, dbytes, dst, rows, par->logn, st);
  if (b->repr == FHE_B200_NTT)     // rq/convert.rs:128-129: p.into_ntt()
    launch_ntt(dst, dst, (u32)rows, ids_of(b), par->d_limbs, par->logn, false, 1, false, st);
  FHE_CUDA(cudaGetLastError());
  API_END
}

int fhe_b200_sync(void* stream) {
  API_BEGIN
  FHE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  API_END
}

// ------------------------------------------------------------------------------ inspection
int fhe_b200_debug_scaler_tables(const fhe_b200_params* p, uint32_t level, int which, uint32_t* n_from, uint32_t* n_to,
                                 uint32_t* shift, uint64_t* gamma, uint64_t* omega, uint64_t* theta_gamma,
                                 uint64_t* theta_omega_lo, uint64_t* theta_omega_hi, uint8_t* theta_omega_sign,
                                 uint64_t* theta_garner_lo, uint64_t* theta_garner_hi) {
  API_BEGIN
  REQUIRE(p, FHE_B200_INVALID_ARGUMENT, "null argument");
  const LevelData& lv = p->level(level);
  const ScalerTablesH& h = which ? lv.down.h : lv.ext.h;
  if (n_from) *n_from = h.n_from;
  if (n_to) *n_to = h.n_to;
  if (shift) *shift = h.shift;
  auto cp = [](uint64_t* dst, const std::vector<u64>& v) {
    if (dst) for (size_t i = 0; i < v.size(); i++) dst[i] = v[i];
  };
  cp(gamma, h.gamma);
  cp(omega, h.omega);
  if (theta_gamma) { theta_gamma[0] = h.theta_gamma_lo; theta_gamma[1] = h.theta_gamma_hi; theta_gamma[2] = h.theta_gamma_sign; }
  cp(theta_omega_lo, h.theta_omega_lo);
  cp(theta_omega_hi, h.theta_omega_hi);
  if (theta_omega_sign) for (size_t i = 0; i < h.theta_omega_sign.size(); i++) theta_omega_sign[i] = h.theta_omega_sign[i];
  cp(theta_garner_lo, h.theta_garner_lo);
  cp(theta_garner_hi, h.theta_garner_hi);
  API_END
}

int fhe_b200_debug_ntt_tables(const fhe_b200_params* p, uint64_t q, uint64_t* omegas, uint64_t* omegas_shoup,
                              uint64_t* zetas_inv, uint64_t* zetas_inv_shoup, uint64_t* size_inv) {
  API_BEGIN
  REQUIRE(p, FHE_B200_INVALID_ARGUMENT, "null argument");
  int i = p->prime_index(q);
  REQUIRE(i >= 0, FHE_B200_INVALID_MODULUS, "prime not in parameter set");
  const NttTablesH& t = p->tables[i];
  auto cp = [](uint64_t* dst, const std::vector<u64>& v) {
    if (dst) for (size_t k = 0; k < v.size(); k++) dst[k] = v[k];
  };
  cp(omegas, t.om);
  cp(omegas_shoup, t.om_s);
  cp(zetas_inv, t.zi);
  cp(zetas_inv_shoup, t.zi_s);
  if (size_inv) *size_inv = t.ninv;
  API_END
}

}  // extern "C"
