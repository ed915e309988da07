// Internal launch interfaces shared by ntt.cu / kernels.cu / capi.cu.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

#include "ntt.cuh"

namespace fhe_b200 {

extern std::atomic<unsigned long long> g_launches;

struct CudaFail {
  cudaError_t err;
  const char* what;
};
#define FHE_CUDA(x)                                         \
  do {                                                      \
    cudaError_t e__ = (x);                                  \
    if (e__ != cudaSuccess) throw CudaFail{e__, #x};        \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remember, per (device, kernel),
// the largest size already granted and raise it when a launch needs more (thread-safe; a parameter set may live on
// any device of the process).  Throws CudaFail when the device refuses.
void ensure_dynamic_smem(const void* kernel, size_t bytes);

// position -> limb id map of the rows of a buffer
struct RowIds {
  u32 limbs_per_poly;
  unsigned short ids[kMaxPos];
};

// ---- NTT (ntt.cu)
// Transforms n_rows rows of N words.  in may differ from out (first pass reads in).
// in_div / reduce_on_load / lazy_out: see NttArgs.
// digit_adjacent (forward, in_div == limbs_per_poly only): polynomial p = (ct, digit d) writes its limb-j row at
// ((ct*limbs_per_poly + j)*n_dig + d)*N instead of (p*limbs_per_poly + j)*N -- the layout the key-switch inner
// product reads fastest (all digits of one (ct, limb) adjacent).
void launch_ntt(const u64* in, u64* out, u32 n_rows, const RowIds& ids, const LimbDev* limbs, u32 logn,
                bool inverse, u32 in_div, bool reduce_on_load, cudaStream_t st, bool lazy_out = false,
                bool digit_adjacent = false, u32 n_dig = 1);

// Tensor product of two 2-part ciphertexts (as launch_tensor with nca == ncb == L) fused with the inverse transform of
// its 3K output rows: T [ct][3][K][N] receives the POWER-BASIS products.  Returns false when the TMA kernels do not
// serve the shape (the caller then runs launch_tensor + launch_ntt).
bool launch_tensor_inverse_ntt(const u64* a, const u64* b, const u64* xa, const u64* xb, u64* T, u32 cts, u32 L, u32 K,
                               const RowIds& mul_ids, const LimbDev* limbs, u32 logn, cudaStream_t st);

// ---- element-wise (kernels.cu)
enum EwOp { EW_ADD = 0, EW_SUB = 1, EW_NEG = 2 };
void launch_ew(EwOp op, u64* a, const u64* b, size_t n_rows, const RowIds& ids, const LimbDev* limbs, u32 logn,
               cudaStream_t st);

// op 0: a[ct][part][limb][:] *= pt[ct % n_pt][limb][:]   (Modulus::mul_vec, zq/mod.rs:332)
// op 1 / 2: a[ct][0][limb][:] +=/-= pt[ct % n_pt][limb][:]   (Ciphertext +=/-= &Plaintext, ops/mod.rs:88-97, :188-197)
void launch_mul_plain(u64* a, const u64* pt, u32 cts, u32 parts, u32 n_pt, const RowIds& ids, const LimbDev* limbs,
                      u32 logn, cudaStream_t st, u32 op = 0);

// dot_product_scalar (bfv/ops/dot_product.rs:55-184): out[g] = sum_{i<n_terms} ct[(g*n+i) % ct_count] (.) pt[(g*n+i) % pt_count]
// ct: [ct_count][parts][limbs][N], pt: [pt_count][limbs][N], out: [groups][parts][limbs][N], all NTT
void launch_dot(const u64* ct, const u64* pt, u64* out, u32 groups, u32 n_terms, u32 parts, u32 ct_count,
                u32 pt_count, const RowIds& ids, const LimbDev* limbs, u32 logn, cudaStream_t st);

// tensor product of two 2-part ciphertexts over the multiplication basis (mul.rs:198-201).
// a,b: [ct][2][L][N] NTT (supply the first nca / ncb mul-basis limbs of their side: the common prefix a factor-one
// extender keeps); xa: [ct][2][K-nca][N], xb: [ct][2][K-ncb][N] (the scaled limbs, NTT); out: [ct][3][K][N].
void launch_tensor(const u64* a, const u64* b, const u64* xa, const u64* xb, u64* out, u32 cts, u32 L, u32 nca,
                   u32 ncb, u32 K, const RowIds& mul_ids, const LimbDev* limbs, u32 logn, cudaStream_t st);

// general part counts (ops/mod.rs:259-358): a [ct][na][L][N], b [ct][nb][L][N], xa [ct][na][E][N], xb [ct][nb][E][N]
// -> out [ct][na+nb-1][K][N], c[k] = sum_{i+j=k} a_i * b_j
void launch_tensor_nm(const u64* a, const u64* b, const u64* xa, const u64* xb, u64* out, u32 cts, u32 L, u32 E, u32 na,
                      u32 nb, const RowIds& mul_ids, const LimbDev* limbs, u32 logn, cudaStream_t st);

// exact RNS scaler (rns/scaler.rs:249-352), tables resident on the device
struct ScalerDev {
  u32 n_from, n_to, is_one, shift;
  u64 tg_lo, tg_hi;
  u32 tg_sign;
  u32 all_solinas;   // every `to` limb is 2^62 - c, c < 2^28 (the persistent TMA kernel's epilogue needs it)
  const u64* gamma;       // [n_to]
  const u64* omega;       // [n_to][n_from]
  const u64* to_lo;       // theta_omega [n_from]
  const u64* to_hi;
  const unsigned char* to_sign;
  const unsigned char* to_order;   // source indices, the theta_omega terms with positive sign first
  u32 n_pos, n_terms;              // positive-sign terms, all non-zero terms (<= n_from)
  const u64* tgar_lo;     // theta_garner [n_from]
  const u64* tgar_hi;
  unsigned short to_ids[kMaxPos];
};
// in: [polys][n_from][N] power basis.  Output rows `start .. start+n_out` of the `to` basis:
//  split3 == 0: out0 + (poly * out_rows_per_poly + row) * N
//  split3 == 1: polys come in triples (c0,c1,c2); c0,c1 -> out0 as [ct][2][n_out][N], c2 -> out1 as [ct][n_out][N]
void launch_scale(const ScalerDev& S, const LimbDev* limbs, const u64* in, u64* out0, u64* out1, u32 polys,
                  u32 out_rows_per_poly, u32 start, u32 n_out, int split3, u32 logn, cudaStream_t st);

// key-switch inner product (key_switching_key.rs:256-268) on already transformed digits:
// inter: [ct][n_dig][Lk][N] NTT values (lazy, any 64-bit word), or [ct][Lk][n_dig][N] when `adjacent`;
// k0,k1: [Lk][n_dig][N] (limb-major: the device copy of a key is transposed once at upload);
// out0/out1 row (ct, j) at out + (ct*out_ct_rows + j)*N ; base0/base1 (nullable) same indexing.
void launch_ksmac(const u64* inter, const u64* k0, const u64* k1, const u64* base0, const u64* base1, u64* out0,
                  u64* out1, u32 cts, u32 n_dig, u32 Lk, u32 out_ct_rows, const RowIds& ids, const LimbDev* limbs,
                  u32 logn, cudaStream_t st, bool adjacent = false);

// whether launch_ntt will take the TMA kernels for this shape (they can write the digit-adjacent layout)
bool ntt_uses_tma(u32 n_rows, const RowIds& ids, u32 logn, u32 in_div, const u64* in, const u64* out);

// base-2^log_base digit decomposition of single-limb polynomials (key_switching_key.rs:339-345):
// in [polys][N] -> out [polys][n_dig][N]
void launch_decompose(const u64* in, u64* out, size_t polys, u32 n_dig, u32 log_base, u32 logn, cudaStream_t st);

// NTT-domain substitution gather (rq/mod.rs:368-377): out[row][t] = in[row][perm[t]]
void launch_gather(const u64* in, u64* out, size_t n_rows, const int* perm, u32 logn, cudaStream_t st);
// Poly<PowerBasis>::substitute (rq/mod.rs:390-408): signed coefficient scatter x^j -> x^(j*exponent)
void launch_substitute_power(const u64* in, u64* out, size_t n_rows, u32 exponent, const RowIds& ids,
                             const LimbDev* limbs, u32 logn, cudaStream_t st);

// Poly<PowerBasis>::switch_down (rq/mod.rs:433-492): in [polys][L][N] -> out [polys][L-1][N]
struct SwitchDownDev {
  u64 q_last, q_last_half;
  const u64* half_mod;  // [L-1]: q_i - (q_last/2 mod q_i)
  const u64* inv;       // [L-1]: q_last^-1 mod q_i
  const u64* inv_s;     // shoup
};
void launch_switch_down(const SwitchDownDev& S, const u64* in, u64* out, u32 polys, u32 L, const RowIds& ids,
                        const LimbDev* limbs, u32 logn, cudaStream_t st);

// bit (un)packing of power-basis rows (fhe-util/src/lib.rs:71-146): row r of `rows` uses nbits[r % limbs] bits per
// coefficient; packed row r starts at byte  (r / limbs) * poly_bytes + offs[r % limbs]
struct PackDev {
  u32 limbs;
  unsigned char nbits[kMaxPos];
  u32 offs[kMaxPos];
  u32 poly_bytes;
};
void launch_pack(const PackDev& P, const u64* words, unsigned char* bytes, size_t n_rows, u32 logn, cudaStream_t st);
void launch_unpack(const PackDev& P, const unsigned char* bytes, u64* words, size_t n_rows, u32 logn, cudaStream_t st);

}  // namespace fhe_b200
