/*
 * fhe_b200.h -- C ABI of the B200-native BFV ciphertext-arithmetic engine.
 *
 * Drop-in boundary for the fhe-math / fhe::bfv hot path of tlepoint/fhe.rs @ e248cd28
 * (pure Rust; it has no FFI of its own -- SURVEY.md section 8b).  Each entry point
 * names the reference item it replaces (file:line under /root/reference/crates).
 * A Rust host binds these with `extern "C"`; in this repository the
 * same symbols are driven from C++ (include/fhe_b200.hpp) and Python ctypes
 * (fhe_rs_b200/_capi.py).
 *
 * Conventions
 *  - plain pointers and sizes only; all polynomial words are u64 residues.
 *  - every call returns 0 on success or a negative fhe_b200_status mirroring the
 *    reference's error enums (fhe-math/src/errors.rs:14-113, fhe/src/errors.rs:17-66);
 *    nothing throws or aborts across the boundary.  fhe_b200_last_error() returns the
 *    message of the calling thread's last failure.  (The reference's operators `+ - *`
 *    panic on mismatched operands, ops/mod.rs:19-29; here the same conditions return
 *    FHE_B200_CONTEXT_MISMATCH / FHE_B200_INVALID_LEVEL.)
 *  - caller owns host memory; the library owns device memory behind opaque handles,
 *    released by the matching *_destroy / *_free.  params and ksk handles are immutable
 *    after creation and may be shared by host threads (like Arc<BfvParameters>,
 *    bfv/parameters.rs:125); a batch handle must not be mutated concurrently.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Work is
 *    enqueued asynchronously; fhe_b200_sync() or a download waits for it.  `stream` is the only
 *    stream the caller has to reason about: a batched operation over more ciphertexts than one chunk
 *    (FHE_B200_CHUNK, default 256) runs its chunks on side streams owned by the parameter set, but
 *    these start behind everything already enqueued on `stream` and `stream` waits for them before
 *    the call returns, so the operation is ordered on `stream` like a single kernel would be.
 *  - host layout of a batch == Vec<u64>::from(&Poly) of the reference
 *    (rq/convert.rs:474-503) concatenated over parts and ciphertexts:
 *    [ciphertext][part][limb][coefficient], row-major, limb i modulo moduli[i].
 *  - there is NO CPU fallback: compute entry points fail with FHE_B200_NO_DEVICE when the
 *    parameter set was created without a CUDA device.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FHE_B200_OK = 0,
  FHE_B200_INVALID_ARGUMENT = -1,       /* null pointer, zero count, bad size                        */
  FHE_B200_INVALID_MODULUS = -2,        /* fhe_math::Error::InvalidModulus / NonCoprimeModuli        */
  FHE_B200_INVALID_DEGREE = -3,         /* fhe_math::Error::InvalidPolynomialDegree                   */
  FHE_B200_NTT_UNAVAILABLE = -4,        /* fhe_math::Error::NttOperatorUnavailable                    */
  FHE_B200_CONTEXT_MISMATCH = -5,       /* PolynomialContextMismatch / fhe::Error::ParameterMismatch  */
  FHE_B200_INVALID_LEVEL = -6,          /* fhe::Error::InvalidLevel / InvalidContextLevel             */
  FHE_B200_BAD_POLY_COUNT = -7,         /* CiphertextError::{MultiplicationPolynomialCount,InvalidPolynomialCount} */
  FHE_B200_INVALID_REPRESENTATION = -8, /* fhe_math::Error::IncorrectRepresentation                   */
  FHE_B200_NO_MORE_CONTEXT = -9,        /* fhe_math::Error::NoMoreContext                             */
  FHE_B200_INVALID_EXPONENT = -10,      /* fhe_math::Error::InvalidSubstitutionExponent               */
  FHE_B200_UNSUPPORTED = -11,           /* feature outside the accelerated path                       */
  FHE_B200_CUDA_ERROR = -20,
  FHE_B200_OUT_OF_MEMORY = -21,
  FHE_B200_NO_DEVICE = -22
} fhe_b200_status;

/* Poly representation tag (rq/mod.rs Representation). */
typedef enum { FHE_B200_POWER_BASIS = 0, FHE_B200_NTT = 1 } fhe_b200_repr;

typedef struct fhe_b200_params fhe_b200_params; /* == Arc<BfvParameters> (bfv/parameters.rs:88-114)            */
typedef struct fhe_b200_batch fhe_b200_batch;   /* == Vec<Ciphertext> of one level (bfv/ciphertext.rs:18-32),
                                                   device resident, [count][parts][limbs][N] u64               */
typedef struct fhe_b200_ksk fhe_b200_ksk;       /* == KeySwitchingKey (bfv/keys/key_switching_key.rs:22-45)     */
typedef struct fhe_b200_multiplicator fhe_b200_multiplicator; /* == Multiplicator with a custom strategy
                                                   (bfv/ops/mul.rs:22-98)                                        */

const char* fhe_b200_version(void);
const char* fhe_b200_last_error(void);

/* ---- parameters ---------------------------------------------------------------------
 * BfvParametersBuilder::build (bfv/parameters.rs:560-738) for explicit moduli
 * (`set_moduli`) or generated ones (`set_moduli_sizes`, parameters.rs:391-431).
 * plaintext_le: plaintext modulus as little-endian bytes (as bfv.proto PlaintextBig).
 * psi: optional 2N-th primitive roots, one per prime in the order
 *      [moduli..., extended_basis...] (n_moduli + n_moduli + 1 entries); NULL selects the
 *      documented default root.  A Rust host passes the reference's roots
 *      (NttOperator.omegas[N/2], ntt/native.rs:50-56) so that NTT-domain data is
 *      interchangeable bit for bit.
 * device: CUDA device ordinal, or -1 for a host-only handle (table inspection only). */
int fhe_b200_params_create(int device, uint32_t degree, const uint64_t* moduli, uint32_t n_moduli,
                           const uint8_t* plaintext_le, uint32_t plaintext_len, const uint64_t* psi,
                           fhe_b200_params** out);
int fhe_b200_params_create_from_sizes(int device, uint32_t degree, const uint32_t* moduli_sizes,
                                      uint32_t n_moduli, const uint8_t* plaintext_le,
                                      uint32_t plaintext_len, fhe_b200_params** out);
int fhe_b200_params_destroy(fhe_b200_params* p);
uint32_t fhe_b200_params_degree(const fhe_b200_params* p);                 /* BfvParameters::degree          */
uint32_t fhe_b200_params_n_moduli(const fhe_b200_params* p);               /* BfvParameters::moduli().len()  */
int fhe_b200_params_moduli(const fhe_b200_params* p, uint64_t* out);       /* BfvParameters::moduli          */
/* multiplication basis of `level`: the level's moduli followed by the extension primes
 * (parameters.rs:660-700); *n receives the count, out may be NULL to query it. */
int fhe_b200_params_mul_basis(const fhe_b200_params* p, uint32_t level, uint64_t* out, uint32_t* n);
/* psi actually used for prime `q` of this parameter set. */
int fhe_b200_params_psi(const fhe_b200_params* p, uint64_t q, uint64_t* psi);

/* ---- batches --------------------------------------------------------------------------
 * parts = polynomials per ciphertext (2 fresh, 3 after `&ct * &ct`); level = modulus-
 * switching level (ciphertext.rs:31); limbs = n_moduli - level. */
int fhe_b200_batch_alloc(const fhe_b200_params* p, uint32_t count, uint32_t parts, uint32_t level,
                         int repr, fhe_b200_batch** out);
int fhe_b200_batch_free(fhe_b200_batch* b);
int fhe_b200_batch_info(const fhe_b200_batch* b, uint32_t* count, uint32_t* parts, uint32_t* level,
                        uint32_t* limbs, int* repr);
/* host <-> device copy of ciphertexts [first, first+n); host may be pageable or pinned.  Uploads (and the host_polys
 * of fhe_b200_mul_plain / fhe_b200_add_plain) are only enqueued: a pinned source must stay valid until `stream` has
 * passed the call (a pageable one has been staged by the CUDA runtime when the call returns). */
int fhe_b200_batch_upload(fhe_b200_batch* b, uint32_t first, uint32_t n, const uint64_t* host, void* stream);
int fhe_b200_batch_download(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint64_t* host, void* stream);
/* same as download, but only enqueues the copy on `stream` (host must be pinned for it to be asynchronous);
 * the caller waits with fhe_b200_sync(stream) before reading `host` */
int fhe_b200_batch_download_async(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint64_t* host, void* stream);
/* Ciphertext::clone: dst <- src (same parameters, shape, level; representation is copied) */
int fhe_b200_batch_copy(fhe_b200_batch* dst, const fhe_b200_batch* src, void* stream);
/* Page-locked host staging memory for asynchronous uploads / downloads (cudaHostAlloc, portable across devices).
 * write_combined != 0 asks for write-combining pages: faster for the device to read over PCIe and invisible to the
 * CPU caches, slow for the CPU to read -- meant for upload-only buffers the host fills once, front to back. */
int fhe_b200_host_alloc(size_t bytes, int write_combined, void** out);
int fhe_b200_host_free(void* p);
/* raw device pointer of the batch storage (for zero-copy producers such as bench.py). */
int fhe_b200_batch_device_ptr(const fhe_b200_batch* b, uint64_t** dptr, size_t* n_words);

/* ---- keys -----------------------------------------------------------------------------
 * KeySwitchingKey (key_switching_key.rs:22-45): c0, c1 are the NTT-domain values of the
 * n_digits key polynomials at the key level, host layout [digit][limb][coeff];
 * n_digits must equal the limb count of ciphertext_level (:113) -- or, when the key level has a
 * single modulus q, ceil(log_modulus / log_base) with log_modulus = ilog2(next_power_of_two(q)),
 * log_base = log_modulus / 2: the base-2^log_base decomposition variant (:92-110), whose key switch
 * (key_switch_decomposition, :323-362) the library then runs.  Shoup companions (Poly<NttShoup>) are
 * not needed by the device inner product. */
int fhe_b200_ksk_upload(const fhe_b200_params* p, uint32_t ciphertext_level, uint32_t ksk_level,
                        const uint64_t* c0, const uint64_t* c1, uint32_t n_digits, fhe_b200_ksk** out);
int fhe_b200_ksk_free(fhe_b200_ksk* k);

/* ---- primitives (each parity-tested one by one) --------------------------------------- */
/* Poly::into_ntt / NttOperator::forward[_vt] on every row (rq/mod.rs:535, ntt/native.rs:77,183) */
int fhe_b200_ntt_forward(fhe_b200_batch* b, void* stream);
/* Poly::into_power_basis / NttOperator::backward[_vt] (rq/mod.rs:590, ntt/native.rs:106,197) */
int fhe_b200_ntt_backward(fhe_b200_batch* b, void* stream);
/* Ciphertext += / -= / unary - (bfv/ops/mod.rs:54, :148, :205; rq/ops.rs:92-172, :354-418) */
int fhe_b200_add(fhe_b200_batch* a, const fhe_b200_batch* b, void* stream);
int fhe_b200_sub(fhe_b200_batch* a, const fhe_b200_batch* b, void* stream);
int fhe_b200_neg(fhe_b200_batch* a, void* stream);
/* Ciphertext *= &Plaintext (bfv/ops/mod.rs:229-238; Poly<Ntt> *= &Poly<Ntt>, rq/ops.rs:174): every part of every
 * ciphertext is multiplied coefficient-wise by an NTT-domain polynomial.  host_polys holds n_polys polynomials of
 * [limbs][N] words (Plaintext::poly_ntt, or a monomial of EvaluationKey::expands); n_polys is 1 (shared) or count. */
int fhe_b200_mul_plain(fhe_b200_batch* a, const uint64_t* host_polys, uint32_t n_polys, void* stream);
/* Ciphertext += &Plaintext / -= &Plaintext (bfv/ops/mod.rs:88-97, :188-197): part 0 of every ciphertext gets
 * +/- Plaintext::to_poly() (the delta-scaled NTT polynomial the host computes, plaintext.rs:172-197); host_polys as for
 * fhe_b200_mul_plain. */
int fhe_b200_add_plain(fhe_b200_batch* a, const uint64_t* host_polys, uint32_t n_polys, int subtract, void* stream);
/* dot_product_scalar (bfv/ops/dot_product.rs:55-184): out[g] = sum_{i < n_terms} cts[g*n_terms + i] (.) pts[g*n_terms + i]
 * for g < out.count.  pts is a batch with one NTT polynomial per entry (Plaintext::poly_ntt); either operand may hold
 * n_terms entries only, shared by every group (the PIR loops of examples/mulpir.rs:153-181 share the expanded query
 * across database columns), or out.count * n_terms.  Errors: EmptyInput / OperandCountMismatch -> INVALID_ARGUMENT,
 * CiphertextPolynomialCountMismatch -> BAD_POLY_COUNT, mixed levels -> INVALID_LEVEL. */
int fhe_b200_dot_product_scalar(const fhe_b200_batch* cts, const fhe_b200_batch* pts, uint32_t n_terms,
                                fhe_b200_batch* out, void* stream);
/* &Ciphertext * &Ciphertext (bfv/ops/mod.rs:259-358): n parts x m parts -> n + m - 1 parts (out3 must have that many;
 * 2 x 2 -> 3 is the fused path) */
int fhe_b200_mul(const fhe_b200_batch* a, const fhe_b200_batch* b, fhe_b200_batch* out3, void* stream);
/* RelinearizationKey::relinearizes: (c0,c1,c2) -> (c0,c1) (keys/relinearization_key.rs:70-103) */
int fhe_b200_relinearize(const fhe_b200_batch* ct3, const fhe_b200_ksk* rk, fhe_b200_batch* out2, void* stream);
/* Multiplicator::default(rk).multiply (bfv/ops/mul.rs:101-138, :165-243); mod_switch != 0
 * additionally applies Ciphertext::switch_down (mul.rs:238) and out2 must be at level+1. */
int fhe_b200_mul_relin(const fhe_b200_batch* a, const fhe_b200_batch* b, const fhe_b200_ksk* rk,
                       int mod_switch, fhe_b200_batch* out2, void* stream);
/* Multiplicator::new / new_leveled (bfv/ops/mul.rs:37-98): custom strategy.  Each ScalingFactor (rns/scaler.rs:20-58)
 * is a numerator / denominator pair of little-endian byte strings (BigUint::to_bytes_le).  extended_basis are the
 * n_basis moduli of the multiplication context (Context::new(extended_basis), mul.rs:82); psi (nullable) gives the
 * 2N-th root per basis prime (default rule of fhe_b200_params_create otherwise; primes shared with the parameter set
 * reuse its tables).  As in rq/scaler.rs:35-43 an extender keeps the common prefix of the two bases only when its
 * factor is one.  Errors: InvalidLevel, DuplicateModuli / InvalidModulus, NTT_UNAVAILABLE. */
int fhe_b200_multiplicator_create(const fhe_b200_params* p, uint32_t level, const uint8_t* lhs_num, uint32_t lhs_num_len,
                                  const uint8_t* lhs_den, uint32_t lhs_den_len, const uint8_t* rhs_num,
                                  uint32_t rhs_num_len, const uint8_t* rhs_den, uint32_t rhs_den_len,
                                  const uint64_t* extended_basis, uint32_t n_basis, const uint64_t* psi,
                                  const uint8_t* post_num, uint32_t post_num_len, const uint8_t* post_den,
                                  uint32_t post_den_len, fhe_b200_multiplicator** out);
int fhe_b200_multiplicator_free(fhe_b200_multiplicator* m);
/* Multiplicator::multiply (mul.rs:165-243) with that strategy.  rk == NULL: no relinearization, out has 3 parts;
 * otherwise enable_relinearization(rk) semantics (mul.rs:141-151: the key must be for the multiplicator's level,
 * ParameterMismatch if not) and out has 2 parts.  mod_switch != 0: enable_mod_switching (mul.rs:155-162,
 * NoMoreContext at the last level), out must be at level+1. */
int fhe_b200_multiplicator_multiply(const fhe_b200_multiplicator* m, const fhe_b200_batch* a, const fhe_b200_batch* b,
                                    const fhe_b200_ksk* rk, int mod_switch, fhe_b200_batch* out, void* stream);
/* GaloisKey::relinearize (keys/galois_key.rs:63-86) for substitution exponent `exponent`
 * (column rotation by i <-> 3^i mod 2N, row swap <-> 2N-1; evaluation_key.rs:118, :278-286) */
int fhe_b200_galois(const fhe_b200_batch* ct, uint32_t exponent, const fhe_b200_ksk* gk,
                    fhe_b200_batch* out, void* stream);
/* Poly::substitute on every row of a batch: the slot permutation of an NTT batch (rq/mod.rs:360-389) or the signed
 * coefficient permutation x^j -> x^(j*exponent) of a power-basis batch (rq/mod.rs:390-408); `out` takes `in`'s
 * representation.  Even exponents: FHE_B200_INVALID_EXPONENT. */
int fhe_b200_substitute(const fhe_b200_batch* in, uint32_t exponent, fhe_b200_batch* out, void* stream);
/* Ciphertext::switch_down: drop the last modulus with rounding (ciphertext.rs:148-161, rq/mod.rs:433-492).
 * Stream-ordered and in place: the batch keeps its allocation (fhe_b200_batch_device_ptr stays valid, the words of the
 * lower level are compacted at its start), nothing is allocated or synchronised. */
int fhe_b200_switch_down(fhe_b200_batch* b, void* stream);
/* KeySwitchingKey::key_switch on part `part` of a POWER_BASIS batch (key_switching_key.rs:241-270):
 * out (2 parts, NTT, ksk level) = (sum_i NTT(d_i) * c0_i, sum_i NTT(d_i) * c1_i) */
int fhe_b200_key_switch(const fhe_b200_batch* pb, uint32_t part, const fhe_b200_ksk* k,
                        fhe_b200_batch* out2, void* stream);
/* rq::scaler::Scaler::scale with the level's multiplication scalers (rq/scaler.rs:55-127):
 * which = 0: extender (level basis -> multiplication basis, factor 1),
 * which = 1: down scaler (multiplication basis -> level basis, factor t/Q).
 * `in` is an NTT batch whose limb count equals the source basis; out gets the target basis. */
int fhe_b200_scale(const fhe_b200_batch* in, int which, fhe_b200_batch* out, void* stream);
/* batch over the multiplication basis of `level` (limbs = L + E), for fhe_b200_scale */
int fhe_b200_batch_alloc_mul_basis(const fhe_b200_params* p, uint32_t count, uint32_t parts, uint32_t level,
                                   int repr, fhe_b200_batch** out);

/* ---- wire format of polynomials -------------------------------------------
 * `Rq.coefficients` of the reference's protobuf message (fhe-math/src/proto/rq.proto:12-17) is, for every limb
 * in order, the power-basis coefficients bit-packed LSB first with ceil(log2 q_i) bits each
 * (Modulus::serialize_vec zq/mod.rs:783-786, fhe_util::transcode_to_bytes fhe-util/src/lib.rs:71-108).
 * The protobuf framing itself (tags, varints, `representation`, `degree`) is host work: a Rust host keeps its prost
 * code; C++ and Python hosts have the same messages in include/fhe_b200_wire.hpp / fhe_rs_b200/wire.py. */
/* Seeded ("compact") ciphertexts and keys are NOT expanded here.  The reference serialises a fresh ciphertext as c0 plus
 * the 32-byte seed of c1 (bfv/ciphertext.rs:231-317; keys: key_switching_key.rs:365-482) and regenerates c1 with
 * Poly::random_from_seed (rq/mod.rs:276-292): ChaCha8Rng::from_seed(seed) driving rand's uniform u64 sampling per
 * limb.  That stream is defined by rand 0.10.2 / rand_chacha 0.10.0, which are not vendored in the reference tree
 * and cannot be run or pinned in this build environment, so a device-side expansion could not be proven identical.
 * The Rust host therefore expands c1 with the reference's own code (ciphertext.rs:287-300) and uploads both halves
 * as words (fhe_b200_batch_upload) or as Rq blobs (fhe_b200_batch_unpack); the device never guesses the RNG. */
/* Modulus::serialization_length summed over the limbs of `level` (rq/convert.rs:78-82): bytes per polynomial */
int fhe_b200_poly_packed_bytes(const fhe_b200_params* p, uint32_t level, size_t* nbytes);
/* the same for the polynomials of one batch -- use this one to size the host buffers of pack / unpack: a batch over
 * the multiplication basis (fhe_b200_batch_alloc_mul_basis) has L + E limbs, not the level's L */
int fhe_b200_batch_packed_bytes(const fhe_b200_batch* b, size_t* nbytes);
/* From<&Poly<R>> for Rq (rq/convert.rs:17-44): polynomials of ciphertexts [first, first+n) -> power basis
 * (if the batch is NTT) -> packed bytes; host_out receives n*parts blobs of packed_bytes each. */
int fhe_b200_batch_pack(const fhe_b200_batch* b, uint32_t first, uint32_t n, uint8_t* host_out, void* stream);
/* TryConvertFrom<&Rq> for Poly<PowerBasis|Ntt> (rq/convert.rs:100-131): unpack n*parts blobs into ciphertexts
 * [first, first+n) of `b`; when `b` is an NTT batch the rows are forward-transformed afterwards (as
 * `p.into_ntt()` does).  The whole batch must be filled with one call per disjoint range before it is used. */
int fhe_b200_batch_unpack(fhe_b200_batch* b, uint32_t first, uint32_t n, const uint8_t* host_in, void* stream);

int fhe_b200_sync(void* stream);
/* kernels launched by this library in the calling process so far (bench.py "gpu_launches") */
uint64_t fhe_b200_launch_count(void);

/* ---- inspection of the host precompute (CPU-only tests of the parameter builder) -------
 * RnsScaler tables (rns/scaler.rs:52-73) of the level's extender (which=0) / down scaler
 * (which=1).  Any output pointer may be NULL.  omega is [n_to][n_from]. */
int fhe_b200_debug_scaler_tables(const fhe_b200_params* p, uint32_t level, int which, uint32_t* n_from,
                                 uint32_t* n_to, uint32_t* shift, uint64_t* gamma, uint64_t* omega,
                                 uint64_t* theta_gamma /* lo,hi,sign */, uint64_t* theta_omega_lo,
                                 uint64_t* theta_omega_hi, uint8_t* theta_omega_sign,
                                 uint64_t* theta_garner_lo, uint64_t* theta_garner_hi);
/* NTT tables of prime q: any of omegas/zetas_inv (N words each) may be NULL. */
int fhe_b200_debug_ntt_tables(const fhe_b200_params* p, uint64_t q, uint64_t* omegas, uint64_t* omegas_shoup,
                              uint64_t* zetas_inv, uint64_t* zetas_inv_shoup, uint64_t* size_inv);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
