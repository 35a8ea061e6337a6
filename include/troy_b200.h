/* troy_b200 — C ABI of the B200-native RNS polynomial engine behind the troy (troyn::) API.
 *
 * The reference (lightbulb128/troy) has no FFI: its boundary is the troyn:: C++ classes and, underneath, the free
 * functions of namespace troy::kernel_util plus RNSToolCuda / GaloisToolCuda methods. This header cuts
 * the boundary exactly at that seam. Every entry point names the reference interface it replaces (file:line under
 * /root/reference). Plain pointers and sizes only; all `d_` pointers are device pointers on the context's GPU;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream). Every function returns 0 on success or a
 * TROYB_E_* code; troyb_last_error() gives the message (the C++ shim maps codes to the reference's exception types).
 *
 * Layouts (identical to the reference):
 *   ciphertext      [poly][limb][coeff] uint64, limb count = level's coeff_modulus size   src/ciphertext_cuda.cuh:62-67
 *   batch           `batch` ciphertexts `*_stride` words apart (stride >= size*L*N); batch = 1 is the reference call
 *   key-switch key  [digit j][component k<2][key limb][coeff], NTT form at the key level   src/kswitchkeys_cuda.cuh:43-56,366
 * Levels: level 0 = first (data) level = all coeff primes but the special one; level l drops l more primes
 * (src/context.cpp:470-524). Key level = level 0 primes + special prime.
 */
#ifndef TROY_B200_H
#define TROY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TROYB_OK 0
#define TROYB_E_INVALID_ARGUMENT 1 /* std::invalid_argument in the reference */
#define TROYB_E_LOGIC 2            /* std::logic_error */
#define TROYB_E_OUT_OF_RANGE 3     /* std::out_of_range */
#define TROYB_E_CUDA 4             /* std::runtime_error("CUDA error.")  src/kernelprovider.cuh:6-14 */
#define TROYB_E_NOT_INITIALIZED 5  /* std::invalid_argument("KernelProvider not initialized.")  src/kernelprovider.cuh:24-27 */

#define TROYB_SCHEME_BFV 1
#define TROYB_SCHEME_CKKS 2
#define TROYB_SCHEME_BGV 3

typedef struct troyb_context troyb_context;

const char *troyb_last_error(void);

/* KernelProvider::initialize()  src/kernelprovider.cuh:29-33  (device index added; the reference hard-codes 0) */
int troyb_initialize(int device);
int troyb_initialized(void);
/* More than one GPU in a process (the reference pins device 0, src/kernelprovider.cuh:30): the CUDA current device is a
 * per-host-thread setting. Calls that take a context make the context's device current for the calling thread; the
 * context-free calls (troyb_malloc, the copies, troyb_context_create) act on the thread's current device, chosen by
 * troyb_initialize / troyb_set_device. */
int troyb_set_device(int device);
int troyb_context_device(const troyb_context *ctx);

/* Number of kernels this library has launched so far in this process (diagnostic used by bench.py). */
uint64_t troyb_launch_count(void);

/* Per-kernel device timing for bench.py's roofline (CUDA events around every launch while enabled; off by default).
 * troyb_profile_report writes JSON lines {"kernel","launches","grid","ms"} aggregated by kernel name and clears the log. */
int troyb_profile_enable(int on);
int troyb_profile_report(char *buf, size_t buf_bytes);

/* CoeffModulus::Create(poly_modulus_degree, bit_sizes)  src/modulus.cpp:80-122 (PlainModulus::Batching is the 1-element case) */
int troyb_coeff_modulus_create(uint64_t poly_modulus_degree, const int *bit_sizes, int count, uint64_t *out);

/* KernelProvider::malloc / free / copy / retrieve / memset  src/kernelprovider.cuh:35-83 (stream-ordered) */
int troyb_malloc(void **d_ptr, size_t bytes, void *stream);
int troyb_free(void *d_ptr, void *stream);
int troyb_copy_to_device(void *d_dst, const void *h_src, size_t bytes, void *stream);
int troyb_copy_to_host(void *h_dst, const void *d_src, size_t bytes, void *stream);
int troyb_copy_on_device(void *d_dst, const void *d_src, size_t bytes, void *stream);
int troyb_memset_zero(void *d_ptr, size_t bytes, void *stream);
int troyb_stream_synchronize(void *stream);

/* SEALContextCuda(EncryptionParameters, expand_mod_chain, SecurityLevel)  src/context_cuda.cuh:146-156, src/context_cuda.cu:5-62
 * Builds every NTT table / RNS constant on the host and uploads it once. Security level is not enforced (none). */
int troyb_context_create(int scheme, uint64_t poly_modulus_degree, const uint64_t *coeff_moduli, int n_moduli,
                         uint64_t plain_modulus, int expand_mod_chain, troyb_context **out);
void troyb_context_destroy(troyb_context *ctx);
int troyb_context_scheme(const troyb_context *ctx);
uint64_t troyb_context_poly_degree(const troyb_context *ctx);
int troyb_context_num_levels(const troyb_context *ctx);           /* data levels in the chain */
int troyb_context_level_limbs(const troyb_context *ctx, int level); /* level = -1: key level */
int troyb_context_level_moduli(const troyb_context *ctx, int level, uint64_t *out);
int troyb_context_using_keyswitching(const troyb_context *ctx);   /* SEALContext::using_keyswitching */
int troyb_context_total_coeff_bits(const troyb_context *ctx, int level); /* ContextData::totalCoeffModulusBitCount */
uint64_t troyb_context_plain_modulus(const troyb_context *ctx);
/* BEHZ auxiliary base of a BFV level (RNSTool::baseBsk, src/utils/rns.h:246): returns |Bsk|, writes the primes */
int troyb_context_bsk(const troyb_context *ctx, int level, uint64_t *out);
/* NTTTablesCuda root powers of one key-level modulus, N pairs {operand, quotient}  src/utils/ntt_cuda.cuh:15-102 */
int troyb_context_root_powers(const troyb_context *ctx, int key_limb, uint64_t *h_out);

/* kernel_util::kNttNegacyclicHarvey / kInverseNttNegacyclicHarvey  src/kernelutils.cuh:578-632
 * In-place transform of n_polys polynomials of the level's L limbs each, [n_polys][L][N]; canonical output. */
int troyb_ntt_forward(troyb_context *ctx, int level, uint64_t *d_data, size_t n_polys, void *stream);
int troyb_ntt_inverse(troyb_context *ctx, int level, uint64_t *d_data, size_t n_polys, void *stream);
/* Same over the key-level base (level limbs + special prime): used for key material and tests. */
int troyb_ntt_forward_key(troyb_context *ctx, uint64_t *d_data, size_t n_polys, void *stream);
int troyb_ntt_inverse_key(troyb_context *ctx, uint64_t *d_data, size_t n_polys, void *stream);

/* kAddPolyCoeffmod / kSubPolyCoeffmod / kNegatePolyCoeffmod  src/kernelutils.cu:48,524,318
 * (EvaluatorCuda::addInplace / subInplace / negateInplace, src/evaluator_cuda.cu:143-260). n_polys = batch*size. */
int troyb_add(troyb_context *ctx, int level, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, size_t n_polys, void *stream);
int troyb_sub(troyb_context *ctx, int level, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, size_t n_polys, void *stream);
int troyb_negate(troyb_context *ctx, int level, const uint64_t *d_a, uint64_t *d_out, size_t n_polys, void *stream);
/* kMultiplyPolyScalarCoeffmod  src/kernelutils.cu:272-298: out = a * (scalar mod q_l) mod q_l on every limb
 * (BGV correction-factor balancing before add / sub, src/evaluator_cuda.cu:171-190) */
int troyb_multiply_scalar(troyb_context *ctx, int level, const uint64_t *d_a, uint64_t scalar, uint64_t *d_out, size_t n_polys,
                          void *stream);

/* kMultiplyInvPolyDegreeCoeffmod  src/kernelutils.cu:578-592 (EvaluatorCuda::divideByPolyModulusDegreeInplace, src/evaluator_cuda.cu:2262):
 * out = a * N^-1 mod q_l on every limb. */
int troyb_multiply_inv_degree(troyb_context *ctx, int level, const uint64_t *d_a, uint64_t *d_out, size_t n_polys, void *stream);
/* kNegacyclicShiftPolyCoeffmod  src/kernelutils.cu:537-576 (EvaluatorCuda::negacyclicShift / extractLWE / assembleLWE / packLWECiphertexts,
 * src/evaluator_cuda.cu:2185-2365): out = in * X^shift modulo X^N + 1 on n_polys coefficient-form polys, shift in [0, 2N); out of place. */
int troyb_negacyclic_shift(troyb_context *ctx, int level, const uint64_t *d_in, uint64_t *d_out, size_t n_polys, uint64_t shift, void *stream);

/* EvaluatorCuda::multiplyInplace → ckksMultiply / bgvMultiply / bfvMultiply  src/evaluator_cuda.cu:262-523
 * (squareInplace :525-701 is the same call with d_b = d_a). Inputs a: a_size polys, b: b_size polys; output
 * a_size+b_size-1 polys. d_out may alias d_a (in-place, as the reference does after resize).
 * CKKS operands are NTT form; BFV/BGV operands are coefficient form. */
int troyb_multiply(troyb_context *ctx, int level, const uint64_t *d_a, size_t a_size, size_t a_stride,
                   const uint64_t *d_b, size_t b_size, size_t b_stride, uint64_t *d_out, size_t out_stride,
                   size_t batch, void *stream);

/* EvaluatorCuda::switchKeyInplace  src/evaluator_cuda.cu:1163-1378 (+ gSwitchKeyInplaceUtilA–G :1012-1161)
 * ct: [2][L][N] updated in place (the switched polynomial is ADDED into c0, c1); target: [L][N];
 * d_key: one key-switching key [Ltop digits][2][Ltop+1][N] (keys[key_index] of RelinKeys/GaloisKeys).
 * d_target must not alias d_ct. */
int troyb_switch_key(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_target,
                     size_t target_stride, const uint64_t *d_key, size_t batch, void *stream);

/* The same key switch in two halves, for sharding the DIGITS of one key switch over several GPUs:
 *   partial : prod [batch][2][L+1][N] = sum over digits j in [digit_begin, digit_end) of NTT_i(digit j) * key[j][k][i]
 *             (canonical residues)                                          src/evaluator_cuda.cu:1222-1297
 *   finish  : mod-down of prod by the special prime, ADDED into ct [2][L][N]   src/evaluator_cuda.cu:1299-1376
 *             reduce_first != 0: prod holds the SUM of several GPUs' partial products (each < 2^60, at most 8 of them),
 *             which is reduced modulo each prime on the way (CKKS: the special-prime rows in place, the data rows inside the
 *             mod-down tail; BFV / BGV: all rows in place).
 * troyb_switch_key == partial over all digits followed by finish. d_key is the whole key; a GPU that holds only its
 * digits passes a pointer offset by -digit_begin digits (only its digits are read). */
int troyb_switch_key_partial(troyb_context *ctx, int level, const uint64_t *d_target, size_t target_stride, const uint64_t *d_key,
                             int digit_begin, int digit_end, uint64_t *d_prod, size_t batch, void *stream);
int troyb_switch_key_finish(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, uint64_t *d_prod, int reduce_first,
                            size_t batch, void *stream);

/* The digit-sharded key switch with the key products kept ROW-major, rows [L+1 output moduli][batch][2][N], so that the rows a
 * GPU owns are one contiguous block: the partial products are summed with ONE reduce-scatter over NVLink, each GPU reduces and
 * mod-downs only its own rows, and the result is all-gathered (the reference has no multi-GPU path).
 *   partial_rows : rows = sum over digits [digit_begin, digit_end) of NTT_i(digit j) * key[j][k][i]   (:1222-1297)
 *   reduce_rows  : rows [row_begin, row_end) hold sums of several GPUs' canonical products: reduce them modulo their primes
 *   special_row  : sp [batch][2][N] = INTT(row L) (+ half of the special prime): the correction every data limb needs (:1299-1331)
 *   moddown_rows : data limbs [limb_begin, limb_end) of ct [batch][2][L][N] += mod-down of their rows by the special prime (:1333-1376)
 * A GPU that holds only the rows [r0, r1) passes d_rows offset by -r0 rows; only rows it names are touched. */
int troyb_switch_key_partial_rows(troyb_context *ctx, int level, const uint64_t *d_target, size_t target_stride, const uint64_t *d_key,
                                  int digit_begin, int digit_end, uint64_t *d_rows, size_t batch, void *stream);
int troyb_switch_key_reduce_rows(troyb_context *ctx, int level, uint64_t *d_rows, int row_begin, int row_end, size_t batch, void *stream);
int troyb_switch_key_special_row(troyb_context *ctx, int level, const uint64_t *d_rows, uint64_t *d_sp, size_t batch, void *stream);
int troyb_switch_key_moddown_rows(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_rows,
                                  const uint64_t *d_sp, int limb_begin, int limb_end, size_t batch, void *stream);

/* EvaluatorCuda::relinearizeInternal, size 3 → 2  src/evaluator_cuda.cu:703-747: key-switch of poly 2 into polys 0,1 */
int troyb_relinearize(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_relin_key,
                      size_t batch, void *stream);

/* GaloisToolCuda::applyGalois / applyGaloisNtt  src/utils/galois_cuda.cu:166-215 — permutation only, n_polys polys */
int troyb_galois_permute(troyb_context *ctx, int level, const uint64_t *d_in, uint64_t *d_out, size_t n_polys,
                         uint32_t galois_elt, int ntt_form, void *stream);
/* GaloisToolCuda::getEltFromStep  src/utils/galois_cuda.cu:44-62 (steps = 0 → column rotation / conjugation element) */
int troyb_galois_elt_from_step(const troyb_context *ctx, int step, uint32_t *out);

/* EvaluatorCuda::applyGaloisInplace  src/evaluator_cuda.cu:2024-2117: permute c0, c1 by X→X^elt, then key-switch c1.
 * d_key = GaloisKeys key for this element (index (elt-1)/2). */
int troyb_apply_galois(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, uint32_t galois_elt,
                       const uint64_t *d_key, size_t batch, void *stream);

/* EvaluatorCuda::applyGalois(encrypted, elt, keys, destination)  src/evaluator_cuda.cuh:275: the same, reading d_in and writing
 * d_out (which may be d_in); the input is not modified otherwise. */
int troyb_apply_galois_to(troyb_context *ctx, int level, const uint64_t *d_in, size_t in_stride, uint64_t *d_out, size_t out_stride,
                          uint32_t galois_elt, const uint64_t *d_key, size_t batch, void *stream);

/* EvaluatorCuda::modSwitchScaleToNext  src/evaluator_cuda.cu:749-825 (rescaleToNext :1380 for CKKS, modSwitchToNext for
 * BFV/BGV): drops the level's last prime with rounding. in: [size][L][N] at `level`, out: [size][L-1][N] at level+1. */
int troyb_mod_switch_scale_to_next(troyb_context *ctx, int level, const uint64_t *d_in, size_t in_stride, uint64_t *d_out,
                                   size_t out_stride, size_t size, size_t batch, void *stream);

/* ---- ciphertext (x) plaintext: the callers right next to the path ------------------------------ */

/* EvaluatorCuda::multiplyPlainNtt  src/evaluator_cuda.cu:1818-1862 (kDyadicProductCoeffmod per poly, src/kernelutils.cu:151):
 * out[b][k] = ct[b][k] * plain[b] limb-wise, both in NTT form over the level's L limbs; plain is [L][N].
 * plain_stride = 0: one plaintext for the whole batch. d_out may alias d_ct. accumulate != 0: out += ct * plain
 * (the multiplyPlain + addInplace pair of MatmulHelper::matmul, app/LinearHelperCKKS.cuh:225-249, in one pass). */
int troyb_multiply_plain_ntt(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                             const uint64_t *d_plain, size_t plain_stride, uint64_t *d_out, size_t out_stride, size_t batch,
                             int accumulate, void *stream);

/* MatmulHelper::matmul  app/LinearHelperCKKS.cuh:225-249: out[b][j] = sum_i ct[b][i] * plain[i][j] over NTT-form operands
 * (the reference: batch*inputs*outputs multiplyPlain + batch*(inputs-1)*outputs addInplace calls), one launch.
 * d_ct: [batch][inputs] ciphertexts of ct_size polys, ct_stride words apart; d_plain: [inputs][outputs][L][N];
 * d_out: [batch][outputs] ciphertexts, out_stride words apart (must not alias d_ct). */
int troyb_matmul_plain_ntt(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                           const uint64_t *d_plain, uint64_t *d_out, size_t out_stride, size_t batch, size_t inputs, size_t outputs,
                           void *stream);

/* Plaintext lift of BFV / BGV: `batch` coefficient-form plaintexts of plain_coeff_count words modulo t (plain_stride words
 * apart) become [batch][L][N] centred residues over the level's limbs; to_ntt != 0 also transforms them.
 * gMultiplyPlainNormalUtilA/B src/evaluator_cuda.cu:1724-1755; EvaluatorCuda::transformToNttInplace(PlaintextCuda&, ParmsID)
 * src/evaluator_cuda.cu:1865-1935 (to_ntt = 1). */
int troyb_plain_lift(troyb_context *ctx, int level, const uint64_t *d_plain, size_t plain_coeff_count, size_t plain_stride,
                     uint64_t *d_out, size_t batch, int to_ntt, void *stream);

/* EvaluatorCuda::multiplyPlainNormal  src/evaluator_cuda.cu:1757-1816 (CPU src/evaluator.cpp:1790-1935): coefficient-form
 * ciphertexts [ct_size][L][N] times ONE coefficient-form plaintext, in place; canonical output (the CPU path's fully
 * reduced inverse transform). The CPU path's monomial shortcut (src/evaluator.cpp:1817-1850) is not reproduced: like the
 * reference's CUDA path this always takes the generic route. */
int troyb_multiply_plain_normal(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                                const uint64_t *d_plain, size_t plain_coeff_count, size_t batch, void *stream);

/* EvaluatorCuda::addPlainInplace / subPlainInplace for BFV and BGV  src/evaluator_cuda.cu:1555-1700:
 *   BFV: multiplyAdd/SubPlainWithScalingVariant  src/utils/scalingvariant_cuda.cu:82-175
 *   BGV: plain * correction_factor mod t, then add/subPlainWithoutScalingVariant  scalingvariant_cuda.cu:21-80
 * d_c0: the first poly of each ciphertext ([L][N], ct_stride words apart). plain_stride = 0: one plaintext for all.
 * (CKKS adds an NTT-form plaintext with troyb_add / troyb_sub on poly 0.) */
int troyb_add_plain(troyb_context *ctx, int level, uint64_t *d_c0, size_t ct_stride, const uint64_t *d_plain,
                    size_t plain_coeff_count, size_t plain_stride, uint64_t correction_factor, int subtract, size_t batch,
                    void *stream);

/* ---- encryption: the step right before the path --------------------------------------------------
 * The reference's CPU encryptor draws from ONE sequential byte stream per call (Blake2xbPRNG, src/randomgen.cpp:184-194:
 * 4096-byte buffers, buffer c = BLAKE2Xb(key = 64-byte seed, input = c)). These entry points reproduce that stream and the
 * samplers' consumption of it on the device, so that for the same seed the ciphertext words equal the CPU path's
 * (the reference's CUDA encryptor, src/utils/rlwe_cuda.cu:193-333, draws from cuRAND and matches neither).
 * The ternary sampler follows libstdc++'s std::uniform_int_distribution (GCC >= 11), which is what the reference CPU library
 * executes when built with this image's toolchain. Seeds: 8 words (64 bytes) per generator, device memory. */

/* UniformRandomGenerator::generate  src/randomgen.cpp:164-180: the first `bytes` (a multiple of 4096) of the stream of each of
 * `batch` generators; seeds seed_stride words apart, outputs out_stride words apart. */
int troyb_prng_fill(const uint64_t *d_seeds, size_t seed_stride, uint64_t *d_out, size_t out_stride, size_t bytes, size_t batch,
                    void *stream);

/* util::samplePolyTernary / samplePolyCbd / samplePolyUniform  src/utils/rlwe.cpp:21-41, 71-105, 123-158 over the base of
 * `level` (-1: key level): d_out [batch][L][N]. The draws are taken from the expanded stream d_stream (stream_bytes per
 * generator, stream_stride words apart) starting at byte start_bytes (uniform: 0); positions past stream_bytes are computed
 * from d_seeds on demand (d_seeds may be NULL when the expanded part is known to suffice). d_end_bytes (optional, [batch]):
 * the generator's position after the ternary sampler (it moves by 4 bytes per draw, rejected draws included). */
int troyb_sample_poly_ternary(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                              const uint64_t *d_seeds, size_t seed_stride, uint64_t start_bytes, uint64_t *d_out, size_t out_stride,
                              uint64_t *d_end_bytes, size_t batch, void *stream);
int troyb_sample_poly_cbd(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                          const uint64_t *d_seeds, size_t seed_stride, uint64_t start_bytes, uint64_t *d_out, size_t out_stride,
                          size_t batch, void *stream);
int troyb_sample_poly_uniform(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                              const uint64_t *d_seeds, size_t seed_stride, uint64_t *d_out, size_t out_stride, size_t batch,
                              void *stream);

/* util::encryptZeroAsymmetric  src/utils/rlwe.cpp:160-243 as Encryptor::encryptZeroInternal calls it (src/encryptor.cpp:88-164):
 * encryption of zero under the public key at `level` (-1: key level), d_ct [batch][2][L][N]; when the chain has a level above
 * `level` the encryption is done there and divided by that level's last prime (the special prime for level 0).
 * d_public_key: [2][key limbs][N], NTT form at the key level. d_seeds: [batch][8], one generator per ciphertext
 * (parms.randomGenerator()->create()). CKKS results are in NTT form, BFV / BGV in coefficient form. */
int troyb_encrypt_zero_asymmetric(troyb_context *ctx, int level, const uint64_t *d_public_key, const uint64_t *d_seeds,
                                  uint64_t *d_ct, size_t ct_stride, size_t batch, void *stream);

/* util::encryptZeroSymmetric  src/utils/rlwe.cpp:245-356 (save_seed = false): (c0, c1) = (-(a s + e), a) at `level` (-1: key
 * level — KeyGenerator's public key). d_secret_key: [key limbs][N], NTT form at the key level. */
int troyb_encrypt_zero_symmetric(troyb_context *ctx, int level, const uint64_t *d_secret_key, const uint64_t *d_seeds, uint64_t *d_ct,
                                 size_t ct_stride, size_t batch, void *stream);

/* Seeded ciphertexts, the reference CUDA path's transport form of a fresh symmetric encryption (half the bytes): c1 is not
 * stored but re-expanded from a 64-bit seed by the reference's cuRAND sampler (coefficient g: XORWOW curand_init(seed + g, 0, 0),
 * limb j: its j-th curand_uniform_double * q_j, truncated; src/utils/rlwe_cuda.cu:25-28, 152-163).
 *   expand_seeded_poly : what CiphertextCuda::load(stream, context) does for seed != 0  src/ciphertext_cuda.cu:158-175: d_poly
 *                        [batch][L][N] from d_ct_seeds [batch], in the sampled (NTT) domain; for BFV the loader then inverts it.
 *   encrypt_zero_symmetric_seeded : encryptZeroSymmetric of the CUDA path  src/utils/rlwe_cuda.cu:261-333: as
 *                        troyb_encrypt_zero_symmetric, but c1 is the expansion of a seed taken from the ciphertext generator
 *                        (written to d_ct_seeds [batch]); the noise still comes from the host generator's stream. */
int troyb_expand_seeded_poly(troyb_context *ctx, int level, const uint64_t *d_ct_seeds, uint64_t *d_poly, size_t poly_stride,
                             size_t batch, void *stream);
int troyb_encrypt_zero_symmetric_seeded(troyb_context *ctx, int level, const uint64_t *d_secret_key, const uint64_t *d_seeds,
                                        uint64_t *d_ct, size_t ct_stride, uint64_t *d_ct_seeds, size_t batch, void *stream);

/* ---- decryption: the step right after the path ------------------------------------------------- */

/* DecryptorCuda::computeSecretKeyArray  src/decryptor_cuda.cu:188-235: d_out[i] = s^(i+1), i < max_power, each
 * [key_limbs][N] in NTT form at the key level (d_secret = the secret key as KeyGenerator stores it). d_out may start at
 * d_secret. */
int troyb_secret_key_powers(troyb_context *ctx, const uint64_t *d_secret, size_t max_power, uint64_t *d_out, void *stream);

/* DecryptorCuda::dotProductCtSkArray  src/decryptor_cuda.cu:237-300: d_phase[b] = c_0 + c_1 s + ... + c_{size-1} s^{size-1}
 * over the level's limbs, [batch][L][N], in the form of the ciphertext (is_ntt_form = 0: coefficient form, BFV / BGV).
 * For CKKS this IS the decrypted plaintext (ckksDecrypt :118-141). d_secret_powers: at least ct_size-1 powers. */
int troyb_decrypt_dot_product(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride, int is_ntt_form,
                              const uint64_t *d_secret_powers, uint64_t *d_phase, size_t batch, void *stream);

/* RNSToolCuda::decryptScaleAndRound  src/utils/rns_cuda.cu:499-577 (BFV, bfvDecrypt src/decryptor_cuda.cu:92-116):
 * phase [batch][L][N] -> plaintext coefficients modulo t, [batch][N]. */
int troyb_decrypt_scale_and_round(troyb_context *ctx, int level, const uint64_t *d_phase, uint64_t *d_plain, size_t batch,
                                  void *stream);

/* RNSToolCuda::decryptModt  src/utils/rns_cuda.cu:617-620 (exact base conversion q -> t, CPU src/utils/rns.cpp:462-538) and
 * the correction-factor fix of bgvDecrypt  src/decryptor_cuda.cu:143-186: phase [batch][L][N] -> plaintext coefficients
 * modulo t, [batch][N]. inv_correction_factor = correction_factor^-1 mod t (1 when the factor is 1). */
int troyb_decrypt_mod_t(troyb_context *ctx, int level, const uint64_t *d_phase, uint64_t *d_plain, uint64_t inv_correction_factor,
                        size_t batch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TROY_B200_H */
