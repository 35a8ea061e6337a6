/* TEST INFRASTRUCTURE ONLY — see troy_oracle.c. */
#ifndef TROY_ORACLE_H
#define TROY_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define ORC_MAX_LIMBS 80
#define ORC_DROP_BFV 1
#define ORC_DROP_CKKS 2
#define ORC_DROP_BGV 3

int orc_is_prime(uint64_t n);
int orc_get_primes(uint64_t factor, int bit_size, int count, uint64_t *out);
int orc_coeff_modulus_create(uint64_t n, const int *bits, int count, uint64_t *out);
uint64_t orc_minimal_primitive_root(uint64_t degree, uint64_t q);
void orc_ntt_tables(uint64_t n, uint64_t q, uint64_t *root_powers, uint64_t *inv_root_powers);
void orc_ntt_forward(uint64_t n, uint64_t q, uint64_t *a);
void orc_ntt_inverse(uint64_t n, uint64_t q, uint64_t *a);
void orc_dyadic_tensor(uint64_t n, uint64_t q, const uint64_t *a, int a_size, const uint64_t *b, int b_size, uint64_t *out,
                       size_t poly_stride_a, size_t poly_stride_b, size_t poly_stride_out);
void orc_galois_coeff(uint64_t n, uint64_t q, const uint64_t *in, uint32_t elt, uint64_t *out);
void orc_galois_ntt(uint64_t n, const uint64_t *in, uint32_t elt, uint64_t *out);
uint32_t orc_galois_elt_from_step(uint64_t n, int step);
void orc_drop_last(int mode, uint64_t n, int L, const uint64_t *q, uint64_t t, uint64_t *poly);
void orc_switch_key(int scheme, uint64_t n, int L, int Ltop, const uint64_t *key_q, uint64_t t, uint64_t *ct,
                    const uint64_t *target, const uint64_t *key);
int orc_behz_base(uint64_t n, int L, const uint64_t *q, uint64_t t, uint64_t *bsk_out);
int orc_bfv_multiply(uint64_t n, int L, const uint64_t *q, uint64_t t, const uint64_t *a, int a_size, const uint64_t *b,
                     int b_size, uint64_t *out);
void orc_multiply_ntt_schemes(int scheme, uint64_t n, int L, const uint64_t *q, const uint64_t *a, int a_size, const uint64_t *b,
                              int b_size, uint64_t *out);
void orc_apply_galois(int scheme, uint64_t n, int L, int Ltop, const uint64_t *key_q, uint64_t t, uint64_t *ct, uint32_t elt,
                      const uint64_t *key);
#endif
