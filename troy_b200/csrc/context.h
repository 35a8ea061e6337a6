// Context: host precompute of every table the RNS polynomial path needs, uploaded once.
//
// Plays the role of SEALContextCuda / ContextDataCuda / NTTTablesCuda / RNSToolCuda
// (src/context_cuda.cu:5-62, src/utils/ntt_cuda.cuh:15-102, src/utils/rns_cuda.cu:206-269),
// but as ONE flat modulus pool plus per-level constant blocks, laid out for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

namespace troyb {

typedef uint64_t u64;
typedef uint32_t u32;

enum Scheme { SCHEME_BFV = 1, SCHEME_CKKS = 2, SCHEME_BGV = 3 };

constexpr int kMaxLimbs = 80; // data limbs + special + BEHZ auxiliary primes

// One entry per distinct modulus in the pool (device resident).
//   fwd[k] = {w, floor(w*2^64/p)} with w = psi^{bitrev_logN(k)}  (same index rule as the reference's
//            root_powers_, src/utils/ntt.cpp:34-43; psi = minimal primitive 2N-th root)
//   inv[k] = {w^-1, shoup}, indexed like fwd (our own layout; the reference's scrambled
//            inv_root_powers_ order, src/utils/ntt.cpp:46-54, is never API-visible)
//   inv[1] is replaced by (w^-1 * N^-1) and ninv holds N^-1 so the last inverse stage applies the scaling.
struct ModDev {
    u64 p;
    u64 ratio0, ratio1;     // floor(2^128/p)
    u64 bar_mu;             // floor(2^(bar_shift+64)/p): single-word Barrett for short sums of products (modarith.cuh reduce_sum)
    u32 bar_shift;          // bit length of p, minus 1
    u64 ninv, ninv_shoup;   // N^-1 mod p
    u64 inv1n, inv1n_shoup; // fwd[1]^-1 * N^-1 mod p
    const u64 *fwd;         // 2*N words
    const u64 *inv;         // 2*N words
    // FP64 butterfly path (primes below 2^50, see ntt.cuh ArithF64 / ArithF64L): the same tables as doubles {w, w/p}; null otherwise
    const double *fwd_d;
    const double *inv_d;
    double pd, pinv_d;            // p and 1/p
    double ninv_d, ninv_dp;       // N^-1 and N^-1 / p
    double inv1n_d, inv1n_dp;     // fwd[1]^-1 * N^-1 and that / p
    int policy;                   // 0 FP64 (p < 2^46), 1 fast integer (p < 2^60), 2 exact integer (p up to 61 bits), 3 FP64 (p < 2^41),
                                  // 4 fast integer without forward range corrections (p < 2^57), 5 FP64 with per-butterfly re-centring (p < 2^50)
};
enum { POLICY_F64 = 0, POLICY_U64_FAST = 1, POLICY_U64_EXACT = 2, POLICY_F64_SMALL = 3, POLICY_U64_LAZY = 4, POLICY_F64_LARGE = 5 };

// Constants for dividing by the last prime of a base {q_0..q_{m-1}, q_last} (rescale, mod-switch, key-switch mod-down)
//   src/utils/rns.cpp:784-803 (inv_q_last_mod_q_, inv_q_last_mod_t_, q_last_mod_t_)
struct DropLimb {
    u64 inv_last, inv_last_shoup;   // q_last^-1 mod q_i
    u64 half_mod;                   // (q_last >> 1) mod q_i
    u64 last_mod, last_mod_shoup;   // q_last mod q_i (BGV correction)
};

struct DropDev {
    const DropLimb *limbs; // device array, one per remaining limb
    u64 q_last, half;      // q_last, q_last >> 1
    u64 last_ratio1;
    u64 inv_last_mod_t;    // q_last^-1 mod t (BGV), 0 otherwise
    u64 inv_last_mod_t_shoup;
    u64 t, t_ratio1;       // plain modulus
    int last_pool;         // pool index of q_last
};

// BEHZ constants for one level (BFV multiply), src/utils/rns.cpp:581-803. All device pointers.
struct BehzDev {
    int L, nB, nBsk;          // |q|, |B|, |Bsk| = |B|+1
    // q -> Bsk ∪ {m_tilde}
    const u64 *mtilde_inv_punct_q; // [L][2]  Shoup pair of (m_tilde * (Q/q_i)^-1) mod q_i
    const u64 *inv_punct_q;        // [L][2]  Shoup pair of (Q/q_i)^-1 mod q_i
    const u64 *q_to_bsk;           // [nBsk][L]  (Q/q_i) mod Bsk_o
    const u64 *q_to_mtilde;        // [L]        (Q/q_i) mod 2^32
    u64 neg_inv_q_mod_mtilde;      // -(Q^-1) mod 2^32
    const u64 *prod_q_mod_bsk;     // [nBsk][2] Shoup
    const u64 *inv_mtilde_mod_bsk; // [nBsk][2] Shoup
    const u64 *inv_prod_q_mod_bsk; // [nBsk][2] Shoup
    // B -> q, B -> m_sk
    const u64 *inv_punct_B;        // [nB][2]
    const u64 *B_to_q;             // [L][nB]
    const u64 *B_to_msk;           // [nB]
    u64 inv_prod_B_mod_msk, inv_prod_B_mod_msk_shoup;
    const u64 *prod_B_mod_q;       // [L][2] Shoup of prod(B) mod q_i
    const u64 *neg_prod_B_mod_q;   // [L][2] Shoup of q_i - prod(B) mod q_i
    const u64 *t_mod;              // [L+nBsk][2] Shoup pair of plain modulus t in each of q ∪ Bsk
};

// Per-level constants of the plaintext space of BFV / BGV (src/context.cpp:296-352), passed by value with a launch.
struct PlainConsts {
    u64 t, threshold;        // plain modulus, plain_upper_half_threshold = (t + 1) / 2
    u64 q_mod_t;             // coeff_modulus_mod_plain_modulus
    u64 inc[kMaxLimbs];      // plain_upper_half_increment: (Q - t) mod q_i
    u64 delta[kMaxLimbs];    // coeff_div_plain_modulus: floor(Q / t) mod q_i
};

// Constants of RNSToolCuda::decryptScaleAndRound (BFV; gamma = 0 otherwise) and decryptModt (BGV) for one level
// (src/utils/rns.cpp:748-775, 462-538), passed by value with the launch.
struct DecryptConsts {
    u64 t, gamma;
    u64 neg_inv_q_mod_t, neg_inv_q_mod_gamma; // -(Q^-1) mod t, mod gamma
    u64 inv_gamma_mod_t;
    u64 tg_mod_q[kMaxLimbs];        // t * gamma mod q_i
    u64 punct_mod_t[kMaxLimbs];     // (Q / q_i) mod t
    u64 punct_mod_gamma[kMaxLimbs]; // (Q / q_i) mod gamma
    u64 inv_punct_q[kMaxLimbs];     // (Q / q_i)^-1 mod q_i
    u64 q_mod_t;                    // Q mod t (BGV decryptModt)
};

struct LimbMap {
    uint8_t idx[kMaxLimbs]; // pool index of launch limb l
};

struct LevelData {
    int L = 0;                       // number of data limbs at this level
    DropDev drop{};                  // dropping q_{L-1}; valid when L >= 2
    DropLimb *d_drop_limbs = nullptr;
    bool has_behz = false;
    BehzDev behz{};
    std::vector<void *> owned;       // device allocations
    double last_prime_as_double = 0; // for scale bookkeeping
    bool has_decrypt = false;        // BFV / BGV
    DecryptConsts decrypt{};
    bool has_plain = false;          // BFV / BGV
    PlainConsts plain{};
};

struct Context {
    int device = 0;
    int sm_count = 148;           // multiprocessors of `device` (grid and prefetch sizing)
    int scheme = 0;
    u64 N = 0;
    int logN = 0;
    u64 plain = 0;
    int Ltop = 0;                 // data limbs at the first level
    bool has_special = false;     // key level has an extra special prime
    std::vector<u64> pool;        // all distinct moduli
    int pool_special = -1;        // pool index of the special prime
    int pool_msk = -1, pool_gamma = -1, pool_B0 = -1;
    int total_bits_first = 0;
    ModDev *d_mods = nullptr;     // device array [pool.size()]
    std::vector<ModDev> h_mods;   // host mirror (with device table pointers)
    std::vector<void *> owned;
    std::vector<LevelData> levels; // levels[0] = first data level
    DropDev key_drop{};           // dropping the special prime at the key level (mod-down)
    DropLimb *d_key_drop_limbs = nullptr;
    std::string error;
    cudaMemPool_t mempool = nullptr;

    ~Context();
    LimbMap map_prefix(int L) const;                 // q_0..q_{L-1}
    LimbMap map_prefix_special(int L) const;         // q_0..q_{L-1}, q_sp
    LimbMap map_q_bsk(int level) const;              // q_0..q_{L-1}, B_0.., m_sk
    int level_total_bits(int level) const;
};

// Throws std::invalid_argument / std::logic_error on bad parameters.
Context *context_create(int device, int scheme, u64 N, const u64 *moduli, int n_moduli, u64 plain_modulus,
                        bool expand_chain);

} // namespace troyb
