// Host-side number theory used to build the device tables (primes, roots, inverses).
//
// Restates the *functions* of the reference's table generators (results must be identical):
//   util::isPrime / getPrimes                src/utils/numth.cpp:215-284
//   util::tryMinimalPrimitiveRoot            src/utils/numth.cpp:335-360
//   util::tryInvertUintMod                   src/utils/numth.h
//   NTTTables::initialize                    src/utils/ntt.cpp:17-66
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace troyb {
namespace host {

typedef uint64_t u64;
typedef unsigned __int128 u128;

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

inline u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

// Deterministic Miller-Rabin for 64-bit inputs.
inline bool is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : small) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (u64 a : small) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; i++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}

// Primes congruent to 1 mod factor, descending from 2^bit_size (same order as the reference's getPrimes).
inline std::vector<u64> get_primes(u64 factor, int bit_size, size_t count) {
    std::vector<u64> out;
    u64 value = ((u64(1) << bit_size) - 1) / factor * factor + 1;
    u64 lower = u64(1) << (bit_size - 1);
    while (count > 0 && value > lower) {
        if (is_prime(value)) { out.push_back(value); count--; }
        value -= factor;
    }
    if (count > 0) throw std::logic_error("failed to find enough qualifying primes");
    return out;
}

// Modular inverse (extended Euclid); works for composite moduli such as m_tilde = 2^32.
inline bool try_invert(u64 a, u64 q, u64 &out) {
    if (a == 0) return false;
    __int128 t = 0, nt = 1;
    __int128 r = q, nr = a % q;
    while (nr != 0) {
        __int128 qt = r / nr;
        __int128 tmp = t - qt * nt; t = nt; nt = tmp;
        tmp = r - qt * nr; r = nr; nr = tmp;
    }
    if (r != 1) return false;
    if (t < 0) t += q;
    out = (u64)t;
    return true;
}

inline u64 invert(u64 a, u64 q) {
    u64 r;
    if (!try_invert(a, q, r)) throw std::logic_error("value is not invertible");
    return r;
}

// The minimal primitive degree-th root of unity mod q (degree a power of two dividing q-1).
// The reference starts from a random generator and then takes the minimum over all odd powers;
// the minimum is independent of the starting point, so a deterministic search is equivalent.
inline bool minimal_primitive_root(u64 degree, u64 q, u64 &out) {
    if ((q - 1) % degree != 0) return false;
    u64 quot = (q - 1) / degree;
    u64 root = 0;
    for (u64 g = 2; g < 1000; g++) {
        u64 c = powmod(g, quot, q);
        if (powmod(c, degree >> 1, q) == q - 1) { root = c; break; }
    }
    if (!root) return false;
    u64 gen_sq = mulmod(root, root, q);
    u64 cur = root, best = root;
    for (u64 i = 0; i < degree; i += 2) {
        if (cur < best) best = cur;
        cur = mulmod(cur, gen_sq, q);
    }
    out = best;
    return true;
}

inline u64 shoup_quotient(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

inline uint32_t reverse_bits(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

inline int bit_count(u64 v) { return v ? 64 - __builtin_clzll(v) : 0; }

} // namespace host
} // namespace troyb
