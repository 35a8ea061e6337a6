// 64-bit modular arithmetic shared by host precompute (context.cpp) and device kernels.
//
// Reference semantics being restated (not copied):
//   Modulus::const_ratio = floor(2^128 / q)                 src/modulus.cpp:28-37
//   barrettReduce128 / barrettReduce64                      src/utils/uintarithsmallmod.h:95-160
//   MultiplyUIntModOperand {operand, quotient=floor(w*2^64/q)}  src/utils/uintarithsmallmod.h:166-185
//   multiplyUintModLazy (Shoup, result in [0,2q))           src/utils/uintarithsmallmod.h:217
//   device twins dMultiplyUintModLazy / dBarrettReduce128   src/kernelutils.cuh:129,161
// All results that leave a kernel are canonical residues, so only the integer function matters.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#define TB_D __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace troyb {

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

// ---- host/device portable 64x64 -> hi/lo -------------------------------------------------------
TB_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((u128)a * b) >> 64);
#endif
}

// x mod q for any 64-bit x, using ratio1 = high word of floor(2^128/q)  (barrettReduce64)
TB_HD u64 barrett64(u64 x, u64 q, u64 ratio1) {
    u64 t = mulhi64(x, ratio1);
    u64 r = x - t * q;
    return r >= q ? r - q : r;
}

// (hi:lo) mod q, ratio = floor(2^128/q) as (ratio0 low, ratio1 high)  (barrettReduce128)
TB_HD u64 barrett128(u64 lo, u64 hi, u64 q, u64 ratio0, u64 ratio1) {
    // floor(x * ratio / 2^128) with the lowest partial product dropped, like the reference
    u64 carry = mulhi64(lo, ratio0);
    u64 t1lo = lo * ratio1, t1hi = mulhi64(lo, ratio1);
    u64 s0 = t1lo + carry;
    u64 c1 = t1hi + (s0 < t1lo);
    u64 t2lo = hi * ratio0, t2hi = mulhi64(hi, ratio0);
    u64 s1 = s0 + t2lo;
    u64 c2 = t2hi + (s1 < t2lo);
    u64 qhat = hi * ratio1 + c1 + c2;
    u64 r = lo - qhat * q;
    return r >= q ? r - q : r;
}

// (hi:lo) mod q for a value below 2^(n+63), n = bit length of q (e.g. a sum of at most 8 products of residues when
// n <= 60, or of 2 products when n = 61): one-word Barrett with mu = floor(2^(n-1+64)/q), shift = n - 1. The quotient
// estimate is at most 2 too small, so the remainder lies in [0, 3q). Our own reduction (the reference reduces such sums
// with barrettReduce128); the canonical result is the same number.
TB_HD u64 reduce_sum(u64 lo, u64 hi, u64 q, u64 mu, u32 shift) {
    const u64 a = (hi << (64 - shift)) | (lo >> shift); // shift in [1, 60]
    const u64 r0 = lo - mulhi64(a, mu) * q;
    const u64 q2 = q << 1;
    const u64 r1 = r0 - (r0 >= q2 ? q2 : 0);
    return r1 - (r1 >= q ? q : 0);
}

TB_HD u64 mulmod_barrett(u64 a, u64 b, u64 q, u64 ratio0, u64 ratio1) {
    return barrett128(a * b, mulhi64(a, b), q, ratio0, ratio1);
}

// Shoup lazy multiply: x * w mod q in [0, 2q) for any 64-bit x; wq = floor(w * 2^64 / q), w < q.
TB_HD u64 shoup_lazy(u64 x, u64 w, u64 wq, u64 q) {
    u64 t = mulhi64(x, wq);
    return x * w - t * q;
}

TB_HD u64 shoup(u64 x, u64 w, u64 wq, u64 q) {
    u64 r = shoup_lazy(x, w, wq, q);
    return r >= q ? r - q : r;
}

TB_HD u64 addmod(u64 a, u64 b, u64 q) {
    u64 s = a + b;
    return s >= q ? s - q : s;
}
TB_HD u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
TB_HD u64 negmod(u64 a, u64 q) { return a == 0 ? 0 : q - a; }

} // namespace troyb
