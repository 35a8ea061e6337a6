// Negacyclic NTT building blocks for sm_100a.
//
// What is computed (bit-exact function of the reference):
//   forward : kNttNegacyclicHarvey      src/kernelutils.cuh:578-603, CPU src/utils/ntt.cpp:158-187,
//             DWTHandler::transformToRev src/utils/dwthandler.h:88-204   (natural in, bit-reversed out)
//   inverse : kInverseNttNegacyclicHarvey src/kernelutils.cuh:607-632, CPU src/utils/ntt.cpp:189-213,
//             DWTHandler::transformFromRev src/utils/dwthandler.h:215-372 (bit-reversed in, natural out, x N^-1)
// How it is computed is new. Instead of one kernel launch per butterfly layer with one thread per coefficient, a CTA
// keeps a tile of 2^LOGT coefficients in padded shared memory and every thread runs radix-32 (or radix-16) register
// butterflies: 5 (4) layers per shared-memory round trip. A tile is either a whole limb (N <= 16384: one HBM read and
// one HBM write per transform) or a 4096-coefficient row of the four-step split used for N = 32768 / 65536 (a strided
// "column" pass does the first log2(N/4096) layers).
//
// The butterfly arithmetic is a policy (measured in tests/microbench):
//   ArithF64      primes below 2^46 (CKKS 40-bit primes): exact FP64-FMA modular multiply on the otherwise idle FP64
//                 pipe, balanced residues, no range corrections in the forward transform
//   ArithU64Fast  primes below 2^60: Shoup multiply with a 3-product quotient estimate, fused accumulate chains
//   ArithU64Exact primes up to 61 bits (BEHZ auxiliary base): textbook Harvey lazy butterfly
// Every policy returns canonical residues at the tile boundary, so results do not depend on the policy.
//
// Twiddle index rule (same table order as the reference's root_powers_): for a tile whose sub-transform has root base R
// (R = 1 for a whole limb, R = N/T + row for row `row` of the four-step split), layer s of the tile (0-based) and lower
// butterfly element i use table entry (R << s) + (i >> (LOGT - s)).
#pragma once
#include "launch.cuh"
#include <type_traits>

#if defined(TROYB_HOST_EMUL)
#include <cmath>
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
#endif

namespace troyb {

// Shared-memory padding: one extra word per 16 keeps every radix round used below and the linear copy phases
// bank-conflict free for 64-bit accesses (checked exhaustively by tools/banksim.py).
__host__ __device__ constexpr u32 pad_idx(u32 i) { return i + (i >> 4); }
__host__ __device__ constexpr u32 padded_words(u32 n) { return n + (n >> 4); }

// ---- 32-bit multiply helpers (the integer pipe's native width) ---------------------------------------------------------
TB_D u32 lo32(u64 x) { return (u32)x; }
TB_D u32 hi32(u64 x) { return (u32)(x >> 32); }
TB_D u64 mk64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
TB_D u64 mulw(u32 a, u32 b) { return (u64)a * b; }
TB_D u64 madw(u32 a, u32 b, u64 c) { return (u64)a * b + c; }
TB_D u32 madlo(u32 a, u32 b, u32 c) { return a * b + c; }

// ---- arithmetic policies --------------------------------------------------------------------------------------------------
struct TwU64 { u64 w, wq; };

// Textbook Harvey lazy butterflies (forward values in [0,4p), inverse in [0,2p)); p up to 61 bits.
struct ArithU64Exact {
    typedef u64 elem;
    typedef TwU64 tw_t;
    static constexpr bool kReduceEachInvRound = false;
    static constexpr bool kReduceInvMidway = false;
    struct Mod {
        u64 p, p2, ratio1;
        TwU64 ninv, inv1n;
    };
    static TB_D Mod make(const ModDev *md) {
        Mod m;
        m.p = md->p; m.p2 = md->p << 1; m.ratio1 = md->ratio1;
        m.ninv = TwU64{md->ninv, md->ninv_shoup};
        m.inv1n = TwU64{md->inv1n, md->inv1n_shoup};
        return m;
    }
    static TB_D const void *fwd_table(const ModDev *md) { return md->fwd; }
    static TB_D const void *inv_table(const ModDev *md) { return md->inv; }
    static TB_D tw_t ld_tw(const void *tab, u32 k) {
        ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(tab) + k);
        return TwU64{v.x, v.y};
    }
    static TB_D elem from_u64(u64 v, const Mod &) { return v; }         // forward accepts [0,4p), inverse [0,2p)
    static TB_D elem from_bits(u64 v) { return v; }
    static TB_D u64 to_bits(elem x) { return x; }
    static TB_D u64 canon(elem x, const Mod &m) { return barrett64(x, m.p, m.ratio1); }
    static TB_D u64 canon_reduced(elem x, const Mod &m) { return x - (x >= m.p ? m.p : 0); } // after gs_last: [0,2p)
    static TB_D void reduce(elem &, const Mod &) {}
    static TB_D void ct(elem &X, elem &Y, tw_t w, const Mod &m, bool = false) {
        u64 x = X - (X >= m.p2 ? m.p2 : 0);
        u64 t = shoup_lazy(Y, w.w, w.wq, m.p);
        X = x + t;
        Y = x - t + m.p2;
    }
    static TB_D void gs(elem &X, elem &Y, tw_t w, const Mod &m) {
        u64 s = X + Y;
        u64 d = X + m.p2 - Y;
        X = s - (s >= m.p2 ? m.p2 : 0);
        Y = shoup_lazy(d, w.w, w.wq, m.p);
    }
    static TB_D void gs_last(elem &X, elem &Y, const Mod &m) {
        u64 s = X + Y;
        u64 d = X + m.p2 - Y;
        X = shoup_lazy(s, m.ninv.w, m.ninv.wq, m.p);
        Y = shoup_lazy(d, m.inv1n.w, m.inv1n.wq, m.p);
    }
};

// Fast integer path for p < 2^60. The quotient of the Shoup multiply is estimated from three 32x32->64 products
// (error <= 2, so the lazy product lies in [0,4p)); the remainder is one fused multiply-accumulate chain modulo 2^64.
// Invariants: forward values < 8p + 2^32 (range correction compares only the high word), inverse values < 4p.
struct ArithU64Fast {
    typedef u64 elem;
    typedef TwU64 tw_t;
    static constexpr bool kReduceEachInvRound = false;
    static constexpr bool kReduceInvMidway = false;
    struct Mod {
        u64 p, p4, ratio1;
        u32 np0, np1, H;
        TwU64 ninv, inv1n;
    };
    static TB_D Mod make(const ModDev *md) {
        Mod m;
        m.p = md->p; m.p4 = md->p << 2; m.ratio1 = md->ratio1;
        u64 np = 0 - md->p;
        m.np0 = lo32(np); m.np1 = hi32(np);
        m.H = hi32(m.p4) + 1;
        m.ninv = TwU64{md->ninv, md->ninv_shoup};
        m.inv1n = TwU64{md->inv1n, md->inv1n_shoup};
        return m;
    }
    static TB_D const void *fwd_table(const ModDev *md) { return md->fwd; }
    static TB_D const void *inv_table(const ModDev *md) { return md->inv; }
    static TB_D tw_t ld_tw(const void *tab, u32 k) { return ArithU64Exact::ld_tw(tab, k); }
    static TB_D elem from_u64(u64 v, const Mod &) { return v; }         // needs v < 4p
    static TB_D elem from_bits(u64 v) { return v; }
    static TB_D u64 to_bits(elem x) { return x; }
    static TB_D u64 canon(elem x, const Mod &m) { return barrett64(x, m.p, m.ratio1); }
    static TB_D u64 canon_reduced(elem x, const Mod &m) { // after gs_last: [0,3p)
        const u64 p2 = m.p << 1;
        x -= (x >= p2 ? p2 : 0);
        return x - (x >= m.p ? m.p : 0);
    }
    static TB_D void reduce(elem &, const Mod &) {}
    // add + Y*w - q*p (mod 2^64) with q in [Q-2, Q], Q = floor(Y*wq / 2^64); equals add + (Y*w mod p) + {0,1,2}*p
    static TB_D u64 mul_acc(u64 add, u64 Y, tw_t t, const Mod &m) {
        u32 y0 = lo32(Y), y1 = hi32(Y), w0 = lo32(t.w), w1 = hi32(t.w), v0 = lo32(t.wq), v1 = hi32(t.wq);
        u64 A = mulw(y1, v0), B = mulw(y0, v1);
        u64 q = madw(y1, v1, (u64)hi32(A) + (u64)hi32(B));
        u32 q0 = lo32(q), q1 = hi32(q);
        u64 r = madw(y0, w0, add);
        r = madw(q0, m.np0, r);
        u32 rh = hi32(r);
        rh = madlo(y1, w0, rh);
        rh = madlo(y0, w1, rh);
        rh = madlo(q1, m.np0, rh);
        rh = madlo(q0, m.np1, rh);
        return mk64(lo32(r), rh);
    }
    static TB_D void ct(elem &X, elem &Y, tw_t w, const Mod &m, bool = false) {
        u64 x = X;
        if (hi32(x) >= m.H) x -= m.p4;          // x < 4p + 2^32
        u64 xn = mul_acc(x, Y, w, m);           // x + T, T in [0,4p)
        X = xn;
        Y = (x + m.p4) + (x - xn);              // x + 4p - T
    }
    static TB_D void gs(elem &X, elem &Y, tw_t w, const Mod &m) {
        u64 s = X + Y;
        u64 d = X + m.p4 - Y;
        u64 t = s - m.p4;
        X = s >= m.p4 ? t : s;
        Y = mul_acc(0, d, w, m);
    }
    static TB_D void gs_last(elem &X, elem &Y, const Mod &m) {
        u64 s = X + Y;
        u64 d = X + m.p4 - Y;
        X = mul_acc(0, s, m.ninv, m);
        Y = mul_acc(0, d, m.inv1n, m);
    }
};

// The fast integer path for p < 2^57 (e.g. the 50-bit primes of the N = 32768 / 65536 parameter sets): a forward
// butterfly adds less than 4p to the magnitude of its inputs, so log2 N <= 16 layers from a start below 4p stay below
// 68p < 2^64 and the forward transform needs NO range correction (5 of its 27 instructions per butterfly); the store's
// Barrett reduction accepts any 64-bit value. The inverse transform is ArithU64Fast's.
struct ArithU64Lazy : ArithU64Fast {
    static TB_D void ct(elem &X, elem &Y, tw_t w, const Mod &m, bool = false) {
        const u64 x = X;
        const u64 xn = mul_acc(x, Y, w, m);     // x + T, T in [0,4p)
        X = xn;
        Y = (x + m.p4) + (x - xn);              // x + 4p - T
    }
};

// FP64 path for p < 2^50: residues are integers held exactly in doubles (|x| < 2^51 throughout). The product y*w is formed
// exactly as h + l with one multiply and one FMA, the quotient q is an estimate of y*w/p rounded to the nearest integer with
// the 1.5*2^52 trick, and h - q*p + l is exact, so the result is congruent to y*w with |r| <= (1/2 + e) p, e = the error of
// the estimate before rounding. Two ways to estimate (the template parameter):
//   COMPACT (p < 2^46): q ~ h * (1/p). Three roundings: e <= 3 * 2^-53 * |y| w / p < 3 * 2^-53 * |y|, i.e. |r| < 0.875 p for
//           |y| <= 2^50 and |r| < p for |y| < 2^50.4. The twiddle is ONE double per table entry (8 bytes): half the
//           registers, half the L2 -> SM bytes and half the load-store-unit wavefronts of the {w, w/p} pairs (the late layers
//           of a transform use one twiddle per butterfly, so twiddle bytes match the data bytes), and the same eight FP64
//           instructions per butterfly.
//   pairs   (p < 2^50): q ~ y * (w/p) with w/p precomputed per twiddle. Two roundings: e <= 2^-52 * |y|; needed where there is
//           no head-room (ArithF64L).
// Forward butterflies need no range correction (growth < 0.875 p per layer: 16 layers from a start below 4p stay under
// 18p < 2^51 for p < 2^46); the inverse doubles the sum path each layer, so every inverse round first re-centres its inputs.
struct TwF64 { double w, wp; };
struct TwF64c { double w; };
template <bool COMPACT> struct ArithF64T {
    typedef double elem;
    typedef typename std::conditional<COMPACT, TwF64c, TwF64>::type tw_t;
    static constexpr bool kReduceEachInvRound = true;
    static constexpr bool kReduceInvMidway = false;
    struct Mod {
        double p, pinv;
        tw_t ninv, inv1n;
    };
    static TB_D tw_t mk_tw(double w, double wp) {
        if constexpr (COMPACT) return TwF64c{w};
        else return TwF64{w, wp};
    }
    static TB_D Mod make(const ModDev *md) {
        Mod m;
        m.p = md->pd; m.pinv = md->pinv_d;
        m.ninv = mk_tw(md->ninv_d, md->ninv_dp);
        m.inv1n = mk_tw(md->inv1n_d, md->inv1n_dp);
        return m;
    }
    static TB_D const void *fwd_table(const ModDev *md) { return md->fwd_d; }
    static TB_D const void *inv_table(const ModDev *md) { return md->inv_d; }
    static TB_D tw_t ld_tw(const void *tab, u32 k) {
        if constexpr (COMPACT) {
            return TwF64c{__ldg(reinterpret_cast<const double *>(tab) + k)};
        } else {
            ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(tab) + k);
            return TwF64{__longlong_as_double((long long)v.x), __longlong_as_double((long long)v.y)};
        }
    }
    static constexpr double kMagic = 6755399441055744.0; // 1.5 * 2^52
    static constexpr double kTwo52 = 4503599627370496.0;
    static TB_D elem from_u64(u64 v, const Mod &) { // exact for v < 2^52
        return __dadd_rn(__longlong_as_double((long long)(v | 0x4330000000000000ull)), -kTwo52);
    }
    static TB_D elem from_bits(u64 v) { return __longlong_as_double((long long)v); }
    static TB_D u64 to_bits(elem x) { return (u64)__double_as_longlong(x); }
    // rint(a * b), the quotient estimate. Two forms: the 1.5 * 2^52 trick (one FMA and one add, both on the FP64 pipe), or a
    // multiply and FRND.F64 (TROYB_F64_FRND: the rounding leaves the FP64 pipe; the product is rounded once more before the
    // rounding to an integer, 2^-12 at most for quotients below 2^41, inside the error bounds stated above).
    static TB_D double rint_mul(double a, double b) {
#if defined(TROYB_F64_FRND) && !defined(TROYB_HOST_EMUL)
        double r;
        asm("cvt.rni.f64.f64 %0, %1;" : "=d"(r) : "d"(__dmul_rn(a, b)));
        return r;
#else
        return __dadd_rn(__fma_rn(a, b, kMagic), -kMagic);
#endif
    }
    static TB_D void reduce(elem &x, const Mod &m) {
        double q = rint_mul(x, m.pinv);
        x = __fma_rn(-q, m.p, x);
    }
    static TB_D u64 canon_reduced(elem x, const Mod &m) { // |x| < p (a mulmod result)
        if (x < 0.0) x = __dadd_rn(x, m.p);
        return (u64)__double_as_longlong(__dadd_rn(x, kTwo52)) & 0x000FFFFFFFFFFFFFull;
    }
    static TB_D u64 canon(elem x, const Mod &m) {
        reduce(x, m);
        return canon_reduced(x, m);
    }
    static TB_D double mulmod(double y, tw_t t, const Mod &m) {
        double h = __dmul_rn(y, t.w);
        double l = __fma_rn(y, t.w, -h);
        double q;
        if constexpr (COMPACT) q = rint_mul(h, m.pinv);
        else q = rint_mul(y, t.wp);
        double r = __fma_rn(-q, m.p, h);
        return __dadd_rn(r, l);
    }
    static TB_D void ct(elem &X, elem &Y, tw_t w, const Mod &m, bool = false) {
        double t = mulmod(Y, w, m);
        double x = X;
        X = __dadd_rn(x, t);
        Y = __dadd_rn(x, -t);
    }
    static TB_D void gs(elem &X, elem &Y, tw_t w, const Mod &m) {
        double s = __dadd_rn(X, Y);
        double d = __dadd_rn(X, -Y);
        X = s;
        Y = mulmod(d, w, m);
    }
    static TB_D void gs_last(elem &X, elem &Y, const Mod &m) {
        double s = __dadd_rn(X, Y);
        double d = __dadd_rn(X, -Y);
        X = mulmod(s, m.ninv, m);
        Y = mulmod(d, m.inv1n, m);
    }
};
typedef ArithF64T<true> ArithF64;

// The same arithmetic for p < 2^41 (the usual 40-bit CKKS scale primes). A mulmod accepts |y| < 2^51 and returns
// |r| < 1.25 p, and an unreduced sum path doubles per layer, so values stay below 320 p < 2^49.4 for 9 layers after a
// reduction (where the estimate's error is below 0.25, so the last layer's outputs are below p as canon_reduced needs): the
// inverse transform re-centres only once, at the start of its second-to-last tile round, instead of in every round.
struct ArithF64S : ArithF64 {
    static constexpr bool kReduceEachInvRound = false;
    static constexpr bool kReduceInvMidway = true;
};

// The same arithmetic for 2^46 <= p < 2^50 (the 50-bit primes of the N = 32768 / 65536 parameter sets). The multiply stays
// exact for |y| <= 2^51 (y w < 2^101: h and l = y w - h are exact, |l| <= 2^47; the quotient estimate is off by at most
// 0.75, so h - q p is an integer below 2^50 and the FMA returns it exactly) and returns |r| <= 0.75 p. There is no head-room
// for unreduced growth, so every butterfly re-centres the operand that is not multiplied (|x| <= p/2 after `reduce`):
// forward outputs are |x +- t| <= 1.25 p < 2^51 (2 p after the one layer per register network that skips it), inverse inputs
// stay below p (sum re-centred, difference multiplied), 10-11 instead of 8 FP64 instructions per butterfly - still fewer issue cycles than the 23-27 instructions of the integer paths.
struct ArithF64L : ArithF64T<false> {
    static constexpr bool kReduceEachInvRound = false;
    static constexpr bool kReduceInvMidway = false;
    static TB_D elem from_u64(u64 v, const Mod &m) { // v < 2^52
        double x = ArithF64T<false>::from_u64(v, m);
        reduce(x, m);
        return x;
    }
    // odd_layer: the second layer of a register network. Its inputs come from a re-centring layer (|.| <= 1.25 p), so it may
    // skip its own: |x +- t| <= 1.25 p + 0.75 p = 2 p < 2^51, which the next layer (always a re-centring one: the third of
    // the network or the first of the next round) accepts both as x and as y.
    static TB_D void ct(elem &X, elem &Y, tw_t w, const Mod &m, bool odd_layer = false) {
        double t = mulmod(Y, w, m);
        double x = X;
        if (!odd_layer) reduce(x, m);
        X = __dadd_rn(x, t);
        Y = __dadd_rn(x, -t);
    }
    static TB_D void gs(elem &X, elem &Y, tw_t w, const Mod &m) {
        double s = __dadd_rn(X, Y);
        double d = __dadd_rn(X, -Y);
        reduce(s, m);
        X = s;
        Y = mulmod(d, w, m);
    }
};

// One output of a forward butterfly (upper = the x - w y side). Used where a CTA recomputes a layer that couples two tiles and
// keeps only its own tile's output (kernels_ntt.cuh, solo tiles).
template <class A> TB_D typename A::elem ct_half(typename A::elem X, typename A::elem Y, typename A::tw_t w, const typename A::Mod &m, bool upper) {
    if constexpr (std::is_same<typename A::elem, double>::value && !std::is_same<A, ArithF64L>::value) {
        // sign flip on the integer pipe: (-t) = t with bit 63 toggled
        const u64 t = A::to_bits(A::mulmod(Y, w, m)) ^ ((u64)upper << 63);
        return __dadd_rn(X, A::from_bits(t));
    } else {
        A::ct(X, Y, w, m);
        return upper ? Y : X;
    }
}

// A loader of a round's inputs may be "staged": raw(i) issues the bare global loads of input i, cook(raw[], x[]) turns the K
// raw inputs of a group into elements. tile_round then requests every input of a group before any of them is used, and the
// loader's mode branches run once per group instead of between the loads.
template <class Ld, class = void> struct IsStagedLd : std::false_type {};
template <class Ld> struct IsStagedLd<Ld, decltype((void)sizeof(typename Ld::Raw))> : std::true_type {};

template <class A, int K, bool PIN, int BLO, class Ld> TB_D void load_group(const Ld &ld, u32 ib, typename A::elem (&x)[K]) {
    if constexpr (IsStagedLd<Ld>::value) {
        typename Ld::Raw r[K];
#pragma unroll
        for (int k = 0; k < K; k++) r[k] = ld.raw(ib + (PIN ? pad_idx((u32)k << BLO) : ((u32)k << BLO)));
        ld.cook(r, x);
    } else {
#pragma unroll
        for (int k = 0; k < K; k++) x[k] = ld(ib + (PIN ? pad_idx((u32)k << BLO) : ((u32)k << BLO)));
    }
}

// ---- register butterfly networks --------------------------------------------------------------------------------------------
// 2^RR coefficients held by one thread; w[(1 << u) - 1 + j] is the twiddle of block j of layer u of the network.
template <class A, int RR> TB_D void radix_fwd(typename A::elem (&x)[1 << RR], const typename A::tw_t (&w)[(1 << RR) - 1 + (RR == 0)],
                                               const typename A::Mod &m) {
#pragma unroll
    for (int u = 0; u < RR; u++) {
        const int half = 1 << (RR - 1 - u);
#pragma unroll
        for (int j = 0; j < (1 << u); j++) {
#pragma unroll
            for (int r = 0; r < half; r++) {
                const int k = (j << (RR - u)) | r;
                A::ct(x[k], x[k + half], w[(1 << u) - 1 + j], m, (u & 1) != 0);
            }
        }
    }
}

// LAST: layer 0 of the whole transform is layer 0 of this network and carries the N^-1 scaling (its twiddle is unused).
template <class A, int RR, bool LAST> TB_D void radix_inv(typename A::elem (&x)[1 << RR],
                                                          const typename A::tw_t (&w)[(1 << RR) - 1 + (RR == 0)],
                                                          const typename A::Mod &m) {
#pragma unroll
    for (int u = RR - 1; u >= 0; u--) {
        const int half = 1 << (RR - 1 - u);
        if (LAST && u == 0) {
#pragma unroll
            for (int r = 0; r < half; r++) A::gs_last(x[r], x[r + half], m);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << u); j++) {
#pragma unroll
                for (int r = 0; r < half; r++) {
                    const int k = (j << (RR - u)) | r;
                    A::gs(x[k], x[k + half], w[(1 << u) - 1 + j], m);
                }
            }
        }
    }
}

// Which group a thread slot handles. Slots are consecutive across lanes; for rounds whose groups are less than 16
// coefficients apart (0 < BLO < 4) consecutive groups would collide in shared-memory banks, so the lane bits [BLO,4) are
// exchanged with the group bits that the padding maps to the missing banks (tools/banksim.py checks every plan).
template <int BLO, int RR> TB_D u32 group_of_slot(u32 w) {
    if constexpr (BLO > 0 && BLO < 4) {
        static_assert(BLO >= RR, "round plan needs a lane rotation that is not implemented");
        constexpr u32 W = 4 - BLO, A0 = BLO, B0 = BLO + 4 - RR;
        constexpr u32 MA = ((1u << W) - 1) << A0, MB = ((1u << W) - 1) << B0;
        return (w & ~(MA | MB)) | ((w & MA) << (B0 - A0)) | ((w & MB) >> (B0 - A0));
    } else {
        return w;
    }
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// Geometry of one radix-2^RR round on a tile of 2^LOGT coefficients, layers S0..S0+RR-1, run by NT cooperating
// threads. The tile splits into NV = 2^(LOGT-RR) independent groups of K = 2^RR coefficients spaced 2^BLO apart;
// group v = (hi, lo) has base index (hi << (BLO+RR)) | lo and uses the twiddles of block hi. A thread takes G groups,
// SH of them at a time from the same block so that one set of 2^RR-1 twiddles serves SH groups:
//   SHARED (SH = G) : a block holds at least 16 G groups;
//   PAIRED (SH = 2) : BLO = 4 and RR = 3: lanes take lo bits 0-2 and hi bit 0 (the padding maps index bit 7 to the bank
//                     bit that lo bit 3 would have supplied), the two groups of a turn differ in lo bit 3;
//   else   (SH = 1) : thread slots are consecutive groups (with the lane rotation of group_of_slot when 0 < BLO < 4);
//                     one group per turn, twiddles requested AHEAD turns early.
template <int LOGT, int NT, int S0, int RR> struct RoundGeom {
    static constexpr int K = 1 << RR, BLO = LOGT - S0 - RR, NV = (1 << LOGT) / K, G = NV / NT, LOGG = ilog2(G);
    static_assert(BLO >= 0 && G >= 1 && NV % NT == 0 && (1 << LOGG) == G, "bad round");
    static constexpr bool SHARED = BLO >= LOGG + 4;
    static constexpr bool PAIRED = !SHARED && BLO == 4 && RR == 3 && G % 2 == 0;
    static constexpr int SH = SHARED ? G : (PAIRED ? 2 : 1);
    static constexpr int TL = SHARED ? BLO - LOGG : 0;
    static constexpr int U = 1;                                          // independent groups in flight (SH == 1)
    // twiddle sets requested ahead of the turn that uses them (radix-4 turns are short: two turns hide an L2 access)
    static constexpr int AHEAD = (!SHARED && RR <= 2 && G >= 4) ? 2 : 0;
    static constexpr int TURNS = G / (SH * U);
    // group handled by thread vt: turn `turn`, in-flight slot q (< U), shared-twiddle member e (< SH)
    static TB_D void locate(u32 vt, int turn, int q, int e, u32 &hi, u32 &base) {
        if constexpr (SHARED) {
            hi = vt >> TL;
            base = (hi << (BLO + RR)) | ((u32)e << TL) | (vt & ((1u << TL) - 1));
        } else if constexpr (PAIRED) {
            hi = ((vt >> 3) & ((NT >> 3) - 1)) | ((u32)turn * (NT >> 3));
            base = (hi << (BLO + RR)) | ((u32)e << 3) | (vt & 7);
        } else {
            const u32 v = group_of_slot<BLO, RR>(vt + (u32)(turn * U + q) * NT);
            hi = v >> BLO;
            base = (hi << (BLO + RR)) | (v & ((1u << BLO) - 1));
        }
    }
};

template <class A, int RR, int S0> TB_D void load_twiddles(typename A::tw_t (&w)[(1 << RR) - 1], const void *__restrict__ tw, u32 R, u32 hi) {
#pragma unroll
    for (int u = 0; u < RR; u++) {
        const u32 tbase = (R << (S0 + u)) + (hi << u);
#pragma unroll
        for (int j = 0; j < (1 << u); j++) w[(1 << u) - 1 + j] = A::ld_tw(tw, tbase + j);
    }
}

// The twiddles of the groups a thread has in flight in one turn of a round. Loading them is split from the round
// itself so that the loads can be issued before the barrier that precedes the round and, inside a round of several
// turns, one turn ahead: their L2 latency then overlaps the barrier wait / the previous turn's butterflies.
template <class A, int LOGT, int NT, int S0, int RR> struct RoundTw {
    typedef RoundGeom<LOGT, NT, S0, RR> Q;
    static constexpr int NW = Q::SHARED ? 1 : Q::U;
    typename A::tw_t w[NW][(1 << RR) - 1];
    TB_D void load(u32 vt, int turn, const void *__restrict__ tab, u32 R) {
#pragma unroll
        for (int q = 0; q < NW; q++) {
            u32 hi, base;
            Q::locate(vt, turn, q, 0, hi, base);
            load_twiddles<A, RR, S0>(w[q], tab, R, hi);
        }
    }
};

// One round of the forward (INVERSE = false) or inverse transform over the tile.
//   PIN / POUT: the source / destination is the padded shared-memory tile (index = pad_idx(i)); the padded offset of
//   coefficient k of a group is a compile-time constant because the group base and k << BLO occupy disjoint bits.
//   LAST   : (inverse) layer 0 of the whole transform is layer S0 = 0 of this round and carries the N^-1 scaling
//   REDUCE : (inverse) re-centre the inputs first (FP64 policies)
//   PRE    : the twiddles of turn 0 were loaded by the caller (*pre), see RoundTw
template <class A, int LOGT, int NT, int S0, int RR, bool INVERSE, bool LAST, bool REDUCE, bool PIN, bool POUT, bool PRE = false, class Ld, class St>
TB_D void tile_round(u32 vt, Ld ld, St st, const void *__restrict__ tw, u32 R, const typename A::Mod &m,
                     const RoundTw<A, LOGT, NT, S0, RR> *pre = nullptr) {
    typedef RoundGeom<LOGT, NT, S0, RR> Q;
    typedef RoundTw<A, LOGT, NT, S0, RR> W;
    constexpr int K = Q::K, BLO = Q::BLO, U = Q::U, SH = Q::SH;
    W cur;
    if constexpr (PRE) cur = *pre;
    else cur.load(vt, 0, tw, R);
    if constexpr (Q::SHARED) {
        // all G groups of the thread share one twiddle set; two groups in flight at a time
        u32 hi, base0;
        Q::locate(vt, 0, 0, 0, hi, base0);
        // rolled: a fully unrolled transform is ~55 KB of code per policy executed once per warp, and the instruction cache
        // (32 KB) then stalls the warps of the other resident CTAs (ncu: stall_no_instruction)
#pragma unroll 1
        for (int e = 0; e < SH; e++) {
            const u32 base = base0 | ((u32)e << Q::TL);
            const u32 ib = PIN ? pad_idx(base) : base, ob = POUT ? pad_idx(base) : base;
            typename A::elem x[K];
            load_group<A, K, PIN, BLO>(ld, ib, x);
            if (REDUCE) {
#pragma unroll
                for (int k = 0; k < K; k++) A::reduce(x[k], m);
            }
            if constexpr (INVERSE) radix_inv<A, RR, LAST>(x, cur.w[0], m);
            else radix_fwd<A, RR>(x, cur.w[0], m);
#pragma unroll
            for (int k = 0; k < K; k++) st(ob + (POUT ? pad_idx((u32)k << BLO) : ((u32)k << BLO)), x[k]);
        }
    } else {
        constexpr int F = U * SH; // groups in flight
        // one turn: the F groups of `turn` with the twiddles wt
        auto do_turn = [&](int turn, const W &wt) {
            u32 ib[F], ob[F];
            typename A::elem x[F][K];
#pragma unroll
            for (int q = 0; q < U; q++) {
#pragma unroll
                for (int e = 0; e < SH; e++) {
                    u32 hi, base;
                    Q::locate(vt, turn, q, e, hi, base);
                    ib[q * SH + e] = PIN ? pad_idx(base) : base;
                    ob[q * SH + e] = POUT ? pad_idx(base) : base;
                }
            }
#pragma unroll
            for (int f = 0; f < F; f++) load_group<A, K, PIN, BLO>(ld, ib[f], x[f]);
#pragma unroll
            for (int f = 0; f < F; f++) {
                if (REDUCE) {
#pragma unroll
                    for (int k = 0; k < K; k++) A::reduce(x[f][k], m);
                }
                if constexpr (INVERSE) radix_inv<A, RR, LAST>(x[f], wt.w[f / SH], m);
                else radix_fwd<A, RR>(x[f], wt.w[f / SH], m);
#pragma unroll
                for (int k = 0; k < K; k++) st(ob[f] + (POUT ? pad_idx((u32)k << BLO) : ((u32)k << BLO)), x[f][k]);
            }
        };
        if constexpr (Q::AHEAD > 0 && Q::TURNS % 2 == 0) {
            // software pipeline of depth one over two twiddle sets, unrolled by two turns (a rolled loop keeps the code small)
            W wa = cur, wb;
#pragma unroll 1
            for (int turn = 0; turn < Q::TURNS; turn += 2) {
                wb.load(vt, turn + 1, tw, R);
                do_turn(turn, wa);
                if (turn + 2 < Q::TURNS) wa.load(vt, turn + 2, tw, R);
                do_turn(turn + 1, wb);
            }
        } else {
            W wa = cur;
#pragma unroll 1
            for (int turn = 0; turn < Q::TURNS; turn++) {
                if (turn > 0) wa.load(vt, turn, tw, R);
                do_turn(turn, wa);
            }
        }
    }
}

// ---- tile plans ------------------------------------------------------------------------------------------------------------
// A tile is 2^LOGT coefficients owned by one CTA: 4096 (N = 4096) or 8192 (N = 8192, and N = 16384 / 32768 / 65536 as a
// thread-block cluster of 2 / 4 / 8 CTAs, see kernels_ntt.cuh). Radix-8 / radix-4 rounds keep the groups in flight plus
// their twiddles under 80 registers, so three 8192-tiles (or six 4096-tiles) are resident per SM and their load,
// butterfly and store phases overlap. Every listed round is bank-conflict free with pad_idx and RoundGeom::locate
// (tools/banksim.py).
template <int LOGT> struct TilePlan;
template <> struct TilePlan<12> { static constexpr int NT = 128, CTAS = 6, R0 = 3, R1 = 3, R2 = 3, R3 = 3, R4 = 0; };
template <> struct TilePlan<13> { static constexpr int NT = 256, CTAS = 3, R0 = 3, R1 = 3, R2 = 3, R3 = 2, R4 = 2; };
template <int LOGT> __host__ __device__ constexpr int tile_threads() { return TilePlan<LOGT>::NT; }

template <class A> struct SmLd { // index is already padded
    const u64 *sm;
    TB_D typename A::elem operator()(u32 i) const { return A::from_bits(sm[i]); }
};
template <class A> struct SmSt {
    u64 *sm;
    TB_D void operator()(u32 i, typename A::elem v) const { sm[i] = A::to_bits(v); }
};

// The tile's own 2^LOGT-point forward transform, in place in shared memory (layers 0..LOGT-1 with root base R).
//   GLOBAL_IN : the first round reads coefficient i of the tile as gld(i) (global memory, prologue fused there).
//               Otherwise the tile holds element bits left by the cluster's scatter stage.
// The results stay in shared memory as element bits.
// On the GPU the twiddles of round r+1 are requested before the barrier that ends round r (RoundTw).
struct NoHook { TB_D void operator()() const {} };
//   loaded()  : called by every thread once the CTA's global inputs have all arrived (after the barrier of the first round)
template <class A, int LOGT, bool GLOBAL_IN, class Gld, class Hook = NoHook>
TB_D void tile_ntt_fwd(u64 *sm, Gld gld, const ModDev *md, u32 R, Hook loaded = Hook()) {
    typedef TilePlan<LOGT> P;
    constexpr int NT = P::NT;
    const typename A::Mod m = A::make(md);
    const void *tw = A::fwd_table(md);
    SmLd<A> sl{sm};
    SmSt<A> ss{sm};
    constexpr int S1 = P::R0, S2 = S1 + P::R1, S3 = S2 + P::R2, S4 = S3 + P::R3;
#if defined(TROYB_HOST_EMUL)
    if constexpr (GLOBAL_IN && IsStagedLd<Gld>::value) {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, 0, P::R0, false, false, false, false, true>(t, gld, ss, tw, R, m));)
    } else if constexpr (GLOBAL_IN) {
        auto rl = [&](u32 i) -> typename A::elem { return A::from_u64(gld(i), m); };
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, 0, P::R0, false, false, false, false, true>(t, rl, ss, tw, R, m));)
    } else {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, 0, P::R0, false, false, false, true, true>(t, sl, ss, tw, R, m));)
    }
    loaded();
    TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S1, P::R1, false, false, false, true, true>(t, sl, ss, tw, R, m));)
    TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S2, P::R2, false, false, false, true, true>(t, sl, ss, tw, R, m));)
    if constexpr (P::R3 > 0) {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S3, P::R3, false, false, false, true, true>(t, sl, ss, tw, R, m));)
    }
    if constexpr (P::R4 > 0) {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S4, P::R4, false, false, false, true, true>(t, sl, ss, tw, R, m));)
    }
#else
    const u32 t = threadIdx.x;
    RoundTw<A, LOGT, NT, 0, P::R0> w0;
    w0.load(t, 0, tw, R);
    if constexpr (GLOBAL_IN && IsStagedLd<Gld>::value) {
        tile_round<A, LOGT, NT, 0, P::R0, false, false, false, false, true, true>(t, gld, ss, tw, R, m, &w0);
    } else if constexpr (GLOBAL_IN) {
        auto rl = [&](u32 i) -> typename A::elem { return A::from_u64(gld(i), m); };
        tile_round<A, LOGT, NT, 0, P::R0, false, false, false, false, true, true>(t, rl, ss, tw, R, m, &w0);
    } else {
        tile_round<A, LOGT, NT, 0, P::R0, false, false, false, true, true, true>(t, sl, ss, tw, R, m, &w0);
    }
    RoundTw<A, LOGT, NT, S1, P::R1> w1;
    w1.load(t, 0, tw, R);
    __syncthreads();
    loaded();
    tile_round<A, LOGT, NT, S1, P::R1, false, false, false, true, true, true>(t, sl, ss, tw, R, m, &w1);
    RoundTw<A, LOGT, NT, S2, P::R2> w2;
    w2.load(t, 0, tw, R);
    __syncthreads();
    tile_round<A, LOGT, NT, S2, P::R2, false, false, false, true, true, true>(t, sl, ss, tw, R, m, &w2);
    if constexpr (P::R3 > 0) {
        RoundTw<A, LOGT, NT, S3, P::R3> w3;
        w3.load(t, 0, tw, R);
        __syncthreads();
        tile_round<A, LOGT, NT, S3, P::R3, false, false, false, true, true, true>(t, sl, ss, tw, R, m, &w3);
    }
    if constexpr (P::R4 > 0) {
        RoundTw<A, LOGT, NT, S4, P::R4> w4;
        w4.load(t, 0, tw, R);
        __syncthreads();
        tile_round<A, LOGT, NT, S4, P::R4, false, false, false, true, true, true>(t, sl, ss, tw, R, m, &w4);
    }
    __syncthreads();
#endif
}

// Canonical store of a finished tile: st(i, v) receives the canonical residue of output i (linear, coalesced order).
template <class A, int LOGT, class St> TB_D void tile_store_canon(const u64 *sm, St st, const ModDev *md) {
    typedef TilePlan<LOGT> P;
    constexpr int NT = P::NT, T = 1 << LOGT;
    const typename A::Mod m = A::make(md);
    // The epilogue's own global operands (St::Pre: accumulator and ciphertext words of the mod-down tail) are requested
    // for UB outputs before any of them is used or stored: one memory latency per UB outputs instead of one per output
    // (the compiler cannot hoist those loads above the stores, which may alias them).
    constexpr int UB = 8;
    static_assert((T / NT) % UB == 0, "bad store batch");
    TB_PHASE(NT, t, {
        _Pragma("unroll 1") for (int k0 = 0; k0 < T / NT; k0 += UB) {
            typename St::Pre pre[UB];
            _Pragma("unroll") for (int k = 0; k < UB; k++) pre[k] = st.pre(t + (k0 + k) * NT);
            _Pragma("unroll") for (int k = 0; k < UB; k++) {
                const u32 i = t + (k0 + k) * NT;
                st(i, A::canon(A::from_bits(sm[pad_idx(i)]), m), pre[k]);
            }
        }
    })
}

// The tile's own inverse transform (layers LOGT-1..0 with root base R); the first round reads the canonical
// coefficient i of the tile as gld(i) straight from global memory.
//   FINAL : the tile is the whole limb: layer 0 carries N^-1 and st(i, v) receives canonical residues; otherwise the
//           element bits stay in shared memory for the cluster's gather stage and st is not called.
template <class A, int LOGT, bool FINAL, class Gld, class St>
TB_D void tile_ntt_inv(u64 *sm, Gld gld, St st, const ModDev *md, u32 R) {
    typedef TilePlan<LOGT> P;
    constexpr int NT = P::NT;
    constexpr bool RE = A::kReduceEachInvRound, RM = A::kReduceInvMidway;
    const typename A::Mod m = A::make(md);
    const void *tw = A::inv_table(md);
    SmLd<A> sl{sm};
    SmSt<A> ss{sm};
    auto rl = [&](u32 i) -> typename A::elem { return A::from_u64(gld(i), m); }; // canonical input: no re-centring needed
    constexpr int S1 = P::R0, S2 = S1 + P::R1, S3 = S2 + P::R2, S4 = S3 + P::R3;
#if defined(TROYB_HOST_EMUL)
    if constexpr (P::R4 > 0) {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S4, P::R4, true, false, false, false, true>(t, rl, ss, tw, R, m));)
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S3, P::R3, true, false, RE, true, true>(t, sl, ss, tw, R, m));)
    } else {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S3, P::R3, true, false, false, false, true>(t, rl, ss, tw, R, m));)
    }
    TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S2, P::R2, true, false, RE, true, true>(t, sl, ss, tw, R, m));)
    TB_PHASE(NT, t, (tile_round<A, LOGT, NT, S1, P::R1, true, false, RE || RM, true, true>(t, sl, ss, tw, R, m));)
    if constexpr (FINAL) {
        auto gs_ = [&](u32 i, typename A::elem x) { st(i, A::canon_reduced(x, m)); };
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, 0, P::R0, true, true, RE, true, false>(t, sl, gs_, tw, R, m));)
    } else {
        TB_PHASE(NT, t, (tile_round<A, LOGT, NT, 0, P::R0, true, false, RE, true, true>(t, sl, ss, tw, R, m));)
    }
#else
    const u32 t = threadIdx.x;
    if constexpr (P::R4 > 0) {
        RoundTw<A, LOGT, NT, S4, P::R4> w4;
        w4.load(t, 0, tw, R);
        tile_round<A, LOGT, NT, S4, P::R4, true, false, false, false, true, true>(t, rl, ss, tw, R, m, &w4);
        RoundTw<A, LOGT, NT, S3, P::R3> w3;
        w3.load(t, 0, tw, R);
        __syncthreads();
        tile_round<A, LOGT, NT, S3, P::R3, true, false, RE, true, true, true>(t, sl, ss, tw, R, m, &w3);
    } else {
        RoundTw<A, LOGT, NT, S3, P::R3> w3;
        w3.load(t, 0, tw, R);
        tile_round<A, LOGT, NT, S3, P::R3, true, false, false, false, true, true>(t, rl, ss, tw, R, m, &w3);
    }
    RoundTw<A, LOGT, NT, S2, P::R2> w2;
    w2.load(t, 0, tw, R);
    __syncthreads();
    tile_round<A, LOGT, NT, S2, P::R2, true, false, RE, true, true, true>(t, sl, ss, tw, R, m, &w2);
    RoundTw<A, LOGT, NT, S1, P::R1> w1;
    w1.load(t, 0, tw, R);
    __syncthreads();
    tile_round<A, LOGT, NT, S1, P::R1, true, false, RE || RM, true, true, true>(t, sl, ss, tw, R, m, &w1);
    RoundTw<A, LOGT, NT, 0, P::R0> w0;
    w0.load(t, 0, tw, R);
    __syncthreads();
    if constexpr (FINAL) {
        auto gs_ = [&](u32 i, typename A::elem x) { st(i, A::canon_reduced(x, m)); };
        tile_round<A, LOGT, NT, 0, P::R0, true, true, RE, true, false, true>(t, sl, gs_, tw, R, m, &w0);
    } else {
        tile_round<A, LOGT, NT, 0, P::R0, true, false, RE, true, true, true>(t, sl, ss, tw, R, m, &w0);
    }
    __syncthreads();
#endif
}

// ---- cluster stages ------------------------------------------------------------------------------------------------------------
// A limb of N = C tiles is one cluster of C CTAs (C = 2, 4, 8). Tile r owns the contiguous block [r T, (r+1) T) of the
// transform's bit-reversed side.
//
// Forward. The first log C layers couple coefficients one tile apart. CTA `rank` loads the C coefficients {v + k T}
// for its share of v, runs the radix-C network in registers and hands output k to tile k (a store into that CTA's
// shared memory through the cluster's distributed shared memory). After the cluster barrier every CTA owns an
// independent 2^LOGT-point transform with root base C + rank.
//   ld(i)            : coefficient i of the limb as u64, a bare global load
//   pro(x)           : prologue applied to a loaded coefficient (kept out of ld: its mode branches between the loads made
//                      ptxas issue one load, wait, reduce, next load - a memory latency per coefficient instead of per batch)
//   st_peer(k, i, b) : store raw element bits b at padded index i of tile k
//   peers_started()  : called once by every thread before the first remote store (the peers' shared memory must exist)
template <class A, int LOGC, int LOGT, int NT, class Ld, class Pro, class StPeer, class Started>
TB_D void cluster_fwd_scatter(u32 vt, u32 rank, Ld ld, Pro pro, StPeer st_peer, Started peers_started, const ModDev *md) {
    constexpr int C = 1 << LOGC, T = 1 << LOGT, G = T / C / NT;
    static_assert(G >= 1, "bad cluster stage");
    // B iterations (8 coefficients; 16 measured 0.4 .. 2 % slower) are loaded before any of them is used: one memory latency per batch, not per
    // iteration
    constexpr int B = (G * C >= 8) ? (C >= 8 ? 1 : 8 / C) : G; // radix-8 holds 7 twiddles: one iteration at a time
    static_assert(G % B == 0, "bad scatter batch");
    const typename A::Mod m = A::make(md);
    const void *tw = A::fwd_table(md);
    typename A::tw_t w[C - 1];
#pragma unroll
    for (int k = 1; k < C; k++) w[k - 1] = A::ld_tw(tw, k); // layer u, block j: table entry 2^u + j
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        u64 raw[B][C];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (T / C) + vt + (g0 + b) * NT;
#pragma unroll
            for (int k = 0; k < C; k++) raw[b][k] = ld(v + (u32)k * T);
        }
        if (g0 == 0) peers_started();
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (T / C) + vt + (g0 + b) * NT;
            typename A::elem x[C];
#pragma unroll
            for (int k = 0; k < C; k++) x[k] = A::from_u64(pro(raw[b][k]), m);
            radix_fwd<A, LOGC>(x, w, m);
            const u32 pv = pad_idx(v);
#pragma unroll
            for (int k = 0; k < C; k++) st_peer((u32)k, pv, A::to_bits(x[k]));
        }
    }
}

// Inverse. Each tile's own inverse transform covers layers log N-1..log C; the last log C layers (layer 0 carries N^-1) couple the tiles: CTA `rank` gathers
// {tile k, position v} for its share of v and stores the canonical results (natural order, coalesced).
//   ld_peer(k, i) : raw element bits at padded index i of tile k
//   st(i, v)      : canonical residue of coefficient i of the limb (epilogue fused there)
template <class A, int LOGC, int LOGT, int NT, class LdPeer, class St>
TB_D void cluster_inv_gather(u32 vt, u32 rank, LdPeer ld_peer, St st, const ModDev *md) {
    constexpr int C = 1 << LOGC, T = 1 << LOGT, G = T / C / NT;
    static_assert(G >= 1, "bad cluster stage");
    constexpr int B = (G * C >= 8) ? 8 / C : G; // distributed-shared-memory loads in flight: B iterations
    static_assert(G % B == 0, "bad gather batch");
    // the tile's last round leaves at most 2^3 x the re-centred magnitude: one more layer needs no re-centring
    constexpr bool RE = A::kReduceEachInvRound && LOGC > 1;
    const typename A::Mod m = A::make(md);
    const void *tw = A::inv_table(md);
    typename A::tw_t w[C - 1];
#pragma unroll
    for (int k = 2; k < C; k++) w[k - 1] = A::ld_tw(tw, k); // entry 1 (layer 0) is folded into gs_last
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        u64 raw[B][C];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 pv = pad_idx(rank * (T / C) + vt + (g0 + b) * NT);
#pragma unroll
            for (int k = 0; k < C; k++) raw[b][k] = ld_peer((u32)k, pv);
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (T / C) + vt + (g0 + b) * NT;
            typename A::elem x[C];
#pragma unroll
            for (int k = 0; k < C; k++) {
                x[k] = A::from_bits(raw[b][k]);
                if (RE) A::reduce(x[k], m);
            }
            radix_inv<A, LOGC, true>(x, w, m);
#pragma unroll
            for (int k = 0; k < C; k++) {
                // outputs of layer 0 are mulmod results; for C > 2 the others passed through it too
                st(v + (u32)k * T, A::canon_reduced(x[k], m));
            }
        }
    }
}

} // namespace troyb
