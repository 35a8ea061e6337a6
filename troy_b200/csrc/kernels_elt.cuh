// Coefficient-wise kernels of the RNS polynomial path (functors for launch_cta).
// Every kernel here is HBM-bound integer work: one thread owns one coefficient index of one limb (or of all
// limbs when the operation couples limbs, as the base conversions do), loads are coalesced along the coefficient
// axis, and everything that the reference does in several passes over global memory is done in one.
#pragma once
#include "launch.cuh"

namespace troyb {

constexpr int kEltThreads = 256;
constexpr u32 kStagedCoeffs = 2 * kEltThreads, kStagedMaxDigits = 8; // KsInnerProductStaged geometry

TB_D void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
    u64 pl = a * b, ph = mulhi64(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

// ---- add / sub / negate / scalar multiply -------------------------------------------------------------------
// kAddPolyCoeffmod / kSubPolyCoeffmod / kNegatePolyCoeffmod   src/kernelutils.cu:30-58, 505-535, 300-328
// kMultiplyPolyScalarCoeffmod                                 src/kernelutils.cu:272-298
struct EltAddSub {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "elt_add_sub_negate_scale"; }
    int op; // 0 add, 1 sub, 2 negate, 3 multiply by `scalar`, 4 reduce modulo the prime, 5 multiply by N^-1 (kMultiplyInvPolyDegreeCoeffmod)
    u64 scalar;
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 total; // words
    const u64 *a, *b;
    u64 *out;
    u64 row_words; // 0: limb = (word / N) % n_limbs (ciphertext layout); else limb = word / row_words (row-major buffers)
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                u32 limb = row_words ? (u32)(idx / row_words) : (u32)((idx >> logN) % n_limbs);
                const ModDev &m = mods[map.idx[limb]];
                u64 p = m.p;
                u64 x = a[idx];
                u64 r;
                if (op == 0) r = addmod(x, b[idx], p);
                else if (op == 1) r = submod(x, b[idx], p);
                else if (op == 2) r = negmod(x, p);
                else if (op == 4) r = barrett64(x, p, m.ratio1);
                else if (op == 5) r = shoup(x, m.ninv, m.ninv_shoup, p);
                else r = mulmod_barrett(x, barrett64(scalar, p, m.ratio1), p, m.ratio0, m.ratio1);
                out[idx] = r;
            }
        })
    }
};

// One row (N words, `offset` into units `unit_stride` words apart) of every unit reduced modulo one prime in place: the
// special-prime rows of unreduced key products (sums of canonical residues) before their inverse transform.
struct EltReduceRow {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "reduce_special_rows"; }
    int logN;
    ModDev mod;
    u64 *base;
    u64 unit_stride, offset;
    u64 total; // units * N / 2 (two adjacent words per thread)
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 w = idx * 2, unit = w >> logN, c = w & ((u64(1) << logN) - 1);
                ulonglong2 *ptr = reinterpret_cast<ulonglong2 *>(base + unit * unit_stride + offset + c);
                ulonglong2 v = *ptr;
                v.x = barrett64(v.x, mod.p, mod.ratio1);
                v.y = barrett64(v.y, mod.p, mod.ratio1);
                *ptr = v;
            }
        })
    }
};

// kNegacyclicShiftPolyCoeffmod  src/kernelutils.cu:537-576: multiplication by X^shift modulo X^N + 1, shift in [0, 2N):
// out[(g + shift) mod N] = in[g], negated when g + shift passes N an odd number of times. Written as a gather.
struct EltNegacyclicShift {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "negacyclic_shift"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 shift;
    u64 total; // polys * n_limbs * N
    const u64 *in;
    u64 *out;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN, i = idx & (N - 1), base = idx - i;
                const u64 g = (i - shift) & (N - 1);          // source coefficient
                const bool flip = ((g + shift) >> logN) & 1;  // crossed X^N = -1 an odd number of times
                const u64 p = mods[map.idx[(u32)((idx >> logN) % n_limbs)]].p;
                const u64 v = in[base + g];
                out[idx] = flip ? negmod(v, p) : v;
            }
        })
    }
};

// c = a * b and c += a * b as ONE 32 x 32 -> 64 multiply(-add) with a 64-bit addend. Written as PTX on the GPU so that the accumulators
// stay 64-bit register pairs (the compiler otherwise splits them into halves and adds the zero initial value separately).
TB_D void madw_u32(u64 &c, u32 a, u32 b) {
#if defined(TROYB_HOST_EMUL)
    c += (u64)a * b;
#else
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
#endif
}
TB_D u64 mulw_u32(u32 a, u32 b) {
#if defined(TROYB_HOST_EMUL)
    return (u64)a * b;
#else
    u64 c;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(c) : "r"(a), "r"(b));
    return c;
#endif
}

// A sum of at most nine products x * k of residues modulo a prime of at most 40 bits (x, k < 2^40; the general form of such sums is
// ColumnAcc below): x is cut into 20-bit halves xa, xb and k into its
// 32-bit low word k0 and the 8-bit rest k1, so that
//   ca = sum xa*k0 (< 9 * 2^52), ca1 = sum xa*k1 (< 9 * 2^28, ONE 32-bit register), cb, cb1 likewise for xb,
//   value = Sa + Sb 2^20,  Sa = ca + ca1 2^32 < 2^64,  Sb = cb + cb1 2^32 < 2^64:
// two wide and two 32-bit multiply-adds per term, six registers of state (the general form: five wide ones and ten), no fold.
// The remainder is taken with a quotient estimated in FP64: value < 2^85 and p > 2^39 give q < 2^46, the estimate
// trunc((Sa + Sb 2^20)~ * (1/p)~) carries a relative error below 2^-51, i.e. less than 2^-5 in absolute terms, so it is q - 1, q or
// q + 1 and value - estimate * p, computed modulo 2^64, lies in [-p, 2p): two corrections give the canonical residue.
struct SmallAcc {
    static constexpr u32 kMaxTerms = 9;
    u64 ca, cb;
    u32 ca1, cb1;
    TB_D void mac_first(u32 xa, u32 xb, u64 k) {
        const u32 k0 = (u32)k, k1 = (u32)(k >> 32);
        ca = mulw_u32(xa, k0);
        cb = mulw_u32(xb, k0);
        ca1 = xa * k1;
        cb1 = xb * k1;
    }
    TB_D void zero() { ca = cb = 0; ca1 = cb1 = 0; }
    // long sums: after eight terms the sum is replaced by its residue (which counts as one term)
    TB_D void renew(u64 p, double pinv) {
        ca = reduce(p, pinv);
        cb = 0; ca1 = cb1 = 0;
    }
    TB_D void mac(u32 xa, u32 xb, u64 k) {
        const u32 k0 = (u32)k, k1 = (u32)(k >> 32);
        madw_u32(ca, xa, k0);
        madw_u32(cb, xb, k0);
        ca1 += xa * k1;
        cb1 += xb * k1;
    }
    TB_D u64 reduce(u64 p, double pinv) const {
        const u64 sa = ca + ((u64)ca1 << 32), sb = cb + ((u64)cb1 << 32);
#if defined(TROYB_HOST_EMUL)
        const double v = __builtin_fma((double)sb, 1048576.0, (double)sa);
        const u64 q = (u64)(v * pinv);
#else
        const double v = __fma_rn(__ull2double_rn(sb), 1048576.0, __ull2double_rn(sa));
        const u64 q = __double2ull_rz(__dmul_rn(v, pinv));
#endif
        u64 r = sa + (sb << 20) - q * p;
        if ((long long)r < 0) r += p;
        return r - (r >= p ? p : 0);
    }
};
// the 20-bit halves of a residue below 2^40
struct XHalves {
    u32 xa, xb;
    TB_D explicit XHalves(u64 x) {
        xa = (u32)x & 0xFFFFFu;
        xb = (u32)(x >> 20);
    }
};

// ---- ciphertext tensor product ----------------------------------------------------------------------------
// gDyadicConvolutionCoeffmod / kDyadicSquareCoeffmod  src/kernelutils.cu:89-199 ; ckksMultiply src/evaluator_cuda.cu:384-433
// out[i] = sum_{j+k=i} a[j]*b[k] mod q, one pass, 128-bit accumulation + one Barrett-128 per output word.
// a, b, out may alias (in-place multiply): a thread reads every input word it needs before it writes.
struct EltTensor {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "ct_tensor_product"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;      // limbs per poly
    u32 a_size, b_size;
    u64 per_batch;    // n_limbs * N
    u64 total;        // batch * n_limbs * N / 2 threads (two adjacent coefficients per thread, 128-bit accesses)
    const u64 *a, *b;
    u64 *out;
    u64 a_batch_stride, b_batch_stride, out_batch_stride; // words
    static TB_D ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
    static TB_D void st2(u64 *p, u64 x, u64 y) {
        ulonglong2 v;
        v.x = x; v.y = y;
        *reinterpret_cast<ulonglong2 *>(p) = v;
    }
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 half = per_batch >> 1;
                u64 bt = idx / half, rem = (idx % half) * 2;
                u32 limb = (u32)(rem >> logN);
                const ModDev &m = mods[map.idx[limb]];
                const u64 p = m.p, r0 = m.ratio0, r1 = m.ratio1, mu = m.bar_mu;
                const u32 sh = m.bar_shift;
                const u64 *pa = a + bt * a_batch_stride + rem;
                const u64 *pb = b + bt * b_batch_stride + rem;
                u64 *po = out + bt * out_batch_stride + rem;
                if (a_size == 2 && b_size == 2 && (p >> 40) == 0) {
                    // primes of at most 40 bits (CTA-uniform: a CTA stays inside one limb for N >= 512, and the result does not depend
                    // on the form taken): 20-bit halves of a, short accumulators, FP64 quotient estimate (SmallAcc)
                    const ulonglong2 a0 = ld2(pa), a1 = ld2(pa + per_batch), b0 = ld2(pb), b1 = ld2(pb + per_batch);
                    const double pinv = m.pinv_d;
                    const XHalves a0x(a0.x), a0y(a0.y), a1x(a1.x), a1y(a1.y);
                    SmallAcc s;
                    s.mac_first(a0x.xa, a0x.xb, b0.x);
                    const u64 c0x = s.reduce(p, pinv);
                    s.mac_first(a0y.xa, a0y.xb, b0.y);
                    const u64 c0y = s.reduce(p, pinv);
                    s.mac_first(a0x.xa, a0x.xb, b1.x);
                    s.mac(a1x.xa, a1x.xb, b0.x);
                    const u64 c1x = s.reduce(p, pinv);
                    s.mac_first(a0y.xa, a0y.xb, b1.y);
                    s.mac(a1y.xa, a1y.xb, b0.y);
                    const u64 c1y = s.reduce(p, pinv);
                    s.mac_first(a1x.xa, a1x.xb, b1.x);
                    const u64 c2x = s.reduce(p, pinv);
                    s.mac_first(a1y.xa, a1y.xb, b1.y);
                    const u64 c2y = s.reduce(p, pinv);
                    st2(po, c0x, c0y);
                    st2(po + per_batch, c1x, c1y);
                    st2(po + 2 * per_batch, c2x, c2y);
                } else if (a_size == 2 && b_size == 2) {
                    const ulonglong2 a0 = ld2(pa), a1 = ld2(pa + per_batch), b0 = ld2(pb), b1 = ld2(pb + per_batch);
                    // residues below p < 2^61: sums of at most two products are below 2^(n+63) (reduce_sum)
                    u64 lo, hi;
                    u64 c0x = reduce_sum(a0.x * b0.x, mulhi64(a0.x, b0.x), p, mu, sh);
                    u64 c0y = reduce_sum(a0.y * b0.y, mulhi64(a0.y, b0.y), p, mu, sh);
                    lo = 0; hi = 0;
                    mac128(lo, hi, a0.x, b1.x);
                    mac128(lo, hi, a1.x, b0.x);
                    u64 c1x = reduce_sum(lo, hi, p, mu, sh);
                    lo = 0; hi = 0;
                    mac128(lo, hi, a0.y, b1.y);
                    mac128(lo, hi, a1.y, b0.y);
                    u64 c1y = reduce_sum(lo, hi, p, mu, sh);
                    u64 c2x = reduce_sum(a1.x * b1.x, mulhi64(a1.x, b1.x), p, mu, sh);
                    u64 c2y = reduce_sum(a1.y * b1.y, mulhi64(a1.y, b1.y), p, mu, sh);
                    st2(po, c0x, c0y);
                    st2(po + per_batch, c1x, c1y);
                    st2(po + 2 * per_batch, c2x, c2y);
                } else {
                    u64 av[8][2], bv[8][2], cv[15][2];
                    for (u32 j = 0; j < a_size; j++) { ulonglong2 v = ld2(pa + j * per_batch); av[j][0] = v.x; av[j][1] = v.y; }
                    for (u32 k = 0; k < b_size; k++) { ulonglong2 v = ld2(pb + k * per_batch); bv[k][0] = v.x; bv[k][1] = v.y; }
                    for (u32 i = 0; i < a_size + b_size - 1; i++) {
                        for (u32 e = 0; e < 2; e++) {
                            u64 lo = 0, hi = 0;
                            for (u32 j = 0; j < a_size; j++) {
                                if (i >= j && i - j < b_size) mac128(lo, hi, av[j][e], bv[i - j][e]);
                            }
                            cv[i][e] = barrett128(lo, hi, p, r0, r1);
                        }
                    }
                    for (u32 i = 0; i < a_size + b_size - 1; i++) st2(po + i * per_batch, cv[i][0], cv[i][1]);
                }
            }
        })
    }
};

// ---- key-switch inner product -----------------------------------------------------------------------------
// gSwitchKeyInplaceUtilB + UtilD  src/evaluator_cuda.cu:1034-1085: acc[k] = sum_j D[i][j] * key[j][k][key_index]
// with lazy accumulators kept in registers, one Barrett-128 at the end.
//   D    : [batch][L+1][nd][N]  mod-up output of digits digit0..digit0+nd-1 (canonical)
//   key  : [digits][2][key_limbs][N]
//   prod : [batch][2][L+1][N]  (t_poly_prod layout, §A.8)
//
// Sum of products of residues below 2^60 (every key-switch modulus is a user prime, at most 60 bits) held as five
// columns of partial products of 16/32-bit limbs of x with 32-bit limbs of k:
//   c0l = sum x0l*k0, c0h = sum x0h*k0 (x0 = x0h 2^16 + x0l), c1a = sum x0*k1, c1b = sum x1*k0, c2 = sum x1*k1,
//   value = c0l + c0h 2^16 + (c1a + c1b) 2^32 + c2 2^64.
// Every product is below 2^60 (x1, k1 < 2^28), so eight terms accumulate in 64 bits through the multiply-add's own
// 64-bit addend (IMAD.WIDE): five multiply instructions per term and nothing else, no carries until fold(), which must
// be called after at most kFoldEvery terms.
struct ColumnAcc {
    static constexpr u32 kFoldEvery = 8;
    u64 c0l, c0h, c1a, c1b, c2;
    u64 lo, hi; // folded 128-bit sum
    TB_D void init() { c0l = c0h = c1a = c1b = c2 = 0; lo = hi = 0; }
    static TB_D void madw(u64 &c, u32 a, u32 b) { madw_u32(c, a, b); }
    static TB_D u64 mulw(u32 a, u32 b) { return mulw_u32(a, b); }
    // first term: plain products, no addend
    TB_D void mac_first(u32 xl, u32 xh, u32 x0, u32 x1, u64 k) {
        const u32 k0 = (u32)k, k1 = (u32)(k >> 32);
        c0l = mulw(xl, k0);
        c0h = mulw(xh, k0);
        c1a = mulw(x0, k1);
        c1b = mulw(x1, k0);
        c2 = mulw(x1, k1);
        lo = hi = 0;
    }
    // xl, xh, x1: the low 16 bits, bits 16-31 and the high word of x
    TB_D void mac(u32 xl, u32 xh, u32 x0, u32 x1, u64 k) {
        const u32 k0 = (u32)k, k1 = (u32)(k >> 32);
        madw(c0l, xl, k0);
        madw(c0h, xh, k0);
        madw(c1a, x0, k1);
        madw(c1b, x1, k0);
        madw(c2, x1, k1);
    }
    TB_D void fold() {
        // bits 0..63 and the carries into bit 64
        u64 l = c0l, h = c2;
        u64 t = c0h << 16;
        l += t; h += (l < t);
        h += c0h >> 48;
        t = c1a << 32;
        l += t; h += (l < t);
        h += c1a >> 32;
        t = c1b << 32;
        l += t; h += (l < t);
        h += c1b >> 32;
        lo += l;
        hi += h + (lo < l);
        c0l = c0h = c1a = c1b = c2 = 0;
    }
};
struct XLimbs {
    u32 xl, xh, x0, x1;
    TB_D explicit XLimbs(u64 x) {
        x0 = (u32)x;
        x1 = (u32)(x >> 32);
#if defined(TROYB_HOST_EMUL)
        xl = x0 & 0xFFFFu;
        xh = x0 >> 16;
#else
        xl = __byte_perm(x0, 0, 0x4410); // one PRMT each
        xh = __byte_perm(x0, 0, 0x4432);
#endif
    }
};

// One face for the two accumulator forms, so that a kernel writes its sum once: Form::X splits an operand, Form::Acc adds the
// products up, finish() returns the canonical residue.
struct GeneralForm {
    typedef XLimbs X;
    struct Acc : ColumnAcc {
        TB_D void first(const X &x, u64 k) { mac_first(x.xl, x.xh, x.x0, x.x1, k); }
        TB_D void add(const X &x, u64 k) { mac(x.xl, x.xh, x.x0, x.x1, k); }
        TB_D u64 finish(u64 p, u64 mu, u32 sh, double) { fold(); return reduce_sum(lo, hi, p, mu, sh); }
    };
};
struct SmallForm {
    typedef XHalves X;
    struct Acc : SmallAcc {
        TB_D void first(const X &x, u64 k) { mac_first(x.xa, x.xb, k); }
        TB_D void add(const X &x, u64 k) { mac(x.xa, x.xb, k); }
        TB_D u64 finish(u64 p, u64, u32, double pinv) { return reduce(p, pinv); }
    };
};

struct KsInnerProduct {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "ks_inner_product"; }
    int logN;
    const ModDev *mods;
    LimbMap map; // output modulus i -> pool index (i == L -> special)
    u32 L;                // data limbs of the level (output moduli: L + 1)
    u32 nd, digit0;       // digits summed by this launch
    u32 key_limbs;
    u32 special_key_limb; // key_limbs - 1
    u64 total;            // batch * (L+1) * N/2 threads (two adjacent coefficients per thread, 128-bit accesses)
    const u64 *D;
    const u64 *key;
    u64 *prod;
    u64 prod_batch_stride, prod_poly_stride, prod_limb_stride; // words: [batch][2][L+1][N] by default, or rows [L+1][batch][2][N]
    const u64 *pass;      // CKKS: the NTT-form target [batch][L][N]; its limb i IS digit i under modulus i (no transform)
    u64 pass_stride;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                const u64 c = (idx & (N / 2 - 1)) * 2;
                const u64 bi = idx >> (logN - 1);
                const u32 i = (u32)(bi % (L + 1));
                const u64 b = bi / (L + 1);
                const ModDev &m = mods[map.idx[i]];
                const u32 key_limb = (i == L) ? special_key_limb : i;
                const u64 *d = D + (bi * nd) * N + c;
                const u64 *kp = key + (u64)key_limb * N + c + (u64)digit0 * 2 * key_limbs * N;
                const u64 key_digit_stride = 2ull * key_limbs * N, key_comp_stride = (u64)key_limbs * N;
                ColumnAcc a0, a1, a2, a3;
                a0.init(); a1.init(); a2.init(); a3.init();
                // digit i - digit0 of output modulus i < L is read from the NTT-form target itself
                const u32 jpass = (pass != nullptr && i < L && i >= digit0) ? i - digit0 : 0xFFFFFFFFu;
                const u64 *ps = pass + b * pass_stride + (u64)i * N + c;
                for (u32 j0 = 0; j0 < nd; j0 += 4) { // four digits' loads in flight
                    ulonglong2 x[4], k0[4], k1[4];
                    _Pragma("unroll") for (u32 q = 0; q < 4; q++) {
                        const u32 j = j0 + q;
                        if (j < nd) {
                            const u64 *src = (j == jpass) ? ps : d + (u64)j * N;
                            x[q] = *reinterpret_cast<const ulonglong2 *>(src);
                            k0[q] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + j * key_digit_stride));
                            k1[q] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + j * key_digit_stride + key_comp_stride));
                        }
                    }
                    _Pragma("unroll") for (u32 q = 0; q < 4; q++) {
                        if (j0 + q < nd) {
                            const XLimbs xa(x[q].x), xb(x[q].y);
                            a0.mac(xa.xl, xa.xh, xa.x0, xa.x1, k0[q].x);
                            a1.mac(xb.xl, xb.xh, xb.x0, xb.x1, k0[q].y);
                            a2.mac(xa.xl, xa.xh, xa.x0, xa.x1, k1[q].x);
                            a3.mac(xb.xl, xb.xh, xb.x0, xb.x1, k1[q].y);
                        }
                    }
                    if (((j0 + 4) % ColumnAcc::kFoldEvery) == 0 || j0 + 4 >= nd) {
                        a0.fold(); a1.fold(); a2.fold(); a3.fold();
                    }
                }
                u64 *po = prod + b * prod_batch_stride + (u64)i * prod_limb_stride + c;
                ulonglong2 r0, r1;
                r0.x = barrett128(a0.lo, a0.hi, m.p, m.ratio0, m.ratio1);
                r0.y = barrett128(a1.lo, a1.hi, m.p, m.ratio0, m.ratio1);
                r1.x = barrett128(a2.lo, a2.hi, m.p, m.ratio0, m.ratio1);
                r1.y = barrett128(a3.lo, a3.hi, m.p, m.ratio0, m.ratio1);
                *reinterpret_cast<ulonglong2 *>(po) = r0;
                *reinterpret_cast<ulonglong2 *>(po + prod_poly_stride) = r1;
            }
        })
    }
};

// The same inner product with the key slice of the CTA staged in shared memory and reused for `bc` ciphertexts of the
// batch: a CTA owns (512 coefficients, output modulus i, a group of bc batch elements), copies key[j][k][i][its 512
// coefficients] for the ND <= 8 digits of the launch once (ND * 8 KiB), then streams the mod-up rows of its
// batch elements (ND is a template parameter so that the digit loops unroll without guards). The key's L2 -> SM traffic (twice the size of D when every ciphertext fetches it again) drops by bc,
// and for keys larger than the L2 (N = 65536, 40 limbs) so does its HBM traffic.
//   accumulate : add the previous contents of prod (launches over successive digit groups of one key switch)
//   SMALL      : CTAs whose row has a prime of at most 40 bits take the short-accumulator form (SmallAcc), the others the general one, in
//                ONE launch: the ALU-bound CTAs of the 60-bit rows then share the SMs with the DRAM-bound ones of the 40-bit rows (0.593 ms;
//                as two launches by row kind 0.624 ms). false: general form everywhere (TROYB_IP_SMALL=0). The rows of a launch are listed in `rows`
template <int ND, bool SMALL = false> struct KsInnerProductStaged {
    static constexpr int kThreads = kEltThreads;
    static constexpr int kMinCtas = 2; // up to 128 registers: a batch element's rows in flight plus the accumulators. The 40-bit form at three CTAs
                                       // (80 registers) spills and measures slower than the general form (0.806 against 0.767 ms), at two it takes 0.624 ms
    static constexpr u32 kCoeffs = 2 * kEltThreads; // per CTA
    static_assert(ND >= 1 && ND <= (int)ColumnAcc::kFoldEvery && ND + 1 <= (int)SmallAcc::kMaxTerms, "one fold per launch");
    const char *name() const { return "ks_inner_product"; }
    u32 n_rows;           // output moduli of this launch ...
    uint8_t rows[kMaxLimbs + 1]; // ... and which (i == L: the special prime)
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 L, digit0, key_limbs, special_key_limb; // ND digits from digit0
    u32 batch, bc;        // batch elements in the launch / per CTA
    u32 d_rows, d_row0;   // rows of D per (batch element, output modulus) and the row of this launch's first digit
    int accumulate;
    const u64 *D;
    const u64 *key;
    u64 *prod;
    u64 prod_batch_stride, prod_poly_stride, prod_limb_stride; // see KsInnerProduct
    const u64 *pass;
    u64 pass_stride;
    static size_t smem_bytes() { return (size_t)ND * 2 * kCoeffs * sizeof(u64); }
    // grid = (N / kCoeffs) * n_rows * ceil(batch / bc), coefficient block fastest
    TB_D void operator()(u32 bid, u64 *sm) const {
        const u64 N = u64(1) << logN;
        const u32 blocks = (u32)(N / kCoeffs);
        const u32 cb = bid % blocks, i = rows[(bid / blocks) % n_rows], bg = bid / blocks / n_rows;
        const u32 key_limb = (i == L) ? special_key_limb : i;
        const u64 key_digit_stride = 2ull * key_limbs * N, key_comp_stride = (u64)key_limbs * N;
        const u64 *kp = key + (u64)key_limb * N + (u64)cb * kCoeffs + (u64)digit0 * key_digit_stride;
        TB_PHASE(kThreads, t, {
            _Pragma("unroll") for (u32 r = 0; r < ND * 2; r++) {
                const u64 *src = kp + (u64)(r >> 1) * key_digit_stride + (u64)(r & 1) * key_comp_stride + 2 * t;
                *reinterpret_cast<ulonglong2 *>(sm + (u64)r * kCoeffs + 2 * t) = __ldg(reinterpret_cast<const ulonglong2 *>(src));
            }
        })
        TB_PHASE(kThreads, t, {
            const ModDev &m = mods[map.idx[i]];
            const u64 p = m.p, mu = m.bar_mu;
            const u32 sh = m.bar_shift;
            const double pinv = m.pinv_d; // SMALL
            (void)mu; (void)sh; (void)pinv;
            const u64 c = (u64)cb * kCoeffs + 2 * t;
            // digit i of output modulus i < L is the NTT-form target itself (pass): its row of D is never written or read
            const u32 jpass = (pass != nullptr && i < L && i >= digit0 && i - digit0 < ND) ? i - digit0 : 0xFFFFFFFFu;
            const u32 b_end = (bg + 1) * bc < batch ? (bg + 1) * bc : batch;
            const u64 d_stride = (u64)(L + 1) * d_rows * N;
            const u64 *d = D + (((u64)(bg * bc) * (L + 1) + i) * d_rows + d_row0) * N + c;
            const u64 *ps = pass + (u64)(bg * bc) * pass_stride + (u64)i * N + c;
            u64 *po = prod + (u64)(bg * bc) * prod_batch_stride + (u64)i * prod_limb_stride + c;
            // Register-rotating pipeline over the batch elements of the CTA: row j of element b + 1 is requested as soon
            // as row j of element b has been copied out of its registers, so ND 16-byte loads per thread stay in flight
            // for a whole iteration (the kernel is bound by HBM latency x bytes in flight otherwise) with ONE set of row
            // registers (a second set for the next element spilled to local memory, and the spill stores waited for the
            // loads: profiles/r1_s6 notes).
            ulonglong2 x[ND], xp;
            _Pragma("unroll") for (u32 j = 0; j < ND; j++) {
                x[j].x = 0; x[j].y = 0;
                if (bg * bc < b_end && j != jpass) x[j] = *reinterpret_cast<const ulonglong2 *>(d + (u64)j * N);
            }
            xp.x = 0; xp.y = 0;
            if (bg * bc < b_end && jpass != 0xFFFFFFFFu) xp = *reinterpret_cast<const ulonglong2 *>(ps);
            for (u32 b = bg * bc; b < b_end; b++, d += d_stride, ps += pass_stride, po += prod_batch_stride) {
                const bool more = b + 1 < b_end;
                ulonglong2 old0, old1;
                if (accumulate) {
                    old0 = *reinterpret_cast<const ulonglong2 *>(po);
                    old1 = *reinterpret_cast<const ulonglong2 *>(po + prod_poly_stride);
                }
                ulonglong2 q0, q1;
                auto sum = [&](auto form) {
                    typedef decltype(form) Form;
                    typedef typename Form::X X;
                    typename Form::Acc a0, a1, a2, a3;
                    _Pragma("unroll") for (u32 j = 0; j < ND; j++) {
                        const ulonglong2 cur = x[j];
                        if (more && j != jpass) x[j] = *reinterpret_cast<const ulonglong2 *>(d + d_stride + (u64)j * N);
                        const ulonglong2 k0 = *reinterpret_cast<const ulonglong2 *>(sm + (u64)(2 * j) * kCoeffs + 2 * t);
                        const ulonglong2 k1 = *reinterpret_cast<const ulonglong2 *>(sm + (u64)(2 * j + 1) * kCoeffs + 2 * t);
                        const X xa(cur.x), xb(cur.y);
                        if (j == 0) {
                            a0.first(xa, k0.x); a1.first(xb, k0.y); a2.first(xa, k1.x); a3.first(xb, k1.y);
                        } else {
                            a0.add(xa, k0.x); a1.add(xb, k0.y); a2.add(xa, k1.x); a3.add(xb, k1.y);
                        }
                    }
                    if (jpass != 0xFFFFFFFFu) { // CTA-uniform
                        const ulonglong2 cur = xp;
                        if (more) xp = *reinterpret_cast<const ulonglong2 *>(ps + pass_stride);
                        const ulonglong2 k0 = *reinterpret_cast<const ulonglong2 *>(sm + (u64)(2 * jpass) * kCoeffs + 2 * t);
                        const ulonglong2 k1 = *reinterpret_cast<const ulonglong2 *>(sm + (u64)(2 * jpass + 1) * kCoeffs + 2 * t);
                        const X xa(cur.x), xb(cur.y);
                        a0.add(xa, k0.x); a1.add(xb, k0.y); a2.add(xa, k1.x); a3.add(xb, k1.y);
                    }
                    q0.x = a0.finish(p, mu, sh, pinv);
                    q0.y = a1.finish(p, mu, sh, pinv);
                    q1.x = a2.finish(p, mu, sh, pinv);
                    q1.y = a3.finish(p, mu, sh, pinv);
                };
                if (SMALL && (p >> 40) == 0) sum(SmallForm{}); // CTA-uniform
                else sum(GeneralForm{});
                if (accumulate) {
                    q0.x = addmod(q0.x, old0.x, p); q0.y = addmod(q0.y, old0.y, p);
                    q1.x = addmod(q1.x, old1.x, p); q1.y = addmod(q1.y, old1.y, p);
                }
                *reinterpret_cast<ulonglong2 *>(po) = q0;
                *reinterpret_cast<ulonglong2 *>(po + prod_poly_stride) = q1;
            }
        })
    }
};

// ---- ciphertext x plaintext, ciphertext +- plaintext -----------------------------------------------
// multiplyPlainNtt: kDyadicProductCoeffmod on every poly   src/evaluator_cuda.cu:1818-1862, src/kernelutils.cu:151-186
// (CPU src/evaluator.cpp:1937-1985). One launch for a whole batch; `accumulate` adds the product to what `out` holds,
// which turns the multiplyPlain + addInplace pairs of MatmulHelper::matmul (app/LinearHelperCKKS.cuh:225-249) into one
// pass over the output.
struct EltMulPlain {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "ct_times_plain"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;      // limbs per poly
    u64 per_ct;       // ct_size * n_limbs * N
    u64 total;        // batch * per_ct / 2 threads (two adjacent coefficients per thread, 128-bit accesses)
    const u64 *ct, *plain;
    u64 *out;
    u64 ct_stride, plain_stride, out_stride; // words between batch elements (plain_stride 0: one plaintext for all)
    int accumulate;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 half = per_ct >> 1;
                const u64 b = idx / half, rem = (idx % half) * 2;
                const u64 per_poly = (u64)n_limbs << logN;
                const u64 in_poly = rem % per_poly;
                const ModDev &m = mods[map.idx[(u32)(in_poly >> logN)]];
                const u64 p = m.p, mu = m.bar_mu;
                const u32 sh = m.bar_shift;
                const ulonglong2 x = EltTensor::ld2(ct + b * ct_stride + rem);
                const ulonglong2 w = EltTensor::ld2(plain + b * plain_stride + in_poly);
                u64 r0, r1;
                if ((p >> 40) == 0) { // SmallAcc: see the tensor product
                    const XHalves xx(x.x), xy(x.y);
                    SmallAcc s;
                    s.mac_first(xx.xa, xx.xb, w.x);
                    r0 = s.reduce(p, m.pinv_d);
                    s.mac_first(xy.xa, xy.xb, w.y);
                    r1 = s.reduce(p, m.pinv_d);
                } else {
                    r0 = reduce_sum(x.x * w.x, mulhi64(x.x, w.x), p, mu, sh);
                    r1 = reduce_sum(x.y * w.y, mulhi64(x.y, w.y), p, mu, sh);
                }
                u64 *po = out + b * out_stride + rem;
                if (accumulate) {
                    const ulonglong2 o = EltTensor::ld2(po);
                    r0 = addmod(r0, o.x, p);
                    r1 = addmod(r1, o.y, p);
                }
                EltTensor::st2(po, r0, r1);
            }
        })
    }
};

// MatmulHelper::matmul (app/LinearHelperCKKS.cuh:225-249): out[b][j] = sum_i a[b][i] * w[i][j] over a grid of NTT-form
// ciphertexts and plaintexts, which the reference runs as B*I*J multiplyPlain calls plus B*(I-1)*J addInplace calls, each
// a pass over global memory. Here a thread owns two adjacent coefficients of one limb of one poly for a 2 x 2 tile of
// (b, j): it streams i once, so every a[b][i] word is read ceil(J/2) times instead of J times, every w[i][j] word
// ceil(B/2) times instead of B, and the partial sums never leave registers (128-bit lazy sums, folded every 32 terms so
// that any inner dimension is safe for 61-bit primes). Canonical output = what the reference's chain of canonical
// products and modular additions yields.
struct EltMatmulPlain {
    static constexpr int kThreads = kEltThreads;
    static constexpr int kMinCtas = 2;
    const char *name() const { return "ct_matmul_plain"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs, ct_size;
    u32 B, I, J;
    u64 total; // ceil(B/2) * ceil(J/2) * ct_size * n_limbs * N/2 threads
    const u64 *a;   // [B][I] ciphertexts, a_ct_stride words apart
    const u64 *w;   // [I][J][n_limbs][N]
    u64 *out;       // [B][J] ciphertexts, out_ct_stride words apart
    u64 a_ct_stride, out_ct_stride;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 c = (idx & ((u64(1) << (logN - 1)) - 1)) * 2;
                u64 r = idx >> (logN - 1);
                const u32 limb = (u32)(r % n_limbs); r /= n_limbs;
                const u32 k = (u32)(r % ct_size); r /= ct_size;
                const u32 JT = (J + 1) / 2;
                const u32 j0 = (u32)(r % JT) * 2, b0 = (u32)(r / JT) * 2;
                const u32 j1 = j0 + 1 < J ? j0 + 1 : j0, b1 = b0 + 1 < B ? b0 + 1 : b0; // clamped: edge tiles recompute
                const ModDev &m = mods[map.idx[limb]];
                const u64 p = m.p, r0 = m.ratio0, r1 = m.ratio1;
                const u64 off = (((u64)k * n_limbs + limb) << logN) + c, woff = ((u64)limb << logN) + c;
                const u64 wplain = (u64)n_limbs << logN;
                const u64 *pa0 = a + (u64)b0 * I * a_ct_stride + off, *pa1 = a + (u64)b1 * I * a_ct_stride + off;
                const u64 *pw0 = w + (u64)j0 * wplain + woff, *pw1 = w + (u64)j1 * wplain + woff;
                if ((p >> 40) == 0) {
                    // primes of at most 40 bits: short accumulators (SmallAcc), renewed every eight terms
                    const double pinv = m.pinv_d;
                    SmallAcc acc[2][2][2]; // [b][j][coefficient]
                    _Pragma("unroll") for (int q = 0; q < 8; q++) (&acc[0][0][0])[q].zero();
                    _Pragma("unroll 2") for (u32 i = 0; i < I; i++) {
                        const ulonglong2 x0 = EltTensor::ld2(pa0 + (u64)i * a_ct_stride), x1 = EltTensor::ld2(pa1 + (u64)i * a_ct_stride);
                        const ulonglong2 w0 = EltTensor::ld2(pw0 + (u64)i * J * wplain), w1 = EltTensor::ld2(pw1 + (u64)i * J * wplain);
                        const XHalves x0x(x0.x), x0y(x0.y), x1x(x1.x), x1y(x1.y);
                        acc[0][0][0].mac(x0x.xa, x0x.xb, w0.x); acc[0][0][1].mac(x0y.xa, x0y.xb, w0.y);
                        acc[0][1][0].mac(x0x.xa, x0x.xb, w1.x); acc[0][1][1].mac(x0y.xa, x0y.xb, w1.y);
                        acc[1][0][0].mac(x1x.xa, x1x.xb, w0.x); acc[1][0][1].mac(x1y.xa, x1y.xb, w0.y);
                        acc[1][1][0].mac(x1x.xa, x1x.xb, w1.x); acc[1][1][1].mac(x1y.xa, x1y.xb, w1.y);
                        if ((i & 7) == 7) {
                            _Pragma("unroll") for (int q = 0; q < 8; q++) (&acc[0][0][0])[q].renew(p, pinv);
                        }
                    }
                    _Pragma("unroll") for (int bb = 0; bb < 2; bb++) {
                        _Pragma("unroll") for (int jj = 0; jj < 2; jj++) {
                            const u32 b = bb ? b0 + 1 : b0, j = jj ? j0 + 1 : j0;
                            if (b < B && j < J) {
                                EltTensor::st2(out + ((u64)b * J + j) * out_ct_stride + off, acc[bb][jj][0].reduce(p, pinv),
                                               acc[bb][jj][1].reduce(p, pinv));
                            }
                        }
                    }
                } else {
                u64 lo[2][2][2], hi[2][2][2]; // [b][j][coefficient]
                _Pragma("unroll") for (int q = 0; q < 8; q++) { (&lo[0][0][0])[q] = 0; (&hi[0][0][0])[q] = 0; }
                _Pragma("unroll 2") for (u32 i = 0; i < I; i++) {
                    const ulonglong2 x0 = EltTensor::ld2(pa0 + (u64)i * a_ct_stride), x1 = EltTensor::ld2(pa1 + (u64)i * a_ct_stride);
                    const ulonglong2 w0 = EltTensor::ld2(pw0 + (u64)i * J * wplain), w1 = EltTensor::ld2(pw1 + (u64)i * J * wplain);
                    mac128(lo[0][0][0], hi[0][0][0], x0.x, w0.x); mac128(lo[0][0][1], hi[0][0][1], x0.y, w0.y);
                    mac128(lo[0][1][0], hi[0][1][0], x0.x, w1.x); mac128(lo[0][1][1], hi[0][1][1], x0.y, w1.y);
                    mac128(lo[1][0][0], hi[1][0][0], x1.x, w0.x); mac128(lo[1][0][1], hi[1][0][1], x1.y, w0.y);
                    mac128(lo[1][1][0], hi[1][1][0], x1.x, w1.x); mac128(lo[1][1][1], hi[1][1][1], x1.y, w1.y);
                    if ((i & 31) == 31) {
                        _Pragma("unroll") for (int q = 0; q < 8; q++) {
                            (&lo[0][0][0])[q] = barrett128((&lo[0][0][0])[q], (&hi[0][0][0])[q], p, r0, r1);
                            (&hi[0][0][0])[q] = 0;
                        }
                    }
                }
                _Pragma("unroll") for (int bb = 0; bb < 2; bb++) {
                    _Pragma("unroll") for (int jj = 0; jj < 2; jj++) {
                        const u32 b = bb ? b0 + 1 : b0, j = jj ? j0 + 1 : j0;
                        if (b < B && j < J) {
                            EltTensor::st2(out + ((u64)b * J + j) * out_ct_stride + off,
                                           barrett128(lo[bb][jj][0], hi[bb][jj][0], p, r0, r1),
                                           barrett128(lo[bb][jj][1], hi[bb][jj][1], p, r0, r1));
                        }
                    }
                }
                } // general form
            }
        })
    }
};

// Plaintext lift: coefficient x of a plaintext modulo t becomes its centred representative modulo every q_i
// (x >= threshold: x + (Q - t)); coefficients past coeff_count are zero. gMultiplyPlainNormalUtilA/B + decomposeArray
// src/evaluator_cuda.cu:1724-1755,1783-1797; gTransformToNttInplace :1865-1935; CPU src/evaluator.cpp:1866-1895. In RNS
// both branches of the reference (using_fast_plain_lift or not) are the same function: (x mod q_i + inc_i) mod q_i.
struct EltPlainLift {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "plain_lift"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 coeff_count;
    u64 total; // batch * n_limbs * N
    const u64 *plain;
    u64 *out;
    u64 plain_stride; // words between the plaintexts of a batch; out is [batch][n_limbs][N]
    PlainConsts pc;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 c = idx & ((u64(1) << logN) - 1);
                const u32 limb = (u32)((idx >> logN) % n_limbs);
                const u64 b = (idx >> logN) / n_limbs;
                u64 r = 0;
                if (c < coeff_count) {
                    const ModDev &m = mods[map.idx[limb]];
                    const u64 x = plain[b * plain_stride + c];
                    r = barrett64(x, m.p, m.ratio1);
                    if (x >= pc.threshold) r = addmod(r, pc.inc[limb], m.p);
                }
                out[idx] = r;
            }
        })
    }
};

// addPlain / subPlain on the first poly of a coefficient-form ciphertext, one thread per coefficient, all limbs:
//   BFV (scaled = 1): multiplyAdd/SubPlainWithScalingVariant  src/utils/scalingvariant_cuda.cu:82-175
//        c0[j][i] +-= (x * floor(Q/t) + floor((x * (Q mod t) + threshold) / t)) mod q_j
//   BGV (scaled = 0): x * correction_factor mod t, then add/subPlainWithoutScalingVariant
//        src/evaluator_cuda.cu:1617-1622,1690-1694, src/utils/scalingvariant_cuda.cu:21-80
struct EltAddPlain {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "ct_plus_plain"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 coeff_count;
    u64 total; // batch * coeff_count
    const u64 *plain;
    u64 *c0;
    u64 plain_stride, ct_stride;
    int scaled, subtract;
    u64 factor; // BGV: the ciphertext's correction factor
    PlainConsts pc;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 b = idx / coeff_count, c = idx % coeff_count;
                u64 x = plain[b * plain_stride + c];
                u64 fix = 0;
                if (scaled) {
                    const unsigned __int128 num = (unsigned __int128)x * pc.q_mod_t + pc.threshold;
                    fix = (u64)(num / pc.t);
                } else {
                    x = (u64)(((unsigned __int128)x * factor) % pc.t);
                }
                u64 *dst = c0 + b * ct_stride + c;
                for (u32 j = 0; j < n_limbs; j++) {
                    const ModDev &m = mods[map.idx[j]];
                    u64 v = barrett64(x, m.p, m.ratio1);
                    if (scaled) v = addmod(mulmod_barrett(v, pc.delta[j], m.p, m.ratio0, m.ratio1), barrett64(fix, m.p, m.ratio1), m.p);
                    u64 *w = dst + ((u64)j << logN);
                    *w = subtract ? submod(*w, v, m.p) : addmod(*w, v, m.p);
                }
            }
        })
    }
};

// ---- decryption --------------------------------------------------------------------------------------
// DecryptorCuda::dotProductCtSkArray  src/decryptor_cuda.cu:237-300 on NTT-form operands: phase = c_0 + sum_{i>=1} c_i s^i,
// one pass with a 128-bit lazy sum (the reference: one product and one addition launch per term through global memory).
struct EltDecryptDot {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "decrypt_dot_product"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs, ct_size, key_limbs;
    u64 total; // batch * n_limbs * N / 2 threads
    const u64 *ct;
    const u64 *sk; // [power][key_limbs][N], NTT form at the key level: power i holds s^(i+1)
    u64 *out;      // [batch][n_limbs][N]
    u64 ct_stride;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 per = ((u64)n_limbs << logN) >> 1;
                const u64 b = idx / per, rem = (idx % per) * 2;
                const ModDev &m = mods[map.idx[(u32)(rem >> logN)]];
                const u64 *pc = ct + b * ct_stride + rem;
                const ulonglong2 c0 = EltTensor::ld2(pc);
                u64 lo0 = c0.x, hi0 = 0, lo1 = c0.y, hi1 = 0;
                for (u32 i = 1; i < ct_size; i++) {
                    const ulonglong2 c = EltTensor::ld2(pc + (u64)i * (per * 2));
                    const ulonglong2 s = EltTensor::ld2(sk + (((u64)(i - 1) * key_limbs) << logN) + rem);
                    mac128(lo0, hi0, c.x, s.x);
                    mac128(lo1, hi1, c.y, s.y);
                }
                EltTensor::st2(out + b * (per * 2) + rem, barrett128(lo0, hi0, m.p, m.ratio0, m.ratio1),
                               barrett128(lo1, hi1, m.p, m.ratio0, m.ratio1));
            }
        })
    }
};

// RNSToolCuda::decryptScaleAndRound  src/utils/rns_cuda.cu:499-577 (CPU src/utils/rns.cpp:1040-1075): the phase modulo q
// becomes the plaintext modulo t. One thread per coefficient does the reference's four launches: x t gamma, fast base
// conversion q -> {t, gamma}, x -(Q^-1), and the gamma-correction with the final x gamma^-1 mod t. Integer only.
struct EltDecryptScaleRound {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "decrypt_scale_round"; }
    int logN;
    const ModDev *mods;
    LimbMap map;          // q_0 .. q_{L-1}
    u32 n_limbs;
    u32 gamma_pool;       // pool index of gamma
    u64 total;            // batch * N
    const u64 *phase;     // [batch][L][N]
    u64 *out;             // [batch][N]
    DecryptConsts dc;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                const u64 b = idx >> logN, c = idx & (N - 1);
                const u64 *ph = phase + ((b * n_limbs) << logN) + c;
                const ModDev &mg = mods[gamma_pool];
                unsigned __int128 st = 0, sg = 0;
                for (u32 i = 0; i < n_limbs; i++) {
                    const ModDev &m = mods[map.idx[i]];
                    u64 y = mulmod_barrett(ph[(u64)i << logN], dc.tg_mod_q[i], m.p, m.ratio0, m.ratio1);
                    y = mulmod_barrett(y, dc.inv_punct_q[i], m.p, m.ratio0, m.ratio1);
                    st += (unsigned __int128)y * dc.punct_mod_t[i];
                    sg += (unsigned __int128)y * dc.punct_mod_gamma[i];
                    if ((i & 31) == 31) {
                        st %= dc.t;
                        sg = barrett128((u64)sg, (u64)(sg >> 64), mg.p, mg.ratio0, mg.ratio1);
                    }
                }
                u64 vt = (u64)(st % dc.t);
                u64 vg = barrett128((u64)sg, (u64)(sg >> 64), mg.p, mg.ratio0, mg.ratio1);
                vt = (u64)(((unsigned __int128)vt * dc.neg_inv_q_mod_t) % dc.t);
                vg = mulmod_barrett(vg, dc.neg_inv_q_mod_gamma, mg.p, mg.ratio0, mg.ratio1);
                u64 d;
                if (vg > (dc.gamma >> 1)) d = addmod(vt, (dc.gamma - vg) % dc.t, dc.t);
                else d = submod(vt, vg % dc.t, dc.t);
                if (d != 0) d = (u64)(((unsigned __int128)d * dc.inv_gamma_mod_t) % dc.t);
                out[idx] = d;
            }
        })
    }
};

// RNSToolCuda::decryptModt  src/utils/rns_cuda.cu:617-620 = BaseConverter::exactConvertArray q -> t (CPU
// src/utils/rns.cpp:462-538), then x correction_factor^-1 (bgvDecrypt src/decryptor_cuda.cu:143-186). The rounding term
// v = floor(sum_i y_i / q_i + 0.5) is formed in double precision exactly as the CPU path does it (IEEE division, the sum
// taken in limb order, truncating conversion), so the result is the same word.
struct EltDecryptModT {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "decrypt_mod_t"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 total;        // batch * N
    const u64 *phase; // [batch][L][N]
    u64 *out;         // [batch][N]
    u64 inv_factor;   // correction_factor^-1 mod t (1: none)
    DecryptConsts dc;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                const u64 b = idx >> logN, c = idx & (N - 1);
                const u64 *ph = phase + ((b * n_limbs) << logN) + c;
                unsigned __int128 st = 0;
                double v = 0.0;
                for (u32 i = 0; i < n_limbs; i++) {
                    const ModDev &m = mods[map.idx[i]];
                    const u64 y = mulmod_barrett(ph[(u64)i << logN], dc.inv_punct_q[i], m.p, m.ratio0, m.ratio1);
#if defined(TROYB_HOST_EMUL)
                    v = v + (double)y / (double)m.p;
#else
                    v = __dadd_rn(v, __ddiv_rn(__ull2double_rn(y), __ull2double_rn(m.p)));
#endif
                    st += (unsigned __int128)y * dc.punct_mod_t[i];
                    if ((i & 31) == 31) st %= dc.t;
                }
#if defined(TROYB_HOST_EMUL)
                const u64 rounded = (u64)(v + 0.5);
#else
                const u64 rounded = __double2ull_rz(__dadd_rn(v, 0.5));
#endif
                const u64 sum = (u64)(st % dc.t);
                const u64 vq = (u64)(((unsigned __int128)rounded * dc.q_mod_t) % dc.t);
                u64 d = submod(sum, vq, dc.t);
                if (inv_factor != 1) d = (u64)(((unsigned __int128)d * inv_factor) % dc.t);
                out[idx] = d;
            }
        })
    }
};

// ---- Galois automorphisms ---------------------------------------------------------------------------------
// gApplyGalois (coefficient form, scatter)  src/utils/galois_cuda.cu:139-164 ; CPU src/utils/galois.cpp:141-175
// gApplyGaloisNtt (NTT form, gather)        src/utils/galois_cuda.cu:181-196 ; table rule src/utils/galois_cuda.cu:26-42
struct EltGalois {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "galois_permute"; }
    int logN;
    int ntt_form;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u32 elt;
    u64 total; // batch * n_limbs * N   (n_limbs = limbs of all polys of one batch element)
    u32 L;     // limbs per poly
    const u64 *in;
    u64 *out;
    u64 in_batch_stride, out_batch_stride;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u32 N = 1u << logN;
                u32 i = (u32)(idx & (N - 1));
                u64 unit = idx >> logN;
                u64 b = unit / n_limbs, l = unit % n_limbs;
                const u64 *src = in + b * in_batch_stride + (l << logN);
                u64 *dst = out + b * out_batch_stride + (l << logN);
                if (ntt_form) {
                    // out[i] = in[ bitrev( ((elt * (2*bitrev(i)+1)) mod 2N) >> 1 ) ]
                    u32 r = __brev(i) >> (32 - logN);
                    u32 raw = (elt * (2 * r + 1)) & (2 * N - 1);
                    u32 srci = __brev(raw >> 1) >> (32 - logN);
                    dst[i] = src[srci];
                } else {
                    u64 p = mods[map.idx[l % L]].p;
                    u32 raw = (u32)(((u64)i * elt) & (2 * N - 1));
                    u64 v = src[i];
                    if (raw >= N) v = negmod(v, p);
                    dst[raw & (N - 1)] = v;
                }
            }
        })
    }
};

// ---- coefficient-form drop-last (BFV / BGV mod-switch-to-next) -----------------------------------------------------
// gDivideAndRoundqLastInplace src/utils/rns_cuda.cu:271-305 ; gModTAndDivideqLastInplace :579-615
//   in  : [batch*polys][L][N]   out : [batch*polys][L-1][N]
struct EltDropLast {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "drop_last_coeff"; }
    int logN;
    int scheme;
    const ModDev *mods;
    DropDev drop;
    u32 L;
    u64 total; // batch * size * (L-1) * N
    u32 size;  // polys per ciphertext
    u64 in_batch_stride, out_batch_stride;
    const u64 *in;
    u64 *out;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                u64 c = idx & (N - 1);
                u64 ul = idx >> logN;
                u32 i = (u32)(ul % (L - 1));
                u64 unit = ul / (L - 1);
                const ModDev &m = mods[i];
                const DropLimb dl = drop.limbs[i];
                u64 p = m.p;
                const u64 *pin = in + (unit / size) * in_batch_stride + (unit % size) * L * N + c;
                u64 last = pin[(u64)(L - 1) * N];
                u64 x = pin[(u64)i * N];
                u64 r;
                if (scheme == SCHEME_BGV) {
                    u64 neg = negmod(barrett64(last, drop.t, drop.t_ratio1), drop.t);
                    neg = shoup(neg, drop.inv_last_mod_t, drop.inv_last_mod_t_shoup, drop.t);
                    u64 delta = shoup(barrett64(neg, p, m.ratio1), dl.last_mod, dl.last_mod_shoup, p);
                    u64 d = x + (p << 1) - barrett64(last, p, m.ratio1) - delta;
                    r = shoup(d, dl.inv_last, dl.inv_last_shoup, p);
                } else {
                    u64 s = addmod(last, drop.half, drop.q_last);
                    u64 tmp = submod(barrett64(s, p, m.ratio1), dl.half_mod, p);
                    r = shoup(submod(x, tmp, p), dl.inv_last, dl.inv_last_shoup, p);
                }
                out[(unit / size) * out_batch_stride + ((unit % size) * (L - 1) + i) * N + c] = r;
            }
        })
    }
};

// ---- BEHZ (BFV multiply) ---------------------------------------------------------------------------------------------------
// Step (1)+(2): fastbconvmTilde + smMrq  src/utils/rns_cuda.cu:500-560, 423-465 ; CPU src/utils/rns.cpp:1012-1040, 943-983
//   in  : [polys][L][N]  (base q, coefficient form)
//   out : [polys][L+nBsk][N]: first L limbs = copy of the input, next nBsk limbs = the value in base Bsk
constexpr int kBehzMax = 40;

struct BehzExtend {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "behz_extend"; }
    int logN;
    const ModDev *mods;
    LimbMap map; // q_0..q_{L-1}, B.., m_sk
    BehzDev z;
    u64 total; // batch * polys_per_batch * N
    u32 polys_per_batch;
    u64 in_batch_stride;
    const u64 *in;
    u64 *out;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                u64 c = idx & (N - 1), poly = idx >> logN;
                const u32 L = z.L, nBsk = z.nBsk;
                const u64 *pi = in + (poly / polys_per_batch) * in_batch_stride + (poly % polys_per_batch) * L * N + c;
                u64 *po = out + poly * (L + nBsk) * N + c;
                u64 y[kBehzMax];
                u64 rm = 0; // value mod m_tilde = 2^32
                for (u32 i = 0; i < L; i++) {
                    u64 x = pi[(u64)i * N];
                    po[(u64)i * N] = x;
                    u64 p = mods[map.idx[i]].p;
                    y[i] = shoup(x, z.mtilde_inv_punct_q[2 * i], z.mtilde_inv_punct_q[2 * i + 1], p);
                    rm += y[i] * z.q_to_mtilde[i];
                }
                rm &= 0xFFFFFFFFull;
                // r_m_tilde = rm * (-Q^-1) mod 2^32
                u64 r = (rm * z.neg_inv_q_mod_mtilde) & 0xFFFFFFFFull;
                for (u32 o = 0; o < nBsk; o++) {
                    const ModDev &m = mods[map.idx[L + o]];
                    u64 lo = 0, hi = 0;
                    for (u32 i = 0; i < L; i++) mac128(lo, hi, y[i], z.q_to_bsk[o * L + i]);
                    u64 conv = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    u64 tmp = r;
                    if (tmp >= 0x80000000ull) tmp += m.p - 0x100000000ull;
                    // (tmp * prod_q + conv) mod p, then * m_tilde^-1
                    u64 l2 = conv, h2 = 0;
                    mac128(l2, h2, tmp, z.prod_q_mod_bsk[2 * o]);
                    u64 s = barrett128(l2, h2, m.p, m.ratio0, m.ratio1);
                    po[(u64)(L + o) * N] = shoup(s, z.inv_mtilde_mod_bsk[2 * o], z.inv_mtilde_mod_bsk[2 * o + 1], m.p);
                }
            }
        })
    }
};

// Steps (6)-(8): multiply by t, fastFloor, fastbconvSk  src/evaluator_cuda.cu:369-381, src/utils/rns_cuda.cu:365-421, 466-498
//   in  : [polys][L+nBsk][N]  (coefficient form, canonical, after INTT)     out : [polys][L][N]
struct BehzFloorSk {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "behz_floor_sk"; }
    int logN;
    const ModDev *mods;
    LimbMap map;
    BehzDev z;
    u64 total;
    const u64 *in;
    u64 *out;
    u32 polys_per_batch;
    u64 out_batch_stride;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 N = u64(1) << logN;
                u64 c = idx & (N - 1), poly = idx >> logN;
                const u32 L = z.L, nB = z.nB, nBsk = z.nBsk;
                const u64 *pi = in + poly * (L + nBsk) * N + c;
                u64 *po = out + (poly / polys_per_batch) * out_batch_stride + (poly % polys_per_batch) * L * N + c;
                u64 y[kBehzMax], w[kBehzMax];
                // fastFloor: q part times t, then * (Q/q_i)^-1
                for (u32 i = 0; i < L; i++) {
                    u64 p = mods[map.idx[i]].p;
                    u64 x = shoup(pi[(u64)i * N], z.t_mod[2 * i], z.t_mod[2 * i + 1], p);
                    y[i] = shoup(x, z.inv_punct_q[2 * i], z.inv_punct_q[2 * i + 1], p);
                }
                for (u32 o = 0; o < nBsk; o++) {
                    const ModDev &m = mods[map.idx[L + o]];
                    u64 lo = 0, hi = 0;
                    for (u32 i = 0; i < L; i++) mac128(lo, hi, y[i], z.q_to_bsk[o * L + i]);
                    u64 conv = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    u64 zt = shoup(pi[(u64)(L + o) * N], z.t_mod[2 * (L + o)], z.t_mod[2 * (L + o) + 1], m.p);
                    w[o] = shoup(zt + (m.p - conv), z.inv_prod_q_mod_bsk[2 * o], z.inv_prod_q_mod_bsk[2 * o + 1], m.p);
                }
                // fastbconvSk: B -> q and B -> m_sk
                const ModDev &msk = mods[map.idx[L + nB]];
                u64 yb[kBehzMax];
                u64 lo = 0, hi = 0;
                for (u32 b = 0; b < nB; b++) {
                    u64 pb = mods[map.idx[L + b]].p;
                    yb[b] = shoup(w[b], z.inv_punct_B[2 * b], z.inv_punct_B[2 * b + 1], pb);
                    mac128(lo, hi, yb[b], z.B_to_msk[b]);
                }
                u64 tsk = barrett128(lo, hi, msk.p, msk.ratio0, msk.ratio1);
                u64 alpha = shoup(tsk + (msk.p - w[nB]), z.inv_prod_B_mod_msk, z.inv_prod_B_mod_msk_shoup, msk.p);
                const bool neg = alpha > (msk.p >> 1);
                for (u32 i = 0; i < L; i++) {
                    const ModDev &m = mods[map.idx[i]];
                    u64 l2 = 0, h2 = 0;
                    for (u32 b = 0; b < nB; b++) mac128(l2, h2, yb[b], z.B_to_q[i * nB + b]);
                    u64 dest = barrett128(l2, h2, m.p, m.ratio0, m.ratio1);
                    u64 l3 = dest, h3 = 0;
                    if (neg) mac128(l3, h3, msk.p - alpha, z.prod_B_mod_q[2 * i]);
                    else mac128(l3, h3, alpha, z.neg_prod_B_mod_q[2 * i]);
                    po[(u64)i * N] = barrett128(l3, h3, m.p, m.ratio0, m.ratio1);
                }
            }
        })
    }
};

} // namespace troyb
