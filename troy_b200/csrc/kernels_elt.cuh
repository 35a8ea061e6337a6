// Coefficient-wise kernels of the RNS polynomial path (functors for launch_cta).
// Every kernel here is HBM-bound integer work: one thread owns one coefficient index of one limb (or of all
// limbs when the operation couples limbs, as the base conversions do), loads are coalesced along the coefficient
// axis, and everything that the reference does in several passes over global memory is done in one.
#pragma once
#include "launch.cuh"

namespace troyb {

constexpr int kEltThreads = 256;

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
    int op; // 0 add, 1 sub, 2 negate, 3 multiply by `scalar`, 4 reduce modulo the prime
    u64 scalar;
    int logN;
    const ModDev *mods;
    LimbMap map;
    u32 n_limbs;
    u64 total; // words
    const u64 *a, *b;
    u64 *out;
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                u32 limb = (u32)((idx >> logN) % n_limbs);
                const ModDev &m = mods[map.idx[limb]];
                u64 p = m.p;
                u64 x = a[idx];
                u64 r;
                if (op == 0) r = addmod(x, b[idx], p);
                else if (op == 1) r = submod(x, b[idx], p);
                else if (op == 2) r = negmod(x, p);
                else if (op == 4) r = barrett64(x, p, m.ratio1);
                else r = mulmod_barrett(x, barrett64(scalar, p, m.ratio1), p, m.ratio0, m.ratio1);
                out[idx] = r;
            }
        })
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
    u64 total;        // batch * n_limbs * N
    const u64 *a, *b;
    u64 *out;
    u64 a_batch_stride, b_batch_stride, out_batch_stride; // words
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                u64 bt = idx / per_batch, rem = idx % per_batch;
                u32 limb = (u32)(rem >> logN);
                const ModDev &m = mods[map.idx[limb]];
                const u64 *pa = a + bt * a_batch_stride + rem;
                const u64 *pb = b + bt * b_batch_stride + rem;
                u64 *po = out + bt * out_batch_stride + rem;
                if (a_size == 2 && b_size == 2) {
                    u64 a0 = pa[0], a1 = pa[per_batch], b0 = pb[0], b1 = pb[per_batch];
                    u64 lo = 0, hi = 0;
                    mac128(lo, hi, a0, b0);
                    u64 c0 = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    lo = 0; hi = 0;
                    mac128(lo, hi, a0, b1);
                    mac128(lo, hi, a1, b0);
                    u64 c1 = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    lo = 0; hi = 0;
                    mac128(lo, hi, a1, b1);
                    u64 c2 = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    po[0] = c0; po[per_batch] = c1; po[2 * per_batch] = c2;
                } else {
                    u64 av[8], bv[8], cv[15];
                    for (u32 j = 0; j < a_size; j++) av[j] = pa[j * per_batch];
                    for (u32 k = 0; k < b_size; k++) bv[k] = pb[k * per_batch];
                    for (u32 i = 0; i < a_size + b_size - 1; i++) {
                        u64 lo = 0, hi = 0;
                        for (u32 j = 0; j < a_size; j++) {
                            if (i >= j && i - j < b_size) mac128(lo, hi, av[j], bv[i - j]);
                        }
                        cv[i] = barrett128(lo, hi, m.p, m.ratio0, m.ratio1);
                    }
                    for (u32 i = 0; i < a_size + b_size - 1; i++) po[i * per_batch] = cv[i];
                }
            }
        })
    }
};

// ---- key-switch inner product -----------------------------------------------------------------------------
// gSwitchKeyInplaceUtilB + UtilD  src/evaluator_cuda.cu:1034-1085: acc[k] = sum_j D[i][j] * key[j][k][key_index]
// with 128-bit lazy accumulators kept in registers, one Barrett-128 at the end.
//   D    : [batch][L+1][nd][N]  mod-up output of digits digit0..digit0+nd-1 (canonical)
//   key  : [digits][2][key_limbs][N]
//   prod : [batch][2][L+1][N]  (t_poly_prod layout, §A.8)
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
                u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0, lo3 = 0, hi3 = 0;
                const u64 *ps = (pass != nullptr && i < L) ? pass + b * pass_stride + (u64)i * N + c : nullptr;
                for (u32 j0 = 0; j0 < nd; j0 += 4) { // four digits' loads in flight
                    ulonglong2 x[4], k0[4], k1[4];
                    _Pragma("unroll") for (u32 q = 0; q < 4; q++) {
                        const u32 j = j0 + q;
                        if (j < nd) {
                            const u64 *src = (ps != nullptr && j + digit0 == i) ? ps : d + (u64)j * N;
                            x[q] = *reinterpret_cast<const ulonglong2 *>(src);
                            k0[q] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + j * key_digit_stride));
                            k1[q] = __ldg(reinterpret_cast<const ulonglong2 *>(kp + j * key_digit_stride + key_comp_stride));
                        }
                    }
                    _Pragma("unroll") for (u32 q = 0; q < 4; q++) {
                        if (j0 + q < nd) {
                            mac128(lo0, hi0, x[q].x, k0[q].x);
                            mac128(lo1, hi1, x[q].y, k0[q].y);
                            mac128(lo2, hi2, x[q].x, k1[q].x);
                            mac128(lo3, hi3, x[q].y, k1[q].y);
                        }
                    }
                }
                u64 *po = prod + (b * 2 * (L + 1) + i) * N + c;
                ulonglong2 r0, r1;
                r0.x = barrett128(lo0, hi0, m.p, m.ratio0, m.ratio1);
                r0.y = barrett128(lo1, hi1, m.p, m.ratio0, m.ratio1);
                r1.x = barrett128(lo2, hi2, m.p, m.ratio0, m.ratio1);
                r1.y = barrett128(lo3, hi3, m.p, m.ratio0, m.ratio1);
                *reinterpret_cast<ulonglong2 *>(po) = r0;
                *reinterpret_cast<ulonglong2 *>(po + (u64)(L + 1) * N) = r1;
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
