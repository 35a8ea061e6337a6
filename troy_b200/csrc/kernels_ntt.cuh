// NTT kernels (functors for launch_cta) with fused prologues / epilogues.
//
// One forward family and one inverse family cover every transform on the hot path; a runtime `mode`
// (uniform per launch) selects how a CTA maps to (source, destination, modulus) and which element-wise
// work is fused into the load and the store:
//
//   FWD_PLAIN     transformToNttInplace / kNttNegacyclicHarvey[Lazy]      src/kernelutils.cuh:578-603
//   FWD_MODUP     key-switch mod-up: digit j reduced mod q_i, then NTT_i   src/evaluator_cuda.cu:1231-1247
//   FWD_MODDOWN   CKKS mod-down / rescale tail: (s' mod q_j) - half, NTT_j, subtract, x q_last^-1, (+ ct)
//                                                                          src/evaluator_cuda.cu:1118-1161 (UtilF/G),
//                                                                          src/utils/rns_cuda.cu:307-363
//   INV_PLAIN     transformFromNttInplace / kInverseNttNegacyclicHarvey   src/kernelutils.cuh:607-632
//   INV_ADD_HALF  INTT of the limb being dropped, + floor(q_last/2) mod q_last   src/evaluator_cuda.cu:1327-1331
//   INV_COMBINE   BFV/BGV mod-down tail in coefficient form               src/evaluator_cuda.cu:1087-1161 (UtilE/F/G)
#pragma once
#include "ntt.cuh"
#if !defined(TROYB_HOST_EMUL)
#include <cuda.h>
#endif

namespace troyb {

enum NttMode {
    FWD_PLAIN = 0, FWD_MODUP = 1, FWD_MODDOWN = 2,
    INV_PLAIN = 10, INV_ADD_HALF = 11, INV_COMBINE = 12,
};

struct NttArgs {
#if !defined(TROYB_HOST_EMUL)
    // Tensor maps of the second-generation kernels (ntt2.cuh): the buffer the inputs live in and the buffer plain outputs go
    // to, each as a 2-D array [rows][16 words] with the 128-byte swizzle; a box is one staging segment / one 1024-coefficient
    // block. Units address them by row: (pointer - base) / 16.
    alignas(64) CUtensorMap tm_src;
    alignas(64) CUtensorMap tm_dst;
    const u64 *tm_src_base;
    const u64 *tm_dst_base;
#endif
    int mode;
    int logN;
    const ModDev *mods;
    LimbMap map;        // launch limb -> pool index
    u32 n_limbs;        // limbs per poly in the launch's unit order
    u64 *dst;
    const u64 *src;
    int canon;          // store canonical residues (else lazy)
    // --- key-switch / drop-last extras
    const u64 *aux0;    // MODUP: original NTT-form target (CKKS) or null; MODDOWN/COMBINE: accumulators (t_poly_prod)
    const u64 *aux1;    // MODDOWN/COMBINE: s' (rounded special limb, coefficient form)
    u64 *ct;            // MODDOWN/COMBINE: ciphertext updated in place (null for rescale -> write dst)
    const u64 *ct_src;  // MODDOWN/COMBINE: when set, ct = ct_src + result for polys k < ct_src_polys and ct = result for the
                        // others (applyGalois: the permuted c0 lives in scratch, c1 starts from zero), instead of ct += result
    u64 ct_src_batch_stride;
    u32 ct_src_polys;
    DropDev drop;       // constants for the prime being dropped
    u32 L;              // data limbs in play (MODUP: digits in this launch)
    u32 digit0;         // MODUP: index of the first digit of this launch
    u32 K;              // polys per batch element handled (2 for key-switch, ct size for rescale)
    u64 src_unit_stride;  // words between batch elements of src
    u64 src_poly_stride;  // words between polys of one batch element in src (INV_ADD_HALF)
    u64 src_offset;       // word offset of the first source unit
    u64 aux0_batch_stride;
    u64 acc_poly_stride;  // words between polys in aux0
    u64 acc_batch_stride;
    u64 acc_limb_stride;  // words between limbs in aux0 (0 = N: [batch][2][L+1][N]; rows [L+1][batch][2][N] otherwise)
    u64 ct_poly_stride;
    u64 ct_batch_stride;
    u64 dst_poly_stride;
    u64 dst_batch_stride;
    int scheme;
    int add_half;       // INV_ADD_HALF: add half (BFV/CKKS) or plain canonical INTT (BGV)
    uint8_t order[kMaxLimbs]; // launch rank -> limb index: integer-pipe and FP64-pipe limbs interleaved (see set_order)
    u32 n_rank;         // limb ranks in the launch; consecutive units cycle through them
    u32 pf_ahead;       // L2 prefetch distance in CTAs (0 = off): CTA b asks the L2 for the source tile of CTA b + pf_ahead
    unsigned long long *dbg; // development builds (-DTROYB_DEV_TIMELINE, TROYB_TIMELINE=file): thread 0 of every CTA records globaltimer at its phase boundaries
    // --- fused mod-up + key inner product (second-generation kernels only): instead of storing the transformed digit, the CTA
    // multiplies it by the key's two components and reduce-adds the products into the accumulators (TMA reduce, u64 sums of
    // canonical residues; the caller zeroes the accumulators first and reduces them modulo each prime afterwards)
    int fused;
    const u64 *key;       // key + digit0 digits: [digit][2][key_limbs][N]
    u64 key_digit_stride, key_comp_stride;
    u32 key_special_limb; // key limb of the special prime (launch limb L_full)
    u64 *prod;            // accumulators [batch][2][L_full + 1][N] (strides below)
    u64 prod_batch_stride, prod_poly_stride, prod_limb_stride;
    int solo;           // forward, limb = 2 tiles: independent CTAs that both recompute layer 0 (FwdSoloLd). Launched without a
                        // cluster when the transform is out of place; in place, as a cluster of two whose only exchange is one
                        // split barrier "my loads are done" before anybody stores (run_fwd)
};

TB_D u64 canon4(u64 v, u64 p) {
    u64 p2 = p << 1;
    v -= (v >= p2 ? p2 : 0);
    v -= (v >= p ? p : 0);
    return v;
}
TB_D u64 canon2(u64 v, u64 p) { return v - (v >= p ? p : 0); }

// Generic small transform (N < 4096; only the reference's toy-parameter tests use it): radix-2 in shared memory.
template <class Ld, class St>
TB_D void small_ntt_fwd(u64 *sm, u32 logn, Ld ld, St st, const u64 *__restrict__ tw, u64 p) {
    const u32 n = 1u << logn;
    const u64 p2 = p << 1;
    TB_PHASE(256, t, for (u32 i = t; i < n; i += 256) sm[i] = ld(i);)
    for (u32 s = 0; s < logn; s++) {
        const u32 lg = logn - 1 - s;
        TB_PHASE(256, t, for (u32 b = t; b < n / 2; b += 256) {
            u32 r = b >> lg, pos = b & ((1u << lg) - 1);
            u32 i = (r << (lg + 1)) | pos;
            ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(tw) + ((1u << s) + r));
            u64 X = sm[i], Y = sm[i + (1u << lg)];
            { u64 x_ = X - (X >= p2 ? p2 : 0); u64 t_ = shoup_lazy(Y, w.x, w.y, p); X = x_ + t_; Y = x_ - t_ + p2; }
            sm[i] = X;
            sm[i + (1u << lg)] = Y;
        })
    }
    TB_PHASE(256, t, for (u32 i = t; i < n; i += 256) st(i, canon4(sm[i], p));)
}

template <class Ld, class St>
TB_D void small_ntt_inv(u64 *sm, u32 logn, Ld ld, St st, const u64 *__restrict__ tw, u64 p, const ModDev *md) {
    const u32 n = 1u << logn;
    const u64 p2 = p << 1;
    TB_PHASE(256, t, for (u32 i = t; i < n; i += 256) sm[i] = ld(i);)
    for (int s = (int)logn - 1; s >= 0; s--) {
        const u32 lg = logn - 1 - s;
        TB_PHASE(256, t, for (u32 b = t; b < n / 2; b += 256) {
            u32 r = b >> lg, pos = b & ((1u << lg) - 1);
            u32 i = (r << (lg + 1)) | pos;
            u64 X = sm[i], Y = sm[i + (1u << lg)];
            if (s == 0) {
                { u64 s_ = X + Y, d_ = X + p2 - Y; X = shoup_lazy(s_, md->ninv, md->ninv_shoup, p); Y = shoup_lazy(d_, md->inv1n, md->inv1n_shoup, p); }
            } else {
                ulonglong2 w = __ldg(reinterpret_cast<const ulonglong2 *>(tw) + ((1u << s) + r));
                { u64 s_ = X + Y, d_ = X + p2 - Y; X = s_ - (s_ >= p2 ? p2 : 0); Y = shoup_lazy(d_, w.x, w.y, p); }
            }
            sm[i] = X;
            sm[i + (1u << lg)] = Y;
        })
    }
    TB_PHASE(256, t, for (u32 i = t; i < n; i += 256) st(i, canon2(sm[i], p));)
}

inline const char *ntt_kernel_name(int mode) {
    switch (mode) {
    case FWD_PLAIN: return "ntt_fwd";
    case FWD_MODUP: return "ntt_fwd_modup";
    case FWD_MODDOWN: return "ntt_fwd_moddown";
    case INV_PLAIN: return "ntt_inv";
    case INV_ADD_HALF: return "ntt_inv_add_half";
    default: return "ntt_inv_combine";
    }
}

// Per-CTA resolved job: where to read / write and with which modulus.
struct NttUnit {
    const ModDev *md;
    const u64 *src;
    u64 *dst;
    const u64 *acc;   // accumulators row
    const u64 *sp;    // s' row
    u64 *ct;          // ciphertext row (written)
    const u64 *ct_in; // the row added to the result (ct itself, a row of NttArgs::ct_src, or null for zero)
    DropLimb dl;
    const u64 *key0;  // fused MODUP: row of key component 0 for (digit, output modulus); component 1 is key_comp_stride further
    u64 *prod0;       // fused MODUP: accumulator row of component 0 for (batch element, output modulus)
    const u64 *pass;  // fused MODUP, pass-through unit: the NTT-form limb that stands in for the transform's output
    bool passthrough; // MODUP with i == j in CKKS: copy the NTT-form limb
    bool reduce_src;  // MODUP: the digit must be reduced mod q_i before the transform
    u32 row;          // four-step row
    u32 rows_log;
    u32 unit;
};

// LOGT == 0 selects the generic small-N path (whole limb, runtime logN).
template <int LOGT> TB_D NttUnit resolve_unit(const NttArgs &a, u32 bid) {
    NttUnit u{};
    const u32 rows_log = LOGT == 0 ? 0 : (u32)(a.logN - LOGT);
    const u32 row = bid & ((1u << rows_log) - 1);
    const u32 unit = bid >> rows_log;
    const u64 N = u64(1) << a.logN;
    u.row = row;
    u.rows_log = rows_log;
    u.unit = unit;
    u.passthrough = false;
    switch (a.mode) {
    case FWD_PLAIN:
    case INV_PLAIN: {
        // storage: unit = b * n_limbs + l, l = poly * L + limb (n_limbs = polys x L); launch order: limb rank fastest
        const u32 polys = a.n_limbs / a.L;
        const u32 limb = a.order[unit % a.n_rank], bp = unit / a.n_rank;
        const u32 l = (bp % polys) * a.L + limb, b = bp / polys;
        u.unit = b * a.n_limbs + l;
        u.md = a.mods + a.map.idx[l % a.L];
        u.src = a.src + (u64)b * a.src_unit_stride + (u64)l * N;
        u.dst = a.dst + (u64)b * a.dst_batch_stride + (u64)l * N;
        break;
    }
    case FWD_MODUP: {
        // storage: unit = (b * (Lfull+1) + i) * L + j; launch order: (b, j, rank of i): the transforms of one digit are
        // neighbours, so the digit is read from HBM once
        const u32 i = a.order[unit % a.n_rank], bj = unit / a.n_rank, j = bj % a.L, b = bj / a.L;
        u.unit = (b * a.n_rank + i) * a.L + j;
        u.md = a.mods + a.map.idx[i];
        u.src = a.src + (u64)b * a.src_unit_stride + (u64)j * N;
        u.dst = a.dst + (u64)u.unit * N;
        if (a.aux0 != nullptr && i == j + a.digit0) u.passthrough = true; // the inner product reads the NTT-form limb in place
        if (a.fused) {
            const u32 key_limb = (i + 1 == a.n_rank) ? a.key_special_limb : i;
            u.key0 = a.key + (u64)j * a.key_digit_stride + (u64)key_limb * N;
            u.prod0 = a.prod + (u64)b * a.prod_batch_stride + (u64)i * a.prod_limb_stride;
            u.pass = u.passthrough ? a.aux0 + (u64)b * a.aux0_batch_stride + (u64)i * N : nullptr;
        }
        // a digit whose own modulus is below 4 q_i is already a valid lazy input of the forward transform (every
        // butterfly policy accepts [0, 4p)): its Barrett reduction mod q_i is skipped (results are canonical either way)
        u.reduce_src = (a.mods[a.map.idx[j + a.digit0]].p >> 2) >= u.md->p;
        break;
    }
    case INV_ADD_HALF: {
        // unit = b * K + k
        u.md = a.mods + a.drop.last_pool;
        u.src = a.src + a.src_offset + (u64)(unit / a.K) * a.src_unit_stride + (u64)(unit % a.K) * a.src_poly_stride;
        u.dst = a.dst + (u64)unit * N;
        break;
    }
    case FWD_MODDOWN:
    case INV_COMBINE: {
        // storage: unit = (b * K + k) * L + j; launch order: (b * K + k, rank of j)
        const u32 j = a.order[unit % a.n_rank], bk = unit / a.n_rank, k = bk % a.K, b = bk / a.K;
        u.unit = bk * a.L + j;
        u.md = a.mods + a.map.idx[j];
        u.sp = a.aux1 + (u64)bk * N;
        u.acc = a.aux0 + (u64)b * a.acc_batch_stride + (u64)k * a.acc_poly_stride + (u64)j * (a.acc_limb_stride ? a.acc_limb_stride : N);
        u.ct = a.ct ? a.ct + (u64)b * a.ct_batch_stride + (u64)k * a.ct_poly_stride + (u64)j * N : nullptr;
        if (a.ct_src == nullptr) u.ct_in = u.ct;
        else u.ct_in = k < a.ct_src_polys ? a.ct_src + (u64)b * a.ct_src_batch_stride + (u64)k * a.ct_poly_stride + (u64)j * N : nullptr;
        u.dst = a.dst ? a.dst + (u64)b * a.dst_batch_stride + (u64)k * a.dst_poly_stride + (u64)j * N : nullptr;
        u.dl = a.drop.limbs[j];
        u.src = (a.mode == FWD_MODDOWN) ? u.sp : u.acc;
        break;
    }
    }
    return u;
}

// L2 prefetch (GPU only; thread 0 of the CTA). The CTA that will run about one wave later finds its source tile in the
// L2 instead of HBM, and this CTA's own epilogue operands (accumulator and ciphertext rows, read in the store phase at
// the very end) are requested now.
template <int LOGT> TB_D void prefetch_tiles(const NttArgs &a, const NttUnit &u, u32 bid, u32 rank, u32 grid) {
#if !defined(TROYB_HOST_EMUL)
    if (a.pf_ahead == 0 || threadIdx.x != 0) return;
    constexpr u32 bytes = (8u << LOGT);
    const u64 off = (u64)rank << LOGT;
    if (a.mode == FWD_MODDOWN || a.mode == INV_COMBINE) {
        if (a.mode == INV_COMBINE) l2_prefetch(u.sp + off, bytes);
        else l2_prefetch(u.acc + off, bytes);
        if (u.ct_in) l2_prefetch(u.ct_in + off, bytes);
    }
    const u32 ahead = bid + a.pf_ahead;
    if (ahead < grid) {
        // sources shared by the n_rank consecutive units of one digit / one poly are requested once
        const bool shared_src = a.mode == FWD_MODUP || a.mode == FWD_MODDOWN;
        if (shared_src && ((ahead >> u.rows_log) % a.n_rank) != 0) return;
        const NttUnit v = resolve_unit<LOGT>(a, ahead);
        l2_prefetch(v.src + off, bytes);
    }
#endif
}

#define TB_WITH_LOGC(logc, ...)                                  \
    switch (logc) {                                              \
    case 1: { constexpr int LOGC = 1; __VA_ARGS__ } break;       \
    case 2: { constexpr int LOGC = 2; __VA_ARGS__ } break;       \
    default: { constexpr int LOGC = 3; __VA_ARGS__ } break;      \
    }

// Prologue of a forward transform applied to the raw coefficient x.
struct FwdPrologue {
    u64 p, ratio1, fix;
    int mode;
    TB_D u64 operator()(u64 x) const {
        if (mode == FWD_MODUP) x = barrett64(x, p, ratio1);
        else if (mode == FWD_MODDOWN) x = barrett64(x, p, ratio1) + fix;
        return x;
    }
};

// Staged loaders (ntt.cuh IsStagedLd) of a forward tile transform's first round.
// The tile is the whole limb: input i is coefficient i.
template <class A> struct FwdTileLd {
    typedef u64 Raw;
    const u64 *src;
    FwdPrologue pro;
    typename A::Mod m;
    TB_D Raw raw(u32 i) const { return src[i]; }
    template <int K> TB_D void cook(const Raw (&r)[K], typename A::elem (&x)[K]) const {
        if (pro.mode != FWD_PLAIN) {
#pragma unroll
            for (int k = 0; k < K; k++) x[k] = A::from_u64(pro(r[k]), m);
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) x[k] = A::from_u64(r[k], m);
        }
    }
};
// Solo tiles (limb = 2 tiles, no cluster): the CTA of tile `upper` recomputes layer 0 of the limb's transform, which couples
// coefficient i with i + T, from both inputs and keeps its own side: one extra modular multiply per coefficient instead of a
// scatter stage through distributed shared memory, a cluster barrier and co-scheduling of the two CTAs.
template <class A, int LOGT> struct FwdSoloLd {
    struct Raw { u64 a, b; };
    const u64 *src; // coefficient 0 of the limb
    FwdPrologue pro;
    typename A::Mod m;
    typename A::tw_t w1;
    bool upper;
    TB_D Raw raw(u32 i) const { return Raw{src[i], src[i + (1u << LOGT)]}; }
    template <int K> TB_D void cook(const Raw (&r)[K], typename A::elem (&x)[K]) const {
        u64 a[K], b[K];
#pragma unroll
        for (int k = 0; k < K; k++) { a[k] = r[k].a; b[k] = r[k].b; }
        if (pro.mode != FWD_PLAIN) {
#pragma unroll
            for (int k = 0; k < K; k++) { a[k] = pro(a[k]); b[k] = pro(b[k]); }
        }
#pragma unroll
        for (int k = 0; k < K; k++) x[k] = ct_half<A>(A::from_u64(a[k], m), A::from_u64(b[k], m), w1, m, upper);
    }
};

// Epilogue of a forward transform: v is the canonical residue of output i of the limb.
struct FwdStore {
    NttUnit u;
    u64 p;
    int mode;
    struct Pre { u64 acc, ct; }; // global operands of output i, loaded ahead of the store loop (tile_store_canon)
    TB_D Pre pre(u32 i) const {
        Pre r{0, 0};
        if (mode == FWD_MODDOWN) {
            r.acc = u.acc[i];
            if (u.ct_in) r.ct = u.ct_in[i];
        }
        return r;
    }
    // (acc - v) * q_last^-1: the Shoup multiply takes any 64-bit operand, so acc may be an unreduced sum of canonical residues
    // (fused key products, sums over GPUs) as long as acc + p fits 64 bits
    TB_D u64 finish(u64 acc, u64 v) const { return shoup(acc + p - v, u.dl.inv_last, u.dl.inv_last_shoup, p); }
    TB_D void operator()(u32 i, u64 v, const Pre &q) const {
        if (mode == FWD_MODDOWN) {
            u64 r = finish(q.acc, v);
            if (u.ct) u.ct[i] = addmod(q.ct, r, p);
            else u.dst[i] = r;
        } else {
            u.dst[i] = v;
        }
    }
    TB_D void operator()(u32 i, u64 v) const { (*this)(i, v, pre(i)); }
};

// Epilogue of a completed inverse transform (v = canonical residue of coefficient idx of the limb).
TB_D void inv_store_final(const NttArgs &a, const NttUnit &u, u64 p, u64 idx, u64 v) {
    const int mode = a.mode;
    const DropDev &dr = a.drop;
    if (mode == INV_PLAIN) {
        u.dst[idx] = v;
    } else if (mode == INV_ADD_HALF) {
        u.dst[idx] = a.add_half ? addmod(v, dr.half, p) : v;
    } else { // INV_COMBINE
        u64 s = u.sp[idx];
        u64 sub;
        const u64 ratio1 = u.md->ratio1;
        if (a.scheme == SCHEME_BGV) {
            // BGV (UtilE): k = -(s mod t) * q_last^-1 mod t ; delta = (k mod q_j) * q_last mod q_j ; c = s mod q_j
            u64 kk = negmod(barrett64(s, dr.t, dr.t_ratio1), dr.t);
            kk = shoup(kk, dr.inv_last_mod_t, dr.inv_last_mod_t_shoup, dr.t);
            u64 delta = shoup(barrett64(kk, p, ratio1), u.dl.last_mod, u.dl.last_mod_shoup, p);
            sub = delta + barrett64(s, p, ratio1);
        } else {
            // BFV (UtilF/G): (s' mod q_j) + q_j - half mod q_j
            sub = barrett64(s, p, ratio1) + (p - u.dl.half_mod);
        }
        u64 d = v + (p << 1) - sub; // v < p, sub <= 2p  ->  d in (0,3p)
        u64 r = shoup(d, u.dl.inv_last, u.dl.inv_last_shoup, p);
        if (u.ct) u.ct[idx] = addmod(u.ct_in ? u.ct_in[idx] : 0, r, p);
        else u.dst[idx] = r;
    }
}

// ---- small N (< 4096; only the reference's toy-parameter tests use it): one CTA per limb, radix-2 -------------------------
struct NttFwdSmall {
    static constexpr int kThreads = 256;
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D void operator()(u32 bid, u64 *sm) const {
        const NttUnit u = resolve_unit<0>(a, bid);
        const ModDev *md = u.md;
        const u32 n = 1u << a.logN;
        if (u.passthrough) return; // digit already in NTT form under this modulus: the inner product reads it in place
        const FwdPrologue pro{md->p, md->ratio1, a.mode == FWD_MODDOWN ? md->p - u.dl.half_mod : 0, a.mode};
        auto ld = [&](u32 i) -> u64 { return pro(u.src[i]); };
        const FwdStore st{u, md->p, a.mode};
        small_ntt_fwd(sm, (u32)a.logN, ld, st, md->fwd, md->p);
    }
};

struct NttInvSmall {
    static constexpr int kThreads = 256;
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D void operator()(u32 bid, u64 *sm) const {
        const NttUnit u = resolve_unit<0>(a, bid);
        const ModDev *md = u.md;
        const u64 p = md->p;
        auto ld = [&](u32 i) -> u64 { return u.src[i]; };
        auto st = [&](u32 i, u64 v) { inv_store_final(a, u, p, i, v); };
        small_ntt_inv(sm, (u32)a.logN, ld, st, md->inv, p, md);
    }
};

// ---- tile / cluster transforms (N >= 4096) ---------------------------------------------------------------------------
// One CTA per tile of 2^LOGT coefficients; a limb of N = C tiles is one thread-block cluster (launch_cluster). See
// ntt.cuh "cluster stages" for how the layers are split between the tiles' own transforms and the exchange stages.
// Runs `body` with the butterfly policy `pol` (warp-uniform branch).
#define TB_WITH_POLICY(pol, ...)                                                    \
    switch (pol) {                                                         \
    case POLICY_F64: { typedef ArithF64 A; __VA_ARGS__ } break;                     \
    case POLICY_F64_SMALL: { typedef ArithF64S A; __VA_ARGS__ } break;              \
    case POLICY_F64_LARGE: { typedef ArithF64L A; __VA_ARGS__ } break;              \
    case POLICY_U64_FAST: { typedef ArithU64Fast A; __VA_ARGS__ } break;            \
    case POLICY_U64_LAZY: { typedef ArithU64Lazy A; __VA_ARGS__ } break;            \
    default: { typedef ArithU64Exact A; __VA_ARGS__ } break;                        \
    }

template <int LOGT> struct NttFwdTile {
    static constexpr int kThreads = TilePlan<LOGT>::NT;
    static constexpr int kMinCtas = TilePlan<LOGT>::CTAS;
    static constexpr int kStages = 3; // scatter | tile transform | canonical store (no cluster barrier between the last two)
    static constexpr bool kTrailingBarrier = false; // peers only write into this CTA's tile, and only in stage 0
    static constexpr u32 T = 1u << LOGT;
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D bool barrier_after(int S) const { return S == 0 && a.solo == 0; } // solo tiles exchange nothing through shared memory

    template <int S> TB_D void stage(u32 bid, u32 grid, u32 rank, u32 nranks, u64 *sm, const Peers &peers) const {
        NttUnit u = resolve_unit<LOGT>(a, bid);
        const ModDev *md = u.md;
        const int pol = md->policy;
        const FwdPrologue pro{md->p, md->ratio1, a.mode == FWD_MODDOWN ? md->p - u.dl.half_mod : 0,
                              (a.mode == FWD_MODUP && !u.reduce_src) ? (int)FWD_PLAIN : a.mode};
        // solo tiles: a limb of two tiles launched without a cluster (run_fwd), layer 0 recomputed by both CTAs (FwdSoloLd)
        const bool solo = a.solo != 0;
        const u32 trank = solo ? u.row : rank;
        if constexpr (S == 0) {
            prefetch_tiles<LOGT>(a, u, bid, trank, grid);
            // scatter stage: the layers that couple the tiles, results into the peers' shared memory
            if (nranks == 1 || solo || u.passthrough) return;
            cluster_arrive_started();
            auto started = [] { cluster_wait_started(); };
            auto ld = [&](u32 i) -> u64 { return u.src[i]; };
            u32 handle[8];
#pragma unroll
            for (int k = 0; k < 8; k++) handle[k] = peers.map((u32)k < nranks ? (u32)k : 0u);
            auto st_peer = [&](u32 k, u32 pi, u64 bits) { peers.st(handle[k], pi, bits); };
            TB_WITH_POLICY(pol, TB_WITH_LOGC(u.rows_log, TB_PHASE(kThreads, t, (cluster_fwd_scatter<A, LOGC, LOGT, kThreads>(t, rank, ld, pro, st_peer, started, md));)))
        } else if constexpr (S == 1) {
            if (u.passthrough) return; // digit already in NTT form under this modulus: the inner product reads it in place
            if (solo) {
                TB_WITH_POLICY(pol, {
                    const FwdSoloLd<A, LOGT> gld{u.src, pro, A::make(md), A::ld_tw(A::fwd_table(md), 1), trank != 0};
                    // in place (cluster of two): the peer may overwrite this CTA's inputs once both have loaded theirs
                    auto loaded = [&] { if (nranks > 1) cluster_arrive_started(); };
                    tile_ntt_fwd<A, LOGT, true>(sm, gld, md, 2u + trank, loaded);
                })
            } else if (nranks == 1) {
                TB_WITH_POLICY(pol, {
                    const FwdTileLd<A> gld{u.src, pro, A::make(md)};
                    tile_ntt_fwd<A, LOGT, true>(sm, gld, md, 1u);
                })
            } else {
                TB_WITH_POLICY(pol, (tile_ntt_fwd<A, LOGT, false>(sm, 0, md, nranks + rank));)
            }
        } else {
            if (u.passthrough) return;
            if (solo && nranks > 1) cluster_wait_started();
            const u64 off = (u64)trank << LOGT;
            // this tile's outputs are coefficients [off, off + T) of the limb's (bit-reversed) output
            if (u.dst) u.dst += off;
            if (u.ct) u.ct += off;
            if (u.ct_in) u.ct_in += off;
            if (u.acc) u.acc += off;
            const FwdStore st{u, md->p, a.mode};
            TB_WITH_POLICY(pol, (tile_store_canon<A, LOGT>(sm, st, md));)
        }
    }
};

template <int LOGT> struct NttInvTile {
    static constexpr int kThreads = TilePlan<LOGT>::NT;
    static constexpr int kMinCtas = TilePlan<LOGT>::CTAS;
    static constexpr int kStages = 2;
    static constexpr bool kTrailingBarrier = true; // peers read this CTA's tile in the gather stage
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D bool barrier_after(int) const { return true; }

    template <int S> TB_D void stage(u32 bid, u32 grid, u32 rank, u32 nranks, u64 *sm, const Peers &peers) const {
        const NttUnit u = resolve_unit<LOGT>(a, bid);
        const ModDev *md = u.md;
        const u64 p = md->p;
        const int pol = md->policy;
        auto st = [&](u32 i, u64 v) { inv_store_final(a, u, p, i, v); };
        if constexpr (S == 0) {
            prefetch_tiles<LOGT>(a, u, bid, rank, grid);
            const u64 *in = u.src + ((u64)rank << LOGT);
            auto gld = [&](u32 i) -> u64 { return in[i]; };
            if (nranks == 1) {
                TB_WITH_POLICY(pol, (tile_ntt_inv<A, LOGT, true>(sm, gld, st, md, 1u));)
            } else {
                TB_WITH_POLICY(pol, (tile_ntt_inv<A, LOGT, false>(sm, gld, 0, md, nranks + rank));)
            }
        } else {
            if (nranks == 1) return;
            u32 handle[8];
#pragma unroll
            for (int k = 0; k < 8; k++) handle[k] = peers.map((u32)k < nranks ? (u32)k : 0u);
            auto ld_peer = [&](u32 k, u32 pi) -> u64 { return peers.ld(handle[k], pi); };
            TB_WITH_POLICY(pol, TB_WITH_LOGC(u.rows_log, TB_PHASE(kThreads, t, (cluster_inv_gather<A, LOGC, LOGT, kThreads>(t, rank, ld_peer, st, md));)))
        }
    }
};

} // namespace troyb
