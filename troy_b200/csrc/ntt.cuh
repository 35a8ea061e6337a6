// Negacyclic NTT building blocks for sm_100a.
//
// What is computed (bit-exact function of the reference):
//   forward : kNttNegacyclicHarvey      src/kernelutils.cuh:578-603, CPU src/utils/ntt.cpp:158-187,
//             DWTHandler::transformToRev src/utils/dwthandler.h:88-204   (natural in, bit-reversed out)
//   inverse : kInverseNttNegacyclicHarvey src/kernelutils.cuh:607-632, CPU src/utils/ntt.cpp:189-213,
//             DWTHandler::transformFromRev src/utils/dwthandler.h:215-372 (bit-reversed in, natural out, x N^-1)
// How it is computed is new: instead of one kernel launch per butterfly layer with one thread per
// coefficient, a CTA keeps a tile of 2^LOGT coefficients in (padded) shared memory and every thread runs
// radix-16 register butterflies (4 layers per shared-memory round trip). A tile is either a whole limb
// (N <= 16384: one HBM read + one HBM write per transform) or a 4096-coefficient row of the four-step
// split used for N = 32768 / 65536 (a strided "column" pass does the first log2(N/4096) layers).
//
// Twiddle index rule (same table order as the reference's root_powers_): for a tile whose sub-transform
// has root base R (R = 1 for a whole limb, R = N/T + row for row `row` of the four-step split), layer s of
// the tile (0-based) and lower butterfly element i use table entry (R << s) + (i >> (LOGT - s)).
#pragma once
#include "launch.cuh"

namespace troyb {

// Shared-memory padding: one extra word per 16 keeps every radix round and the linear copy phases
// bank-conflict free for 64-bit accesses.
__host__ __device__ constexpr u32 pad_idx(u32 i) { return i + (i >> 4); }
__host__ __device__ constexpr u32 padded_words(u32 n) { return n + (n >> 4); }

// Cooley-Tukey (forward) Harvey butterfly. X in [0,4p), Y in [0,4p) -> both in [0,4p).
TB_D void ct_bfly(u64 &X, u64 &Y, u64 w, u64 wq, u64 p, u64 p2) {
    u64 x = X - (X >= p2 ? p2 : 0);
    u64 t = shoup_lazy(Y, w, wq, p);
    X = x + t;
    Y = x - t + p2;
}

// Gentleman-Sande (inverse) Harvey butterfly. X, Y in [0,2p) -> both in [0,2p).
TB_D void gs_bfly(u64 &X, u64 &Y, u64 w, u64 wq, u64 p, u64 p2) {
    u64 s = X + Y;
    u64 d = X + p2 - Y;
    X = s - (s >= p2 ? p2 : 0);
    Y = shoup_lazy(d, w, wq, p);
}

// Last inverse layer with the N^-1 scaling folded in (like the reference CPU path, dwthandler.h:289-330).
TB_D void gs_bfly_last(u64 &X, u64 &Y, u64 ninv, u64 ninvq, u64 w, u64 wq, u64 p, u64 p2) {
    u64 s = X + Y;
    u64 d = X + p2 - Y;
    X = shoup_lazy(s, ninv, ninvq, p); // s < 4p < 2^64
    Y = shoup_lazy(d, w, wq, p);
}

TB_D ulonglong2 ld_tw(const u64 *tw, u32 k) { return __ldg(reinterpret_cast<const ulonglong2 *>(tw) + k); }

// One radix-2^RR round of the forward transform on a tile of 2^LOGT coefficients, layers S0..S0+RR-1.
// `vt` is this thread's index among NT cooperating threads (NT * 16 == 2^LOGT); each thread owns
// G = 16 >> RR groups of 2^RR coefficients spaced 2^BLO apart.
template <int LOGT, int S0, int RR, int NT, class Ld, class St>
TB_D void fwd_round(u32 vt, Ld ld, St st, const u64 *__restrict__ tw, u32 R, u64 p, u64 p2) {
    constexpr int E = 16, G = E >> RR, K = 1 << RR, BLO = LOGT - S0 - RR;
    static_assert(NT * E == (1 << LOGT), "tile/threads mismatch");
    static_assert(BLO >= 0, "bad round");
    u64 x[E];
#pragma unroll
    for (int g = 0; g < G; g++) {
        u32 v = vt + g * NT;
        u32 lo = v & ((1u << BLO) - 1), hi = v >> BLO;
        u32 base = (hi << (BLO + RR)) | lo;
#pragma unroll
        for (int k = 0; k < K; k++) x[g * K + k] = ld(base + (k << BLO));
    }
#pragma unroll
    for (int u = 0; u < RR; u++) {
        const int half = 1 << (RR - 1 - u);
#pragma unroll
        for (int g = 0; g < G; g++) {
            u32 hi = (vt + g * NT) >> BLO;
            u32 tbase = (R << (S0 + u)) + (hi << u);
#pragma unroll
            for (int j = 0; j < (1 << u); j++) {
                ulonglong2 w = ld_tw(tw, tbase + j);
#pragma unroll
                for (int r = 0; r < half; r++) {
                    int k = (j << (RR - u)) | r;
                    ct_bfly(x[g * K + k], x[g * K + k + half], w.x, w.y, p, p2);
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        u32 v = vt + g * NT;
        u32 lo = v & ((1u << BLO) - 1), hi = v >> BLO;
        u32 base = (hi << (BLO + RR)) | lo;
#pragma unroll
        for (int k = 0; k < K; k++) st(base + (k << BLO), x[g * K + k]);
    }
}

// One radix-2^RR round of the inverse transform: layers S0+RR-1 down to S0 of the tile.
// FINAL: layer 0 of the whole transform is in this round and carries the N^-1 scaling.
template <int LOGT, int S0, int RR, int NT, bool FINAL, class Ld, class St>
TB_D void inv_round(u32 vt, Ld ld, St st, const u64 *__restrict__ tw, u32 R, u64 p, u64 p2, const ModDev *md) {
    constexpr int E = 16, G = E >> RR, K = 1 << RR, BLO = LOGT - S0 - RR;
    static_assert(NT * E == (1 << LOGT), "tile/threads mismatch");
    static_assert(BLO >= 0, "bad round");
    u64 x[E];
#pragma unroll
    for (int g = 0; g < G; g++) {
        u32 v = vt + g * NT;
        u32 lo = v & ((1u << BLO) - 1), hi = v >> BLO;
        u32 base = (hi << (BLO + RR)) | lo;
#pragma unroll
        for (int k = 0; k < K; k++) x[g * K + k] = ld(base + (k << BLO));
    }
#pragma unroll
    for (int u = RR - 1; u >= 0; u--) {
        const int half = 1 << (RR - 1 - u);
#pragma unroll
        for (int g = 0; g < G; g++) {
            u32 hi = (vt + g * NT) >> BLO;
            u32 tbase = (R << (S0 + u)) + (hi << u);
            if (FINAL && S0 == 0 && u == 0) {
                u64 ninv = md->ninv, ninvq = md->ninv_shoup, w1 = md->inv1n, w1q = md->inv1n_shoup;
#pragma unroll
                for (int r = 0; r < half; r++) gs_bfly_last(x[g * K + r], x[g * K + r + half], ninv, ninvq, w1, w1q, p, p2);
            } else {
#pragma unroll
                for (int j = 0; j < (1 << u); j++) {
                    ulonglong2 w = ld_tw(tw, tbase + j);
#pragma unroll
                    for (int r = 0; r < half; r++) {
                        int k = (j << (RR - u)) | r;
                        gs_bfly(x[g * K + k], x[g * K + k + half], w.x, w.y, p, p2);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        u32 v = vt + g * NT;
        u32 lo = v & ((1u << BLO) - 1), hi = v >> BLO;
        u32 base = (hi << (BLO + RR)) | lo;
#pragma unroll
        for (int k = 0; k < K; k++) st(base + (k << BLO), x[g * K + k]);
    }
}

// ---- shared-memory accessors ---------------------------------------------------------------------
struct SmLd {
    const u64 *sm;
    TB_D u64 operator()(u32 i) const { return sm[pad_idx(i)]; }
};
struct SmSt {
    u64 *sm;
    TB_D void operator()(u32 i, u64 v) const { sm[pad_idx(i)] = v; }
};

// ---- whole-tile transforms --------------------------------------------------------------------------
// Forward transform of a tile held by one CTA of NT = 2^LOGT/16 threads.
//   ld(i)    : coefficient i of the tile, value in [0,4p)  (prologue fused here)
//   st(i, v) : receives output i, v in [0,4p)              (epilogue fused here; called in linear, coalesced order)
// The last round's results pass through shared memory once more so that global stores are coalesced.
template <int LOGT, class Ld, class St>
TB_D void tile_ntt_fwd(u64 *sm, Ld ld, St st, const u64 *__restrict__ tw, u32 R, u64 p) {
    constexpr int NT = (1 << LOGT) / 16;
    constexpr int NR = (LOGT + 3) / 4;      // rounds
    constexpr int RL = LOGT - 4 * (NR - 1); // layers in the last round
    const u64 p2 = p << 1;
    SmLd sl{sm};
    SmSt ss{sm};
    if constexpr (NR == 1) {
        TB_PHASE(NT, t, fwd_round<LOGT, 0, RL, NT>(t, ld, ss, tw, R, p, p2);)
    } else {
        TB_PHASE(NT, t, fwd_round<LOGT, 0, 4, NT>(t, ld, ss, tw, R, p, p2);)
        if constexpr (NR >= 3) {
            TB_PHASE(NT, t, fwd_round<LOGT, 4, 4, NT>(t, sl, ss, tw, R, p, p2);)
        }
        if constexpr (NR >= 4) {
            TB_PHASE(NT, t, fwd_round<LOGT, 8, 4, NT>(t, sl, ss, tw, R, p, p2);)
        }
        TB_PHASE(NT, t, fwd_round<LOGT, 4 * (NR - 1), RL, NT>(t, sl, ss, tw, R, p, p2);)
    }
    TB_PHASE(NT, t, {
        _Pragma("unroll") for (int k = 0; k < 16; k++) {
            u32 i = t + k * NT;
            st(i, sm[pad_idx(i)]);
        }
    })
}

// Inverse transform of a tile. ld(i) must return values in [0,2p); st(i, v) gets v in [0,2p).
// FINAL = the tile contains layer 0 of the whole transform (N^-1 folded into that layer).
template <int LOGT, bool FINAL, class Ld, class St>
TB_D void tile_ntt_inv(u64 *sm, Ld ld, St st, const u64 *__restrict__ tw, u32 R, u64 p, const ModDev *md) {
    constexpr int NT = (1 << LOGT) / 16;
    constexpr int NR = (LOGT + 3) / 4;
    constexpr int RL = LOGT - 4 * (NR - 1); // layers in the FIRST inverse round (deepest layers)
    const u64 p2 = p << 1;
    SmLd sl{sm};
    SmSt ss{sm};
    TB_PHASE(NT, t, {
        _Pragma("unroll") for (int k = 0; k < 16; k++) {
            u32 i = t + k * NT;
            sm[pad_idx(i)] = ld(i);
        }
    })
    if constexpr (NR == 1) {
        TB_PHASE(NT, t, (inv_round<LOGT, 0, RL, NT, FINAL>(t, sl, st, tw, R, p, p2, md));)
    } else {
        TB_PHASE(NT, t, (inv_round<LOGT, 4 * (NR - 1), RL, NT, false>(t, sl, ss, tw, R, p, p2, md));)
        if constexpr (NR >= 4) {
            TB_PHASE(NT, t, (inv_round<LOGT, 8, 4, NT, false>(t, sl, ss, tw, R, p, p2, md));)
        }
        if constexpr (NR >= 3) {
            TB_PHASE(NT, t, (inv_round<LOGT, 4, 4, NT, false>(t, sl, ss, tw, R, p, p2, md));)
        }
        TB_PHASE(NT, t, (inv_round<LOGT, 0, 4, NT, FINAL>(t, sl, st, tw, R, p, p2, md));)
    }
}

} // namespace troyb
