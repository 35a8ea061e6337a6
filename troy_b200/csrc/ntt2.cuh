// Tile transforms, second generation (GPU only): one warp owns a block of 1024 coefficients.
//
// Same function as ntt.cuh (kNttNegacyclicHarvey / kInverseNttNegacyclicHarvey, src/kernelutils.cuh:578-632; table order of
// src/utils/ntt.cpp:34-54) and the same butterfly policies; what changes is who synchronises with whom and how the data gets
// on chip. A limb of N = 2^logN coefficients is C = N / 8192 tiles (one CTA of 256 threads each, C > 1: a thread-block
// cluster); a tile is eight blocks of 1024 coefficients and warp w owns block w.
//
//   forward : staging      - the 2^LOGT inputs of the CTA's radix groups are copied global -> shared memory with cp.async
//                            (16 bytes per lane and copy, all in flight at once: ONE memory latency per tile, no registers)
//             first stage  - the logN - 10 layers that couple the blocks, one radix-16 (8, 4) network per group in registers;
//                            output k goes to the block that owns it (its own tile or a peer's through distributed shared
//                            memory). N >= 32768: the cluster layers only; the three layers that couple the blocks of a
//                            tile are a second, CTA-wide round in shared memory.
//             ONE barrier (cluster-wide for C > 1)
//             block stage  - the last 10 layers inside the warp's own block as a radix-16 round (twiddles uniform over the
//                            warp, fetched before the barrier) and two radix-8 rounds whose twiddles are fetched one group
//                            ahead: only __syncwarp() between them, so the warps of a CTA drift apart and their
//                            shared-memory, twiddle and butterfly phases overlap instead of queueing up behind CTA barriers.
//                            The last round reads 8 consecutive coefficients per lane (128-bit accesses) and leaves
//                            canonical residues; the warp then copies its block out with 128-bit coalesced accesses
//                            (epilogue fused there).
//   inverse : the mirror image - the warp stages its block (cp.async), runs radix-8, radix-8, radix-16, ONE barrier, and
//             the last stage gathers across blocks / tiles, applies N^-1 in the last layer and stores.
//
// Shared-memory layout: two pad words after every 16 (pad2). Every access pattern above is conflict free with it (lanes
// read consecutive words inside 16-word rows; the stride-8 round splits a half warp over the two halves of the banks; the
// last round's 64-byte groups sit 64 or 80 bytes apart, which spreads a quarter warp over all 32 banks), and all offsets
// inside a group are compile-time constants.
#pragma once
#include "ntt.cuh"

namespace troyb {

__host__ __device__ constexpr u32 pad2(u32 i) { return i + ((i >> 4) << 1); }
__host__ __device__ constexpr u32 padded2_words(u32 n) { return n + (n >> 3); }

#if defined(TROYB_DEV_TIMELINE)
TB_D unsigned long long tb_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// timeline slot k of CTA bid (development build, see NttArgs::dbg)
#define TB_MARK(dbg, bid, k) do { if ((dbg) != nullptr && threadIdx.x == 0) (dbg)[(size_t)(bid) * 16 + (k)] = tb_now(); } while (0)
#else
#define TB_MARK(dbg, bid, k) do { } while (0)
#endif

TB_D void lds128(const u64 *sm, u32 w, u64 &a, u64 &b) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sm + w);
    a = v.x;
    b = v.y;
}
TB_D void sts128(u64 *sm, u32 w, u64 a, u64 b) {
    ulonglong2 v;
    v.x = a;
    v.y = b;
    *reinterpret_cast<ulonglong2 *>(sm + w) = v;
}
TB_D void cp_async16(u64 *smem_dst, const u64 *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
TB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
TB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// NP groups that share their twiddles run the same network; tw(u, j) is the twiddle of block j of layer u of the network.
template <class A, int RR, int NP, class TwF>
TB_D void radix_fwd_tw(typename A::elem (&x)[NP][1 << RR], const TwF &tw, const typename A::Mod &m) {
#pragma unroll
    for (int u = 0; u < RR; u++) {
        const int half = 1 << (RR - 1 - u);
#pragma unroll
        for (int j = 0; j < (1 << u); j++) {
            const typename A::tw_t w = tw(u, j);
#pragma unroll
            for (int r = 0; r < half; r++) {
                const int k = (j << (RR - u)) | r;
#pragma unroll
                for (int q = 0; q < NP; q++) A::ct(x[q][k], x[q][k + half], w, m, (u & 1) != 0);
            }
        }
    }
}

// LAST: layer 0 of the whole transform is layer 0 of this network and carries the N^-1 scaling (tw(0, 0) is not used).
template <class A, int RR, int NP, bool LAST, class TwF>
TB_D void radix_inv_tw(typename A::elem (&x)[NP][1 << RR], const TwF &tw, const typename A::Mod &m) {
#pragma unroll
    for (int u = RR - 1; u >= 0; u--) {
        const int half = 1 << (RR - 1 - u);
        if (LAST && u == 0) {
#pragma unroll
            for (int r = 0; r < half; r++) {
#pragma unroll
                for (int q = 0; q < NP; q++) A::gs_last(x[q][r], x[q][r + half], m);
            }
        } else {
#pragma unroll
            for (int j = 0; j < (1 << u); j++) {
                const typename A::tw_t w = tw(u, j);
#pragma unroll
                for (int r = 0; r < half; r++) {
                    const int k = (j << (RR - u)) | r;
#pragma unroll
                    for (int q = 0; q < NP; q++) A::gs(x[q][k], x[q][k + half], w, m);
                }
            }
        }
    }
}

// Twiddles of one radix-2^RR group in registers: layer s0 + u of the enclosing transform (root base R) uses the 2^u
// consecutive table entries (R << (s0 + u)) + (hi << u) + j.
template <class A, int RR> struct TwSet {
    typename A::tw_t w[(1 << RR) - 1];
    TB_D void load(const void *__restrict__ tw, u32 R, u32 s0, u32 hi) {
#pragma unroll
        for (int u = 0; u < RR; u++) {
            const u32 t0 = (R << (s0 + (u32)u)) + (hi << u);
            if constexpr (sizeof(typename A::tw_t) == 8) {
                // one double per entry: the entries of a layer as 16-byte loads (t0 is even for u >= 1)
                if (u == 0) {
                    w[0] = A::ld_tw(tw, t0);
                } else {
#pragma unroll
                    for (int j = 0; j < (1 << u); j += 2) {
                        const double2 v = __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(tw) + t0 + j));
                        w[(1 << u) - 1 + j] = typename A::tw_t{v.x};
                        w[(1 << u) + j] = typename A::tw_t{v.y};
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < (1 << u); j++) w[(1 << u) - 1 + j] = A::ld_tw(tw, t0 + (u32)j);
            }
        }
    }
    TB_D typename A::tw_t operator()(int u, int j) const { return w[(1 << u) - 1 + j]; }
};
// The same twiddles fetched where they are used (policies whose twiddles are 16 bytes: a set in registers would not fit).
template <class A> struct TwJit {
    const void *tw;
    u32 R, s0, hi;
    TB_D typename A::tw_t operator()(int u, int j) const { return A::ld_tw(tw, (R << (s0 + (u32)u)) + (hi << u) + (u32)j); }
};
// Policies with 8-byte twiddles keep whole sets in registers and fetch them ahead of their use.
template <class A> struct Prefetching { static constexpr bool value = sizeof(typename A::tw_t) == 8; };

// ---- CTA-wide radix-8 round (N >= 32768: the three layers that couple the blocks of a tile) -----------------------------------
// Groups {g + (k << 10)} of the tile, root base R, one set of 7 twiddles for everybody; two neighbouring groups per thread
// and access where an element is one register pair.
template <class A, int NTHR, bool INVERSE, bool REDUCE>
TB_D void round_tile(u64 *sm, u32 tid, const void *__restrict__ tw, u32 R, const typename A::Mod &m) {
    constexpr int NP = Prefetching<A>::value ? 2 : 1;
    constexpr u32 KS = 1024u + 128u;
    constexpr int IT = 1024 / (NP * NTHR);
    TwSet<A, 3> w;
    w.load(tw, R, 0, 0);
#pragma unroll 1
    for (int it = 0; it < IT; it++) {
        const u32 a0 = pad2((u32)NP * tid + (u32)(NP * NTHR * it));
        typename A::elem x[NP][8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if constexpr (NP == 2) {
                u64 a, b;
                lds128(sm, a0 + KS * (u32)k, a, b);
                x[0][k] = A::from_bits(a);
                x[1][k] = A::from_bits(b);
            } else {
                x[0][k] = A::from_bits(sm[a0 + KS * (u32)k]);
            }
        }
        if (REDUCE) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int q = 0; q < NP; q++) A::reduce(x[q][k], m);
            }
        }
        if constexpr (INVERSE) radix_inv_tw<A, 3, NP, false>(x, w, m);
        else radix_fwd_tw<A, 3, NP>(x, w, m);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if constexpr (NP == 2) sts128(sm, a0 + KS * (u32)k, A::to_bits(x[0][k]), A::to_bits(x[1][k]));
            else sm[a0 + KS * (u32)k] = A::to_bits(x[0][k]);
        }
    }
}

// ---- the three rounds of a block (warp-private; base = unpadded tile index of the block, Rb = its root base) ------------------
// B16 : layers 0..3, groups {g + 64 k} (k < 16), twiddles uniform over the warp
// M8  : layers 4..6, groups {64 h + l + 8 k}, the 8 lanes with the same h share their twiddles
// L8  : layers 7..9, groups of 8 consecutive coefficients
template <class A, bool INVERSE, bool REDUCE, class TwF>
TB_D void block_b16_group(u64 *sm, u32 a0, const TwF &w, const typename A::Mod &m) {
    typename A::elem x[1][16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[0][k] = A::from_bits(sm[a0 + 72u * (u32)k]);
    if (REDUCE) {
#pragma unroll
        for (int k = 0; k < 16; k++) A::reduce(x[0][k], m);
    }
    if constexpr (INVERSE) radix_inv_tw<A, 4, 1, false>(x, w, m);
    else radix_fwd_tw<A, 4, 1>(x, w, m);
#pragma unroll
    for (int k = 0; k < 16; k++) sm[a0 + 72u * (u32)k] = A::to_bits(x[0][k]);
}

template <class A, bool INVERSE, bool REDUCE, class TwF>
TB_D void block_m8_group(u64 *sm, u32 a0, const TwF &w, const typename A::Mod &m) {
    typename A::elem x[1][8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[0][k] = A::from_bits(sm[a0 + 8u * (u32)k + 2u * (u32)(k >> 1)]);
    if (REDUCE) {
#pragma unroll
        for (int k = 0; k < 8; k++) A::reduce(x[0][k], m);
    }
    if constexpr (INVERSE) radix_inv_tw<A, 3, 1, false>(x, w, m);
    else radix_fwd_tw<A, 3, 1>(x, w, m);
#pragma unroll
    for (int k = 0; k < 8; k++) sm[a0 + 8u * (u32)k + 2u * (u32)(k >> 1)] = A::to_bits(x[0][k]);
}

// forward: leaves canonical residues; inverse: reads canonical residues
template <class A, bool INVERSE, class TwF>
TB_D void block_l8_group(u64 *sm, u32 a0, const TwF &w, const typename A::Mod &m) {
    typename A::elem x[1][8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u64 a, b;
        lds128(sm, a0 + 2u * (u32)q, a, b);
        if constexpr (INVERSE) {
            x[0][2 * q] = A::from_u64(a, m);
            x[0][2 * q + 1] = A::from_u64(b, m);
        } else {
            x[0][2 * q] = A::from_bits(a);
            x[0][2 * q + 1] = A::from_bits(b);
        }
    }
    if constexpr (INVERSE) {
        radix_inv_tw<A, 3, 1, false>(x, w, m);
#pragma unroll
        for (int q = 0; q < 4; q++) sts128(sm, a0 + 2u * (u32)q, A::to_bits(x[0][2 * q]), A::to_bits(x[0][2 * q + 1]));
    } else {
        radix_fwd_tw<A, 3, 1>(x, w, m);
#pragma unroll
        for (int q = 0; q < 4; q++) sts128(sm, a0 + 2u * (u32)q, A::canon(x[0][2 * q], m), A::canon(x[0][2 * q + 1], m));
    }
}

TB_D u32 m8_addr(u32 base, u32 lane, int it) { // group (h, l) of lane in turn `it`: h = lane / 8 + 4 it, l = lane % 8
    const u32 h = (lane >> 3) + 4u * (u32)it;
    return base + (base >> 3) + 72u * h + (lane & 7u);
}
TB_D u32 l8_addr(u32 base, u32 lane, int it) { // group g = lane + 32 it
    const u32 g = lane + 32u * (u32)it;
    return base + (base >> 3) + 8u * g + 2u * (g >> 1);
}

// Twiddles of the block's radix-16 round (fetched by the caller BEFORE the barrier that precedes the block stage).
template <class A> struct BlockTw {
    TwSet<A, Prefetching<A>::value ? 4 : 1> b16;
    TB_D void load(const void *__restrict__ tw, u32 Rb) {
        if constexpr (Prefetching<A>::value) b16.load(tw, Rb, 0, 0);
    }
};

template <class A> TB_D void fwd_block_stage(u64 *sm, u32 base, u32 lane, const void *__restrict__ tw, u32 Rb,
                                             const typename A::Mod &m, const BlockTw<A> &pre) {
    const u32 pb = base + (base >> 3); // pad2(base): base is a multiple of 1024
    if constexpr (Prefetching<A>::value) {
        TwSet<A, 3> wc, wn;
        block_b16_group<A, false, false>(sm, pb + pad2(lane), pre.b16, m);
        wc.load(tw, Rb, 4, lane >> 3); // first M8 group's twiddles: in flight during the second radix-16 group
        block_b16_group<A, false, false>(sm, pb + pad2(lane) + 36u, pre.b16, m); // group lane + 32: pad2(32) = 36
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)(it + 1));
            else wn.load(tw, Rb, 7, lane);
            block_m8_group<A, false, false>(sm, m8_addr(base, lane, it), wc, m);
            wc = wn;
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 7, lane + 32u * (u32)(it + 1));
            block_l8_group<A, false>(sm, l8_addr(base, lane, it), wc, m);
            wc = wn;
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, false, false>(sm, pb + pad2(lane) + 36u * (u32)it, TwJit<A>{tw, Rb, 0, 0}, m);
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)it);
            block_m8_group<A, false, false>(sm, m8_addr(base, lane, it), w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 7, lane + 32u * (u32)it);
            block_l8_group<A, false>(sm, l8_addr(base, lane, it), w, m);
        }
    }
    __syncwarp();
}

// Inverse block stage: the block holds canonical residues (staged by the warp itself); leaves element bits.
//   first : twiddles of the first L8 group, fetched by the caller while the staging copies were in flight
template <class A, class First> TB_D void inv_block_stage(u64 *sm, u32 base, u32 lane, const void *__restrict__ tw, u32 Rb,
                                                          const typename A::Mod &m, const First &first) {
    constexpr bool RE = A::kReduceEachInvRound, RM = A::kReduceInvMidway;
    const u32 pb = base + (base >> 3);
    if constexpr (Prefetching<A>::value) {
        TwSet<A, 3> wc = first, wn;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 7, lane + 32u * (u32)(it + 1));
            else wn.load(tw, Rb, 4, lane >> 3);
            block_l8_group<A, true>(sm, l8_addr(base, lane, it), wc, m);
            wc = wn;
        }
        __syncwarp();
        TwSet<A, 4> wb;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)(it + 1));
            else wb.load(tw, Rb, 0, 0);
            block_m8_group<A, true, RE>(sm, m8_addr(base, lane, it), wc, m);
            wc = wn;
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, true, RE || RM>(sm, pb + pad2(lane) + 36u * (u32)it, wb, m);
    } else {
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 7, lane + 32u * (u32)it);
            block_l8_group<A, true>(sm, l8_addr(base, lane, it), w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)it);
            block_m8_group<A, true, RE>(sm, m8_addr(base, lane, it), w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, true, RE || RM>(sm, pb + pad2(lane) + 36u * (u32)it, TwJit<A>{tw, Rb, 0, 0}, m);
    }
}

// Geometry of a limb of 2^(LOGT + LOGC) coefficients in tiles of 2^LOGT.
//   F1   : layers of the first forward / last inverse stage (registers)
//   MID  : a CTA-wide radix-8 round couples the blocks of a tile (N >= 32768)
template <int LOGT, int LOGC> struct LimbPlan {
    static constexpr bool MID = LOGC + LOGT - 10 > 4;
    static constexpr int F1 = MID ? LOGC : LOGC + LOGT - 10;
    static constexpr int VLOG = LOGT + LOGC - F1;          // log2 of the number of radix groups of the limb
    static constexpr int NT = LOGT == 13 ? 256 : 128;      // one warp per block
    static_assert(!MID || LOGT == 13, "the CTA-wide round is written for 8192-coefficient tiles");
    static_assert(F1 >= 2 && F1 <= 4, "first stage radix out of range");
};

// ---- forward first stage fed by asynchronous copies ----------------------------------------------------------------------------
// The inputs are copied straight INTO the tile: input k of group v'' (v'' counts the CTA's own groups) lives at
// pad2(k * VC + v''), VC = groups per CTA. That is exactly where one of the outputs of the same thread position lands - output
// k of group v'' of CTA r' goes to word pad2((k mod KT) * 2^VLOG + r' * VC + v'') of tile k / KT, the home of input
// (k mod KT) * C + r' of group v'' there - so the only hazard is between the threads that handle the same v'' in the CTAs of
// a cluster: a thread stores the outputs of a batch of groups only after every CTA has read that batch's inputs (a split
// cluster barrier: arrive after the reads, wait before the stores, the butterflies in between).
template <int LOGT, int LOGC> TB_D void fwd_stage_inputs(u64 *sm, u32 tid, u32 rank, const u64 *__restrict__ src) {
    typedef LimbPlan<LOGT, LOGC> P;
    constexpr int F1 = P::F1, VLOG = P::VLOG, NT = P::NT, LOGVC = LOGT - F1;
    static_assert((1 << LOGT) / 2 == NT * 16, "16 copies per thread");
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const u32 pr = tid + (u32)(NT * j);                       // pair index: k * (VC / 2) + pp
        const u32 k = pr >> (LOGVC - 1), v2 = (pr & ((1u << (LOGVC - 1)) - 1)) << 1;
        cp_async16(sm + pad2((k << LOGVC) + v2), src + ((u64)k << VLOG) + (rank << LOGVC) + v2);
    }
    cp_async_commit();
}

//   pro(x)                 : prologue on a staged coefficient
//   st_peer(tile, w, bits) : element bits to padded word w of tile `tile`
template <class A, int LOGT, int LOGC, bool CLUSTER, class Pro, class StPeer>
TB_D void fwd_first_stage_staged(const u64 *sm, u32 tid, u32 rank, Pro pro, StPeer st_peer, const ModDev *md) {
    typedef LimbPlan<LOGT, LOGC> P;
    constexpr int F1 = P::F1, K = 1 << F1, VLOG = P::VLOG, NT = P::NT, LOGVC = LOGT - F1, VC = 1 << LOGVC, G = VC / NT;
    constexpr int B = 16 / K;                       // groups per batch: 16 inputs per thread and batch
    constexpr u32 KS = VC + (VC >> 3);              // padded distance between the homes of inputs k and k + 1
    static_assert(G == 2 * B, "two batches per tile");
    const typename A::Mod m = A::make(md);
    const void *tw = A::fwd_table(md);
    const TwJit<A> twf{tw, 1u, 0u, 0u}; // uniform: layer u, block j of the limb = entry 2^u + j
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        typename A::elem x[B][K];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 pv = pad2(tid + (u32)((g0 + b) * NT));
#pragma unroll
            for (int k = 0; k < K; k++) x[b][k] = A::from_u64(pro(sm[pv + KS * (u32)k]), m);
        }
        // every input of the batch has been issued ahead of the arrive (in order, to the same pipe), and the stores it
        // guards are a barrier round trip and a trip across the cluster away: the homes may be overwritten
        if (CLUSTER) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        radix_fwd_tw<A, F1, B>(x, twf, m);
        if (CLUSTER) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 pv = pad2((rank << LOGVC) + tid + (u32)((g0 + b) * NT));
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 local = ((u32)k & ((1u << (LOGT - VLOG)) - 1)) << VLOG;
                st_peer((u32)k >> (LOGT - VLOG), pad2(local) + pv, A::to_bits(x[b][k]));
            }
        }
    }
}

// Inverse last stage: the mirror of the forward first stage. Thread `tid` of CTA `rank` takes the groups
// v = rank * VC + tid + g * NT: inputs k at position (k << VLOG) + v of the limb's block-ordered side, outputs to
// coefficients (k << VLOG) + v of the limb (natural order; consecutive lanes store consecutive coefficients). Layer 0 carries
// N^-1; st(i, v) receives the canonical residue of coefficient i.
//   ld_peer(tile, w) : element bits at padded word w of tile `tile`
template <class A, int LOGT, int LOGC, class LdPeer, class St>
TB_D void inv_last_stage(u32 tid, u32 rank, LdPeer ld_peer, St st, const ModDev *md) {
    typedef LimbPlan<LOGT, LOGC> P;
    constexpr int F1 = P::F1, K = 1 << F1, VLOG = P::VLOG, NT = P::NT, G = ((1 << LOGT) / K) / NT;
    constexpr int B = K >= 16 ? 1 : (16 / K < G ? 16 / K : G);
    static_assert(G >= 1 && G % B == 0, "bad last stage");
    // FP64 policies re-centre here as well (the network would otherwise start above the range its multiply accepts)
    constexpr bool RE = A::kReduceEachInvRound || A::kReduceInvMidway;
    const typename A::Mod m = A::make(md);
    const void *tw = A::inv_table(md);
    const TwJit<A> twf{tw, 1u, 0u, 0u};
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        u64 raw[B][K];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (u32)((1 << LOGT) / K) + tid + (u32)((g0 + b) * NT);
            const u32 pv = pad2(v & ((1u << VLOG) - 1));
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 local = ((u32)k & ((1u << (LOGT - VLOG)) - 1)) << VLOG;
                raw[b][k] = ld_peer((u32)k >> (LOGT - VLOG), pad2(local) + pv);
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (u32)((1 << LOGT) / K) + tid + (u32)((g0 + b) * NT);
            typename A::elem x[1][K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                x[0][k] = A::from_bits(raw[b][k]);
                if (RE) A::reduce(x[0][k], m);
            }
            radix_inv_tw<A, F1, 1, true>(x, twf, m);
#pragma unroll
            for (int k = 0; k < K; k++) st(((u32)k << VLOG) + v, A::canon_reduced(x[0][k], m));
        }
    }
}

} // namespace troyb
