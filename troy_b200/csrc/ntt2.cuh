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
// Shared-memory layout: the tile is exactly 2^LOGT words with the 128-byte swizzle of the tensor-map unit: word i lives at
// swz(i) = i ^ ((i >> 3) & 0xE), i.e. inside every 128-byte row the 16-byte chunk index is XOR-ed with the low three bits of
// the row number. Every access pattern above is conflict free with it (lanes on consecutive words of one row, on rows
// 64 bytes apart, or on whole 64-byte groups of neighbouring rows), and it is the layout cp.async.bulk.tensor reads and
// writes, so the tile is filled by TMA loads (one per 4 .. 16 KiB segment, issued by one thread, no registers and no
// load-store-unit wavefronts) and, where the epilogue is a plain store, drained by one TMA store per block.
#pragma once
#include "ntt.cuh"
#include <cuda.h>

namespace troyb {

__host__ __device__ constexpr u32 swz(u32 i) { return i ^ ((i >> 3) & 0xEu); }

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
TB_D u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
TB_D void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); // visible to the asynchronous proxy
}
TB_D void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TB_D void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box (0, row) of the tensor map -> shared memory at dst; completion counted in bytes on bar
TB_D void tma_load_rows(void *dst, const CUtensorMap *tm, u32 row, u64 *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}
TB_D void tma_store_rows(const CUtensorMap *tm, u32 row, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(0), "r"(row), "r"(smem_u32(src)) : "memory");
}
TB_D void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}
// The same box as tma_store_rows, ADDED (64-bit integer add, performed at the L2) into global memory instead of stored.
TB_D void tma_reduce_add_rows(const CUtensorMap *tm, u32 row, const void *src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tm), "r"(0), "r"(row), "r"(smem_u32(src)) : "memory");
}
// `bytes` (a multiple of 16) of consecutive 64-bit words at src (shared memory, linear) added into dst (global memory)
TB_D void bulk_reduce_add_u64(u64 *dst, const void *src, u32 bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
TB_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING> TB_D void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
// orders this thread's shared-memory writes before later accesses of the asynchronous proxy (TMA store)
TB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
    constexpr u32 KS = 1024u;
    constexpr int IT = 1024 / (NP * NTHR);
    TwSet<A, 3> w;
    w.load(tw, R, 0, 0);
#pragma unroll 1
    for (int it = 0; it < IT; it++) {
        const u32 a0 = swz((u32)NP * tid + (u32)(NP * NTHR * it));
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
// a0 = address of element 0 of the group, a1 = a0 ^ 8: elements g + 64 k sit in rows 4 k apart, whose parity flips the swizzle
template <class A, bool INVERSE, bool REDUCE, class TwF>
TB_D void block_b16_group(u64 *sm, u32 a0, const TwF &w, const typename A::Mod &m) {
    const u32 a1 = a0 ^ 8u;
    typename A::elem x[1][16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[0][k] = A::from_bits(sm[((k & 1) ? a1 : a0) + 64u * (u32)k]);
    if (REDUCE) {
#pragma unroll
        for (int k = 0; k < 16; k++) A::reduce(x[0][k], m);
    }
    if constexpr (INVERSE) radix_inv_tw<A, 4, 1, false>(x, w, m);
    else radix_fwd_tw<A, 4, 1>(x, w, m);
#pragma unroll
    for (int k = 0; k < 16; k++) sm[((k & 1) ? a1 : a0) + 64u * (u32)k] = A::to_bits(x[0][k]);
}

// rb = base + 64 h (row of element 0), lh = l ^ ((h & 1) << 3): element k = 64 h + l + 8 k sits in row rb / 16 + (k >> 1) at
// in-row word (l + 8 (k & 1)) ^ (((4 (h & 1) + (k >> 1)) << 1) = lh ^ c_k with the constant c_k = 8 (k & 1) ^ ((k >> 1) << 1)
template <class A, bool INVERSE, bool REDUCE, class TwF>
TB_D void block_m8_group(u64 *sm, u32 rb, u32 lh, const TwF &w, const typename A::Mod &m) {
    typename A::elem x[1][8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[0][k] = A::from_bits(sm[rb + 16u * (u32)(k >> 1) + (lh ^ (u32)((8 * (k & 1)) ^ ((k >> 1) << 1)))]);
    if (REDUCE) {
#pragma unroll
        for (int k = 0; k < 8; k++) A::reduce(x[0][k], m);
    }
    if constexpr (INVERSE) radix_inv_tw<A, 3, 1, false>(x, w, m);
    else radix_fwd_tw<A, 3, 1>(x, w, m);
#pragma unroll
    for (int k = 0; k < 8; k++) sm[rb + 16u * (u32)(k >> 1) + (lh ^ (u32)((8 * (k & 1)) ^ ((k >> 1) << 1)))] = A::to_bits(x[0][k]);
}

// rb = first word of the row of group g (base + 16 (g >> 1)), c0 = 8 (g & 1) ^ (((g >> 1) & 7) << 1): the pair (2 q, 2 q + 1) of
// the group sits at in-row word c0 ^ 2 q.   forward: leaves canonical residues; inverse: reads canonical residues
template <class A, bool INVERSE, class TwF>
TB_D void block_l8_group(u64 *sm, u32 rb, u32 c0, const TwF &w, const typename A::Mod &m) {
    typename A::elem x[1][8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u64 a, b;
        lds128(sm, rb + (c0 ^ (u32)(2 * q)), a, b);
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
        for (int q = 0; q < 4; q++) sts128(sm, rb + (c0 ^ (u32)(2 * q)), A::to_bits(x[0][2 * q]), A::to_bits(x[0][2 * q + 1]));
    } else {
        radix_fwd_tw<A, 3, 1>(x, w, m);
#pragma unroll
        for (int q = 0; q < 4; q++) sts128(sm, rb + (c0 ^ (u32)(2 * q)), A::canon(x[0][2 * q], m), A::canon(x[0][2 * q + 1], m));
    }
}

// group of lane in turn `it` of the radix-16 round: g = lane + 32 it; address of its element 0
TB_D u32 b16_addr(u32 base, u32 lane, int it) {
    const u32 g = lane + 32u * (u32)it;
    return base + (g ^ (((g >> 4) & 3u) << 1));
}
// group (h, l) of lane in turn `it` of the stride-8 round: h = lane / 8 + 4 it, l = lane % 8
TB_D void m8_addr(u32 base, u32 lane, int it, u32 &rb, u32 &lh) {
    const u32 h = (lane >> 3) + 4u * (u32)it;
    rb = base + 64u * h;
    lh = (lane & 7u) ^ ((h & 1u) << 3);
}
// group g = lane + 32 it of the last round
TB_D void l8_addr(u32 base, u32 lane, int it, u32 &rb, u32 &c0) {
    const u32 g = lane + 32u * (u32)it;
    rb = base + 16u * (g >> 1);
    c0 = (8u * (g & 1u)) ^ (((g >> 1) & 7u) << 1);
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
    u32 r0, r1;
    if constexpr (Prefetching<A>::value) {
        TwSet<A, 3> wc, wn;
        block_b16_group<A, false, false>(sm, b16_addr(base, lane, 0), pre.b16, m);
        wc.load(tw, Rb, 4, lane >> 3); // first M8 group's twiddles: in flight during the second radix-16 group
        block_b16_group<A, false, false>(sm, b16_addr(base, lane, 1), pre.b16, m);
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)(it + 1));
            else wn.load(tw, Rb, 7, lane);
            m8_addr(base, lane, it, r0, r1);
            block_m8_group<A, false, false>(sm, r0, r1, wc, m);
            wc = wn;
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 7, lane + 32u * (u32)(it + 1));
            l8_addr(base, lane, it, r0, r1);
            block_l8_group<A, false>(sm, r0, r1, wc, m);
            wc = wn;
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, false, false>(sm, b16_addr(base, lane, it), TwJit<A>{tw, Rb, 0, 0}, m);
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)it);
            m8_addr(base, lane, it, r0, r1);
            block_m8_group<A, false, false>(sm, r0, r1, w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 7, lane + 32u * (u32)it);
            l8_addr(base, lane, it, r0, r1);
            block_l8_group<A, false>(sm, r0, r1, w, m);
        }
    }
    __syncwarp();
}

// Inverse block stage: the block holds canonical residues (staged by the warp itself); leaves element bits.
//   first : twiddles of the first L8 group, fetched by the caller while the staging copies were in flight
template <class A, class First> TB_D void inv_block_stage(u64 *sm, u32 base, u32 lane, const void *__restrict__ tw, u32 Rb,
                                                          const typename A::Mod &m, const First &first) {
    constexpr bool RE = A::kReduceEachInvRound, RM = A::kReduceInvMidway;
    u32 r0, r1;
    if constexpr (Prefetching<A>::value) {
        TwSet<A, 3> wc = first, wn;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 7, lane + 32u * (u32)(it + 1));
            else wn.load(tw, Rb, 4, lane >> 3);
            l8_addr(base, lane, it, r0, r1);
            block_l8_group<A, true>(sm, r0, r1, wc, m);
            wc = wn;
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (it < 3) wn.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)(it + 1));
            m8_addr(base, lane, it, r0, r1);
            block_m8_group<A, true, RE>(sm, r0, r1, wc, m);
            wc = wn;
        }
        TwSet<A, 4> wb; // uniform over the warp; requested before the warp synchronises
        wb.load(tw, Rb, 0, 0);
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, true, RE || RM>(sm, b16_addr(base, lane, it), wb, m);
    } else {
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 7, lane + 32u * (u32)it);
            l8_addr(base, lane, it, r0, r1);
            block_l8_group<A, true>(sm, r0, r1, w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            TwSet<A, 3> w;
            w.load(tw, Rb, 4, (lane >> 3) + 4u * (u32)it);
            m8_addr(base, lane, it, r0, r1);
            block_m8_group<A, true, RE>(sm, r0, r1, w, m);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 2; it++) block_b16_group<A, true, RE || RM>(sm, b16_addr(base, lane, it), TwJit<A>{tw, Rb, 0, 0}, m);
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

// ---- forward first stage fed by the tensor-map unit -------------------------------------------------------------------------
// The inputs are copied straight INTO the tile: input k of group v'' (v'' counts the CTA's own groups) lives at word
// k * VC + v'' (VC = groups per CTA), one TMA box of VC / 16 rows per k. That is exactly where one of the outputs of the same
// thread position lands - output k of group v'' of CTA r' goes to word (k mod KT) * 2^VLOG + r' * VC + v'' of tile k / KT, the
// home of input (k mod KT) * C + r' of group v'' there - so the only hazard is between the threads that handle the same v''
// in the CTAs of a cluster: a thread stores the outputs of a batch of groups only after every CTA has read that batch's
// inputs (a split cluster barrier: arrive after the reads, wait before the stores, the butterflies in between).
//   row0 : row (16 words) of coefficient 0 of the limb in the tensor map; called by ONE thread, after mbar_init
template <int LOGT, int LOGC> TB_D void fwd_stage_inputs(u64 *sm, u32 rank, const CUtensorMap *tm, u32 row0, u64 *bar) {
    typedef LimbPlan<LOGT, LOGC> P;
    constexpr int F1 = P::F1, K = 1 << F1, VLOG = P::VLOG, LOGVC = LOGT - F1;
    mbar_expect_tx(bar, 8u << LOGT);
#pragma unroll 1
    for (u32 k = 0; k < (u32)K; k++) tma_load_rows(sm + (k << LOGVC), tm, row0 + (k << (VLOG - 4)) + (rank << (LOGVC - 4)), bar);
}

//   pro(x)                 : prologue on a staged coefficient
//   st_peer(tile, w, bits) : element bits to padded word w of tile `tile`
template <class A, int LOGT, int LOGC, bool CLUSTER, class Pro, class StPeer>
TB_D void fwd_first_stage_staged(const u64 *sm, u32 tid, u32 rank, Pro pro, StPeer st_peer, const ModDev *md) {
    typedef LimbPlan<LOGT, LOGC> P;
    constexpr int F1 = P::F1, K = 1 << F1, VLOG = P::VLOG, NT = P::NT, LOGVC = LOGT - F1, VC = 1 << LOGVC, G = VC / NT;
    constexpr int B = 16 / K;                       // groups per batch: 16 inputs per thread and batch
    constexpr u32 KS = VC;                          // distance between the homes of inputs k and k + 1
    static_assert(G == 2 * B, "two batches per tile");
    const typename A::Mod m = A::make(md);
    const void *tw = A::fwd_table(md);
    const TwJit<A> twf{tw, 1u, 0u, 0u}; // uniform: layer u, block j of the limb = entry 2^u + j
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        typename A::elem x[B][K];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 pv = swz(tid + (u32)((g0 + b) * NT));
            u64 raw[K];
#pragma unroll
            for (int k = 0; k < K; k++) raw[k] = sm[pv + KS * (u32)k];
            if (pro.mode != FWD_PLAIN) { // the prologue's mode branches stay out of the load sequence
#pragma unroll
                for (int k = 0; k < K; k++) raw[k] = pro(raw[k]);
            }
#pragma unroll
            for (int k = 0; k < K; k++) x[b][k] = A::from_u64(raw[k], m);
        }
        // every input of the batch has been issued ahead of the arrive (in order, to the same pipe), and the stores it
        // guards are a barrier round trip and a trip across the cluster away: the homes may be overwritten
        if (CLUSTER) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        radix_fwd_tw<A, F1, B>(x, twf, m);
        if (CLUSTER) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 pv = swz((rank << LOGVC) + tid + (u32)((g0 + b) * NT));
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 local = ((u32)k & ((1u << (LOGT - VLOG)) - 1)) << VLOG;
                st_peer((u32)k >> (LOGT - VLOG), local + pv, A::to_bits(x[b][k]));
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
    // FP64 policies that re-centre in every round do it here as well; those that re-centre once, at the block's radix-16
    // round, stay inside the 9 layers their multiply tolerates unless 4 + 3 + 3 layers follow it (N = 65536)
    constexpr bool RE = A::kReduceEachInvRound || (A::kReduceInvMidway && P::MID && F1 >= 3);
    const typename A::Mod m = A::make(md);
    const void *tw = A::inv_table(md);
    const TwJit<A> twf{tw, 1u, 0u, 0u};
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += B) {
        u64 raw[B][K];
#pragma unroll
        for (int b = 0; b < B; b++) {
            const u32 v = rank * (u32)((1 << LOGT) / K) + tid + (u32)((g0 + b) * NT);
            const u32 pv = swz(v & ((1u << VLOG) - 1));
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 local = ((u32)k & ((1u << (LOGT - VLOG)) - 1)) << VLOG;
                raw[b][k] = ld_peer((u32)k >> (LOGT - VLOG), local + pv);
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
