// NTT kernels on warp-owned blocks (ntt2.cuh): the same launch interface (NttArgs, modes, fused prologues / epilogues) as
// kernels_ntt.cuh, the transforms re-cut so that a CTA meets ONE barrier per transform and its inputs arrive by asynchronous
// copies. GPU only: the host emulation (tests/emul) keeps the first-generation kernels.
#pragma once
#include "kernels_ntt.cuh"
#include "ntt2.cuh"

namespace troyb {

// 128-bit face of FwdStore: two consecutive outputs (i even).
struct FwdStore2 {
    FwdStore s;
    struct Pre { ulonglong2 acc, ct; };
    TB_D Pre pre(u32 i) const {
        Pre r;
        r.acc = *reinterpret_cast<const ulonglong2 *>(s.u.acc + i);
        r.ct.x = r.ct.y = 0;
        if (s.u.ct_in) r.ct = *reinterpret_cast<const ulonglong2 *>(s.u.ct_in + i);
        return r;
    }
    // mod-down tail (FWD_MODDOWN)
    TB_D void operator()(u32 i, u64 a, u64 b, const Pre &q) const {
        ulonglong2 o;
        o.x = s.finish(q.acc.x, a);
        o.y = s.finish(q.acc.y, b);
        if (s.u.ct) {
            o.x = addmod(q.ct.x, o.x, s.p);
            o.y = addmod(q.ct.y, o.y, s.p);
            *reinterpret_cast<ulonglong2 *>(s.u.ct + i) = o;
        } else {
            *reinterpret_cast<ulonglong2 *>(s.u.dst + i) = o;
        }
    }
};

// Mod-down tail: the warp reads the canonical residues of its block back (128 bits per lane and access, coalesced on the
// global side) and applies the epilogue; the epilogue's own global operands are requested 8 accesses ahead (one memory
// latency per batch, not per access). Plain stores do not come here: one TMA store per block drains them.
TB_D void block_copy_out(const u64 *sm, u32 base, u32 lane, const FwdStore2 &st) {
    constexpr int UB = 8;
    const u32 g = 2u * lane, a0 = base + (g ^ (((g >> 4) & 3u) << 1)), a1 = a0 ^ 8u; // pair 2 lane + 64 k: see block_b16_group
#pragma unroll 1
    for (int k0 = 0; k0 < 16; k0 += UB) {
        FwdStore2::Pre pre[UB];
#pragma unroll
        for (int k = 0; k < UB; k++) pre[k] = st.pre(base + g + 64u * (u32)(k0 + k));
#pragma unroll
        for (int k = 0; k < UB; k++) {
            u64 a, b;
            lds128(sm, ((k & 1) ? a1 : a0) + 64u * (u32)(k0 + k), a, b);
            st(base + g + 64u * (u32)(k0 + k), a, b, pre[k]);
        }
    }
}

// Fused mod-up + key inner product (NttArgs::fused): the warp multiplies its finished block (canonical residues x of digit j
// under output modulus i, still in the tile) by the key's two components and ADDS the products to the accumulators of (batch
// element, output modulus i) with bulk reduce-adds (cp.reduce.async.bulk ... .add.u64: 64-bit integer adds performed at the
// L2; the sums of the digits of one key switch stay below 2^64 and are reduced by the consumers). The block is handled in two
// halves of 512 coefficients: row by row (64 coefficients) the lanes read x from the tile's swizzle, form y0 = x k0 and
// y1 = x k1, put y0 back into the row in LINEAR order and keep y1 in registers; one 4 KiB reduce-add sends the half's y0, and
// once the unit has read it the same half takes y1 and a second reduce-add. A pass-through unit (CKKS, digit i under modulus
// i) takes x from the NTT-form target instead of the tile. Key operands are requested four rows ahead.
// F64: the limb's prime is below 2^46 and the products are formed on the FP64 pipe.
struct FusedKey {
    const u64 *key0, *key1; // rows of the two key components, at the tile's first coefficient
    u64 *prod0, *prod1;     // accumulator rows of the two components, at the tile's first coefficient
    const u64 *pass;        // pass-through source row at the tile's first coefficient, or null
    const ModDev *md;
};
template <bool F64> TB_D void block_key_product(u64 *sm, u32 base, u32 lane, const FusedKey &f) {
    const u32 g = 2u * lane, a0 = base + (g ^ (((g >> 4) & 3u) << 1)), a1 = a0 ^ 8u; // as block_copy_out
    const ModDev *md = f.md;
    const u64 p = md->p, mu = md->bar_mu;
    const u32 sh = md->bar_shift;
    typedef ArithF64T<true> AF;
    const AF::Mod m = AF::make(md);
    const double lift = AF::kTwo52 + m.p; // r + p as the low bits of r + p + 2^52: a residue in (0, 2p), one instruction
    auto mul = [&](u64 x, u64 k) -> u64 {
        if constexpr (F64) {
            const double r = AF::mulmod(AF::from_u64(x, m), AF::mk_tw(AF::from_u64(k, m), 0.0), m); // |r| < p
            return (u64)__double_as_longlong(__dadd_rn(r, lift)) & 0x000FFFFFFFFFFFFFull;
        } else {
            return reduce_sum(x * k, mulhi64(x, k), p, mu, sh); // canonical
        }
    };
#pragma unroll 1
    for (u32 h = 0; h < 2; h++) {
        ulonglong2 y1[8];
#pragma unroll
        for (int q4 = 0; q4 < 8; q4 += 4) {
            ulonglong2 ka[4], kb[4], xs[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 i = base + g + 64u * (8u * h + (u32)(q4 + k));
                ka[k] = __ldg(reinterpret_cast<const ulonglong2 *>(f.key0 + i));
                kb[k] = __ldg(reinterpret_cast<const ulonglong2 *>(f.key1 + i));
                if (f.pass) xs[k] = *reinterpret_cast<const ulonglong2 *>(f.pass + i);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 kk = 8u * h + (u32)(q4 + k), addr = ((kk & 1) ? a1 : a0) + 64u * kk;
                u64 xa, xb;
                if (f.pass) { xa = xs[k].x; xb = xs[k].y; }
                else lds128(sm, addr, xa, xb);
                ulonglong2 y0;
                y0.x = mul(xa, ka[k].x); y0.y = mul(xb, ka[k].y);
                y1[q4 + k].x = mul(xa, kb[k].x); y1[q4 + k].y = mul(xb, kb[k].y);
                __syncwarp(); // every lane has read its part of row kk
                *reinterpret_cast<ulonglong2 *>(sm + base + 64u * kk + g) = y0;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_reduce_add_u64(f.prod0 + base + 512u * h, sm + base + 512u * h, 4096u);
            bulk_commit();
            bulk_wait_read<0>(); // the half is free again once the unit has read it
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) *reinterpret_cast<ulonglong2 *>(sm + base + 64u * (8u * h + (u32)k) + g) = y1[k];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_reduce_add_u64(f.prod1 + base + 512u * h, sm + base + 512u * h, 4096u);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_read<0>(); // the tile must outlive the unit's reads
}

template <int LOGT> struct NttFwdTile2 {
    static constexpr int kThreads = LimbPlan<LOGT, 0>::NT;
    static constexpr int kMinCtas = LOGT == 13 ? 3 : 6;
    static constexpr int kStages = 1;
    static constexpr bool kTrailingBarrier = false; // peers only write into this CTA's tile, before the kernel's own barrier
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D bool barrier_after(int) const { return false; }

    template <int LOGC> TB_D void run(NttUnit &u, u32 bid, u32 grid, u32 rank, u64 *sm, const Peers &peers) const {
        typedef LimbPlan<LOGT, LOGC> P;
        const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
        TB_MARK(a.dbg, bid, 0);
        u64 *bar_in = sm + (1u << LOGT); // barrier words live behind the tile (tile_smem2)
        if (t == 0) {
            mbar_init(bar_in, 1);
            fwd_stage_inputs<LOGT, LOGC>(sm, rank, &a.tm_src, (u32)((u.src - a.tm_src_base) >> 4), bar_in);
        }
        const ModDev *md = u.md;
        const int pol = md->policy;
        prefetch_tiles<LOGT>(a, u, bid, rank, grid);
        const FwdPrologue pro{md->p, md->ratio1, a.mode == FWD_MODDOWN ? md->p - u.dl.half_mod : 0,
                              (a.mode == FWD_MODUP && !u.reduce_src) ? (int)FWD_PLAIN : a.mode};
        u32 handle[1 << LOGC];
        if constexpr (LOGC > 0) {
#pragma unroll
            for (int k = 0; k < (1 << LOGC); k++) handle[k] = peers.map((u32)k);
        }
        const u32 rt = (1u << LOGC) + rank;              // root base of this tile's own transform
        const u32 rb = (rt << (LOGT - 10)) + warp;       // ... and of this warp's block
        TB_MARK(a.dbg, bid, 1);
        __syncthreads();                                 // the barrier word is initialised
        mbar_wait(bar_in, 0);
        TB_MARK(a.dbg, bid, 2);
        TB_WITH_POLICY(pol, {
            if constexpr (LOGC == 0) {
                auto st_own = [&](u32, u32 w, u64 bits) { sm[w] = bits; };
                fwd_first_stage_staged<A, LOGT, LOGC, false>(sm, t, 0u, pro, st_own, md);
            } else {
                auto st_peer = [&](u32 tile, u32 w, u64 bits) { peers.st(handle[tile], w, bits); };
                fwd_first_stage_staged<A, LOGT, LOGC, true>(sm, t, rank, pro, st_peer, md);
            }
            TB_MARK(a.dbg, bid, 4);
            const typename A::Mod m = A::make(md);
            const void *tw = A::fwd_table(md);
            BlockTw<A> pre;
            if constexpr (!P::MID) pre.load(tw, rb); // in flight across the barrier
            if constexpr (LOGC == 0) __syncthreads();
            else cluster_barrier();
            if constexpr (P::MID) {
                round_tile<A, kThreads, false, false>(sm, t, tw, rt, m); // the three layers that couple the blocks of the tile
                pre.load(tw, rb);
                __syncthreads();
            }
            TB_MARK(a.dbg, bid, 5);
            fwd_block_stage<A>(sm, warp << 10, lane, tw, rb, m, pre);
            TB_MARK(a.dbg, bid, 8);
        })
        // this tile's outputs are coefficients [off, off + T) of the limb's (bit-reversed) output
        const u64 off = (u64)rank << LOGT;
        if (a.mode == FWD_MODUP && a.fused) {
            fused_tail(u, sm, off, warp, lane, pol);
        } else if (a.mode != FWD_MODDOWN) {
            // plain store: the block leaves through the tensor-map unit (the residues were written by this warp's lanes)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_rows(&a.tm_dst, (u32)(((u.dst - a.tm_dst_base) + off) >> 4) + 64u * warp, sm + (warp << 10));
                tma_store_commit_and_wait_read(); // the tile must outlive the unit's reads
            }
        } else {
            if (u.dst) u.dst += off;
            if (u.ct) u.ct += off;
            if (u.ct_in) u.ct_in += off;
            if (u.acc) u.acc += off;
            const FwdStore2 st{FwdStore{u, md->p, a.mode}};
            block_copy_out(sm, warp << 10, lane, st);
        }
#if defined(TROYB_DEV_TIMELINE)
        if (a.dbg != nullptr && threadIdx.x == 0) {
            u32 smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.dbg[(size_t)bid * 16 + 9] = tb_now();
            a.dbg[(size_t)bid * 16 + 10] = smid;
        }
#endif
    }

    // Fused mod-up: key products of this warp's block into the accumulators (block_key_product).
    TB_D void fused_tail(const NttUnit &u, u64 *sm, u64 off, u32 warp, u32 lane, int pol) const {
        const FusedKey f{u.key0 + off, u.key0 + a.key_comp_stride + off, u.prod0 + off, u.prod0 + a.prod_poly_stride + off,
                         u.pass ? u.pass + off : nullptr, u.md};
        if (pol == POLICY_F64 || pol == POLICY_F64_SMALL) block_key_product<true>(sm, warp << 10, lane, f);
        else block_key_product<false>(sm, warp << 10, lane, f);
    }

    template <int S> TB_D void stage(u32 bid, u32 grid, u32 rank, u32 nranks, u64 *sm, const Peers &peers) const {
        NttUnit u = resolve_unit<LOGT>(a, bid);
        if (u.passthrough) {
            // digit already in NTT form under this modulus: unfused, the inner product reads it in place; fused, its key
            // products are formed here from the NTT-form limb (no transform, no barriers: every CTA of the cluster takes this path)
            if (a.fused) fused_tail(u, sm, (u64)rank << LOGT, threadIdx.x >> 5, threadIdx.x & 31u, u.md->policy);
            return;
        }
        if constexpr (LOGT == 12) {
            run<0>(u, bid, grid, rank, sm, peers);
        } else {
            switch (u.rows_log) {
            case 0: run<0>(u, bid, grid, rank, sm, peers); break;
            case 1: run<1>(u, bid, grid, rank, sm, peers); break;
            case 2: run<2>(u, bid, grid, rank, sm, peers); break;
            default: run<3>(u, bid, grid, rank, sm, peers); break;
            }
        }
    }
};

template <int LOGT> struct NttInvTile2 {
    static constexpr int kThreads = LimbPlan<LOGT, 0>::NT;
    static constexpr int kMinCtas = LOGT == 13 ? 3 : 6;
    static constexpr int kStages = 1;
    static constexpr bool kTrailingBarrier = true; // peers read this CTA's tile in the last stage
    NttArgs a;
    const char *name() const { return ntt_kernel_name(a.mode); }
    TB_D bool barrier_after(int) const { return true; }

    template <int LOGC> TB_D void run(const NttUnit &u, u32 bid, u32 grid, u32 rank, u64 *sm, const Peers &peers) const {
        typedef LimbPlan<LOGT, LOGC> P;
        const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u, base = warp << 10;
        // the warp stages its own block: canonical residues, one TMA box of 64 rows
        u64 *bar_in = sm + (1u << LOGT); // barrier words live behind the tile (tile_smem2)
        if (lane == 0) {
            mbar_init(&bar_in[warp], 1);
            mbar_expect_tx(&bar_in[warp], 8192u);
            tma_load_rows(sm + base, &a.tm_src, (u32)(((u.src - a.tm_src_base) + ((u64)rank << LOGT) + base) >> 4), &bar_in[warp]);
        }
        const ModDev *md = u.md;
        const u64 p = md->p;
        const int pol = md->policy;
        prefetch_tiles<LOGT>(a, u, bid, rank, grid);
        u32 handle[1 << LOGC];
        if constexpr (LOGC > 0) {
#pragma unroll
            for (int k = 0; k < (1 << LOGC); k++) handle[k] = peers.map((u32)k);
        }
        const u32 rt = (1u << LOGC) + rank;
        const u32 rb = (rt << (LOGT - 10)) + warp;
        // the epilogue's mode is resolved once per launch, not once per stored word
        const int mode = a.mode;
        auto st_plain = [&](u32 i, u64 v) { u.dst[i] = v; };
        auto st_half = [&](u32 i, u64 v) { u.dst[i] = a.add_half ? addmod(v, a.drop.half, p) : v; };
        auto st_tail = [&](u32 i, u64 v) { inv_store_final(a, u, p, i, v); };
        TB_WITH_POLICY(pol, {
            const typename A::Mod m = A::make(md);
            const void *tw = A::inv_table(md);
            TwSet<A, Prefetching<A>::value ? 3 : 1> first;
            if constexpr (Prefetching<A>::value) first.load(tw, rb, 7, lane);
            __syncwarp();                                // the barrier word is initialised
            mbar_wait(&bar_in[warp], 0);
            inv_block_stage<A>(sm, base, lane, tw, rb, m, first);
            if constexpr (P::MID) {
                __syncthreads();
                round_tile<A, kThreads, true, A::kReduceEachInvRound>(sm, t, tw, rt, m);
            }
            if constexpr (LOGC == 0) {
                __syncthreads();
                auto ld_own = [&](u32, u32 w) -> u64 { return sm[w]; };
                if (mode == INV_PLAIN) inv_last_stage<A, LOGT, LOGC>(t, 0u, ld_own, st_plain, md);
                else if (mode == INV_ADD_HALF) inv_last_stage<A, LOGT, LOGC>(t, 0u, ld_own, st_half, md);
                else inv_last_stage<A, LOGT, LOGC>(t, 0u, ld_own, st_tail, md);
            } else {
                cluster_barrier();
                auto ld_peer = [&](u32 tile, u32 w) -> u64 { return peers.ld(handle[tile], w); };
                if (mode == INV_PLAIN) inv_last_stage<A, LOGT, LOGC>(t, rank, ld_peer, st_plain, md);
                else if (mode == INV_ADD_HALF) inv_last_stage<A, LOGT, LOGC>(t, rank, ld_peer, st_half, md);
                else inv_last_stage<A, LOGT, LOGC>(t, rank, ld_peer, st_tail, md);
            }
        })
    }

    template <int S> TB_D void stage(u32 bid, u32 grid, u32 rank, u32 nranks, u64 *sm, const Peers &peers) const {
        const NttUnit u = resolve_unit<LOGT>(a, bid);
        if constexpr (LOGT == 12) {
            run<0>(u, bid, grid, rank, sm, peers);
        } else {
            switch (u.rows_log) {
            case 0: run<0>(u, bid, grid, rank, sm, peers); break;
            case 1: run<1>(u, bid, grid, rank, sm, peers); break;
            case 2: run<2>(u, bid, grid, rank, sm, peers); break;
            default: run<3>(u, bid, grid, rank, sm, peers); break;
            }
        }
    }
};

} // namespace troyb
