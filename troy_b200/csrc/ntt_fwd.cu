// Forward tile transforms: the only translation unit that instantiates the NttFwd* kernels (ntt_launch.h).
#include "ntt_launch.h"
#include <cstdio>
#include <vector>

namespace troyb {

#if !defined(TROYB_HOST_EMUL)
// Tensor maps of a forward launch: inputs (the coefficient-form source; the rounded special limb s' for the mod-down tail) and,
// for plain stores, outputs. False when a stride of the launch is not a whole number of 16-word rows (first-generation kernels).
static bool fwd_maps(NttArgs &a, int logN) {
    const u64 N = u64(1) << logN;
    const u64 *src = a.mode == FWD_MODDOWN ? a.aux1 : a.src;
    if (!rows_ok({N, a.src_unit_stride, a.dst_batch_stride, a.dst_poly_stride, (u64)(uintptr_t)src >> 3, (u64)(uintptr_t)a.dst >> 3})) return false;
    a.tm_src_base = src;
    a.tm_dst_base = a.fused ? a.prod : a.dst; // fused mod-up: no stores through the map (bulk reduce-adds into prod instead)
    if (a.fused && !rows_ok({a.prod_batch_stride, a.prod_poly_stride, a.prod_limb_stride, a.key_digit_stride, a.key_comp_stride,
                             (u64)(uintptr_t)a.prod >> 3, (u64)(uintptr_t)a.key >> 3, a.aux0_batch_stride, (u64)(uintptr_t)a.aux0 >> 3}))
        return false;
    if (make_rows_map(&a.tm_src, src, stage_box_rows(logN)) != cudaSuccess) return false;
    if (a.mode != FWD_MODDOWN && make_rows_map(&a.tm_dst, a.tm_dst_base, 64) != cudaSuccess) return false;
    return true;
}
#endif

cudaError_t launch_ntt_fwd(const Context *ctx, NttArgs a, u32 units, cudaStream_t s) {
    const int logN = ctx->logN;
    a.logN = logN;
    a.pf_ahead = prefetch_distance(ctx, tile_log(logN));
    a.mods = ctx->d_mods;
    a.solo = 0;
    if (logN < 12) return launch_cta(NttFwdSmall{a}, units, (size_t)ctx->N * sizeof(u64), s);
#if defined(TROYB_DEV_TIMELINE)
    static const char *timeline = getenv("TROYB_TIMELINE");
    if (use_ntt2() && timeline && logN >= 13) {
        // development build: one instrumented launch, timeline dumped as text (CTA, smid, nanoseconds since the first mark)
        const u32 c2 = 1u << (logN - 13), grid = units * c2;
        unsigned long long *d = nullptr;
        cudaMalloc(&d, (size_t)grid * 16 * 8);
        cudaMemset(d, 0, (size_t)grid * 16 * 8);
        a.dbg = d;
        static const int extra = [] { const char *x = getenv("TROYB_SMEM_EXTRA"); return x ? atoi(x) : 0; }();
        if (!fwd_maps(a, logN)) return cudaErrorInvalidValue;
        cudaError_t e = launch_cluster(NttFwdTile2<13>{a}, grid, c2, tile_smem2(13) + (size_t)extra, s);
        cudaStreamSynchronize(s);
        std::vector<unsigned long long> h((size_t)grid * 16);
        cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (FILE *f = fopen(timeline, "w")) {
            unsigned long long t0 = ~0ull;
            for (u32 b = 0; b < grid; b++) if (h[(size_t)b * 16] && h[(size_t)b * 16] < t0) t0 = h[(size_t)b * 16];
            for (u32 b = 0; b < grid; b++) {
                fprintf(f, "%u %llu", b, h[(size_t)b * 16 + 10]);
                for (int k = 0; k < 10; k++) fprintf(f, " %lld", h[(size_t)b * 16 + k] ? (long long)(h[(size_t)b * 16 + k] - t0) : -1ll);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        return e;
    }
#endif
#if !defined(TROYB_HOST_EMUL)
    if (use_ntt2() && fwd_maps(a, logN)) {
        if (logN == 12) return launch_cluster(NttFwdTile2<12>{a}, units, 1u, tile_smem2(12), s);
        const u32 c2 = 1u << (logN - 13);
        return launch_cluster(NttFwdTile2<13>{a}, units * c2, c2, tile_smem2(13), s);
    }
    if (a.fused) return cudaErrorNotSupported; // the caller checks fused_modup_possible() first
#endif
    const u32 c = tiles_per_limb(logN);
    if (tile_log(logN) == 12) return launch_cluster(NttFwdTile<12>{a}, units * c, c, tile_smem(12), s);
    // a limb of two tiles runs as two independent CTAs that both recompute layer 0 (kernels_ntt.cuh FwdSoloLd);
    // TROYB_SOLO=0 (development knob) keeps the cluster of two with its scatter stage
    static const bool solo = [] { const char *e = getenv("TROYB_SOLO"); return !e || atoi(e) != 0; }();
    a.solo = (c == 2 && solo) ? 1 : 0;
    // in place the two CTAs of a limb must both have loaded before either stores: a cluster of two for its barrier
    const bool in_place = a.mode == FWD_PLAIN && a.src == a.dst;
    return launch_cluster(NttFwdTile<13>{a}, units * c, (a.solo && !in_place) ? 1u : c, tile_smem(13), s);
}

} // namespace troyb
