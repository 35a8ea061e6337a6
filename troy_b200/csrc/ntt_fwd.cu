// Forward tile transforms: the only translation unit that instantiates the NttFwd* kernels (ntt_launch.h).
#include "ntt_launch.h"

namespace troyb {

cudaError_t launch_ntt_fwd(const Context *ctx, NttArgs a, u32 units, cudaStream_t s) {
    const int logN = ctx->logN;
    a.logN = logN;
    a.pf_ahead = prefetch_distance(ctx, tile_log(logN));
    a.mods = ctx->d_mods;
    a.solo = 0;
    if (logN < 12) return launch_cta(NttFwdSmall{a}, units, (size_t)ctx->N * sizeof(u64), s);
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
