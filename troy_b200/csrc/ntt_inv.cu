// Inverse tile transforms: the only translation unit that instantiates the NttInv* kernels (ntt_launch.h).
#include "ntt_launch.h"

namespace troyb {

cudaError_t launch_ntt_inv(const Context *ctx, NttArgs a, u32 units, cudaStream_t s) {
    const int logN = ctx->logN;
    a.logN = logN;
    a.pf_ahead = prefetch_distance(ctx, tile_log(logN));
    a.mods = ctx->d_mods;
    a.solo = 0;
    if (logN < 12) return launch_cta(NttInvSmall{a}, units, (size_t)ctx->N * sizeof(u64), s);
#if !defined(TROYB_HOST_EMUL)
    if (use_ntt2()) {
        if (logN == 12) return launch_cluster(NttInvTile2<12>{a}, units, 1u, tile_smem2(12), s);
        const u32 c2 = 1u << (logN - 13);
        return launch_cluster(NttInvTile2<13>{a}, units * c2, c2, tile_smem2(13), s);
    }
#endif
    const u32 c = tiles_per_limb(logN);
    if (tile_log(logN) == 12) return launch_cluster(NttInvTile<12>{a}, units * c, c, tile_smem(12), s);
    return launch_cluster(NttInvTile<13>{a}, units * c, c, tile_smem(13), s);
}

} // namespace troyb
