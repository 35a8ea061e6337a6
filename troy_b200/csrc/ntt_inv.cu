// Inverse tile transforms: the only translation unit that instantiates the NttInv* kernels (ntt_launch.h).
#include "ntt_launch.h"

namespace troyb {

#if !defined(TROYB_HOST_EMUL)
// Tensor map of an inverse launch's inputs (NTT-form source; the key products for the BFV / BGV mod-down tail): a box is one
// block of 1024 coefficients. False when a stride is not a whole number of 16-word rows (first-generation kernels).
static bool inv_maps(NttArgs &a, int logN) {
    const u64 N = u64(1) << logN;
    const u64 *src = a.mode == INV_COMBINE ? a.aux0 : a.src;
    if (!rows_ok({N, a.src_unit_stride, a.src_poly_stride, a.src_offset, a.acc_batch_stride, a.acc_poly_stride, a.acc_limb_stride,
                  (u64)(uintptr_t)src >> 3})) return false;
    a.tm_src_base = src;
    a.tm_dst_base = nullptr;
    return make_rows_map(&a.tm_src, src, 64) == cudaSuccess;
}
#endif

cudaError_t launch_ntt_inv(const Context *ctx, NttArgs a, u32 units, cudaStream_t s) {
    const int logN = ctx->logN;
    a.logN = logN;
    a.pf_ahead = prefetch_distance(ctx, tile_log(logN));
    a.mods = ctx->d_mods;
    a.solo = 0;
    if (logN < 12) return launch_cta(NttInvSmall{a}, units, (size_t)ctx->N * sizeof(u64), s);
#if !defined(TROYB_HOST_EMUL)
    if (use_ntt2() && inv_maps(a, logN)) {
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
