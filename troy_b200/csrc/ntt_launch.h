// Launchers of the tile transforms. The NTT kernel templates are by far the largest part of the library's device code, so
// they are compiled in translation units of their own (ntt_fwd.cu, ntt_inv.cu) next to api.cu.
#pragma once
#if defined(TROYB_HOST_EMUL)
#include "kernels_ntt.cuh"
#else
#include "kernels_ntt2.cuh"
#endif

namespace troyb {

// Tile size of a transform of 2^logN coefficients: 4096-coefficient tiles (four radix-8 rounds, six CTAs per SM) for
// N = 4096 and, as clusters of 4 / 8 tiles, for N up to 2^TROYB_TILE12_MAXLOGN (development knob for A/B runs);
// 8192-coefficient tiles otherwise.
inline int tile_log(int logN) {
    static const int max12 = [] { const char *e = getenv("TROYB_TILE12_MAXLOGN"); return e ? atoi(e) : 12; }();
    return (logN == 12 || (logN > 13 && logN <= max12 && logN <= 15)) ? 12 : 13;
}
inline u32 tiles_per_limb(int logN) { const int lt = tile_log(logN); return logN <= lt ? 1u : (1u << (logN - lt)); }
inline size_t tile_smem(int logt) { return (size_t)padded_words(1u << logt) * sizeof(u64); }
// Warp-owned blocks (ntt2.cuh) are what the library runs; TROYB_NTT2=0 (development knob for A/B runs) selects the
// first-generation tiles, which are also what the host emulation (tests/emul) executes.
#if defined(TROYB_HOST_EMUL)
inline bool use_ntt2() { return false; }
#else
inline size_t tile_smem2(int logt) { return (size_t)padded2_words(1u << logt) * sizeof(u64); }
inline bool use_ntt2() {
    static const bool on = [] { const char *e = getenv("TROYB_NTT2"); return !e || atoi(e) != 0; }();
    return on;
}
#endif

// L2 prefetch distance of the tile transforms: about one wave of resident CTAs (SMs x CTAs per SM, rounded down to
// whole clusters). TROYB_PF_WAVES_X4 (development knob) scales it in quarter waves; 0 switches the prefetch off.
inline u32 prefetch_distance(const Context *ctx, int logt) {
    static const int quarters = [] { const char *e = getenv("TROYB_PF_WAVES_X4"); return e ? atoi(e) : 4; }();
    const u32 wave = (u32)ctx->sm_count * (logt == 12 ? 6u : 3u);
    return (wave * (u32)quarters / 4u) & ~7u;
}

// `units` limb transforms described by `a` (a.logN, a.mods, a.pf_ahead and a.solo are filled in here).
cudaError_t launch_ntt_fwd(const Context *ctx, NttArgs a, u32 units, cudaStream_t s);
cudaError_t launch_ntt_inv(const Context *ctx, NttArgs a, u32 units, cudaStream_t s);

} // namespace troyb
