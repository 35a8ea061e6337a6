// Launchers of the tile transforms. The NTT kernel templates are by far the largest part of the library's device code, so
// they are compiled in translation units of their own (ntt_fwd.cu, ntt_inv.cu) next to api.cu.
#pragma once
#if defined(TROYB_HOST_EMUL)
#include "kernels_ntt.cuh"
#else
#include "kernels_ntt2.cuh"
#include <initializer_list>
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
inline size_t tile_smem2(int logt) { return (sizeof(u64) << logt) + 128; } // the tile + the mbarrier words
inline bool use_ntt2() {
    static const bool on = [] { const char *e = getenv("TROYB_NTT2"); return !e || atoi(e) != 0; }();
    return on;
}
// Geometry of the second-generation plans on the host (mirrors LimbPlan in ntt2.cuh): tile size and rows per staging box.
inline int tile2_log(int logN) { return logN == 12 ? 12 : 13; }
inline u32 stage_box_rows(int logN) {
    const int logt = tile2_log(logN), logc = logN - logt;
    const bool mid = logc + logt - 10 > 4;
    const int f1 = mid ? logc : logc + logt - 10;
    return (1u << (logt - f1)) >> 4;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda).
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
    static const TensorMapEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}
// The buffer at `base` as [rows][16 words] with the 128-byte swizzle; a box is `box_rows` whole rows.
inline cudaError_t make_rows_map(CUtensorMap *tm, const u64 *base, u32 box_rows) {
    const TensorMapEncodeFn fn = tensor_map_encoder();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {16, 1ull << 31};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, box_rows}, elem[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<u64 *>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
// Units address the maps by 16-word rows: every stride that separates units must be a multiple of 16 words.
inline bool rows_ok(std::initializer_list<u64> strides) {
    for (u64 v : strides) if (v & 15u) return false;
    return true;
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
