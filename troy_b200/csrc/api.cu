// C ABI of troy_b200 (include/troy_b200.h): argument checks, scratch management and kernel orchestration.
// The host control flow restates the reference's evaluator methods (cited per function); the kernels do not.
#include "../../include/troy_b200.h"
#include "context.h"
#include "kernels_elt.cuh"
#include "kernels_ntt.cuh"
#include "kernels_rlwe.cuh"
#include "ntt_launch.h"
#include "hostmath.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

using namespace troyb;

struct troyb_context {
    Context *impl;
};

namespace {

thread_local std::string g_error;
std::atomic<bool> g_initialized{false};
// The CUDA "current device" is a per-host-thread setting. One process may drive several GPUs (one context per device,
// typically one host thread per device): every entry point that takes a context makes the context's device current for the
// calling thread first, and the context-free calls (troyb_malloc, copies) act on the thread's current device, which
// troyb_initialize / troyb_set_device choose.
thread_local int t_device = -1;
void use_device(int device) {
    if (t_device == device) return;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e));
    t_device = device;
}

struct CudaError : std::runtime_error {
    explicit CudaError(cudaError_t e) : std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e)) {}
};

inline void ck(cudaError_t e) {
    if (e != cudaSuccess) throw CudaError(e);
}

int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

#define API_BEGIN try {
#define API_END                                                                                       \
    return TROYB_OK;                                                                                  \
    }                                                                                                 \
    catch (const CudaError &e) { return fail(TROYB_E_CUDA, e.what()); }                               \
    catch (const std::invalid_argument &e) { return fail(TROYB_E_INVALID_ARGUMENT, e.what()); }       \
    catch (const std::out_of_range &e) { return fail(TROYB_E_OUT_OF_RANGE, e.what()); }               \
    catch (const std::logic_error &e) { return fail(TROYB_E_LOGIC, e.what()); }                       \
    catch (const std::exception &e) { return fail(TROYB_E_CUDA, e.what()); }

void need_init() {
    if (!g_initialized) throw std::invalid_argument("KernelProvider not initialized.");
}

inline cudaStream_t S(void *s) { return static_cast<cudaStream_t>(s); }

// Stream-ordered scratch buffer (RAII).
struct Scratch {
    u64 *p = nullptr;
    cudaStream_t s;
    Scratch(size_t words, cudaStream_t stream) : s(stream) {
        if (words) ck(cudaMallocAsync(reinterpret_cast<void **>(&p), words * sizeof(u64), s));
    }
    ~Scratch() {
        if (p) cudaFreeAsync(p, s);
    }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};

Context *C(troyb_context *c) {
    if (!c || !c->impl) throw std::invalid_argument("context is null");
    use_device(c->impl->device);
    return c->impl;
}
const Context *C(const troyb_context *c) {
    if (!c || !c->impl) throw std::invalid_argument("context is null");
    use_device(c->impl->device);
    return c->impl;
}

int level_limbs(const Context *ctx, int level) {
    if (level == -1) return ctx->Ltop + (ctx->has_special ? 1 : 0);
    if (level < 0 || level >= (int)ctx->levels.size()) throw std::invalid_argument("level is not valid for encryption parameters");
    return ctx->levels[level].L;
}

u32 grid_for(u64 total, int threads) { return (u32)((total + threads - 1) / threads); }

// ---- NTT dispatch (the transforms are compiled in ntt_fwd.cu / ntt_inv.cu, see ntt_launch.h) -------------------------
// Launch order of the limbs of a launch. Consecutive CTAs cycle through the limbs; limbs whose butterflies run on the
// integer pipe (60-bit primes) are spread evenly between those on the FP64 pipe (primes below 2^46), so that the CTAs
// resident on one SM at any time keep both pipes busy.
void set_order(const Context *ctx, NttArgs &a, u32 n_limbs_in_map) {
    std::vector<int> heavy, light;
    for (u32 i = 0; i < n_limbs_in_map; i++) {
        const int pol = ctx->h_mods[a.map.idx[i]].policy;
        (pol == POLICY_U64_FAST || pol == POLICY_U64_EXACT || pol == POLICY_U64_LAZY ? heavy : light).push_back((int)i);
    }
    size_t ih = 0, il = 0;
    for (u32 r = 0; r < n_limbs_in_map; r++) {
        // Bresenham merge: heavy limb whenever its share of the ranks so far falls behind
        const bool take_heavy = il >= light.size() || (ih < heavy.size() && ih * n_limbs_in_map <= r * heavy.size());
        a.order[r] = (uint8_t)(take_heavy ? heavy[ih++] : light[il++]);
    }
    a.n_rank = n_limbs_in_map;
}

void run_fwd(const Context *ctx, const NttArgs &a, u32 units, cudaStream_t s) {
    if (ctx->logN > 16) throw std::invalid_argument("poly_modulus_degree above 65536 is not supported");
    ck(launch_ntt_fwd(ctx, a, units, s));
}
void run_inv(const Context *ctx, const NttArgs &a, u32 units, cudaStream_t s) {
    if (ctx->logN > 16) throw std::invalid_argument("poly_modulus_degree above 65536 is not supported");
    ck(launch_ntt_inv(ctx, a, units, s));
}

// Plain batched transform: `batch` elements of `limbs_per_elem` limbs (polys x L), moduli map[l % L].
void ntt_plain(const Context *ctx, bool inverse, const LimbMap &map, u32 L, const u64 *src, u64 src_stride, u64 *dst,
               u64 dst_stride, u32 limbs_per_elem, u32 batch, bool canon, cudaStream_t s) {
    NttArgs a{};
    a.mode = inverse ? INV_PLAIN : FWD_PLAIN;
    a.map = map;
    a.n_limbs = limbs_per_elem;
    a.L = L;
    a.src = src;
    a.dst = dst;
    a.src_unit_stride = src_stride;
    a.dst_batch_stride = dst_stride;
    a.canon = canon ? 1 : 0;
    a.scheme = ctx->scheme;
    set_order(ctx, a, L);
    if (inverse) run_inv(ctx, a, batch * limbs_per_elem, s);
    else run_fwd(ctx, a, batch * limbs_per_elem, s);
}

// ---- key switching ---------------------------------------------------------------------------------------------------
// Host control flow of EvaluatorCuda::switchKeyInplace (src/evaluator_cuda.cu:1163-1378) over a batch, in two halves:
//   partial : (CKKS) INTT the target digits  →  mod-up NTTs for all (output modulus i, digit j)  →  inner product with
//             the key, for the digits [digit0, digit0 + nd) only: prod = sum over those digits        (:1222-1297)
//   finish  : INTT the special-prime rows (+ half)  →  mod-down tail fused into the last transform of every data limb,
//             added into the ciphertext                                                                 (:1299-1376)
// One GPU runs partial over all digits and then finish. With the digits of one key-switch sharded over several GPUs
// (troy_b200/distributed.py) every GPU runs partial on its digits, the products are summed across GPUs (the only
// exchange of the path), and finish first reduces the sums modulo each prime.
// Where the key inner products live. The default is the reference's t_poly_prod layout [batch][2][L+1][N]; `rows` is
// [L+1 output moduli][batch][2][N], in which the rows a GPU owns are one contiguous block (reduce-scatter over NVLink).
struct ProdLayout {
    u64 batch_stride, poly_stride, limb_stride; // words
    static ProdLayout standard(u32 Lp, u64 N) { return ProdLayout{2ull * Lp * N, (u64)Lp * N, N}; }
    static ProdLayout rows(u32 batch, u64 N) { return ProdLayout{2ull * N, N, 2ull * batch * N}; }
};

// Returns true when prod holds UNREDUCED sums of canonical residues (the fused path): the caller reduces them modulo each
// prime before use (switch_key_finish(reduce_first)).
bool switch_key_partial(Context *ctx, int level, const u64 *target, u64 target_stride, const u64 *key, u32 digit0, u32 nd,
                        u64 *prod, u32 batch, cudaStream_t s, const ProdLayout *layout = nullptr) {
    if (!ctx->has_special) throw std::logic_error("keyswitching is not supported by the context");
    const u32 L = (u32)ctx->levels[level].L;
    const u64 N = ctx->N;
    const u32 Lp = L + 1;
    const u32 key_limbs = (u32)ctx->Ltop + 1;
    const bool ckks = ctx->scheme == SCHEME_CKKS;
    if (nd == 0 || digit0 + nd > L) throw std::invalid_argument("digit range is not valid");
    const ProdLayout pl = layout ? *layout : ProdLayout::standard(Lp, N);
    LimbMap kmap = ctx->map_prefix(L);
    kmap.idx[L] = (uint8_t)ctx->pool_special;
    LimbMap dmap{}; // moduli of the digits in play
    for (u32 j = 0; j < nd; j++) dmap.idx[j] = kmap.idx[digit0 + j];

    // Only the mod-up output D ([L+1][nd] limbs per ciphertext) is large. It is produced and consumed in chunks of at
    // most 4 GiB (TROYB_KS_CHUNK_MB overrides): measured on B200, few large launches beat L2-sized chunks (the kernels are
    // bound by the FP64 / integer pipes, not by the HBM round trip of D, and small launches lose to partial waves).
    const u64 per_ct_words = (u64)Lp * nd * N;
    const char *chunk_env = getenv("TROYB_KS_CHUNK_MB");
    const u64 chunk_bytes = (chunk_env ? (u64)atoi(chunk_env) : 4096ull) << 20;
    const u32 max_chunk = (u32)std::max<u64>(1, chunk_bytes / (per_ct_words * 8));
    const u32 n_chunks = (batch + max_chunk - 1) / max_chunk;
    const u32 chunk = (batch + n_chunks - 1) / n_chunks; // equal chunks

    Scratch t_target(ckks ? (size_t)batch * nd * N : 0, s);
    if (ckks) ntt_plain(ctx, true, dmap, nd, target + (u64)digit0 * N, target_stride, t_target.p, (u64)nd * N, nd, batch, true, s);

#if !defined(TROYB_HOST_EMUL)
    // Fused mod-up + key inner product (TROYB_KS_FUSED=1, off by default: measured 10 % slower on B200): every
    // transform multiplies its output by the key's two components and reduce-adds the products into prod at the L2 (bulk
    // reduce-adds), so D is never allocated and the inner-product kernel is not launched. Needs the standard product layout,
    // the second-generation transforms, and sums of nd + 2 canonical residues that fit 64 bits. The switch is read per call.
    const char *fe = getenv("TROYB_KS_FUSED");
    const bool fused_env = fe && atoi(fe) != 0;
    if (fused_env && layout == nullptr && use_ntt2() && ctx->logN >= 12) {
        // every accumulator takes nd + 1 products - canonical, or below 2 p on the FP64 limbs - and the mod-down tail adds one prime
        bool fits = true;
        for (u32 i = 0; i < Lp; i++) {
            const ModDev &md = ctx->h_mods[kmap.idx[i]];
            const u32 lazy = (md.policy == POLICY_F64 || md.policy == POLICY_F64_SMALL) ? 2 : 1;
            fits = fits && (unsigned __int128)md.p * (lazy * (nd + 1) + 1) < ((unsigned __int128)1 << 64);
        }
        if (fits) {
            ck(cudaMemsetAsync(prod, 0, (size_t)batch * pl.batch_stride * sizeof(u64), s));
            NttArgs a{};
            a.mode = FWD_MODUP;
            a.map = kmap;
            a.L = nd;
            a.digit0 = digit0;
            a.src = ckks ? t_target.p : target + (u64)digit0 * N;
            a.src_unit_stride = ckks ? (u64)nd * N : target_stride;
            a.aux0 = ckks ? target : nullptr;
            a.aux0_batch_stride = target_stride;
            a.dst = nullptr;
            a.scheme = ctx->scheme;
            a.fused = 1;
            a.key = key + (u64)digit0 * 2 * key_limbs * N;
            a.key_digit_stride = 2ull * key_limbs * N;
            a.key_comp_stride = (u64)key_limbs * N;
            a.key_special_limb = key_limbs - 1;
            a.prod = prod;
            a.prod_batch_stride = pl.batch_stride;
            a.prod_poly_stride = pl.poly_stride;
            a.prod_limb_stride = pl.limb_stride;
            set_order(ctx, a, Lp);
            run_fwd(ctx, a, batch * Lp * nd, s);
            return true;
        }
    }
#endif
    Scratch D((size_t)chunk * per_ct_words, s);

    for (u32 b0 = 0; b0 < batch; b0 += chunk) {
        const u32 nb = std::min(chunk, batch - b0);
        const u64 *tgt = target + (u64)b0 * target_stride;
        {
            NttArgs a{};
            a.mode = FWD_MODUP;
            a.map = kmap;
            a.L = nd;
            a.digit0 = digit0;
            a.src = ckks ? t_target.p + (u64)b0 * nd * N : tgt + (u64)digit0 * N;
            a.src_unit_stride = ckks ? (u64)nd * N : target_stride;
            a.aux0 = ckks ? tgt : nullptr;
            a.aux0_batch_stride = target_stride;
            a.dst = D.p;
            a.scheme = ctx->scheme;
            set_order(ctx, a, Lp);
            run_fwd(ctx, a, nb * Lp * nd, s);
        }
        if (N >= kStagedCoeffs) {
            // key slices staged in shared memory, digits in groups of at most 8, partial sums added up in prod
            // batch elements per CTA (each CTA stages its key slice once): about eight waves of CTAs
            const u64 units = (u64)nb * Lp * (N / kStagedCoeffs);
            static const u32 bc_max = [] { const char *e = getenv("TROYB_IP_BC"); return e ? (u32)std::max(1, atoi(e)) : 32u; }(); // cap; 4 / 8 / 16 / 32 / 64 measure 0.672 / 0.627 / 0.594 / 0.588 / 0.588 ms at the bench batch
            const u32 bc = (u32)std::min<u64>(bc_max, std::max<u64>(1, units / ((u64)ctx->sm_count * 3 * 8)));
            // every row (output modulus) in one launch; CTAs of rows whose prime has at most 40 bits take the short-accumulator form
            // (SmallAcc) unless TROYB_IP_SMALL=0
            static const bool use_small = [] { const char *e = getenv("TROYB_IP_SMALL"); return !(e && e[0] == '0'); }();
            std::vector<uint8_t> rows_all;
            for (u32 i = 0; i < Lp; i++) rows_all.push_back((uint8_t)i);
            for (u32 g0 = 0; g0 < nd; g0 += kStagedMaxDigits) {
                auto launch = [&](auto k, const std::vector<uint8_t> &rows) {
                    k.logN = ctx->logN;
                    k.mods = ctx->d_mods;
                    k.map = kmap;
                    k.L = L;
                    k.digit0 = digit0 + g0;
                    k.key_limbs = key_limbs;
                    k.special_key_limb = key_limbs - 1;
                    k.batch = nb;
                    k.bc = bc;
                    k.d_rows = nd;
                    k.d_row0 = g0;
                    k.accumulate = g0 > 0;
                    k.D = D.p;
                    k.key = key;
                    k.prod = prod + (u64)b0 * pl.batch_stride;
                    k.prod_batch_stride = pl.batch_stride;
                    k.prod_poly_stride = pl.poly_stride;
                    k.prod_limb_stride = pl.limb_stride;
                    k.pass = ckks ? tgt : nullptr;
                    k.pass_stride = target_stride;
                    k.n_rows = (u32)rows.size();
                    std::copy(rows.begin(), rows.end(), k.rows);
                    const u32 grid = (u32)(N / kStagedCoeffs) * k.n_rows * ((nb + bc - 1) / bc);
                    ck(launch_cta(k, grid, decltype(k)::smem_bytes(), s));
                };
                const u32 ndg = std::min(kStagedMaxDigits, nd - g0);
                if (!use_small) {
                    switch (ndg) {
                    case 1: launch(KsInnerProductStaged<1>{}, rows_all); break;
                    case 2: launch(KsInnerProductStaged<2>{}, rows_all); break;
                    case 3: launch(KsInnerProductStaged<3>{}, rows_all); break;
                    case 4: launch(KsInnerProductStaged<4>{}, rows_all); break;
                    case 5: launch(KsInnerProductStaged<5>{}, rows_all); break;
                    case 6: launch(KsInnerProductStaged<6>{}, rows_all); break;
                    case 7: launch(KsInnerProductStaged<7>{}, rows_all); break;
                    default: launch(KsInnerProductStaged<8>{}, rows_all); break;
                    }
                } else {
                    switch (ndg) {
                    case 1: launch(KsInnerProductStaged<1, true>{}, rows_all); break;
                    case 2: launch(KsInnerProductStaged<2, true>{}, rows_all); break;
                    case 3: launch(KsInnerProductStaged<3, true>{}, rows_all); break;
                    case 4: launch(KsInnerProductStaged<4, true>{}, rows_all); break;
                    case 5: launch(KsInnerProductStaged<5, true>{}, rows_all); break;
                    case 6: launch(KsInnerProductStaged<6, true>{}, rows_all); break;
                    case 7: launch(KsInnerProductStaged<7, true>{}, rows_all); break;
                    default: launch(KsInnerProductStaged<8, true>{}, rows_all); break;
                    }
                }
            }
        } else {
            KsInnerProduct k{};
            k.logN = ctx->logN;
            k.mods = ctx->d_mods;
            k.map = kmap;
            k.L = L;
            k.nd = nd;
            k.digit0 = digit0;
            k.key_limbs = key_limbs;
            k.special_key_limb = key_limbs - 1;
            k.total = (u64)nb * Lp * (N / 2);
            k.D = D.p;
            k.key = key;
            k.prod = prod + (u64)b0 * pl.batch_stride;
            k.prod_batch_stride = pl.batch_stride;
            k.prod_poly_stride = pl.poly_stride;
            k.prod_limb_stride = pl.limb_stride;
            k.pass = ckks ? tgt : nullptr;
            k.pass_stride = target_stride;
            ck(launch_cta(k, grid_for(k.total, KsInnerProduct::kThreads), 0, s));
        }
    }
    return false;
}

// The second half of a key switch in three phases, each over a range of rows (output moduli) so that GPUs that own
// different rows can run them side by side (troy_b200/distributed.py):
//   reduce  : rows [r0, r1) hold sums of several GPUs' canonical partial products: reduce modulo each row's prime
//   special : INTT of the special-prime row (+ half) -> sp [batch][2][N]                       (:1299-1331)
//   moddown : for the data limbs [j0, j1): mod-down tail fused into the last transform, added into ct   (:1333-1376)
void switch_key_reduce_rows(Context *ctx, int level, u64 *prod, const ProdLayout &pl, u32 r0, u32 r1, u32 batch, cudaStream_t s) {
    const u32 L = (u32)ctx->levels[level].L;
    const u64 N = ctx->N;
    LimbMap kmap = ctx->map_prefix(L);
    kmap.idx[L] = (uint8_t)ctx->pool_special;
    EltAddSub k{};
    k.op = 4;
    k.logN = ctx->logN;
    k.mods = ctx->d_mods;
    if (pl.limb_stride == N) { // standard layout: every row of every ciphertext in one pass
        if (r0 != 0 || r1 != L + 1) throw std::invalid_argument("row ranges need the row-major product layout");
        k.map = kmap;
        k.n_limbs = L + 1;
        k.total = (u64)batch * 2 * (L + 1) * N;
        k.a = k.b = k.out = prod;
    } else {
        for (u32 r = r0; r < r1; r++) k.map.idx[r - r0] = kmap.idx[r];
        k.n_limbs = r1 - r0;
        k.row_words = pl.limb_stride;
        k.total = (u64)(r1 - r0) * pl.limb_stride;
        k.a = k.b = k.out = prod + (u64)r0 * pl.limb_stride;
    }
    ck(launch_cta(k, grid_for(k.total, EltAddSub::kThreads), 0, s));
}

void switch_key_special(Context *ctx, int level, const u64 *prod, const ProdLayout &pl, u64 *sp, u32 batch, cudaStream_t s) {
    const u32 L = (u32)ctx->levels[level].L;
    LimbMap kmap = ctx->map_prefix(L);
    kmap.idx[L] = (uint8_t)ctx->pool_special;
    NttArgs a{};
    a.mode = INV_ADD_HALF;
    a.map = kmap;
    a.L = L;
    a.K = 2;
    a.src = prod;
    a.src_offset = (u64)L * pl.limb_stride;
    a.src_unit_stride = pl.batch_stride;
    a.src_poly_stride = pl.poly_stride;
    a.dst = sp;
    a.drop = ctx->key_drop;
    a.add_half = ctx->scheme != SCHEME_BGV;
    a.scheme = ctx->scheme;
    run_inv(ctx, a, batch * 2, s);
}

// ct_src (optional): the tail writes ct = ct_src + result for the first ct_src_polys polys and ct = result for the rest,
// instead of ct += result (applyGalois: permuted c0 in scratch, c1 from zero; saves a copy and a memset pass over ct).
void switch_key_moddown(Context *ctx, int level, u64 *ct, u64 ct_stride, const u64 *prod, const ProdLayout &pl, const u64 *sp,
                        u32 j0, u32 j1, u32 batch, cudaStream_t s, const u64 *ct_src = nullptr, u64 ct_src_stride = 0,
                        u32 ct_src_polys = 0) {
    const u32 L = (u32)ctx->levels[level].L;
    const bool ckks = ctx->scheme == SCHEME_CKKS;
    if (j1 > L) j1 = L;
    if (j0 >= j1) return;
    LimbMap kmap = ctx->map_prefix(L);
    kmap.idx[L] = (uint8_t)ctx->pool_special;
    NttArgs a{};
    a.mode = ckks ? FWD_MODDOWN : INV_COMBINE;
    a.map = kmap;
    a.L = L;
    a.K = 2;
    a.aux0 = prod;
    a.acc_batch_stride = pl.batch_stride;
    a.acc_poly_stride = pl.poly_stride;
    a.acc_limb_stride = pl.limb_stride;
    a.aux1 = sp;
    a.ct = ct;
    a.ct_batch_stride = ct_stride;
    a.ct_poly_stride = (u64)L * ctx->N;
    a.ct_src = ct_src;
    a.ct_src_batch_stride = ct_src_stride;
    a.ct_src_polys = ct_src_polys;
    a.drop = ctx->key_drop;
    a.scheme = ctx->scheme;
    set_order(ctx, a, L);
    if (j0 != 0 || j1 != L) { // a subset of the limbs: the launch's ranks are the owned limbs only
        u32 n = 0;
        uint8_t keep[kMaxLimbs];
        for (u32 r = 0; r < L; r++)
            if (a.order[r] >= j0 && a.order[r] < j1) keep[n++] = a.order[r];
        for (u32 r = 0; r < n; r++) a.order[r] = keep[r];
        a.n_rank = n;
    }
    if (ckks) run_fwd(ctx, a, batch * 2 * a.n_rank, s);
    else run_inv(ctx, a, batch * 2 * a.n_rank, s);
}

void switch_key_finish(Context *ctx, int level, u64 *ct, u64 ct_stride, u64 *prod, bool reduce_first, u32 batch, cudaStream_t s,
                       const u64 *ct_src = nullptr, u64 ct_src_stride = 0, u32 ct_src_polys = 0) {
    if (!ctx->has_special) throw std::logic_error("keyswitching is not supported by the context");
    const u32 L = (u32)ctx->levels[level].L;
    const ProdLayout pl = ProdLayout::standard(L + 1, ctx->N);
    Scratch sp((size_t)batch * 2 * ctx->N, s);
    if (reduce_first) {
        if (ctx->scheme == SCHEME_CKKS) {
            // the mod-down tail takes unreduced accumulators as they are (FwdStore::finish); only the special-prime rows, which
            // feed an inverse transform, are reduced here
            EltReduceRow k{};
            k.logN = ctx->logN;
            k.mod = ctx->h_mods[ctx->pool_special];
            k.base = prod;
            k.unit_stride = pl.poly_stride;
            k.offset = (u64)L * pl.limb_stride;
            k.total = (u64)batch * 2 * ctx->N / 2;
            ck(launch_cta(k, grid_for(k.total, EltReduceRow::kThreads), 0, s));
        } else {
            switch_key_reduce_rows(ctx, level, prod, pl, 0, L + 1, batch, s);
        }
    }
    switch_key_special(ctx, level, prod, pl, sp.p, batch, s);
    switch_key_moddown(ctx, level, ct, ct_stride, prod, pl, sp.p, 0, L, batch, s, ct_src, ct_src_stride, ct_src_polys);
}

void switch_key_impl(Context *ctx, int level, u64 *ct, u64 ct_stride, const u64 *target, u64 target_stride, const u64 *key,
                     u32 batch, cudaStream_t s, const u64 *ct_src = nullptr, u64 ct_src_stride = 0, u32 ct_src_polys = 0) {
    if (!ctx->has_special) throw std::logic_error("keyswitching is not supported by the context");
    const u32 L = (u32)ctx->levels[level].L;
    Scratch prod((size_t)batch * 2 * (L + 1) * ctx->N, s);
    const bool unreduced = switch_key_partial(ctx, level, target, target_stride, key, 0, L, prod.p, batch, s);
    switch_key_finish(ctx, level, ct, ct_stride, prod.p, unreduced, batch, s, ct_src, ct_src_stride, ct_src_polys);
}

// Division by the last prime of a base with rounding (the level's last prime, or the special prime at the key level):
// divideAndRoundqLastNttInplace (CKKS, NTT form)  src/utils/rns_cuda.cu:348-363; divideAndRoundqLastInplace (BFV) and
// modTAndDivideqLastInplace (BGV) in coefficient form. in: [size][L][N], out: [size][L-1][N].
void drop_last(Context *c, const DropDev &drop, const LimbMap &map, u32 L, bool ntt_form, const u64 *d_in, u64 in_stride, u64 *d_out,
               u64 out_stride, u32 size, u32 batch, cudaStream_t s) {
    const u64 N = c->N;
    if (ntt_form) {
        Scratch sp((size_t)batch * size * N, s);
        NttArgs a{};
        a.mode = INV_ADD_HALF;
        a.map = map;
        a.L = L;
        a.K = size;
        a.src = d_in;
        a.src_offset = (u64)(L - 1) * N;
        a.src_unit_stride = in_stride;
        a.src_poly_stride = (u64)L * N;
        a.dst = sp.p;
        a.drop = drop;
        a.add_half = 1;
        a.scheme = c->scheme;
        run_inv(c, a, batch * size, s);
        NttArgs f{};
        f.mode = FWD_MODDOWN;
        f.map = c->map_prefix(L - 1);
        f.L = L - 1;
        f.K = size;
        f.aux0 = d_in;
        f.acc_batch_stride = in_stride;
        f.acc_poly_stride = (u64)L * N;
        f.aux1 = sp.p;
        f.ct = nullptr;
        f.dst = d_out;
        f.dst_batch_stride = out_stride;
        f.dst_poly_stride = (u64)(L - 1) * N;
        f.drop = drop;
        f.scheme = c->scheme;
        set_order(c, f, L - 1);
        run_fwd(c, f, batch * size * (L - 1), s);
    } else {
        EltDropLast k{};
        k.logN = c->logN;
        k.scheme = c->scheme;
        k.mods = c->d_mods;
        k.drop = drop;
        k.L = L;
        k.size = size;
        k.total = (u64)batch * size * (L - 1) * N;
        k.in_batch_stride = in_stride;
        k.out_batch_stride = out_stride;
        k.in = d_in;
        k.out = d_out;
        ck(launch_cta(k, grid_for(k.total, EltDropLast::kThreads), 0, s));
    }
}

void check_batch_stride(size_t stride, size_t min_words, const char *what) {
    if (stride < min_words) throw std::invalid_argument(std::string(what) + " stride is smaller than one ciphertext");
    if (stride & 1) throw std::invalid_argument(std::string(what) + " stride must be even (128-bit accesses)");
}

// The kernels use 128-bit accesses: every buffer must be 16-byte aligned (anything from cudaMalloc / troyb_malloc is).
void check_aligned(const void *p, const char *what) {
    if (reinterpret_cast<uintptr_t>(p) & 15) throw std::invalid_argument(std::string(what) + " must be 16-byte aligned");
}

} // namespace

extern "C" {

const char *troyb_last_error(void) { return g_error.c_str(); }

int troyb_initialize(int device) {
    API_BEGIN
    ck(cudaSetDevice(device));
    ck(cudaFree(0));
    t_device = device;
    g_initialized = true;
    API_END
}

int troyb_set_device(int device) {
    API_BEGIN
    need_init();
    t_device = -1;
    use_device(device);
    API_END
}

int troyb_context_device(const troyb_context *ctx) { return ctx && ctx->impl ? ctx->impl->device : -1; }

int troyb_initialized(void) { return g_initialized ? 1 : 0; }

uint64_t troyb_launch_count(void) { return launch_counter(); }

int troyb_profile_enable(int on) {
    profiler().enabled = on != 0;
    return TROYB_OK;
}

int troyb_profile_report(char *buf, size_t buf_bytes) {
    API_BEGIN
    Profiler &p = profiler();
    ck(cudaDeviceSynchronize());
    struct Agg { std::string name; unsigned long long launches = 0, grid = 0; double ms = 0; };
    std::vector<Agg> aggs;
    for (auto &r : p.records) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        Agg *a = nullptr;
        for (auto &x : aggs) if (x.name == r.name) a = &x;
        if (!a) { aggs.push_back(Agg{}); a = &aggs.back(); a->name = r.name; }
        a->launches++; a->grid += r.grid; a->ms += ms;
    }
    p.records.clear();
    std::string out;
    for (auto &a : aggs) {
        char line[256];
        snprintf(line, sizeof line, "{\"kernel\": \"%s\", \"launches\": %llu, \"grid\": %llu, \"ms\": %.6f}\n", a.name.c_str(), a.launches, a.grid, a.ms);
        out += line;
    }
    if (buf && buf_bytes) {
        size_t n = std::min(out.size(), buf_bytes - 1);
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    API_END
}

int troyb_coeff_modulus_create(uint64_t n, const int *bit_sizes, int count, uint64_t *out) {
    API_BEGIN
    if (!bit_sizes || !out || count < 0 || count > 64) throw std::invalid_argument("bit_sizes is invalid");
    if (n < 2 || n > 131072 || (n & (n - 1))) throw std::invalid_argument("poly_modulus_degree is invalid");
    std::vector<std::vector<u64>> lists(64);
    for (int i = 0; i < count; i++) {
        int b = bit_sizes[i];
        if (b < 2 || b > 60) throw std::invalid_argument("bit_sizes is invalid");
        if (lists[b].empty()) {
            size_t need = 0;
            for (int j = 0; j < count; j++) need += bit_sizes[j] == b;
            lists[b] = host::get_primes(2 * n, b, need);
        }
    }
    for (int i = 0; i < count; i++) {
        out[i] = lists[bit_sizes[i]].back();
        lists[bit_sizes[i]].pop_back();
    }
    API_END
}

int troyb_malloc(void **d_ptr, size_t bytes, void *stream) {
    API_BEGIN
    need_init();
    ck(cudaMallocAsync(d_ptr, bytes ? bytes : 8, S(stream)));
    API_END
}
int troyb_free(void *d_ptr, void *stream) {
    API_BEGIN
    if (d_ptr) ck(cudaFreeAsync(d_ptr, S(stream)));
    API_END
}
int troyb_copy_to_device(void *d_dst, const void *h_src, size_t bytes, void *stream) {
    API_BEGIN
    need_init();
    ck(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, S(stream)));
    API_END
}
int troyb_copy_to_host(void *h_dst, const void *d_src, size_t bytes, void *stream) {
    API_BEGIN
    need_init();
    ck(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, S(stream)));
    ck(cudaStreamSynchronize(S(stream)));
    API_END
}
int troyb_copy_on_device(void *d_dst, const void *d_src, size_t bytes, void *stream) {
    API_BEGIN
    need_init();
    ck(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, S(stream)));
    API_END
}
int troyb_memset_zero(void *d_ptr, size_t bytes, void *stream) {
    API_BEGIN
    need_init();
    ck(cudaMemsetAsync(d_ptr, 0, bytes, S(stream)));
    API_END
}
int troyb_stream_synchronize(void *stream) {
    API_BEGIN
    ck(cudaStreamSynchronize(S(stream)));
    API_END
}

int troyb_context_create(int scheme, uint64_t n, const uint64_t *moduli, int n_moduli, uint64_t plain_modulus,
                         int expand_mod_chain, troyb_context **out) {
    API_BEGIN
    need_init();
    if (!out || !moduli) throw std::invalid_argument("null argument");
    int device = t_device;
    if (device < 0) ck(cudaGetDevice(&device));
    Context *impl = context_create(device, scheme, n, moduli, n_moduli, plain_modulus, expand_mod_chain != 0);
    t_device = device;
    *out = new troyb_context{impl};
    API_END
}

void troyb_context_destroy(troyb_context *ctx) {
    if (!ctx) return;
    delete ctx->impl;
    delete ctx;
}

int troyb_context_scheme(const troyb_context *ctx) { return ctx && ctx->impl ? ctx->impl->scheme : 0; }
uint64_t troyb_context_poly_degree(const troyb_context *ctx) { return ctx && ctx->impl ? ctx->impl->N : 0; }
int troyb_context_num_levels(const troyb_context *ctx) { return ctx && ctx->impl ? (int)ctx->impl->levels.size() : 0; }
int troyb_context_level_limbs(const troyb_context *ctx, int level) {
    try {
        return level_limbs(C(ctx), level);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}
int troyb_context_level_moduli(const troyb_context *ctx, int level, uint64_t *out) {
    API_BEGIN
    const Context *c = C(ctx);
    int L = level_limbs(c, level);
    if (level == -1) {
        for (int i = 0; i < c->Ltop; i++) out[i] = c->pool[i];
        if (c->has_special) out[c->Ltop] = c->pool[c->pool_special];
    } else {
        for (int i = 0; i < L; i++) out[i] = c->pool[i];
    }
    API_END
}
int troyb_context_using_keyswitching(const troyb_context *ctx) { return ctx && ctx->impl && ctx->impl->has_special ? 1 : 0; }
int troyb_context_total_coeff_bits(const troyb_context *ctx, int level) {
    try {
        const Context *c = C(ctx);
        level_limbs(c, level);
        // the key level (-1) is not an entry of levels[]: its modulus is every prime of the pool's chain
        return c->level_total_bits(level);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}
uint64_t troyb_context_plain_modulus(const troyb_context *ctx) { return ctx && ctx->impl ? ctx->impl->plain : 0; }

int troyb_context_bsk(const troyb_context *ctx, int level, uint64_t *out) {
    try {
        const Context *c = C(ctx);
        level_limbs(c, level);
        if (level < 0) return 0; // the key level has no BEHZ base
        const LevelData &lv = c->levels[level];
        if (!lv.has_behz) return 0;
        if (out) {
            for (int b = 0; b < lv.behz.nB; b++) out[b] = c->pool[c->pool_B0 + b];
            out[lv.behz.nB] = c->pool[c->pool_msk];
        }
        return lv.behz.nBsk;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

int troyb_context_root_powers(const troyb_context *ctx, int key_limb, uint64_t *h_out) {
    API_BEGIN
    const Context *c = C(ctx);
    int n = c->Ltop + (c->has_special ? 1 : 0);
    if (key_limb < 0 || key_limb >= n) throw std::out_of_range("key_limb");
    int pool = key_limb < c->Ltop ? key_limb : c->pool_special;
    ck(cudaMemcpy(h_out, c->h_mods[pool].fwd, 2 * c->N * sizeof(u64), cudaMemcpyDeviceToHost));
    API_END
}

int troyb_ntt_forward(troyb_context *ctx, int level, uint64_t *d_data, size_t n_polys, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    u32 L = (u32)level_limbs(c, level);
    check_aligned(d_data, "data");
    ntt_plain(c, false, c->map_prefix(L), L, d_data, (u64)L * c->N, d_data, (u64)L * c->N, L, (u32)n_polys, true, S(stream));
    API_END
}
int troyb_ntt_inverse(troyb_context *ctx, int level, uint64_t *d_data, size_t n_polys, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    u32 L = (u32)level_limbs(c, level);
    check_aligned(d_data, "data");
    ntt_plain(c, true, c->map_prefix(L), L, d_data, (u64)L * c->N, d_data, (u64)L * c->N, L, (u32)n_polys, true, S(stream));
    API_END
}
int troyb_ntt_forward_key(troyb_context *ctx, uint64_t *d_data, size_t n_polys, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    u32 L = (u32)level_limbs(c, -1);
    LimbMap m = c->has_special ? c->map_prefix_special(c->Ltop) : c->map_prefix(c->Ltop);
    check_aligned(d_data, "data");
    ntt_plain(c, false, m, L, d_data, (u64)L * c->N, d_data, (u64)L * c->N, L, (u32)n_polys, true, S(stream));
    API_END
}
int troyb_ntt_inverse_key(troyb_context *ctx, uint64_t *d_data, size_t n_polys, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    u32 L = (u32)level_limbs(c, -1);
    LimbMap m = c->has_special ? c->map_prefix_special(c->Ltop) : c->map_prefix(c->Ltop);
    check_aligned(d_data, "data");
    ntt_plain(c, true, m, L, d_data, (u64)L * c->N, d_data, (u64)L * c->N, L, (u32)n_polys, true, S(stream));
    API_END
}

static int addsub(troyb_context *ctx, int level, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n_polys,
                  void *stream, uint64_t scalar = 0) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    u32 L = (u32)level_limbs(c, level);
    EltAddSub k{};
    k.op = op;
    k.scalar = scalar;
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.total = (u64)n_polys * L * c->N;
    k.a = a;
    k.b = b;
    k.out = out;
    ck(launch_cta(k, grid_for(k.total, EltAddSub::kThreads), 0, S(stream)));
    API_END
}
int troyb_add(troyb_context *ctx, int level, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, void *st) {
    return addsub(ctx, level, 0, a, b, out, n, st);
}
int troyb_sub(troyb_context *ctx, int level, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, void *st) {
    return addsub(ctx, level, 1, a, b, out, n, st);
}
int troyb_negate(troyb_context *ctx, int level, const uint64_t *a, uint64_t *out, size_t n, void *st) {
    return addsub(ctx, level, 2, a, a, out, n, st);
}
int troyb_multiply_scalar(troyb_context *ctx, int level, const uint64_t *a, uint64_t scalar, uint64_t *out, size_t n, void *st) {
    return addsub(ctx, level, 3, a, a, out, n, st, scalar);
}

int troyb_multiply_inv_degree(troyb_context *ctx, int level, const uint64_t *a, uint64_t *out, size_t n, void *st) {
    return addsub(ctx, level, 5, a, nullptr, out, n, st);
}
int troyb_negacyclic_shift(troyb_context *ctx, int level, const uint64_t *d_in, uint64_t *d_out, size_t n_polys, uint64_t shift, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    if (shift >= 2 * c->N) throw std::invalid_argument("shift must be below 2 N");
    if (d_in == d_out) throw std::invalid_argument("negacyclic shift is out of place");
    EltNegacyclicShift k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.shift = shift;
    k.total = (u64)n_polys * L * c->N;
    k.in = d_in;
    k.out = d_out;
    ck(launch_cta(k, grid_for(k.total, EltNegacyclicShift::kThreads), 0, S(stream)));
    API_END
}

int troyb_multiply(troyb_context *ctx, int level, const uint64_t *d_a, size_t a_size, size_t a_stride, const uint64_t *d_b,
                   size_t b_size, size_t b_stride, uint64_t *d_out, size_t out_stride, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    const u64 N = c->N;
    cudaStream_t s = S(stream);
    if (a_size < 1 || b_size < 1 || a_size > 8 || b_size > 8) throw std::invalid_argument("invalid ciphertext size");
    const size_t out_size = a_size + b_size - 1;
    check_batch_stride(a_stride, a_size * L * N, "a");
    check_batch_stride(b_stride, b_size * L * N, "b");
    check_batch_stride(out_stride, out_size * L * N, "out");
    check_aligned(d_a, "a");
    check_aligned(d_b, "b");
    check_aligned(d_out, "out");

    auto tensor = [&](const LimbMap &map, u32 limbs, const u64 *a, u64 as, const u64 *b, u64 bs, u64 *o, u64 os) {
        EltTensor k{};
        k.logN = c->logN;
        k.mods = c->d_mods;
        k.map = map;
        k.n_limbs = limbs;
        k.a_size = (u32)a_size;
        k.b_size = (u32)b_size;
        k.per_batch = (u64)limbs * N;
        k.total = (u64)batch * limbs * N / 2;
        k.a = a; k.b = b; k.out = o;
        k.a_batch_stride = as; k.b_batch_stride = bs; k.out_batch_stride = os;
        ck(launch_cta(k, grid_for(k.total, EltTensor::kThreads), 0, s));
    };

    if (c->scheme == SCHEME_CKKS) {
        // ckksMultiply  src/evaluator_cuda.cu:384-433
        tensor(c->map_prefix(L), L, d_a, a_stride, d_b, b_stride, d_out, out_stride);
    } else if (c->scheme == SCHEME_BGV) {
        // bgvMultiply  src/evaluator_cuda.cu:435-523: NTT both operands, tensor, INTT
        const bool square = (d_a == d_b && a_size == b_size && a_stride == b_stride);
        LimbMap map = c->map_prefix(L);
        Scratch ta((size_t)batch * a_size * L * N, s);
        Scratch tb(square ? 0 : (size_t)batch * b_size * L * N, s);
        ntt_plain(c, false, map, L, d_a, a_stride, ta.p, (u64)a_size * L * N, (u32)a_size * L, (u32)batch, true, s);
        if (!square) ntt_plain(c, false, map, L, d_b, b_stride, tb.p, (u64)b_size * L * N, (u32)b_size * L, (u32)batch, true, s);
        tensor(map, L, ta.p, (u64)a_size * L * N, square ? ta.p : tb.p, (u64)b_size * L * N, d_out, out_stride);
        ntt_plain(c, true, map, L, d_out, out_stride, d_out, out_stride, (u32)out_size * L, (u32)batch, true, s);
    } else {
        // bfvMultiply (BEHZ)  src/evaluator_cuda.cu:283-382
        const LevelData &lv = c->levels[level];
        if (!lv.has_behz) throw std::logic_error("BEHZ tables missing");
        const u32 E = L + (u32)lv.behz.nBsk;
        if (E > (u32)kBehzMax) throw std::invalid_argument("too many limbs for BFV multiply");
        LimbMap map = c->map_q_bsk(level);
        const bool square = (d_a == d_b && a_size == b_size && a_stride == b_stride);
        Scratch ea((size_t)batch * a_size * E * N, s);
        Scratch eb(square ? 0 : (size_t)batch * b_size * E * N, s);
        Scratch prod((size_t)batch * out_size * E * N, s);
        auto extend = [&](const u64 *in, size_t size, u64 stride, u64 *out) {
            BehzExtend k{};
            k.logN = c->logN;
            k.mods = c->d_mods;
            k.map = map;
            k.z = lv.behz;
            k.total = (u64)batch * size * N;
            k.polys_per_batch = (u32)size;
            k.in_batch_stride = stride;
            k.in = in;
            k.out = out;
            ck(launch_cta(k, grid_for(k.total, BehzExtend::kThreads), 0, s));
            ntt_plain(c, false, map, E, out, (u64)size * E * N, out, (u64)size * E * N, (u32)size * E, (u32)batch, true, s);
        };
        extend(d_a, a_size, a_stride, ea.p);
        if (!square) extend(d_b, b_size, b_stride, eb.p);
        tensor(map, E, ea.p, (u64)a_size * E * N, square ? ea.p : eb.p, (u64)b_size * E * N, prod.p, (u64)out_size * E * N);
        ntt_plain(c, true, map, E, prod.p, (u64)out_size * E * N, prod.p, (u64)out_size * E * N, (u32)out_size * E, (u32)batch,
                  true, s);
        BehzFloorSk k{};
        k.logN = c->logN;
        k.mods = c->d_mods;
        k.map = map;
        k.z = lv.behz;
        k.total = (u64)batch * out_size * N;
        k.in = prod.p;
        k.out = d_out;
        k.polys_per_batch = (u32)out_size;
        k.out_batch_stride = out_stride;
        ck(launch_cta(k, grid_for(k.total, BehzFloorSk::kThreads), 0, s));
    }
    API_END
}

// ---- ciphertext (x) plaintext -------------------------------------------------------------------------------------------
namespace {
void mul_plain(Context *c, u32 L, const u64 *ct, size_t ct_size, u64 ct_stride, const u64 *plain, u64 plain_stride, u64 *out,
               u64 out_stride, size_t batch, int accumulate, cudaStream_t s) {
    EltMulPlain k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.per_ct = (u64)ct_size * L * c->N;
    k.total = (u64)batch * k.per_ct / 2;
    k.ct = ct; k.plain = plain; k.out = out;
    k.ct_stride = ct_stride; k.plain_stride = plain_stride; k.out_stride = out_stride;
    k.accumulate = accumulate;
    ck(launch_cta(k, grid_for(k.total, EltMulPlain::kThreads), 0, s));
}

void plain_lift(Context *c, int level, u32 L, const u64 *plain, size_t coeff_count, u64 plain_stride, u64 *out, size_t batch,
                bool to_ntt, cudaStream_t s) {
    const LevelData &lv = c->levels[level];
    if (!lv.has_plain) throw std::invalid_argument("unsupported scheme");
    if (coeff_count > c->N) throw std::invalid_argument("plain is not valid for encryption parameters");
    EltPlainLift k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.coeff_count = coeff_count;
    k.total = (u64)batch * L * c->N;
    k.plain = plain; k.out = out;
    k.plain_stride = plain_stride;
    k.pc = lv.plain;
    ck(launch_cta(k, grid_for(k.total, EltPlainLift::kThreads), 0, s));
    if (to_ntt) ntt_plain(c, false, k.map, L, out, (u64)L * c->N, out, (u64)L * c->N, L, (u32)batch, true, s);
}
} // namespace

int troyb_multiply_plain_ntt(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                             const uint64_t *d_plain, size_t plain_stride, uint64_t *d_out, size_t out_stride, size_t batch,
                             int accumulate, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (ct_size < 1) throw std::invalid_argument("invalid ciphertext size");
    check_batch_stride(ct_stride, ct_size * L * c->N, "ct");
    check_batch_stride(out_stride, ct_size * L * c->N, "out");
    if (plain_stride != 0) check_batch_stride(plain_stride, (size_t)L * c->N, "plain");
    check_aligned(d_ct, "ct");
    check_aligned(d_plain, "plain");
    check_aligned(d_out, "out");
    mul_plain(c, L, d_ct, ct_size, ct_stride, d_plain, plain_stride, d_out, out_stride, batch, accumulate, S(stream));
    API_END
}

int troyb_matmul_plain_ntt(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                           const uint64_t *d_plain, uint64_t *d_out, size_t out_stride, size_t batch, size_t inputs, size_t outputs,
                           void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (ct_size < 1) throw std::invalid_argument("invalid ciphertext size");
    if (inputs < 1) throw std::invalid_argument("inner dimension cannot be zero");
    check_batch_stride(ct_stride, ct_size * L * c->N, "ct");
    check_batch_stride(out_stride, ct_size * L * c->N, "out");
    check_aligned(d_ct, "ct");
    check_aligned(d_plain, "plain");
    check_aligned(d_out, "out");
    if (batch == 0 || outputs == 0) return 0;
    EltMatmulPlain k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.ct_size = (u32)ct_size;
    k.B = (u32)batch; k.I = (u32)inputs; k.J = (u32)outputs;
    k.total = (u64)((batch + 1) / 2) * ((outputs + 1) / 2) * ct_size * L * (c->N / 2);
    k.a = d_ct; k.w = d_plain; k.out = d_out;
    k.a_ct_stride = ct_stride; k.out_ct_stride = out_stride;
    ck(launch_cta(k, grid_for(k.total, EltMatmulPlain::kThreads), 0, S(stream)));
    API_END
}

int troyb_plain_lift(troyb_context *ctx, int level, const uint64_t *d_plain, size_t plain_coeff_count, size_t plain_stride,
                     uint64_t *d_out, size_t batch, int to_ntt, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    check_aligned(d_out, "out");
    plain_lift(c, level, L, d_plain, plain_coeff_count, plain_stride, d_out, batch, to_ntt != 0, S(stream));
    API_END
}

int troyb_multiply_plain_normal(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_size, size_t ct_stride,
                                const uint64_t *d_plain, size_t plain_coeff_count, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    if (ct_size < 1) throw std::invalid_argument("invalid ciphertext size");
    check_batch_stride(ct_stride, ct_size * L * c->N, "ct");
    check_aligned(d_ct, "ct");
    cudaStream_t s = S(stream);
    // lift + NTT of the plaintext once, then NTT -> dyadic product -> INTT over all polys of the batch (3 launches)
    Scratch temp((size_t)L * c->N, s);
    plain_lift(c, level, L, d_plain, plain_coeff_count, 0, temp.p, 1, true, s);
    const LimbMap map = c->map_prefix(L);
    ntt_plain(c, false, map, L, d_ct, ct_stride, d_ct, ct_stride, (u32)ct_size * L, (u32)batch, true, s);
    mul_plain(c, L, d_ct, ct_size, ct_stride, temp.p, 0, d_ct, ct_stride, batch, 0, s);
    ntt_plain(c, true, map, L, d_ct, ct_stride, d_ct, ct_stride, (u32)ct_size * L, (u32)batch, true, s);
    API_END
}

int troyb_add_plain(troyb_context *ctx, int level, uint64_t *d_c0, size_t ct_stride, const uint64_t *d_plain,
                    size_t plain_coeff_count, size_t plain_stride, uint64_t correction_factor, int subtract, size_t batch,
                    void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    const LevelData &lv = c->levels[level];
    if (!lv.has_plain) throw std::invalid_argument("unsupported scheme");
    if (plain_coeff_count > c->N) throw std::invalid_argument("plain is not valid for encryption parameters");
    check_batch_stride(ct_stride, (size_t)L * c->N, "ct");
    if (plain_coeff_count == 0 || batch == 0) return 0;
    EltAddPlain k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.coeff_count = plain_coeff_count;
    k.total = (u64)batch * plain_coeff_count;
    k.plain = d_plain; k.c0 = d_c0;
    k.plain_stride = plain_stride; k.ct_stride = ct_stride;
    k.scaled = c->scheme == SCHEME_BFV ? 1 : 0;
    k.subtract = subtract;
    k.factor = correction_factor;
    k.pc = lv.plain;
    ck(launch_cta(k, grid_for(k.total, EltAddPlain::kThreads), 0, S(stream)));
    API_END
}

// ---- decryption ------------------------------------------------------------------------------------------------------------
int troyb_secret_key_powers(troyb_context *ctx, const uint64_t *d_secret, size_t max_power, uint64_t *d_out, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 KL = (u32)level_limbs(c, -1);
    if (max_power < 1) throw std::invalid_argument("max_power cannot be zero");
    check_aligned(d_secret, "secret");
    check_aligned(d_out, "out");
    cudaStream_t s = S(stream);
    const u64 words = (u64)KL * c->N;
    if (d_out != d_secret) ck(cudaMemcpyAsync(d_out, d_secret, words * 8, cudaMemcpyDeviceToDevice, s));
    for (size_t i = 1; i < max_power; i++) {
        // computeSecretKeyArray  src/decryptor_cuda.cu:188-235: s^(i+1) = s^i * s, limb-wise over the key-level base
        EltMulPlain k{};
        k.logN = c->logN;
        k.mods = c->d_mods;
        k.map = c->has_special ? c->map_prefix_special(c->Ltop) : c->map_prefix(c->Ltop);
        k.n_limbs = KL;
        k.per_ct = words;
        k.total = words / 2;
        k.ct = d_out + (i - 1) * words; k.plain = d_out; k.out = d_out + i * words;
        ck(launch_cta(k, grid_for(k.total, EltMulPlain::kThreads), 0, s));
    }
    API_END
}

int troyb_decrypt_dot_product(troyb_context *ctx, int level, const uint64_t *d_ct, size_t ct_size, size_t ct_stride, int is_ntt_form,
                              const uint64_t *d_secret_powers, uint64_t *d_phase, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    if (ct_size < 2 || ct_size > 8) throw std::invalid_argument("encrypted is empty");
    check_batch_stride(ct_stride, ct_size * L * c->N, "ct");
    check_aligned(d_ct, "ct");
    check_aligned(d_secret_powers, "secret");
    check_aligned(d_phase, "phase");
    if (batch == 0) return 0;
    cudaStream_t s = S(stream);
    const LimbMap map = c->map_prefix(L);
    const u64 per_ct = (u64)ct_size * L * c->N;
    // coefficient form (BFV / BGV): the transform is linear, so NTT every poly (c_0 included), take the dot product in the
    // NTT domain and transform back: the canonical phase c_0 + INTT(sum NTT(c_i) s^i) of src/decryptor_cuda.cu:262-268
    Scratch tmp(is_ntt_form ? 0 : (size_t)batch * per_ct, s);
    const u64 *src = d_ct;
    u64 src_stride = ct_stride;
    if (!is_ntt_form) {
        ntt_plain(c, false, map, L, d_ct, ct_stride, tmp.p, per_ct, (u32)ct_size * L, (u32)batch, true, s);
        src = tmp.p;
        src_stride = per_ct;
    }
    EltDecryptDot k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = map;
    k.n_limbs = L;
    k.ct_size = (u32)ct_size;
    k.key_limbs = (u32)level_limbs(c, -1);
    k.total = (u64)batch * L * c->N / 2;
    k.ct = src; k.sk = d_secret_powers; k.out = d_phase;
    k.ct_stride = src_stride;
    ck(launch_cta(k, grid_for(k.total, EltDecryptDot::kThreads), 0, s));
    if (!is_ntt_form) ntt_plain(c, true, map, L, d_phase, (u64)L * c->N, d_phase, (u64)L * c->N, L, (u32)batch, true, s);
    API_END
}

int troyb_decrypt_scale_and_round(troyb_context *ctx, int level, const uint64_t *d_phase, uint64_t *d_plain, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    const LevelData &lv = c->levels[level];
    if (!lv.has_decrypt || c->scheme != SCHEME_BFV) throw std::invalid_argument("unsupported scheme");
    if (batch == 0) return 0;
    EltDecryptScaleRound k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.gamma_pool = (u32)c->pool_gamma;
    k.total = (u64)batch * c->N;
    k.phase = d_phase; k.out = d_plain;
    k.dc = lv.decrypt;
    ck(launch_cta(k, grid_for(k.total, EltDecryptScaleRound::kThreads), 0, S(stream)));
    API_END
}

int troyb_decrypt_mod_t(troyb_context *ctx, int level, const uint64_t *d_phase, uint64_t *d_plain, uint64_t inv_correction_factor,
                        size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level is not valid for encryption parameters");
    const LevelData &lv = c->levels[level];
    if (!lv.has_decrypt) throw std::invalid_argument("unsupported scheme");
    if (batch == 0) return 0;
    EltDecryptModT k{};
    k.logN = c->logN;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.total = (u64)batch * c->N;
    k.phase = d_phase; k.out = d_plain;
    k.inv_factor = inv_correction_factor;
    k.dc = lv.decrypt;
    ck(launch_cta(k, grid_for(k.total, EltDecryptModT::kThreads), 0, S(stream)));
    API_END
}

int troyb_switch_key(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_target,
                     size_t target_stride, const uint64_t *d_key, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_ct || !d_target || !d_key) throw std::invalid_argument("target_iter");
    check_batch_stride(ct_stride, 2ull * L * c->N, "ct");
    check_batch_stride(target_stride, (u64)L * c->N, "target");
    check_aligned(d_ct, "ct");
    check_aligned(d_target, "target");
    check_aligned(d_key, "key");
    switch_key_impl(c, level, d_ct, ct_stride, d_target, target_stride, d_key, (u32)batch, S(stream));
    API_END
}

int troyb_switch_key_partial(troyb_context *ctx, int level, const uint64_t *d_target, size_t target_stride, const uint64_t *d_key,
                             int digit_begin, int digit_end, uint64_t *d_prod, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_target || !d_key || !d_prod) throw std::invalid_argument("null argument");
    if (digit_begin < 0 || digit_end <= digit_begin || (u32)digit_end > L) throw std::invalid_argument("digit range is not valid");
    check_batch_stride(target_stride, (u64)L * c->N, "target");
    check_aligned(d_target, "target");
    check_aligned(d_key, "key");
    check_aligned(d_prod, "prod");
    if (switch_key_partial(c, level, d_target, target_stride, d_key, (u32)digit_begin, (u32)(digit_end - digit_begin), d_prod, (u32)batch,
                           S(stream))) // the contract of this entry point is canonical residues
        switch_key_reduce_rows(c, level, d_prod, ProdLayout::standard(L + 1, c->N), 0, L + 1, (u32)batch, S(stream));
    API_END
}

int troyb_switch_key_finish(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, uint64_t *d_prod, int reduce_first,
                            size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_ct || !d_prod) throw std::invalid_argument("null argument");
    check_batch_stride(ct_stride, 2ull * L * c->N, "ct");
    check_aligned(d_ct, "ct");
    check_aligned(d_prod, "prod");
    switch_key_finish(c, level, d_ct, ct_stride, d_prod, reduce_first != 0, (u32)batch, S(stream));
    API_END
}

int troyb_switch_key_partial_rows(troyb_context *ctx, int level, const uint64_t *d_target, size_t target_stride,
                                  const uint64_t *d_key, int digit_begin, int digit_end, uint64_t *d_rows, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_target || !d_key || !d_rows) throw std::invalid_argument("null argument");
    if (digit_begin < 0 || digit_end <= digit_begin || (u32)digit_end > L) throw std::invalid_argument("digit range is not valid");
    check_batch_stride(target_stride, (u64)L * c->N, "target");
    check_aligned(d_target, "target");
    check_aligned(d_key, "key");
    check_aligned(d_rows, "rows");
    const ProdLayout pl = ProdLayout::rows((u32)batch, c->N);
    switch_key_partial(c, level, d_target, target_stride, d_key, (u32)digit_begin, (u32)(digit_end - digit_begin), d_rows, (u32)batch,
                       S(stream), &pl);
    API_END
}

int troyb_switch_key_reduce_rows(troyb_context *ctx, int level, uint64_t *d_rows, int row_begin, int row_end, size_t batch,
                                 void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_rows) throw std::invalid_argument("null argument");
    if (row_begin < 0 || row_end < row_begin || (u32)row_end > L + 1) throw std::invalid_argument("row range is not valid");
    if (row_end > row_begin)
        switch_key_reduce_rows(c, level, d_rows, ProdLayout::rows((u32)batch, c->N), (u32)row_begin, (u32)row_end, (u32)batch, S(stream));
    API_END
}

int troyb_switch_key_special_row(troyb_context *ctx, int level, const uint64_t *d_rows, uint64_t *d_sp, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_rows || !d_sp) throw std::invalid_argument("null argument");
    if (!c->has_special) throw std::logic_error("keyswitching is not supported by the context");
    switch_key_special(c, level, d_rows, ProdLayout::rows((u32)batch, c->N), d_sp, (u32)batch, S(stream));
    API_END
}

int troyb_switch_key_moddown_rows(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_rows,
                                  const uint64_t *d_sp, int limb_begin, int limb_end, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_ct || !d_rows || !d_sp) throw std::invalid_argument("null argument");
    if (limb_begin < 0 || limb_end < limb_begin || (u32)limb_end > L + 1) throw std::invalid_argument("limb range is not valid");
    check_batch_stride(ct_stride, 2ull * L * c->N, "ct");
    check_aligned(d_ct, "ct");
    if (!c->has_special) throw std::logic_error("keyswitching is not supported by the context");
    switch_key_moddown(c, level, d_ct, ct_stride, d_rows, ProdLayout::rows((u32)batch, c->N), d_sp, (u32)limb_begin, (u32)limb_end,
                       (u32)batch, S(stream));
    API_END
}

int troyb_relinearize(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, const uint64_t *d_relin_key,
                      size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (level < 0) throw std::invalid_argument("level");
    if (!d_ct || !d_relin_key) throw std::invalid_argument("null argument");
    check_batch_stride(ct_stride, 3ull * L * c->N, "ct");
    check_aligned(d_ct, "ct");
    check_aligned(d_relin_key, "key");
    switch_key_impl(c, level, d_ct, ct_stride, d_ct + 2ull * L * c->N, ct_stride, d_relin_key, (u32)batch, S(stream));
    API_END
}

int troyb_galois_permute(troyb_context *ctx, int level, const uint64_t *d_in, uint64_t *d_out, size_t n_polys,
                         uint32_t galois_elt, int ntt_form, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    if (!(galois_elt & 1) || galois_elt >= 2 * c->N) throw std::invalid_argument("Galois element is not valid");
    if (d_in == d_out) throw std::invalid_argument("operand cannot point to the same value as result");
    EltGalois k{};
    k.logN = c->logN;
    k.ntt_form = ntt_form;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = L;
    k.L = L;
    k.elt = galois_elt;
    k.total = (u64)n_polys * L * c->N;
    k.in = d_in;
    k.out = d_out;
    k.in_batch_stride = (u64)L * c->N;
    k.out_batch_stride = (u64)L * c->N;
    ck(launch_cta(k, grid_for(k.total, EltGalois::kThreads), 0, S(stream)));
    API_END
}

int troyb_galois_elt_from_step(const troyb_context *ctx, int step, uint32_t *out) {
    API_BEGIN
    // GaloisTool::getEltFromStep  src/utils/galois.cpp:58-92 (generator 3, row size N/2)
    const Context *c = C(ctx);
    const u32 n = (u32)c->N, m = 2 * n;
    if (step == 0) {
        *out = m - 1;
    } else {
        bool sign = step < 0;
        u32 pos = (u32)(sign ? -(int64_t)step : step);
        if (pos >= (n >> 1)) throw std::invalid_argument("step count too large");
        pos = sign ? (n >> 1) - pos : pos;
        u64 g = 1;
        for (u32 i = 0; i < pos; i++) g = (g * 3) & (m - 1);
        *out = (u32)g;
    }
    API_END
}

int troyb_apply_galois(troyb_context *ctx, int level, uint64_t *d_ct, size_t ct_stride, uint32_t galois_elt,
                       const uint64_t *d_key, size_t batch, void *stream) {
    return troyb_apply_galois_to(ctx, level, d_ct, ct_stride, d_ct, ct_stride, galois_elt, d_key, batch, stream);
}

int troyb_apply_galois_to(troyb_context *ctx, int level, const uint64_t *d_in, size_t in_stride, uint64_t *d_ct, size_t ct_stride,
                          uint32_t galois_elt, const uint64_t *d_key, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    const u64 N = c->N;
    cudaStream_t s = S(stream);
    if (level < 0) throw std::invalid_argument("level");
    if (!(galois_elt & 1) || galois_elt >= 2 * N) throw std::invalid_argument("Galois element is not valid");
    if (!d_ct || !d_in || !d_key) throw std::invalid_argument("Galois key not present");
    check_batch_stride(ct_stride, 2ull * L * N, "ct");
    check_batch_stride(in_stride, 2ull * L * N, "in");
    check_aligned(d_ct, "ct");
    check_aligned(d_in, "in");
    check_aligned(d_key, "key");
    Scratch tmp((size_t)batch * 2 * L * N, s);
    EltGalois k{};
    k.logN = c->logN;
    k.ntt_form = c->scheme == SCHEME_CKKS;
    k.mods = c->d_mods;
    k.map = c->map_prefix(L);
    k.n_limbs = 2 * L;
    k.L = L;
    k.elt = galois_elt;
    k.total = (u64)batch * 2 * L * N;
    k.in = d_in;
    k.out = tmp.p;
    k.in_batch_stride = in_stride;
    k.out_batch_stride = 2ull * L * N;
    ck(launch_cta(k, grid_for(k.total, EltGalois::kThreads), 0, s));
    // c0 <- permuted c0 ; c1 <- 0 ; then key-switch the permuted c1 into (c0, c1)   src/evaluator_cuda.cu:2104-2116.
    // The copy and the memset are folded into the key switch's tail: it writes ct = (permuted c0, 0) + result.
    switch_key_impl(c, level, d_ct, ct_stride, tmp.p + (u64)L * N, 2ull * L * N, d_key, (u32)batch, s, tmp.p, 2ull * L * N, 1);
    API_END
}

int troyb_mod_switch_scale_to_next(troyb_context *ctx, int level, const uint64_t *d_in, size_t in_stride, uint64_t *d_out,
                                   size_t out_stride, size_t size, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const u32 L = (u32)level_limbs(c, level);
    const u64 N = c->N;
    if (level < 0 || level + 1 >= (int)c->levels.size()) throw std::invalid_argument("end of modulus switching chain reached");
    if (d_in == d_out) throw std::invalid_argument("in-place call must go through a temporary (as the reference does)");
    check_batch_stride(in_stride, size * L * N, "in");
    check_batch_stride(out_stride, size * (L - 1) * N, "out");
    check_aligned(d_in, "in");
    check_aligned(d_out, "out");
    drop_last(c, c->levels[level].drop, c->map_prefix(L), L, c->scheme == SCHEME_CKKS, d_in, in_stride, d_out, out_stride, (u32)size,
              (u32)batch, S(stream));
    API_END
}

// ---- encrypt-zero with the host generator's semantics (kernels_rlwe.cuh) ------------------------------
namespace {

struct LevelBase {
    u32 L;
    LimbMap map;
};
LevelBase level_base(const Context *c, int level) {
    LevelBase b;
    b.L = (u32)level_limbs(c, level);
    b.map = level == -1 && c->has_special ? c->map_prefix_special(c->Ltop) : c->map_prefix((int)b.L);
    return b;
}

// Expands `bytes` (rounded up to whole 4096-byte buffers) of `batch` generators; returns the words per generator.
u64 prng_expand(const u64 *d_seeds, u64 seed_stride, u64 *d_out, u64 out_stride, u64 bytes, u32 batch, cudaStream_t s) {
    PrngExpand k{};
    k.seeds = d_seeds;
    k.seed_stride = seed_stride;
    k.out = d_out;
    k.out_stride = out_stride;
    k.buffers = (u32)((bytes + Blake2::kBufferBytes - 1) / Blake2::kBufferBytes);
    k.ctas_per_stream = (k.buffers + PrngExpand::kBuffersPerCta - 1) / PrngExpand::kBuffersPerCta;
    ck(launch_cta(k, batch * k.ctas_per_stream, PrngExpand::kBuffersPerCta * 8 * sizeof(u64), s));
    return (u64)k.buffers * (Blake2::kBufferBytes / 8);
}
u64 buffers_words(u64 bytes) { return (bytes + Blake2::kBufferBytes - 1) / Blake2::kBufferBytes * (Blake2::kBufferBytes / 8); }

PrngView make_view(const u64 *d_stream, u64 bytes, const u64 *d_seeds) {
    PrngView v;
    v.buf = d_stream;
    v.bytes = bytes & ~63ull;
    v.seed = d_seeds;
    return v;
}

// samplePolyTernary (n_tern = 1) followed by n_cbd x samplePolyCbd from one stream; out [batch][n_tern + n_cbd][L][N].
// d_end (optional, [batch]): stream position in bytes after the last draw.
void sample_small(Context *c, const LevelBase &lb, const PrngView &view, u64 view_stride, u64 seed_stride, u64 start_bytes,
                  u32 n_tern, u32 n_cbd, u64 *d_out, u64 out_stride, u64 *d_end, u32 batch, cudaStream_t s) {
    const u32 N = (u32)c->N, nseg = (N + kSampleSeg - 1) / kSampleSeg;
    if (n_tern && (start_bytes & 3)) throw std::invalid_argument("stream position must be a multiple of 4");
    if (start_bytes & 1) throw std::invalid_argument("stream position must be even");
    Scratch seg(n_tern ? (size_t)batch * (nseg + 2) : 0, s);
    if (n_tern) {
        SampleTernaryScan k{};
        k.view = view;
        k.view_stride = view_stride;
        k.seed_stride = seed_stride;
        k.start_bytes = start_bytes;
        k.N = N;
        k.nseg = nseg;
        k.seg = seg.p;
        ck(launch_cta(k, batch, 0, s));
    }
    SampleExpand e{};
    e.view = view;
    e.view_stride = view_stride;
    e.seed_stride = seed_stride;
    e.start_bytes = start_bytes & ~3ull; // ternary draws (multiple of 4 when n_tern)
    e.cbd_bytes = start_bytes;           // noise draws when there is no ternary polynomial before them
    e.N = N;
    e.nseg = nseg;
    e.n_tern = n_tern;
    e.n_cbd = n_cbd;
    e.n_limbs = lb.L;
    e.seg = n_tern ? seg.p : nullptr;
    e.mods = c->d_mods;
    e.map = lb.map;
    e.out = d_out;
    e.out_stride = out_stride;
    ck(launch_cta(e, batch * nseg, 0, s));
    if (d_end) {
        if (n_tern) {
            // end of the ternary draws (+ the rejection-free noise draws)
            ck(cudaMemcpy2DAsync(d_end, sizeof(u64), seg.p + nseg + 1, (nseg + 2) * sizeof(u64), sizeof(u64), batch,
                                 cudaMemcpyDeviceToDevice, s));
            if (n_cbd) throw std::logic_error("end position is reported for single samplers only");
        } else {
            std::vector<u64> end(batch, start_bytes + 6ull * n_cbd * N);
            ck(cudaMemcpyAsync(d_end, end.data(), batch * sizeof(u64), cudaMemcpyHostToDevice, s));
            ck(cudaStreamSynchronize(s)); // `end` is a host temporary
        }
    }
}

// samplePolyUniform over the level's base from the start of each stream; out [batch][L][N] (out_stride words apart).
void sample_uniform(Context *c, const LevelBase &lb, const PrngView &view, u64 view_stride, u64 seed_stride, u64 *d_out,
                    u64 out_stride, u32 batch, cudaStream_t s) {
    const u64 N = c->N;
    if (view.bytes < (u64)lb.L * N * 8) throw std::invalid_argument("stream shorter than the polynomial");
    SampleUniformFill f{};
    f.view = view;
    f.view_stride = view_stride;
    f.logN = c->logN;
    f.n_limbs = lb.L;
    f.mods = c->d_mods;
    f.map = lb.map;
    f.out = d_out;
    f.out_stride = out_stride;
    f.total = (u64)batch * lb.L * N;
    ck(launch_cta(f, grid_for(f.total, SampleUniformFill::kThreads), 0, s));
    Scratch repl((size_t)batch * N, s);
    SampleUniformResolve r{};
    r.view = view;
    r.view_stride = view_stride;
    r.seed_stride = seed_stride;
    r.N = (u32)N;
    r.n_limbs = lb.L;
    r.mods = c->d_mods;
    r.map = lb.map;
    r.out = d_out;
    r.out_stride = out_stride;
    r.repl = repl.p;
    ck(launch_cta(r, batch, SampleUniformResolve::kSmemWords * sizeof(u64), s));
}

// c1 of a seeded ciphertext from its 64-bit seed (NTT domain for every scheme; see SeededUniformExpand)
void seeded_expand(Context *c, const LevelBase &lb, const u64 *d_ct_seeds, u64 *d_poly, u64 poly_stride, u32 batch, cudaStream_t s) {
#if defined(TROYB_HOST_EMUL)
    (void)c; (void)lb; (void)d_ct_seeds; (void)d_poly; (void)poly_stride; (void)batch; (void)s;
    throw std::logic_error("seeded ciphertexts need the cuRAND device generator (not available in the host emulation)");
#else
    SeededUniformExpand k{};
    k.seeds = d_ct_seeds;
    k.N = (u32)c->N;
    k.n_limbs = lb.L;
    k.mods = c->d_mods;
    k.map = lb.map;
    k.out = d_poly;
    k.out_stride = poly_stride;
    k.total = (u64)batch * c->N;
    ck(launch_cta(k, grid_for(k.total, SeededUniformExpand::kThreads), 0, s));
#endif
}

void enc_combine(Context *c, const LevelBase &lb, int op, u32 polys, u32 batch, const u64 *x, u64 x_batch, u64 x_poly, const u64 *y,
                 u64 y_batch, u64 y_poly, const u64 *z, u64 z_batch, u64 z_poly, u64 *out, u64 out_batch, u64 out_poly, u64 f,
                 cudaStream_t s) {
    EncCombine k{};
    k.op = op;
    k.logN = c->logN;
    k.n_limbs = lb.L;
    k.polys = polys;
    k.total = (u64)batch * polys * lb.L * c->N / 2;
    k.mods = c->d_mods;
    k.map = lb.map;
    k.f = f;
    k.x = x; k.x_batch = x_batch; k.x_poly = x_poly;
    k.y = y; k.y_batch = y_batch; k.y_poly = y_poly;
    k.z = z; k.z_batch = z_batch; k.z_poly = z_poly;
    k.out = out; k.out_batch = out_batch; k.out_poly = out_poly;
    ck(launch_cta(k, grid_for(k.total, EncCombine::kThreads), 0, s));
}

} // namespace

int troyb_prng_fill(const uint64_t *d_seeds, size_t seed_stride, uint64_t *d_out, size_t out_stride, size_t bytes, size_t batch,
                    void *stream) {
    API_BEGIN
    need_init();
    if (bytes % Blake2::kBufferBytes) throw std::invalid_argument("bytes must be a multiple of 4096 (the generator's buffer)");
    if (out_stride * 8 < bytes) throw std::invalid_argument("out_stride shorter than the stream");
    if (batch && bytes) prng_expand(d_seeds, seed_stride, d_out, out_stride, bytes, (u32)batch, S(stream));
    API_END
}

int troyb_sample_poly_ternary(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                              const uint64_t *d_seeds, size_t seed_stride, uint64_t start_bytes, uint64_t *d_out, size_t out_stride,
                              uint64_t *d_end_bytes, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const LevelBase lb = level_base(c, level);
    check_batch_stride(out_stride, (size_t)lb.L * c->N, "out");
    sample_small(c, lb, make_view(d_stream, stream_bytes, d_seeds), stream_stride, seed_stride, start_bytes, 1, 0, d_out, out_stride,
                 d_end_bytes, (u32)batch, S(stream));
    API_END
}

int troyb_sample_poly_cbd(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                          const uint64_t *d_seeds, size_t seed_stride, uint64_t start_bytes, uint64_t *d_out, size_t out_stride,
                          size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const LevelBase lb = level_base(c, level);
    check_batch_stride(out_stride, (size_t)lb.L * c->N, "out");
    sample_small(c, lb, make_view(d_stream, stream_bytes, d_seeds), stream_stride, seed_stride, start_bytes, 0, 1, d_out,
                 out_stride, nullptr, (u32)batch, S(stream));
    API_END
}

int troyb_sample_poly_uniform(troyb_context *ctx, int level, const uint64_t *d_stream, size_t stream_stride, size_t stream_bytes,
                              const uint64_t *d_seeds, size_t seed_stride, uint64_t *d_out, size_t out_stride, size_t batch,
                              void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const LevelBase lb = level_base(c, level);
    check_batch_stride(out_stride, (size_t)lb.L * c->N, "out");
    sample_uniform(c, lb, make_view(d_stream, stream_bytes, d_seeds), stream_stride, seed_stride, d_out, out_stride, (u32)batch,
                   S(stream));
    API_END
}

int troyb_encrypt_zero_asymmetric(troyb_context *ctx, int level, const uint64_t *d_public_key, const uint64_t *d_seeds,
                                  uint64_t *d_ct, size_t ct_stride, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    cudaStream_t s = S(stream);
    const u64 N = c->N;
    const bool ntt_form = c->scheme == SCHEME_CKKS;
    const LevelBase out_lb = level_base(c, level);
    // Encryptor::encryptZeroInternal  src/encryptor.cpp:119-156: encrypt one level up the chain (the key level for the first
    // data level) and divide by the extra prime, when the chain has such a level
    const bool switch_down = level == 0 ? c->has_special : level > 0;
    const int enc_level = switch_down ? level - 1 : level;
    const LevelBase lb = level_base(c, enc_level);
    const u64 pk_poly = (u64)level_limbs(c, -1) * N; // the key's polynomials keep the key level's stride
    check_batch_stride(ct_stride, 2 * (size_t)out_lb.L * N, "ct");
    check_aligned(d_ct, "ct");
    check_aligned(d_public_key, "public key");
    const u32 B = (u32)batch;
    if (!B) return TROYB_OK;
    // the generator's stream: N ternary draws (4 bytes each, + the rare rejected ones), then 2 x N noise draws of 6 bytes
    const u64 stream_bytes = 16 * N + 1024, sw = buffers_words(stream_bytes);
    Scratch st((size_t)B * sw, s);
    prng_expand(d_seeds, 8, st.p, sw, stream_bytes, B, s);
    const u64 poly = (u64)lb.L * N;
    Scratch smp((size_t)B * 3 * poly, s); // u, e_0, e_1
    sample_small(c, lb, make_view(st.p, sw * 8, d_seeds), sw, 8, 0, 1, 2, smp.p, 3 * poly, nullptr, B, s);
    Scratch wide(switch_down ? (size_t)B * 2 * poly : 0, s);
    u64 *enc = switch_down ? wide.p : d_ct;
    const u64 enc_stride = switch_down ? 2 * poly : ct_stride;
    if (ntt_form) {
        ntt_plain(c, false, lb.map, lb.L, smp.p, 3 * poly, smp.p, 3 * poly, 3 * lb.L, B, true, s);
        enc_combine(c, lb, 1, 2, B, smp.p, 3 * poly, 0, d_public_key, 0, pk_poly, smp.p + poly, 3 * poly, poly, enc, enc_stride, poly,
                    1, s);
    } else {
        ntt_plain(c, false, lb.map, lb.L, smp.p, 3 * poly, smp.p, 3 * poly, lb.L, B, true, s);
        enc_combine(c, lb, 0, 2, B, smp.p, 3 * poly, 0, d_public_key, 0, pk_poly, nullptr, 0, 0, enc, enc_stride, poly, 1, s);
        ntt_plain(c, true, lb.map, lb.L, enc, enc_stride, enc, enc_stride, 2 * lb.L, B, true, s);
        enc_combine(c, lb, 2, 2, B, nullptr, 0, 0, nullptr, 0, 0, smp.p + poly, 3 * poly, poly, enc, enc_stride, poly,
                    c->scheme == SCHEME_BGV ? c->plain : 1, s);
    }
    if (switch_down)
        drop_last(c, enc_level == -1 ? c->key_drop : c->levels[enc_level].drop, lb.map, lb.L, ntt_form, enc, enc_stride, d_ct, ct_stride,
                  2, B, s);
    API_END
}

namespace {
// encryptZeroSymmetric. d_ct_seeds == null: c1 is drawn from the ciphertext generator by samplePolyUniform (the CPU path,
// src/utils/rlwe.cpp:245-356); else c1 is the cuRAND expansion of a 64-bit seed taken from that generator, which is written to
// d_ct_seeds [batch] (the CUDA path's seeded form, src/utils/rlwe_cuda.cu:261-333).
void encrypt_zero_symmetric(Context *c, int level, const u64 *d_secret_key, const u64 *d_seeds, u64 *d_ct, u64 ct_stride,
                            u64 *d_ct_seeds, u32 B, cudaStream_t s) {
    const u64 N = c->N;
    const bool ntt_form = c->scheme == SCHEME_CKKS;
    const LevelBase lb = level_base(c, level);
    const u64 poly = (u64)lb.L * N;
    check_batch_stride(ct_stride, 2 * (size_t)poly, "ct");
    check_aligned(d_ct, "ct");
    check_aligned(d_secret_key, "secret key");
    if (!B) return;
    // bootstrap generator: 64 bytes = the seed of the generator that expands c1, then N noise draws of 6 bytes
    const u64 boot_bytes = 64 + 6 * N, bw = buffers_words(boot_bytes);
    Scratch boot((size_t)B * bw, s);
    prng_expand(d_seeds, 8, boot.p, bw, boot_bytes, B, s);
    u64 *c1 = d_ct + poly;
    if (d_ct_seeds) {
        const u64 uw = buffers_words(Blake2::kBufferBytes);
        Scratch uni((size_t)B * uw, s);
        prng_expand(boot.p, bw, uni.p, uw, Blake2::kBufferBytes, B, s);
        SeedFromStream k{};
        k.view = make_view(uni.p, uw * 8, boot.p);
        k.view_stride = uw;
        k.seed_stride = bw;
        k.batch = B;
        k.out = d_ct_seeds;
        ck(launch_cta(k, grid_for(B, SeedFromStream::kThreads), 0, s));
        seeded_expand(c, lb, d_ct_seeds, c1, ct_stride, B, s);
    } else {
        // ciphertext generator: L * N words + the redraws. A word is rejected with probability (2^64 mod q + 1) / 2^64 (tiny for
        // the usual primes just below a power of two): expand twice the expected number + a buffer; the rest is computed on demand
        double redraws = 0;
        for (u32 l = 0; l < lb.L; l++) {
            const u64 q = c->pool[lb.map.idx[l]];
            redraws += (double)N * ((double)((~0ull) % q + 2) / 18446744073709551616.0);
        }
        const u64 uni_bytes = poly * 8 + 16 * (u64)redraws + 4096, uw = buffers_words(uni_bytes);
        Scratch uni((size_t)B * uw, s);
        prng_expand(boot.p, bw, uni.p, uw, uni_bytes, B, s);
        sample_uniform(c, lb, make_view(uni.p, uw * 8, boot.p), uw, bw, c1, ct_stride, B, s);
    }
    Scratch noise((size_t)B * poly, s);
    sample_small(c, lb, make_view(boot.p, bw * 8, d_seeds), bw, 8, 64, 0, 1, noise.p, poly, nullptr, B, s);
    const u64 f = c->scheme == SCHEME_BGV ? c->plain : 1;
    if (ntt_form) {
        ntt_plain(c, false, lb.map, lb.L, noise.p, poly, noise.p, poly, lb.L, B, true, s);
        enc_combine(c, lb, 4, 1, B, c1, ct_stride, 0, d_secret_key, 0, 0, noise.p, poly, 0, d_ct, ct_stride, 0, f, s);
    } else {
        enc_combine(c, lb, 0, 1, B, c1, ct_stride, 0, d_secret_key, 0, 0, nullptr, 0, 0, d_ct, ct_stride, 0, 1, s);
        ntt_plain(c, true, lb.map, lb.L, d_ct, ct_stride, d_ct, ct_stride, 2 * lb.L, B, true, s);
        enc_combine(c, lb, 3, 1, B, nullptr, 0, 0, nullptr, 0, 0, noise.p, poly, 0, d_ct, ct_stride, 0, f, s);
    }
}
} // namespace

int troyb_encrypt_zero_symmetric(troyb_context *ctx, int level, const uint64_t *d_secret_key, const uint64_t *d_seeds, uint64_t *d_ct,
                                 size_t ct_stride, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    encrypt_zero_symmetric(C(ctx), level, d_secret_key, d_seeds, d_ct, ct_stride, nullptr, (u32)batch, S(stream));
    API_END
}

int troyb_encrypt_zero_symmetric_seeded(troyb_context *ctx, int level, const uint64_t *d_secret_key, const uint64_t *d_seeds,
                                        uint64_t *d_ct, size_t ct_stride, uint64_t *d_ct_seeds, size_t batch, void *stream) {
    API_BEGIN
    need_init();
    if (!d_ct_seeds) throw std::invalid_argument("ct_seeds is null");
    encrypt_zero_symmetric(C(ctx), level, d_secret_key, d_seeds, d_ct, ct_stride, d_ct_seeds, (u32)batch, S(stream));
    API_END
}

int troyb_expand_seeded_poly(troyb_context *ctx, int level, const uint64_t *d_ct_seeds, uint64_t *d_poly, size_t poly_stride,
                             size_t batch, void *stream) {
    API_BEGIN
    need_init();
    Context *c = C(ctx);
    const LevelBase lb = level_base(c, level);
    check_batch_stride(poly_stride, (size_t)lb.L * c->N, "poly");
    seeded_expand(c, lb, d_ct_seeds, d_poly, poly_stride, (u32)batch, S(stream));
    API_END
}

} // extern "C"
