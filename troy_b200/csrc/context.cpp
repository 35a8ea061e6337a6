// Host precompute + upload of all tables. See context.h for the layout and the reference citations.
#include "context.h"
#include "hostmath.h"

#include <algorithm>
#include <memory>
#include <cstring>
#include <cstdlib>
#include <stdexcept>

namespace troyb {

using namespace host;

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_)); \
    } while (0)

namespace {

template <typename T> T *upload(std::vector<void *> &owned, const std::vector<T> &v) {
    if (v.empty()) return nullptr;
    void *d = nullptr;
    CUDA_OK(cudaMalloc(&d, v.size() * sizeof(T)));
    CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    owned.push_back(d);
    return static_cast<T *>(d);
}

void push_shoup(std::vector<u64> &v, u64 w, u64 q) {
    v.push_back(w);
    v.push_back(shoup_quotient(w, q));
}

// prod_{k in idx, k != skip} (base[k] mod m) mod m  — (Q / q_skip) mod m without big integers
u64 punctured_prod_mod(const std::vector<u64> &base, int skip, u64 m) {
    u64 r = 1 % m;
    for (int k = 0; k < (int)base.size(); k++) {
        if (k == skip) continue;
        r = mulmod(r, base[k] % m, m);
    }
    return r;
}

u64 ratio_hi(u64 q) { return (u64)((((u128)1 << 127) / q * 2 + ((((u128)1 << 127) % q) * 2) / q) >> 64); }

void barrett_ratio(u64 q, u64 &r0, u64 &r1) {
    // floor(2^128 / q): long division of (1 << 128) by q
    u128 hi = ((u128)1 << 64) / q;           // floor(2^64 / q) -> contributes to the high word
    u128 rem = ((u128)1 << 64) % q;          // remainder
    u128 lo = (rem << 64) / q;               // next 64 bits
    r1 = (u64)hi;
    r0 = (u64)lo;
    (void)ratio_hi;
}

// Plaintext-space constants of one level (src/context.cpp:296-352): floor(Q / t) and Q mod t need the full product
// Q = q_0 ... q_{L-1}, formed here as a little-endian multi-word integer.
void build_plain(const Context &ctx, LevelData &lv) {
    const u64 t = ctx.plain;
    std::vector<u64> Q{1};
    for (int i = 0; i < lv.L; i++) {
        u64 carry = 0;
        for (size_t k = 0; k < Q.size(); k++) {
            u128 v = (u128)Q[k] * ctx.pool[i] + carry;
            Q[k] = (u64)v;
            carry = (u64)(v >> 64);
        }
        if (carry) Q.push_back(carry);
    }
    std::vector<u64> quo(Q.size());
    u64 rem = 0;
    for (size_t k = Q.size(); k-- > 0;) {
        u128 v = ((u128)rem << 64) | Q[k];
        quo[k] = (u64)(v / t);
        rem = (u64)(v % t);
    }
    PlainConsts &pc = lv.plain;
    pc.t = t;
    pc.threshold = (t + 1) >> 1;
    pc.q_mod_t = rem;
    for (int i = 0; i < lv.L; i++) {
        const u64 qi = ctx.pool[i];
        u64 d = 0;
        for (size_t k = quo.size(); k-- > 0;) d = (u64)((((u128)d << 64) | quo[k]) % qi);
        pc.delta[i] = d;
        pc.inc[i] = (qi - t % qi) % qi;
    }
    lv.has_plain = true;
}

// RNSTool's constants for BFV decryption ({t, gamma} base, src/utils/rns.cpp:748-775) and BGV decryption (exact conversion
// q -> t, src/utils/rns.cpp:462-538)
void build_decrypt(const Context &ctx, LevelData &lv) {
    const u64 t = ctx.plain;
    std::vector<u64> q(ctx.pool.begin(), ctx.pool.begin() + lv.L);
    DecryptConsts &d = lv.decrypt;
    d.t = t;
    d.q_mod_t = punctured_prod_mod(q, -1, t);
    u64 inv;
    for (int i = 0; i < lv.L; i++) {
        d.punct_mod_t[i] = punctured_prod_mod(q, i, t);
        d.inv_punct_q[i] = invert(punctured_prod_mod(q, i, q[i]), q[i]);
    }
    if (ctx.pool_gamma >= 0) {
        const u64 g = ctx.pool[ctx.pool_gamma];
        d.gamma = g;
        if (!try_invert(g % t, t, inv)) throw std::logic_error("invalid rns bases");
        d.inv_gamma_mod_t = inv;
        if (!try_invert(d.q_mod_t, t, inv)) throw std::logic_error("invalid rns bases");
        d.neg_inv_q_mod_t = (t - inv) % t;
        if (!try_invert(punctured_prod_mod(q, -1, g), g, inv)) throw std::logic_error("invalid rns bases");
        d.neg_inv_q_mod_gamma = (g - inv) % g;
        for (int i = 0; i < lv.L; i++) {
            d.tg_mod_q[i] = mulmod(t % q[i], g % q[i], q[i]);
            d.punct_mod_gamma[i] = punctured_prod_mod(q, i, g);
        }
    }
    lv.has_decrypt = true;
}

void build_drop(Context &ctx, const std::vector<u64> &base, std::vector<void *> &owned, DropDev &out, DropLimb *&d_limbs,
                int last_pool) {
    int m = (int)base.size() - 1;
    u64 qlast = base.back();
    std::vector<DropLimb> limbs(m);
    for (int i = 0; i < m; i++) {
        u64 qi = base[i];
        u64 inv = invert(qlast % qi, qi);
        limbs[i].inv_last = inv;
        limbs[i].inv_last_shoup = shoup_quotient(inv, qi);
        limbs[i].half_mod = (qlast >> 1) % qi;
        limbs[i].last_mod = qlast % qi;
        limbs[i].last_mod_shoup = shoup_quotient(qlast % qi, qi);
    }
    d_limbs = upload(owned, limbs);
    out.limbs = d_limbs;
    out.q_last = qlast;
    out.half = qlast >> 1;
    u64 r0, r1;
    barrett_ratio(qlast, r0, r1);
    out.last_ratio1 = r1;
    out.t = ctx.plain;
    out.t_ratio1 = 0;
    out.inv_last_mod_t = 0;
    out.inv_last_mod_t_shoup = 0;
    out.last_pool = last_pool;
    if (ctx.plain > 1) {
        u64 t0, t1;
        barrett_ratio(ctx.plain, t0, t1);
        out.t_ratio1 = t1;
        u64 inv;
        if (try_invert(qlast % ctx.plain, ctx.plain, inv)) {
            out.inv_last_mod_t = inv;
            out.inv_last_mod_t_shoup = shoup_quotient(inv, ctx.plain);
        }
    }
}

void build_behz(Context &ctx, LevelData &lv) {
    // src/utils/rns.cpp:581-803 (RNSTool::initialize), restated over the flat modulus pool
    int L = lv.L;
    std::vector<u64> q(ctx.pool.begin(), ctx.pool.begin() + L);
    int total_bits = ctx.level_total_bits((int)(&lv - ctx.levels.data()));
    int nB = L;
    if (32 + bit_count(ctx.plain) + total_bits >= 61 * L + 61) nB++;
    int nBsk = nB + 1;
    if (ctx.pool_B0 + nB > (int)ctx.pool.size()) throw std::logic_error("BEHZ auxiliary base not available");
    std::vector<u64> B(ctx.pool.begin() + ctx.pool_B0, ctx.pool.begin() + ctx.pool_B0 + nB);
    u64 msk = ctx.pool[ctx.pool_msk];
    std::vector<u64> Bsk = B;
    Bsk.push_back(msk);
    const u64 mtilde = u64(1) << 32;

    BehzDev d{};
    d.L = L; d.nB = nB; d.nBsk = nBsk;
    auto &own = lv.owned;

    std::vector<u64> v_mt_inv, v_inv, v_q2bsk(nBsk * L), v_q2mt(L);
    for (int i = 0; i < L; i++) {
        u64 punct = punctured_prod_mod(q, i, q[i]);
        u64 inv = invert(punct, q[i]);
        push_shoup(v_inv, inv, q[i]);
        push_shoup(v_mt_inv, mulmod(mtilde % q[i], inv, q[i]), q[i]);
        for (int o = 0; o < nBsk; o++) v_q2bsk[o * L + i] = punctured_prod_mod(q, i, Bsk[o]);
        v_q2mt[i] = punctured_prod_mod(q, i, mtilde);
    }
    d.mtilde_inv_punct_q = upload(own, v_mt_inv);
    d.inv_punct_q = upload(own, v_inv);
    d.q_to_bsk = upload(own, v_q2bsk);
    d.q_to_mtilde = upload(own, v_q2mt);

    u64 prod_q_mod_mt = punctured_prod_mod(q, -1, mtilde);
    u64 inv_q_mt = invert(prod_q_mod_mt, mtilde);
    d.neg_inv_q_mod_mtilde = (mtilde - inv_q_mt) % mtilde;

    std::vector<u64> v_pq, v_imt, v_ipq;
    for (int o = 0; o < nBsk; o++) {
        u64 pq = punctured_prod_mod(q, -1, Bsk[o]);
        push_shoup(v_pq, pq, Bsk[o]);
        push_shoup(v_imt, invert(mtilde % Bsk[o], Bsk[o]), Bsk[o]);
        push_shoup(v_ipq, invert(pq, Bsk[o]), Bsk[o]);
    }
    d.prod_q_mod_bsk = upload(own, v_pq);
    d.inv_mtilde_mod_bsk = upload(own, v_imt);
    d.inv_prod_q_mod_bsk = upload(own, v_ipq);

    std::vector<u64> v_ipB, v_B2q(L * nB), v_B2msk(nB);
    for (int b = 0; b < nB; b++) {
        u64 punct = punctured_prod_mod(B, b, B[b]);
        push_shoup(v_ipB, invert(punct, B[b]), B[b]);
        for (int i = 0; i < L; i++) v_B2q[i * nB + b] = punctured_prod_mod(B, b, q[i]);
        v_B2msk[b] = punctured_prod_mod(B, b, msk);
    }
    d.inv_punct_B = upload(own, v_ipB);
    d.B_to_q = upload(own, v_B2q);
    d.B_to_msk = upload(own, v_B2msk);
    u64 ipb = invert(punctured_prod_mod(B, -1, msk), msk);
    d.inv_prod_B_mod_msk = ipb;
    d.inv_prod_B_mod_msk_shoup = shoup_quotient(ipb, msk);

    std::vector<u64> v_pB, v_npB, v_t;
    for (int i = 0; i < L; i++) {
        u64 pb = punctured_prod_mod(B, -1, q[i]);
        push_shoup(v_pB, pb, q[i]);
        push_shoup(v_npB, q[i] - pb, q[i]);
    }
    for (int i = 0; i < L; i++) push_shoup(v_t, ctx.plain % q[i], q[i]);
    for (int o = 0; o < nBsk; o++) push_shoup(v_t, ctx.plain % Bsk[o], Bsk[o]);
    d.prod_B_mod_q = upload(own, v_pB);
    d.neg_prod_B_mod_q = upload(own, v_npB);
    d.t_mod = upload(own, v_t);

    lv.behz = d;
    lv.has_behz = true;
}

} // namespace

Context::~Context() {
    cudaSetDevice(device);
    for (void *p : owned) cudaFree(p);
    for (auto &lv : levels)
        for (void *p : lv.owned) cudaFree(p);
}

LimbMap Context::map_prefix(int L) const {
    LimbMap m{};
    for (int i = 0; i < L; i++) m.idx[i] = (uint8_t)i;
    return m;
}

LimbMap Context::map_prefix_special(int L) const {
    LimbMap m = map_prefix(L);
    m.idx[L] = (uint8_t)pool_special;
    return m;
}

LimbMap Context::map_q_bsk(int level) const {
    const LevelData &lv = levels[level];
    LimbMap m = map_prefix(lv.L);
    for (int b = 0; b < lv.behz.nB; b++) m.idx[lv.L + b] = (uint8_t)(pool_B0 + b);
    m.idx[lv.L + lv.behz.nB] = (uint8_t)pool_msk;
    return m;
}

int Context::level_total_bits(int level) const {
    // bit count of prod(q_0..q_{L-1}) computed exactly with a small big-integer; level -1 is the key level: the data
    // primes of the first level and the special prime
    const bool key_level = level < 0;
    int L = key_level ? Ltop : levels[level].L;
    std::vector<u64> acc{1};
    for (int i = 0; i < L; i++) {
        u64 carry = 0;
        for (auto &w : acc) {
            u128 t = (u128)w * pool[i] + carry;
            w = (u64)t;
            carry = (u64)(t >> 64);
        }
        if (carry) acc.push_back(carry);
    }
    if (key_level && has_special) {
        u64 carry = 0;
        for (auto &w : acc) {
            u128 t = (u128)w * pool[pool_special] + carry;
            w = (u64)t;
            carry = (u64)(t >> 64);
        }
        if (carry) acc.push_back(carry);
    }
    return (int)(acc.size() - 1) * 64 + bit_count(acc.back());
}

Context *context_create(int device, int scheme, u64 N, const u64 *moduli, int n_moduli, u64 plain_modulus,
                        bool expand_chain) {
    if (scheme != SCHEME_BFV && scheme != SCHEME_CKKS && scheme != SCHEME_BGV) throw std::invalid_argument("invalid scheme");
    int logN = 0;
    while ((u64(1) << logN) < N) logN++;
    if ((u64(1) << logN) != N || N < 2 || N > 131072) throw std::invalid_argument("poly_modulus_degree is invalid");
    // the transforms are planned for N <= 65536 (a cluster of eight 8192-coefficient tiles); say so now, not at the first NTT
    if (N > 65536) throw std::invalid_argument("poly_modulus_degree above 65536 is not supported");
    if (n_moduli < 1 || n_moduli > 64) throw std::invalid_argument("invalid coeff modulus size");
    for (int i = 0; i < n_moduli; i++) {
        if ((moduli[i] >> 60) || moduli[i] < 2) throw std::invalid_argument("invalid coeff modulus bit count");
        if (!is_prime(moduli[i]) || (moduli[i] - 1) % (2 * N) != 0) throw std::invalid_argument("coeff modulus does not support NTT");
        for (int j = 0; j < i; j++)
            if (moduli[i] == moduli[j]) throw std::invalid_argument("coeff moduli are not distinct");
    }
    if (scheme == SCHEME_CKKS) {
        if (plain_modulus != 0) throw std::invalid_argument("plain_modulus is not zero");
    } else {
        if (plain_modulus < 2 || (plain_modulus >> 60)) throw std::invalid_argument("invalid plain modulus bit count");
        for (int i = 0; i < n_moduli; i++)
            if (std::__gcd(moduli[i], plain_modulus) != 1) throw std::invalid_argument("plain_modulus is not coprime to coeff_modulus");
    }

    CUDA_OK(cudaSetDevice(device));
    std::unique_ptr<Context> ctx(new Context());
    ctx->device = device;
    {
        int sms = 0;
        CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        ctx->sm_count = sms > 0 ? sms : 1;
    }
    ctx->scheme = scheme;
    ctx->N = N;
    ctx->logN = logN;
    ctx->plain = plain_modulus;
    ctx->has_special = n_moduli > 1;
    ctx->Ltop = ctx->has_special ? n_moduli - 1 : n_moduli;
    ctx->pool.assign(moduli, moduli + n_moduli);
    ctx->pool_special = ctx->has_special ? n_moduli - 1 : -1;

    // BEHZ auxiliary primes: getPrimes(2N, 61, |Bsk ∪ m_tilde|) = m_sk, gamma, B_0, B_1, ... (src/utils/rns.cpp:629-635)
    if (scheme == SCHEME_BFV) {
        int need = ctx->Ltop + 1 + 2; // up to |B| = L+1, plus m_sk and gamma
        std::vector<u64> aux = get_primes(2 * N, 61, need);
        ctx->pool_msk = (int)ctx->pool.size();
        ctx->pool_gamma = ctx->pool_msk + 1;
        ctx->pool_B0 = ctx->pool_msk + 2;
        for (u64 p : aux) ctx->pool.push_back(p);
    }
    if ((int)ctx->pool.size() > kMaxLimbs) throw std::invalid_argument("too many moduli");

    // ---- NTT tables for every modulus in the pool (src/utils/ntt.cpp:17-66)
    ctx->h_mods.resize(ctx->pool.size());
    std::vector<u64> fwd(2 * N), inv(2 * N);
    for (size_t m = 0; m < ctx->pool.size(); m++) {
        u64 p = ctx->pool[m];
        ModDev &md = ctx->h_mods[m];
        md.p = p;
        barrett_ratio(p, md.ratio0, md.ratio1);
        md.bar_shift = 63u - (u32)__builtin_clzll(p);
        md.bar_mu = (u64)(((unsigned __int128)1 << (md.bar_shift + 64)) / p);
        u64 psi;
        if (!minimal_primitive_root(2 * N, p, psi)) throw std::invalid_argument("invalid modulus");
        u64 power = 1, ipower = 1;
        const u64 psi_inv = invert(psi, p);
        for (u64 i = 0; i < N; i++) {
            u32 k = reverse_bits((u32)i, logN);
            u64 wi = ipower;
            ipower = mulmod(ipower, psi_inv, p);
            fwd[2 * k] = power;
            fwd[2 * k + 1] = shoup_quotient(power, p);
            inv[2 * k] = wi;
            inv[2 * k + 1] = shoup_quotient(wi, p);
            power = mulmod(power, psi, p);
        }
        md.ninv = invert(N % p, p);
        md.ninv_shoup = shoup_quotient(md.ninv, p);
        md.inv1n = mulmod(inv[2], md.ninv, p);
        md.inv1n_shoup = shoup_quotient(md.inv1n, p);
        md.fwd = upload(ctx->owned, fwd);
        md.inv = upload(ctx->owned, inv);
        // butterfly policy (ntt.cuh). TROYB_NTT_POLICY=exact|fast forces the integer paths, nof64l only for 46..50-bit primes (A/B measurements).
        const char *force = getenv("TROYB_NTT_POLICY");
        const bool force_exact = force && std::string(force) == "exact";
        const bool force_fast = force && std::string(force) == "fast";
        md.policy = (p >> 60) || force_exact ? POLICY_U64_EXACT : ((p >> 57) || force_fast ? POLICY_U64_FAST : POLICY_U64_LAZY);
        md.fwd_d = md.inv_d = nullptr;
        md.pd = (double)p;
        md.pinv_d = 1.0 / (double)p;
        md.ninv_d = (double)md.ninv;
        md.ninv_dp = (double)md.ninv / (double)p;
        md.inv1n_d = (double)md.inv1n;
        md.inv1n_dp = (double)md.inv1n / (double)p;
        const bool no_f64l = force && std::string(force) == "nof64l"; // A/B: 46..50-bit primes on the integer path
        if (!(p >> 50) && !force_exact && !force_fast && !(no_f64l && (p >> 46))) {
            md.policy = (p >> 46) ? POLICY_F64_LARGE : ((p >> 41) ? POLICY_F64 : POLICY_F64_SMALL);
            // p < 2^46: one double per twiddle (the quotient is estimated from the product, ntt.cuh ArithF64T<true>);
            // 46..50 bits: pairs {w, w/p}
            const u64 per = md.policy == POLICY_F64_LARGE ? 2 : 1;
            std::vector<double> fd(per * N), id(per * N);
            for (u64 k = 0; k < N; k++) {
                fd[per * k] = (double)fwd[2 * k];
                id[per * k] = (double)inv[2 * k];
                if (per == 2) {
                    fd[2 * k + 1] = (double)fwd[2 * k] / (double)p;
                    id[2 * k + 1] = (double)inv[2 * k] / (double)p;
                }
            }
            md.fwd_d = upload(ctx->owned, fd);
            md.inv_d = upload(ctx->owned, id);
        }
    }
    ctx->d_mods = upload(ctx->owned, ctx->h_mods);

    // ---- modulus-switching chain (src/context.cpp:470-524): level l keeps q_0..q_{Ltop-1-l}
    int max_levels = expand_chain ? ctx->Ltop : 1;
    for (int l = 0; l < max_levels; l++) {
        LevelData lv;
        lv.L = ctx->Ltop - l;
        ctx->levels.push_back(std::move(lv));
        if (scheme != SCHEME_CKKS) {
            // plain_modulus must stay smaller than the total coeff modulus (src/context.cpp:277-285)
            int bits = ctx->level_total_bits(l);
            bool ok = bits > bit_count(plain_modulus);
            if (!ok && bits == bit_count(plain_modulus)) {
                ok = ctx->levels[l].L == 1 && ctx->pool[0] > plain_modulus;
            }
            if (!ok) {
                ctx->levels.pop_back();
                break;
            }
        }
    }
    if (ctx->levels.empty()) throw std::invalid_argument("plain_modulus is not smaller than coeff_modulus");
    ctx->total_bits_first = ctx->level_total_bits(0);

    for (size_t l = 0; l < ctx->levels.size(); l++) {
        LevelData &lv = ctx->levels[l];
        if (lv.L >= 2) {
            std::vector<u64> base(ctx->pool.begin(), ctx->pool.begin() + lv.L);
            build_drop(*ctx, base, lv.owned, lv.drop, lv.d_drop_limbs, lv.L - 1);
            lv.last_prime_as_double = (double)base.back();
        }
        if (scheme == SCHEME_BFV) build_behz(*ctx, lv);
        if (scheme != SCHEME_CKKS && plain_modulus > 1) build_decrypt(*ctx, lv);
        if (scheme != SCHEME_CKKS && plain_modulus > 1) build_plain(*ctx, lv);
    }
    if (ctx->has_special) {
        std::vector<u64> base(ctx->pool.begin(), ctx->pool.begin() + ctx->Ltop);
        base.push_back(ctx->pool[ctx->pool_special]);
        build_drop(*ctx, base, ctx->owned, ctx->key_drop, ctx->d_key_drop_limbs, ctx->pool_special);
    }

    // Stream-ordered scratch pool that keeps its memory (replaces MemoryPoolCuda, src/utils/memorypool_cuda.cuh)
    CUDA_OK(cudaDeviceGetDefaultMemPool(&ctx->mempool, device));
    uint64_t threshold = UINT64_MAX;
    CUDA_OK(cudaMemPoolSetAttribute(ctx->mempool, cudaMemPoolAttrReleaseThreshold, &threshold));
    return ctx.release();
}

} // namespace troyb
