// TEST INFRASTRUCTURE ONLY — the product (troy_b200/) never links or loads this.
//
// Thin extern "C" adaptor over the UNMODIFIED reference troy:: CPU library (compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/libtroyref.so).  It lets the Python
// tests and bench.py's reference arm drive the reference's own public API
// (SEALContext / KeyGenerator / Encryptor / Evaluator / Decryptor / encoders, and the
// util::NTTTables / RNSTool / GaloisTool members underneath) through ctypes with plain
// pointers.  No reference source is copied: this file only *calls* the reference.
//
// Ciphertext word layout is the reference's own: [poly][limb][coeff] (src/ciphertext.h:300-348).
// Key-switching key layout exported by ref_kswitch_key_words: [digit j][component k][limb][coeff]
// at the key level (src/kswitchkeys.h:88-130, src/publickey.h).

#include "troy_cpu.h"
#include "utils/rns.h"
#include "utils/ntt.h"
#include "utils/galois.h"
#include "utils/polyarithsmallmod.h"
#include "utils/rlwe.h"

#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <complex>
#include <chrono>
#include <thread>
#include <atomic>

using namespace troy;
using troy::util::HostPointer;
using troy::util::ConstHostPointer;

namespace {

thread_local std::string g_last_error;

struct RefCtx {
    std::shared_ptr<SEALContext> context;
    std::unique_ptr<KeyGenerator> keygen;
    std::unique_ptr<Evaluator> evaluator;
    std::unique_ptr<Encryptor> encryptor;
    std::unique_ptr<Decryptor> decryptor;
    std::unique_ptr<BatchEncoder> batch;
    std::unique_ptr<CKKSEncoder> ckks;
    PublicKey pk;
    SchemeType scheme;
    size_t n;
    std::vector<ParmsID> chain; // chain[0] = first (data) level, ... ; key level separate
};

struct RefKeys { // RelinKeys or GaloisKeys
    KSwitchKeys keys;
    bool galois;
};

std::shared_ptr<const SEALContext::ContextData> level_data(RefCtx *c, int level) {
    if (level < 0) return c->context->keyContextData();
    return c->context->getContextData(c->chain.at(level));
}

#define REF_TRY try {
#define REF_CATCH(ret_err)                                   \
    } catch (const std::invalid_argument &e) { g_last_error = std::string("invalid_argument: ") + e.what(); return ret_err; } \
      catch (const std::out_of_range &e)     { g_last_error = std::string("out_of_range: ") + e.what(); return ret_err; }     \
      catch (const std::logic_error &e)      { g_last_error = std::string("logic_error: ") + e.what(); return ret_err; }      \
      catch (const std::exception &e)        { g_last_error = std::string("exception: ") + e.what(); return ret_err; }

} // namespace

extern "C" {

const char *ref_last_error() { return g_last_error.c_str(); }

// CoeffModulus::Create (src/modulus.cpp:80-122)
int ref_coeff_modulus_create(uint64_t n, const int *bits, int count, uint64_t *out) {
    REF_TRY
    auto v = CoeffModulus::Create(n, std::vector<int>(bits, bits + count));
    for (int i = 0; i < count; i++) out[i] = v[i].value();
    return 0;
    REF_CATCH(-1)
}

// scheme: 1 bfv, 2 ckks, 3 bgv (src/encryptionparams.h SchemeType)
void *ref_ctx_create(int scheme, uint64_t n, const uint64_t *moduli, int n_moduli, uint64_t plain_modulus,
                     const uint64_t *seed8) {
    REF_TRY
    EncryptionParameters parms(static_cast<uint8_t>(scheme));
    parms.setPolyModulusDegree(n);
    std::vector<Modulus> q;
    for (int i = 0; i < n_moduli; i++) q.emplace_back(moduli[i]);
    parms.setCoeffModulus(q);
    if (scheme != 2) parms.setPlainModulus(plain_modulus);
    if (seed8) {
        PRNGSeed seed;
        for (size_t i = 0; i < seed.size(); i++) seed[i] = seed8[i];
        parms.setRandomGenerator(std::make_shared<Blake2xbPRNGFactory>(seed));
    }
    auto c = std::make_unique<RefCtx>();
    c->context = std::make_shared<SEALContext>(parms, true, SecurityLevel::none);
    if (!c->context->parametersSet()) {
        g_last_error = std::string("parameters not set: ") + c->context->parameterErrorMessage();
        return nullptr;
    }
    c->scheme = parms.scheme();
    c->n = n;
    auto d = c->context->firstContextData();
    while (d) { c->chain.push_back(d->parmsID()); d = d->nextContextData(); }
    c->evaluator = std::make_unique<Evaluator>(*c->context);
    return c.release();
    REF_CATCH(nullptr)
}

void ref_ctx_destroy(void *h) { delete static_cast<RefCtx *>(h); }

int ref_ctx_num_levels(void *h) { return (int)static_cast<RefCtx *>(h)->chain.size(); }

// level = -1 → key level. Returns limb count; writes moduli if out != null.
int ref_ctx_level_moduli(void *h, int level, uint64_t *out) {
    REF_TRY
    auto d = level_data(static_cast<RefCtx *>(h), level);
    auto &q = d->parms().coeffModulus();
    if (out) for (size_t i = 0; i < q.size(); i++) out[i] = q[i].value();
    return (int)q.size();
    REF_CATCH(-1)
}

int ref_ctx_using_keyswitching(void *h) { return static_cast<RefCtx *>(h)->context->using_keyswitching() ? 1 : 0; }

// ---- tables (src/utils/ntt.h NTTTables) -------------------------------------------------
// out_root/out_inv: N entries of {operand, quotient}; scalars: [root, inv_root, inv_degree.operand, inv_degree.quotient]
int ref_ntt_tables(void *h, int level, int limb, uint64_t *out_root, uint64_t *out_inv, uint64_t *scalars) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto d = level_data(c, level);
    const util::NTTTables &t = d->smallNTTTables()[limb];
    for (size_t i = 0; i < c->n; i++) {
        if (out_root) { out_root[2 * i] = t.getFromRootPowers()[i].operand; out_root[2 * i + 1] = t.getFromRootPowers()[i].quotient; }
        if (out_inv)  { out_inv[2 * i] = t.getFromInvRootPowers()[i].operand; out_inv[2 * i + 1] = t.getFromInvRootPowers()[i].quotient; }
    }
    if (scalars) {
        scalars[0] = t.getRoot();
        scalars[1] = 0;
        scalars[2] = t.invDegreeModulo().operand;
        scalars[3] = t.invDegreeModulo().quotient;
    }
    return 0;
    REF_CATCH(-1)
}

// Standalone NTT on one limb with a modulus of the given level (util::nttNegacyclicHarvey, src/utils/ntt.cpp:158-213)
int ref_ntt(void *h, int level, int limb, uint64_t *data, int inverse) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto d = level_data(c, level);
    const util::NTTTables &t = d->smallNTTTables()[limb];
    if (inverse) util::inverseNttNegacyclicHarvey(HostPointer<uint64_t>(data), t);
    else util::nttNegacyclicHarvey(HostPointer<uint64_t>(data), t);
    return 0;
    REF_CATCH(-1)
}

// NTT under an arbitrary prime (used for the standalone sweep with 60-bit primes and for Bsk tables)
int ref_ntt_prime(uint64_t n, uint64_t prime, uint64_t *data, int inverse, int reps) {
    REF_TRY
    int logn = util::getPowerOfTwo(n);
    util::NTTTables t(logn, Modulus(prime));
    for (int r = 0; r < reps; r++) {
        if (inverse) util::inverseNttNegacyclicHarvey(HostPointer<uint64_t>(data), t);
        else util::nttNegacyclicHarvey(HostPointer<uint64_t>(data), t);
    }
    return 0;
    REF_CATCH(-1)
}

// Multiplication by X^shift modulo X^n + 1 under one prime (src/utils/polyarithsmallmod.cpp:128-152), shift in [0, 2n),
// and multiplication by n^-1 (what EvaluatorCuda::divideByPolyModulusDegreeInplace does per limb, src/evaluator_cuda.cu:2262):
// the CPU statements of the two kernels behind the LWE extraction / packing methods.
int ref_negacyclic_shift(uint64_t n, uint64_t prime, uint64_t shift, const uint64_t *in, uint64_t *out) {
    REF_TRY
    util::negacyclicShiftPolyCoeffmod(ConstHostPointer<uint64_t>(in), (size_t)n, (size_t)shift, Modulus(prime), HostPointer<uint64_t>(out));
    return 0;
    REF_CATCH(-1)
}
int ref_multiply_inv_degree(uint64_t n, uint64_t prime, const uint64_t *in, uint64_t *out) {
    REF_TRY
    Modulus q(prime);
    uint64_t inv = 0;
    if (!util::tryInvertUintMod(n, q, inv)) throw std::invalid_argument("n is not invertible");
    util::multiplyPolyScalarCoeffmod(ConstHostPointer<uint64_t>(in), (size_t)n, inv, q, HostPointer<uint64_t>(out));
    return 0;
    REF_CATCH(-1)
}

// ---- RNSTool pieces (src/utils/rns.cpp) ----------------------------------------------------
// info: [|q|, |B|, |Bsk|, |Bsk∪m̃|, m_sk, gamma, m_tilde, t]; bsk_out: Bsk primes
int ref_rns_info(void *h, int level, uint64_t *info, uint64_t *bsk_out) {
    REF_TRY
    auto d = level_data(static_cast<RefCtx *>(h), level);
    auto r = d->rnsTool();
    info[0] = r->baseq()->size(); info[1] = r->baseB()->size(); info[2] = r->baseBsk()->size();
    info[3] = r->baseBskmTilde()->size(); info[4] = r->msk().value(); info[5] = r->gamma().value();
    info[6] = r->mTilde().value(); info[7] = r->t().value();
    if (bsk_out) for (size_t i = 0; i < r->baseBsk()->size(); i++) bsk_out[i] = r->baseBsk()->base()[i].value();
    return 0;
    REF_CATCH(-1)
}

// which: 0 fastbconvmTilde (q → Bsk∪m̃), 1 smMrq (Bsk∪m̃ → Bsk), 2 fastFloor (q∪Bsk → Bsk), 3 fastbconvSk (Bsk → q),
//        4 divideAndRoundqLastInplace (in-place on `in`), 5 divideAndRoundqLastNttInplace, 6 modTAndDivideqLastInplace
int ref_rns_op(void *h, int level, int which, uint64_t *in, uint64_t *out) {
    REF_TRY
    auto d = level_data(static_cast<RefCtx *>(h), level);
    auto r = d->rnsTool();
    switch (which) {
    case 0: r->fastbconvmTilde(ConstHostPointer<uint64_t>(in), HostPointer<uint64_t>(out)); break;
    case 1: r->smMrq(ConstHostPointer<uint64_t>(in), HostPointer<uint64_t>(out)); break;
    case 2: r->fastFloor(ConstHostPointer<uint64_t>(in), HostPointer<uint64_t>(out)); break;
    case 3: r->fastbconvSk(ConstHostPointer<uint64_t>(in), HostPointer<uint64_t>(out)); break;
    case 4: r->divideAndRoundqLastInplace(HostPointer<uint64_t>(in)); break;
    case 5: r->divideAndRoundqLastNttInplace(HostPointer<uint64_t>(in), d->smallNTTTables()); break;
    case 6: r->modTAndDivideqLastInplace(HostPointer<uint64_t>(in)); break;
    default: g_last_error = "bad op"; return -1;
    }
    return 0;
    REF_CATCH(-1)
}

// ---- Galois (src/utils/galois.cpp) ---------------------------------------------------------
uint32_t ref_galois_elt_from_step(void *h, int step) {
    REF_TRY
    return static_cast<RefCtx *>(h)->context->keyContextData()->galoisTool()->getEltFromStep(step);
    REF_CATCH(0)
}

int ref_apply_galois_poly(void *h, int level, const uint64_t *in, int n_limbs, uint32_t elt, uint64_t *out, int ntt_form) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto d = level_data(c, level);
    auto g = d->galoisTool();
    auto &q = d->parms().coeffModulus();
    for (int l = 0; l < n_limbs; l++) {
        if (ntt_form) g->applyGaloisNtt(ConstHostPointer<uint64_t>(in + l * c->n), elt, HostPointer<uint64_t>(out + l * c->n));
        else g->applyGalois(ConstHostPointer<uint64_t>(in + l * c->n), elt, q[l], HostPointer<uint64_t>(out + l * c->n));
    }
    return 0;
    REF_CATCH(-1)
}

// ---- keys (src/keygenerator.cpp) -----------------------------------------------------------
int ref_keygen(void *h) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    c->keygen = std::make_unique<KeyGenerator>(*c->context);
    c->keygen->createPublicKey(c->pk);
    c->encryptor = std::make_unique<Encryptor>(*c->context, c->pk);
    c->decryptor = std::make_unique<Decryptor>(*c->context, c->keygen->secretKey());
    if (c->scheme == SchemeType::ckks) c->ckks = std::make_unique<CKKSEncoder>(*c->context);
    else {
        try { c->batch = std::make_unique<BatchEncoder>(*c->context); } catch (...) {}
    }
    return 0;
    REF_CATCH(-1)
}

void *ref_relin_keys(void *h) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = std::make_unique<RefKeys>();
    RelinKeys rk; c->keygen->createRelinKeys(rk);
    k->keys = rk; k->galois = false;
    return k.release();
    REF_CATCH(nullptr)
}

void *ref_galois_keys_elts(void *h, const uint32_t *elts, int count) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = std::make_unique<RefKeys>();
    GaloisKeys gk; c->keygen->createGaloisKeys(std::vector<uint32_t>(elts, elts + count), gk);
    k->keys = gk; k->galois = true;
    return k.release();
    REF_CATCH(nullptr)
}

void *ref_galois_keys_steps(void *h, const int *steps, int count) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = std::make_unique<RefKeys>();
    GaloisKeys gk; c->keygen->createGaloisKeys(std::vector<int>(steps, steps + count), gk);
    k->keys = gk; k->galois = true;
    return k.release();
    REF_CATCH(nullptr)
}

void *ref_galois_keys_all(void *h) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = std::make_unique<RefKeys>();
    GaloisKeys gk; c->keygen->createGaloisKeys(gk);
    k->keys = gk; k->galois = true;
    return k.release();
    REF_CATCH(nullptr)
}

// Synthetic key-switching keys: container shaped exactly like KeyGenerator output
// (key level, NTT form, [n_keys][digits] of size-2 ciphertexts) but filled by the caller
// via ref_kswitch_key_set_words — used for throughput runs where key *values* are irrelevant.
void *ref_keys_synthetic(void *h, int galois, int n_slots) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = std::make_unique<RefKeys>();
    k->galois = galois != 0;
    k->keys.parmsID() = c->context->keyParmsID();
    k->keys.data().resize(n_slots);
    return k.release();
    REF_CATCH(nullptr)
}

int ref_kswitch_key_set_words(void *h, void *kh, int index, const uint64_t *words) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = static_cast<RefKeys *>(kh);
    auto kd = c->context->keyContextData();
    size_t digits = c->context->firstContextData()->parms().coeffModulus().size();
    size_t klimbs = kd->parms().coeffModulus().size();
    auto &vec = k->keys.data().at(index);
    vec.resize(digits);
    size_t per = 2 * klimbs * c->n;
    for (size_t j = 0; j < digits; j++) {
        Ciphertext &ct = vec[j].data();
        ct.resize(*c->context, c->context->keyParmsID(), 2);
        ct.isNttForm() = true;
        std::memcpy(ct.data(), words + j * per, per * sizeof(uint64_t));
    }
    return 0;
    REF_CATCH(-1)
}

void ref_keys_destroy(void *kh) { delete static_cast<RefKeys *>(kh); }

int ref_keys_num_slots(void *kh) { return (int)static_cast<RefKeys *>(kh)->keys.data().size(); }

// Returns number of digits at `index` (0 if no key there).
int ref_kswitch_key_digits(void *kh, int index) {
    auto k = static_cast<RefKeys *>(kh);
    if (index < 0 || (size_t)index >= k->keys.data().size()) return 0;
    return (int)k->keys.data()[index].size();
}

// Copies digits × 2 × key_limbs × N words.
int ref_kswitch_key_words(void *h, void *kh, int index, uint64_t *out) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto k = static_cast<RefKeys *>(kh);
    auto &vec = k->keys.data().at(index);
    size_t off = 0;
    for (auto &pk : vec) {
        const Ciphertext &ct = pk.data();
        size_t words = ct.size() * ct.coeffModulusSize() * c->n;
        std::memcpy(out + off, ct.data(), words * sizeof(uint64_t));
        off += words;
    }
    return 0;
    REF_CATCH(-1)
}

// ---- ciphertexts -------------------------------------------------------------------------
void *ref_ct_new(void *h, int level, int size, int is_ntt, double scale, uint64_t correction, const uint64_t *words) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto ct = std::make_unique<Ciphertext>();
    ct->resize(*c->context, c->chain.at(level), size);
    ct->isNttForm() = is_ntt != 0;
    ct->scale() = scale;
    ct->correctionFactor() = correction;
    if (words) std::memcpy(ct->data(), words, sizeof(uint64_t) * size * ct->coeffModulusSize() * c->n);
    return ct.release();
    REF_CATCH(nullptr)
}

void *ref_ct_clone(void *cth) { return new Ciphertext(*static_cast<Ciphertext *>(cth)); }
void ref_ct_destroy(void *cth) { delete static_cast<Ciphertext *>(cth); }

// meta: [level, size, limbs, is_ntt, correction_factor]; returns 0
int ref_ct_meta(void *h, void *cth, uint64_t *meta, double *scale) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto ct = static_cast<Ciphertext *>(cth);
    int level = -1;
    for (size_t i = 0; i < c->chain.size(); i++) if (c->chain[i] == ct->parmsID()) level = (int)i;
    meta[0] = (uint64_t)(int64_t)level; meta[1] = ct->size(); meta[2] = ct->coeffModulusSize();
    meta[3] = ct->isNttForm() ? 1 : 0; meta[4] = ct->correctionFactor();
    *scale = ct->scale();
    return 0;
    REF_CATCH(-1)
}

int ref_ct_words(void *h, void *cth, uint64_t *out) {
    auto c = static_cast<RefCtx *>(h);
    auto ct = static_cast<Ciphertext *>(cth);
    std::memcpy(out, ct->data(), sizeof(uint64_t) * ct->size() * ct->coeffModulusSize() * c->n);
    return 0;
}

// ---- evaluator ops (src/evaluator.cpp) ----------------------------------------------------
// op codes
enum {
    OP_MULTIPLY = 1, OP_SQUARE, OP_RELIN, OP_RESCALE_NEXT, OP_MODSWITCH_NEXT, OP_ROTATE_ROWS, OP_ROTATE_COLUMNS,
    OP_ROTATE_VECTOR, OP_CONJUGATE, OP_APPLY_GALOIS, OP_TO_NTT, OP_FROM_NTT, OP_ADD, OP_SUB, OP_NEGATE
};

int ref_eval(void *h, int op, void *cta, void *ctb, void *keys, int64_t arg) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto &ev = *c->evaluator;
    auto a = static_cast<Ciphertext *>(cta);
    auto b = static_cast<Ciphertext *>(ctb);
    auto k = static_cast<RefKeys *>(keys);
    switch (op) {
    case OP_MULTIPLY: ev.multiplyInplace(*a, *b); break;
    case OP_SQUARE: ev.squareInplace(*a); break;
    case OP_RELIN: ev.relinearizeInplace(*a, static_cast<const RelinKeys &>(k->keys)); break;
    case OP_RESCALE_NEXT: ev.rescaleToNextInplace(*a); break;
    case OP_MODSWITCH_NEXT: ev.modSwitchToNextInplace(*a); break;
    case OP_ROTATE_ROWS: ev.rotateRowsInplace(*a, (int)arg, static_cast<const GaloisKeys &>(k->keys)); break;
    case OP_ROTATE_COLUMNS: ev.rotateColumnsInplace(*a, static_cast<const GaloisKeys &>(k->keys)); break;
    case OP_ROTATE_VECTOR: ev.rotateVectorInplace(*a, (int)arg, static_cast<const GaloisKeys &>(k->keys)); break;
    case OP_CONJUGATE: ev.complexConjugateInplace(*a, static_cast<const GaloisKeys &>(k->keys)); break;
    case OP_APPLY_GALOIS: ev.applyGaloisInplace(*a, (uint32_t)arg, static_cast<const GaloisKeys &>(k->keys)); break;
    case OP_TO_NTT: ev.transformToNttInplace(*a); break;
    case OP_FROM_NTT: ev.transformFromNttInplace(*a); break;
    case OP_ADD: ev.addInplace(*a, *b); break;
    case OP_SUB: ev.subInplace(*a, *b); break;
    case OP_NEGATE: ev.negateInplace(*a); break;
    default: g_last_error = "bad op"; return -1;
    }
    return 0;
    REF_CATCH(-1)
}

// ---- plaintexts and ciphertext (x) plaintext ops (src/evaluator.cpp:1603-1985, 1987-2060) ------------------
// ntt_level < 0: coefficient-form plaintext modulo t (parms_id zero); otherwise `count` = L*N words in NTT form at that level.
void *ref_pt_new(void *h, const uint64_t *words, uint64_t count, int ntt_level, double scale) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto pt = std::make_unique<Plaintext>();
    pt->resize(count);
    std::memcpy(pt->data(), words, count * sizeof(uint64_t));
    if (ntt_level >= 0) pt->parmsID() = c->chain.at(ntt_level);
    pt->scale() = scale;
    return pt.release();
    REF_CATCH(nullptr)
}
void ref_pt_destroy(void *pth) { delete static_cast<Plaintext *>(pth); }
uint64_t ref_pt_count(void *pth) { return static_cast<Plaintext *>(pth)->coeffCount(); }
int ref_pt_words(void *pth, uint64_t *out) {
    auto pt = static_cast<Plaintext *>(pth);
    std::memcpy(out, pt->data(), pt->coeffCount() * sizeof(uint64_t));
    return 0;
}
// Evaluator::transformToNttInplace(Plaintext &, ParmsID)
int ref_pt_to_ntt(void *h, void *pth, int level) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    c->evaluator->transformToNttInplace(*static_cast<Plaintext *>(pth), c->chain.at(level));
    return 0;
    REF_CATCH(-1)
}
// op: 1 multiplyPlainInplace, 2 addPlainInplace, 3 subPlainInplace
int ref_eval_plain(void *h, int op, void *cta, void *pth) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto &ev = *c->evaluator;
    auto a = static_cast<Ciphertext *>(cta);
    auto p = static_cast<Plaintext *>(pth);
    switch (op) {
    case 1: ev.multiplyPlainInplace(*a, *p); break;
    case 2: ev.addPlainInplace(*a, *p); break;
    case 3: ev.subPlainInplace(*a, *p); break;
    default: g_last_error = "bad op"; return -1;
    }
    return 0;
    REF_CATCH(-1)
}

// ParmsID (4 words) of a level of the chain (-1: key level)
int ref_parms_id(void *h, int level, uint64_t *out) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    const ParmsID &id = level < 0 ? c->context->keyParmsID() : c->chain.at(level);
    for (int i = 0; i < 4; i++) out[i] = id[i];
    return 0;
    REF_CATCH(-1)
}

// The secret key as KeyGenerator stores it: [key_limbs][N] words, NTT form at the key level.
int ref_secret_key_words(void *h, uint64_t *out) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    const SecretKey &sk = c->keygen->secretKey();
    std::memcpy(out, sk.data().data(), sk.data().coeffCount() * sizeof(uint64_t));
    return (int)sk.data().coeffCount();
    REF_CATCH(-1)
}
// Decryptor::decrypt (src/decryptor.cpp): returns a new plaintext handle (ref_pt_*)
void *ref_decrypt_plain(void *h, void *cth) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto pt = std::make_unique<Plaintext>();
    c->decryptor->decrypt(*static_cast<Ciphertext *>(cth), *pt);
    return pt.release();
    REF_CATCH(nullptr)
}

// ---- encode / encrypt / decrypt (functional round trips; host-side, seeded PRNG) -------------
void *ref_encrypt_batch(void *h, const uint64_t *values, int count) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    Plaintext p; c->batch->encode(std::vector<uint64_t>(values, values + count), p);
    auto ct = std::make_unique<Ciphertext>();
    c->encryptor->encrypt(p, *ct);
    return ct.release();
    REF_CATCH(nullptr)
}

int ref_decrypt_batch(void *h, void *cth, uint64_t *out) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    Plaintext p; c->decryptor->decrypt(*static_cast<Ciphertext *>(cth), p);
    std::vector<uint64_t> v; c->batch->decode(p, v);
    std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
    return (int)v.size();
    REF_CATCH(-1)
}

void *ref_encrypt_ckks(void *h, const double *re_im, int count, double scale) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    std::vector<std::complex<double>> v(count);
    for (int i = 0; i < count; i++) v[i] = {re_im[2 * i], re_im[2 * i + 1]};
    Plaintext p; c->ckks->encode(v, c->context->firstParmsID(), scale, p);
    auto ct = std::make_unique<Ciphertext>();
    c->encryptor->encrypt(p, *ct);
    return ct.release();
    REF_CATCH(nullptr)
}

int ref_decrypt_ckks(void *h, void *cth, double *re_im) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    Plaintext p; c->decryptor->decrypt(*static_cast<Ciphertext *>(cth), p);
    std::vector<std::complex<double>> v; c->ckks->decode(p, v);
    for (size_t i = 0; i < v.size(); i++) { re_im[2 * i] = v[i].real(); re_im[2 * i + 1] = v[i].imag(); }
    return (int)v.size();
    REF_CATCH(-1)
}

// ---- CPU baseline timing (bench.py --impl reference / cpu_baseline) --------------------------
// Runs `reps` repetitions of an op sequence on clones of the given ciphertexts on `threads` host
// threads (one independent Evaluator per thread, as BASELINE.md §3 prescribes) and returns the
// wall-clock seconds.  seq: 0 = multiply+relinearize, 1 = rotate (rows for bfv/bgv, vector for ckks) by `arg`,
// 2 = multiply+relinearize+rescale+rotate, 3 = key-switch only (relinearize of a size-3 input)
double ref_time_ops(void *h, int seq, void *cta, void *ctb, void *rlk, void *gk, int arg, int reps, int threads) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto a0 = static_cast<Ciphertext *>(cta);
    auto b0 = static_cast<Ciphertext *>(ctb);
    auto rk = static_cast<RefKeys *>(rlk);
    auto g = static_cast<RefKeys *>(gk);
    std::atomic<int> failed{0};
    auto body = [&](int nreps) {
        try {
            Evaluator ev(*c->context);
            for (int r = 0; r < nreps; r++) {
                Ciphertext a(*a0);
                switch (seq) {
                case 0:
                    ev.multiplyInplace(a, *b0);
                    ev.relinearizeInplace(a, static_cast<const RelinKeys &>(rk->keys));
                    break;
                case 1:
                    if (c->scheme == SchemeType::ckks) ev.rotateVectorInplace(a, arg, static_cast<const GaloisKeys &>(g->keys));
                    else ev.rotateRowsInplace(a, arg, static_cast<const GaloisKeys &>(g->keys));
                    break;
                case 2:
                    ev.multiplyInplace(a, *b0);
                    ev.relinearizeInplace(a, static_cast<const RelinKeys &>(rk->keys));
                    if (c->scheme == SchemeType::ckks) {
                        ev.rescaleToNextInplace(a);
                        ev.rotateVectorInplace(a, arg, static_cast<const GaloisKeys &>(g->keys));
                    } else {
                        ev.rotateRowsInplace(a, arg, static_cast<const GaloisKeys &>(g->keys));
                    }
                    break;
                case 3:
                    ev.relinearizeInplace(a, static_cast<const RelinKeys &>(rk->keys));
                    break;
                }
            }
        } catch (...) { failed = 1; }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 1) body(reps);
    else {
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; t++) ts.emplace_back(body, reps);
        for (auto &t : ts) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (failed) { g_last_error = "op failed in timing loop"; return -1.0; }
    return std::chrono::duration<double>(t1 - t0).count();
    REF_CATCH(-1.0)
}

// NTT timing: `reps` forward(+inverse if both) transforms of one N-coefficient limb per thread.
double ref_time_ntt(uint64_t n, uint64_t prime, int inverse, int reps, int threads) {
    REF_TRY
    int logn = util::getPowerOfTwo(n);
    util::NTTTables t(logn, Modulus(prime));
    auto body = [&](int nreps) {
        std::vector<uint64_t> data(n);
        for (size_t i = 0; i < n; i++) data[i] = (i * 0x9E3779B97F4A7C15ull) % prime;
        for (int r = 0; r < nreps; r++) {
            if (inverse) util::inverseNttNegacyclicHarvey(HostPointer<uint64_t>(data.data()), t);
            else util::nttNegacyclicHarvey(HostPointer<uint64_t>(data.data()), t);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 1) body(reps);
    else {
        std::vector<std::thread> ts;
        for (int i = 0; i < threads; i++) ts.emplace_back(body, reps);
        for (auto &x : ts) x.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
    REF_CATCH(-1.0)
}

} // extern "C"

// ---- generator, samplers, encrypt-zero (src/randomgen.cpp, src/utils/rlwe.cpp, src/encryptor.cpp) ---------------------------
namespace {
// A Blake2xbPRNG that knows how many bytes have been drawn from it (the stream position the device samplers must reproduce).
struct CountingBlake2xb : Blake2xbPRNG {
    using Blake2xbPRNG::Blake2xbPRNG;
    uint64_t refills = 0;
    void refill_buffer() override { refills++; Blake2xbPRNG::refill_buffer(); }
    uint64_t position() const { return refills ? (refills - 1) * buffer_size_ + buffer_current_ : 0; }
};
// A generator that replays a given byte stream (zeros past its end): lets the tests put rejected draws where they want them.
struct ReplayPRNG : UniformRandomGenerator {
    const uint8_t *bytes;
    uint64_t len, refills = 0;
    ReplayPRNG(const uint8_t *b, uint64_t n) : UniformRandomGenerator(PRNGSeed{}), bytes(b), len(n) {}
    PRNGType type() const noexcept override { return PRNGType::unknown; }
    void refill_buffer() override {
        const uint64_t at = refills++ * buffer_size_;
        for (uint64_t i = 0; i < buffer_size_; i++) buffer_[i] = std::byte(at + i < len ? bytes[at + i] : 0);
    }
    uint64_t position() const { return refills ? (refills - 1) * buffer_size_ + buffer_current_ : 0; }
};
PRNGSeed seed_of(const uint64_t *seed8) {
    PRNGSeed seed;
    for (size_t i = 0; i < seed.size(); i++) seed[i] = seed8[i];
    return seed;
}
} // namespace
extern "C" {

// UniformRandomGenerator::generate on a fresh Blake2xbPRNG(seed)
int ref_prng_bytes(const uint64_t *seed8, uint64_t nbytes, uint8_t *out) {
    REF_TRY
    Blake2xbPRNG g(seed_of(seed8));
    g.generate(nbytes, reinterpret_cast<std::byte *>(out));
    return 0;
    REF_CATCH(-1)
}

// which: 0 samplePolyTernary, 1 samplePolyCbd, 2 samplePolyUniform over the base of `level`, from a fresh Blake2xbPRNG(seed)
// after discarding skip_bytes; *pos_out = stream position afterwards.
int ref_sample_poly(void *h, int level, int which, const uint64_t *seed8, uint64_t skip_bytes, uint64_t *out, uint64_t *pos_out) {
    REF_TRY
    auto d = level_data(static_cast<RefCtx *>(h), level);
    auto g = std::make_shared<CountingBlake2xb>(seed_of(seed8));
    std::vector<std::byte> skip(skip_bytes);
    if (skip_bytes) g->generate(skip_bytes, skip.data());
    switch (which) {
    case 0: util::samplePolyTernary(g, d->parms(), out); break;
    case 1: util::samplePolyCbd(g, d->parms(), out); break;
    case 2: util::samplePolyUniform(g, d->parms(), out); break;
    default: g_last_error = "bad sampler"; return -1;
    }
    if (pos_out) *pos_out = g->position();
    return 0;
    REF_CATCH(-1)
}

// The same samplers drawing from a caller-supplied byte stream.
int ref_sample_poly_replay(void *h, int level, int which, const uint8_t *stream, uint64_t stream_bytes, uint64_t skip_bytes,
                           uint64_t *out, uint64_t *pos_out) {
    REF_TRY
    auto d = level_data(static_cast<RefCtx *>(h), level);
    auto g = std::make_shared<ReplayPRNG>(stream, stream_bytes);
    std::vector<std::byte> skip(skip_bytes);
    if (skip_bytes) g->generate(skip_bytes, skip.data());
    switch (which) {
    case 0: util::samplePolyTernary(g, d->parms(), out); break;
    case 1: util::samplePolyCbd(g, d->parms(), out); break;
    case 2: util::samplePolyUniform(g, d->parms(), out); break;
    default: g_last_error = "bad sampler"; return -1;
    }
    if (pos_out) *pos_out = g->position();
    return 0;
    REF_CATCH(-1)
}

// Public key as KeyGenerator made it: [2][key limbs][N], NTT form. Returns the word count.
int64_t ref_public_key_words(void *h, uint64_t *out) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    const Ciphertext &k = c->pk.data();
    size_t words = k.size() * k.coeffModulusSize() * k.polyModulusDegree();
    if (out) std::memcpy(out, k.data(), words * sizeof(uint64_t));
    return (int64_t)words;
    REF_CATCH(-1)
}

// Encryptor::encryptZero / encryptZeroSymmetric at `level` (-1: key level) with the context's (seeded) generator factory.
void *ref_encrypt_zero(void *h, int level, int asymmetric) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    auto d = level_data(c, level);
    Encryptor enc(*c->context, c->pk, c->keygen->secretKey());
    auto ct = std::make_unique<Ciphertext>();
    if (asymmetric) enc.encryptZero(d->parmsID(), *ct);
    else enc.encryptZeroSymmetric(d->parmsID(), *ct);
    return ct.release();
    REF_CATCH(nullptr)
}

// Encryptor::encrypt / encryptSymmetric of a plaintext handle
void *ref_encrypt_plain(void *h, void *pth, int asymmetric) {
    REF_TRY
    auto c = static_cast<RefCtx *>(h);
    Encryptor enc(*c->context, c->pk, c->keygen->secretKey());
    auto ct = std::make_unique<Ciphertext>();
    if (asymmetric) enc.encrypt(*static_cast<Plaintext *>(pth), *ct);
    else enc.encryptSymmetric(*static_cast<Plaintext *>(pth), *ct);
    return ct.release();
    REF_CATCH(nullptr)
}

} // extern "C"
