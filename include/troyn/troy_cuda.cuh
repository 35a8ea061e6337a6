// troyn drop-in host layer: the reference's `troyn::` / `troy::*Cuda` C++ surface for the RNS polynomial hot path,
// implemented over the C ABI of libtroy_b200.so (include/troy_b200.h).
//
// A user of lightbulb128/troy includes this header INSTEAD of the reference's src/troy_cuda.cuh (same class names in
// the same namespaces, same method names, argument meaning and exception types) and links libtroy_b200.so plus the
// reference's own host (CPU) classes, which this header reuses exactly where the reference's CUDA classes do:
//   * KeyGeneratorCuda wraps the host KeyGenerator          (reference: src/keygenerator_cuda.cuh:36-60)
//   * encoders and plaintexts are host work that is out of the hot path: they delegate to the
//     host classes and move words over the C ABI's copy calls.
// Everything arithmetic on ciphertexts (Evaluator, Encryptor, Decryptor) runs in the CUDA library; nothing here computes on the CPU.
//
// Needs the reference's host headers on the include path (-I<troy>/src). Header-only; C++17; no CUDA headers needed.
//
// Reference interfaces mirrored (file:line under the reference tree):
//   KernelProvider            src/kernelprovider.cuh:18-85
//   EncryptionParametersCuda  src/encryptionparams_cuda.cuh
//   SEALContextCuda           src/context_cuda.cuh:146-186
//   CiphertextCuda            src/ciphertext_cuda.cuh:12-268
//   KSwitchKeysCuda / RelinKeysCuda / GaloisKeysCuda   src/kswitchkeys_cuda.cuh:31-366, relinkeys_cuda.cuh, galoiskeys_cuda.cuh
//   EvaluatorCuda             src/evaluator_cuda.cuh:17-440
//   KeyGeneratorCuda / EncryptorCuda / DecryptorCuda    src/keygenerator_cuda.cuh, encryptor_cuda.cuh, decryptor_cuda.cuh
#pragma once

#include "troy_cpu.h" // the reference's host classes (troy::)

#include "../troy_b200.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <istream>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace troy {

namespace b200 {

// C ABI status -> the reference's exception types (SURVEY.md §8b "Errors")
inline void check(int rc) {
    if (rc == TROYB_OK) return;
    const std::string msg = troyb_last_error();
    switch (rc) {
    case TROYB_E_INVALID_ARGUMENT:
    case TROYB_E_NOT_INITIALIZED: throw std::invalid_argument(msg);
    case TROYB_E_LOGIC: throw std::logic_error(msg);
    case TROYB_E_OUT_OF_RANGE: throw std::out_of_range(msg);
    default: throw std::runtime_error("CUDA error.");
    }
}

// The CUDA stream every call of the calling host thread is enqueued on (allocations, copies and kernels: one stream keeps them
// ordered without host synchronisation). Default: the legacy default stream, like the reference (which has no stream
// arguments at all). troyn::b200::setStream(s) / StreamGuard select another one, e.g. one stream per worker thread.
inline void *&stream_slot() { static thread_local void *s = nullptr; return s; }
inline void *stream() { return stream_slot(); }
inline void setStream(void *cuda_stream) { stream_slot() = cuda_stream; }
inline void synchronize() { check(troyb_stream_synchronize(stream())); }
struct StreamGuard {
    explicit StreamGuard(void *s) : prev_(stream()) { setStream(s); }
    ~StreamGuard() { setStream(prev_); }
    StreamGuard(const StreamGuard &) = delete;
    StreamGuard &operator=(const StreamGuard &) = delete;
private:
    void *prev_;
};

// Owning device buffer of 64-bit words (util::DeviceArray<uint64_t>, src/utils/devicearray.cuh:90-200)
class DeviceWords {
public:
    DeviceWords() = default;
    explicit DeviceWords(std::size_t words) { allocate(words); }
    DeviceWords(const DeviceWords &o) { *this = o; }
    DeviceWords(DeviceWords &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DeviceWords &operator=(const DeviceWords &o) {
        if (this == &o) return *this;
        allocate(o.n_);
        if (n_) check(troyb_copy_on_device(p_, o.p_, n_ * 8, b200::stream()));
        return *this;
    }
    DeviceWords &operator=(DeviceWords &&o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DeviceWords() { release(); }
    void allocate(std::size_t words) {
        if (words == n_ && p_) return;
        release();
        if (words) { void *p = nullptr; check(troyb_malloc(&p, words * 8, b200::stream())); p_ = static_cast<std::uint64_t *>(p); n_ = words; }
    }
    void upload(const std::uint64_t *h, std::size_t words, std::size_t at = 0) {
        if (words) { check(troyb_copy_to_device(p_ + at, h, words * 8, b200::stream()));  }
    }
    void download(std::uint64_t *h, std::size_t words, std::size_t at = 0) const {
        if (words) check(troyb_copy_to_host(h, p_ + at, words * 8, b200::stream()));
    }
    std::uint64_t *get() const { return p_; }
    std::size_t size() const { return n_; }

private:
    void release() { if (p_) troyb_free(p_, b200::stream()); p_ = nullptr; n_ = 0; }
    std::uint64_t *p_ = nullptr;
    std::size_t n_ = 0;
};

} // namespace b200

class KernelProvider {
public:
    static void initialize(int device = 0) { b200::check(troyb_initialize(device)); }
    static bool initialized() { return troyb_initialized() != 0; }
};

class EncryptionParametersCuda {
public:
    EncryptionParametersCuda(SchemeType scheme = SchemeType::none) : host_(scheme) {}
    EncryptionParametersCuda(const EncryptionParameters &host) : host_(host) {}
    void setPolyModulusDegree(std::size_t n) { host_.setPolyModulusDegree(n); }
    void setCoeffModulus(const std::vector<Modulus> &q) { host_.setCoeffModulus(q); }
    void setPlainModulus(const Modulus &t) { host_.setPlainModulus(t); }
    void setPlainModulus(std::uint64_t t) { host_.setPlainModulus(t); }
    void setRandomGenerator(std::shared_ptr<UniformRandomGeneratorFactory> f) { host_.setRandomGenerator(std::move(f)); }
    SchemeType scheme() const { return host_.scheme(); }
    std::size_t polyModulusDegree() const { return host_.polyModulusDegree(); }
    const std::vector<Modulus> &coeffModulus() const { return host_.coeffModulus(); }
    const Modulus &plainModulus() const { return host_.plainModulus(); }
    const ParmsID &parmsID() const { return host_.parmsID(); }
    const EncryptionParameters &host() const { return host_; }
    operator const EncryptionParameters &() const { return host_; }

private:
    EncryptionParameters host_;
};

class SEALContextCuda {
public:
    SEALContextCuda(const EncryptionParametersCuda &parms, bool expand_mod_chain = true, SecurityLevel sec = SecurityLevel::tc128)
        : host_(std::make_shared<SEALContext>(parms.host(), expand_mod_chain, sec)) {
        if (!host_->parametersSet()) throw std::invalid_argument("encryption parameters are not set correctly");
        const auto &p = host_->keyContextData()->parms();
        std::vector<std::uint64_t> q;
        for (const auto &m : p.coeffModulus()) q.push_back(m.value());
        const int scheme = p.scheme() == SchemeType::bfv ? TROYB_SCHEME_BFV : p.scheme() == SchemeType::ckks ? TROYB_SCHEME_CKKS : TROYB_SCHEME_BGV;
        troyb_context *h = nullptr;
        b200::check(troyb_context_create(scheme, p.polyModulusDegree(), q.data(), (int)q.size(),
                                         p.scheme() == SchemeType::ckks ? 0 : p.plainModulus().value(), expand_mod_chain ? 1 : 0, &h));
        handle_ = std::shared_ptr<troyb_context>(h, troyb_context_destroy);
        for (auto cd = host_->firstContextData(); cd; cd = cd->nextContextData()) levels_.push_back(cd->parmsID());
    }
    // chain index of a parms_id (the C ABI's `level`); throws like the reference when the id is not in the chain
    int levelOf(const ParmsID &id) const {
        for (std::size_t i = 0; i < levels_.size(); i++) if (levels_[i] == id) return (int)i;
        throw std::invalid_argument("encrypted is not valid for encryption parameters");
    }
    const ParmsID &parmsIDOf(int level) const { return levels_.at((std::size_t)level); }
    std::size_t chainLength() const { return levels_.size(); }
    auto getContextData(const ParmsID &id) const { return host_->getContextData(id); }
    auto keyContextData() const { return host_->keyContextData(); }
    auto firstContextData() const { return host_->firstContextData(); }
    auto lastContextData() const { return host_->lastContextData(); }
    const ParmsID &keyParmsID() const { return host_->keyParmsID(); }
    const ParmsID &firstParmsID() const { return host_->firstParmsID(); }
    const ParmsID &lastParmsID() const { return host_->lastParmsID(); }
    bool parametersSet() const { return host_->parametersSet(); }
    bool using_keyswitching() const { return host_->using_keyswitching(); }
    const SEALContext &host() const { return *host_; }
    operator const SEALContext &() const { return *host_; } // host encoders / key generator take the host context
    troyb_context *handle() const { return handle_.get(); }

private:
    std::shared_ptr<SEALContext> host_;
    std::shared_ptr<troyb_context> handle_;
    std::vector<ParmsID> levels_;
};

// Host-side objects that never enter the hot path keep their host representation (the reference's PlaintextCuda /
// PublicKeyCuda / SecretKeyCuda are device mirrors of data that only encode / encrypt / decrypt touch).
class PlaintextCuda : public Plaintext {
public:
    using Plaintext::Plaintext;
    PlaintextCuda() = default;
    PlaintextCuda(const Plaintext &p) : Plaintext(p) {}
    PlaintextCuda &operator=(const std::string &hex_poly) { Plaintext::operator=(hex_poly); return *this; }
    PlaintextCuda &operator=(std::uint64_t c) { Plaintext::operator=(c); return *this; }
    const Plaintext &cpu() const { return *this; }
    // PlaintextCuda::save / load  src/plaintext_cuda.cu:7-27: parms_id, coeff_count, scale, data_size, data
    void save(std::ostream &stream) const {
        const std::size_t count = coeffCount();
        stream.write(reinterpret_cast<const char *>(&parmsID()), sizeof(ParmsID));
        stream.write(reinterpret_cast<const char *>(&count), sizeof(count));
        stream.write(reinterpret_cast<const char *>(&scale()), sizeof(double));
        stream.write(reinterpret_cast<const char *>(&count), sizeof(count));
        stream.write(reinterpret_cast<const char *>(data()), count * sizeof(std::uint64_t));
    }
    void load(std::istream &stream) {
        ParmsID id;
        std::size_t count = 0, data_size = 0;
        double sc = 1.0;
        stream.read(reinterpret_cast<char *>(&id), sizeof(ParmsID));
        stream.read(reinterpret_cast<char *>(&count), sizeof(count));
        stream.read(reinterpret_cast<char *>(&sc), sizeof(double));
        stream.read(reinterpret_cast<char *>(&data_size), sizeof(data_size));
        std::vector<std::uint64_t> w(data_size);
        stream.read(reinterpret_cast<char *>(w.data()), data_size * sizeof(std::uint64_t));
        parmsID() = parmsIDZero;
        resize(count);
        std::memcpy(data(), w.data(), std::min(count, data_size) * sizeof(std::uint64_t));
        parmsID() = id;
        scale() = sc;
    }
};
class PublicKeyCuda : public PublicKey {
public:
    PublicKeyCuda() = default;
    PublicKeyCuda(const PublicKey &k) : PublicKey(k) {}
    // PublicKeyCuda::save / load: the key's ciphertext in CiphertextCuda's format (src/publickey_cuda.cuh, src/ciphertext_cuda.cu:16-43)
    inline void save(std::ostream &stream) const;
    inline void load(std::istream &stream);
};
class SecretKeyCuda : public SecretKey {
public:
    SecretKeyCuda() = default;
    SecretKeyCuda(const SecretKey &k) : SecretKey(k) {}
    // SecretKeyCuda::save / load  src/secretkey_cuda.cuh:292-297: the key's plaintext in PlaintextCuda's format
    void save(std::ostream &stream) const { PlaintextCuda(data()).save(stream); }
    void load(std::istream &stream) { PlaintextCuda p; p.load(stream); data() = p; }
};
// Encoders are host work in this build: the host classes, constructed from the device context.
class CKKSEncoderCuda : public CKKSEncoder {
public:
    CKKSEncoderCuda(const SEALContextCuda &context) : CKKSEncoder(context.host()), host_context_(context.host()) {}

    // encodePolynomial / decodePolynomial exist only in the reference's CUDA encoder (src/ckks_cuda.cu:455-571, 983-1055):
    // coefficient i of the plaintext polynomial is round(values[i] * scale). Host work here, like the rest of the encoder.
    void encodePolynomial(const std::vector<double> &values, ParmsID parms_id, double scale, PlaintextCuda &destination) {
        auto cd = host_context_.getContextData(parms_id);
        if (!cd) throw std::invalid_argument("parms_id is not valid for encryption parameters");
        const auto &q = cd->parms().coeffModulus();
        const std::size_t n = cd->parms().polyModulusDegree(), L = q.size();
        if (values.size() > n) throw std::invalid_argument("values_size is too large");
        if (scale <= 0 || (int)std::log2(scale) + 1 >= cd->totalCoeffModulusBitCount()) throw std::invalid_argument("scale out of bounds");
        double max_coeff = 0;
        for (double v : values) max_coeff = std::max(max_coeff, std::fabs(v * scale));
        const int bits = (int)std::ceil(std::log2(std::max(max_coeff, 1.0))) + 1;
        if (bits >= cd->totalCoeffModulusBitCount()) throw std::invalid_argument("encoded values are too large");
        if (bits > 64) throw std::logic_error("encodePolynomial: coefficients above 64 bits are outside the ported path");
        destination.parmsID() = parmsIDZero;
        destination.resize(n * L);
        for (std::size_t i = 0; i < n; i++) {
            const double c = i < values.size() ? std::round(values[i] * scale) : 0.0;
            const bool neg = c < 0;
            const std::uint64_t u = (std::uint64_t)std::fabs(c);
            for (std::size_t j = 0; j < L; j++) {
                const std::uint64_t r = util::barrettReduce64(u, q[j]);
                destination.data()[j * n + i] = neg ? util::negateUintMod(r, q[j]) : r;
            }
        }
        util::nttNegacyclicHarvey(util::HostPointer<std::uint64_t>(destination.data()), L, cd->smallNTTTables());
        destination.parmsID() = parms_id;
        destination.scale() = scale;
    }
    void encodePolynomial(const std::vector<double> &values, double scale, PlaintextCuda &destination) {
        encodePolynomial(values, host_context_.firstParmsID(), scale, destination);
    }
    void decodePolynomial(const PlaintextCuda &plain, std::vector<double> &destination) {
        if (!plain.isNttForm()) throw std::invalid_argument("plain is not in NTT form");
        auto cd = host_context_.getContextData(plain.parmsID());
        if (!cd) throw std::invalid_argument("plain is not valid for encryption parameters");
        const std::size_t n = cd->parms().polyModulusDegree(), L = cd->parms().coeffModulus().size();
        if (plain.scale() <= 0 || (int)std::log2(plain.scale()) >= cd->totalCoeffModulusBitCount()) throw std::invalid_argument("scale out of bounds");
        std::vector<std::uint64_t> w(plain.data(), plain.data() + n * L);
        util::inverseNttNegacyclicHarvey(util::HostPointer<std::uint64_t>(w.data()), L, cd->smallNTTTables());
        cd->rnsTool()->baseq()->composeArray(w.data(), n);
        const std::uint64_t *Q = cd->totalCoeffModulus(), *half = cd->upperHalfThreshold();
        const double inv_scale = 1.0 / plain.scale(), two64 = std::pow(2.0, 64);
        destination.assign(n, 0.0);
        for (std::size_t i = 0; i < n; i++) { // gDecodeInternal, src/ckks_cuda.cu:793-831
            const std::uint64_t *x = w.data() + i * L;
            double r = 0, f = inv_scale;
            if (util::isGreaterThanOrEqualUint(x, half, L)) {
                for (std::size_t j = 0; j < L; j++, f *= two64) {
                    if (x[j] > Q[j]) r += (double)(x[j] - Q[j]) * f;
                    else r -= (double)(Q[j] - x[j]) * f;
                }
            } else {
                for (std::size_t j = 0; j < L; j++, f *= two64) r += (double)x[j] * f;
            }
            destination[i] = r;
        }
    }

private:
    SEALContext host_context_;
};
class BatchEncoderCuda : public BatchEncoder {
public:
    BatchEncoderCuda(const SEALContextCuda &context)
        : BatchEncoder(context.host()), t_(context.firstContextData()->parms().plainModulus().value()) {}
    // encodePolynomial / decodePolynomial  src/batchencoder_cuda.cu:124-171, 267-286: coefficients modulo t, no slot transform
    void encodePolynomial(const std::vector<std::uint64_t> &values, PlaintextCuda &destination) const {
        if (values.size() > slotCount()) throw std::invalid_argument("values_matrix size is too large");
        destination.parmsID() = parmsIDZero;
        destination.resize(values.size());
        for (std::size_t i = 0; i < values.size(); i++) destination.data()[i] = values[i] % t_;
    }
    void encodePolynomial(const std::vector<std::int64_t> &values, PlaintextCuda &destination) const {
        if (values.size() > slotCount()) throw std::invalid_argument("values_matrix size is too large");
        destination.parmsID() = parmsIDZero;
        destination.resize(slotCount());
        for (std::size_t i = 0; i < slotCount(); i++) {
            const std::int64_t v = i < values.size() ? values[i] : 0;
            destination.data()[i] = v < 0 ? t_ - (static_cast<std::uint64_t>(-v) % t_) : static_cast<std::uint64_t>(v) % t_;
        }
    }
    void decodePolynomial(const PlaintextCuda &plain, std::vector<std::uint64_t> &destination) const {
        destination.assign(plain.data(), plain.data() + std::min(plain.coeffCount(), slotCount()));
    }
    void decodePolynomial(const PlaintextCuda &plain, std::vector<std::int64_t> &destination) const {
        destination.assign(slotCount(), 0);
        for (std::size_t i = 0; i < std::min(plain.coeffCount(), slotCount()); i++) {
            const std::uint64_t r = plain.data()[i];
            destination[i] = r > (t_ >> 1) ? static_cast<std::int64_t>(r) - static_cast<std::int64_t>(t_) : static_cast<std::int64_t>(r);
        }
    }

private:
    std::uint64_t t_;
};

class EvaluatorCuda;

class CiphertextCuda {
public:
    CiphertextCuda() = default;
    // upload (CiphertextCuda(const Ciphertext&), src/ciphertext_cuda.cuh:20-36)
    CiphertextCuda(const Ciphertext &h)
        : size_(h.size()), n_(h.polyModulusDegree()), limbs_(h.coeffModulusSize()), parms_id_(h.parmsID()), is_ntt_(h.isNttForm()),
          scale_(h.scale()), correction_(h.correctionFactor()), data_(h.size() * h.polyModulusDegree() * h.coeffModulusSize()) {
        data_.upload(h.data(), data_.size());
    }
    // capacity-only constructors (src/ciphertext_cuda.cuh:102-123): parameters are recorded, size stays 0
    explicit CiphertextCuda(const SEALContextCuda &context) : CiphertextCuda(context, context.firstParmsID()) {}
    explicit CiphertextCuda(const SEALContextCuda &context, const ParmsID &parms_id, std::size_t = 2) {
        auto cd = context.getContextData(parms_id);
        if (!cd) throw std::invalid_argument("parms_id is not valid for encryption parameters");
        parms_id_ = parms_id;
        n_ = cd->parms().polyModulusDegree();
        limbs_ = cd->parms().coeffModulus().size();
    }
    // download (cpu / toHost, src/ciphertext_cuda.cuh:38-51); Ciphertext befriends CiphertextCuda (src/ciphertext.h:55)
    Ciphertext cpu() const {
        Ciphertext h;
        h.parms_id_ = parms_id_;
        h.is_ntt_form_ = is_ntt_;
        h.size_ = size_;
        h.poly_modulus_degree_ = n_;
        h.coeff_modulus_size_ = limbs_;
        h.scale_ = scale_;
        h.correction_factor_ = correction_;
        h.data_ = util::HostDynamicArray<std::uint64_t>(size_ * n_ * limbs_);
        data_.download(h.data(), size_ * n_ * limbs_);
        return h;
    }
    Ciphertext toHost() const { return cpu(); }
    bool isTransparent() const { return cpu().isTransparent(); }

    std::size_t size() const { return size_; }
    std::size_t polyModulusDegree() const { return n_; }
    std::size_t coeffModulusSize() const { return limbs_; }
    std::size_t polyCoeffSize() const { return n_ * limbs_; }
    ParmsID &parmsID() { return parms_id_; }
    const ParmsID &parmsID() const { return parms_id_; }
    bool &isNttForm() { return is_ntt_; }
    bool isNttForm() const { return is_ntt_; }
    double &scale() { return scale_; }
    const double &scale() const { return scale_; }
    std::uint64_t &correctionFactor() { return correction_; }
    const std::uint64_t &correctionFactor() const { return correction_; }
    // CiphertextCuda::seed  src/ciphertext_cuda.cuh:189: nonzero for a fresh symmetric encryption whose c1 is the cuRAND
    // expansion of the seed (save then stores c0 only). Any mutable access to the words clears it, which covers the
    // `seed() = 0` of every evaluator method of the reference.
    std::uint64_t seed() const { return seed_; }
    std::uint64_t &seed() { return seed_; }
    std::uint64_t *data() { seed_ = 0; return data_.get(); }
    const std::uint64_t *data() const { return data_.get(); }
    std::uint64_t *data(std::size_t poly) { seed_ = 0; return data_.get() + poly * polyCoeffSize(); }
    const std::uint64_t *data(std::size_t poly) const { return data_.get() + poly * polyCoeffSize(); }

    // resize / reserve keep the first min(old, new) polynomials (src/ciphertext_cuda.cuh:68-152)
    void resize(const SEALContextCuda &context, const ParmsID &parms_id, std::size_t size) {
        auto cd = context.getContextData(parms_id);
        if (!cd) throw std::invalid_argument("parms_id is not valid for encryption parameters");
        const auto &p = cd->parms();
        parms_id_ = parms_id;
        resizeInternal(size, p.polyModulusDegree(), p.coeffModulus().size());
    }
    void resize(std::size_t size) { resizeInternal(size, n_, limbs_); }
    void release() { *this = CiphertextCuda(); }

    // Wire format of the reference (src/ciphertext_cuda.cu:16-194): parms_id, is_ntt_form, size, N, limbs, scale,
    // correction_factor, seed, terms, data_size, data. A seeded ciphertext (a fresh symmetric encryption) stores c0 only; its
    // c1 is re-expanded from the seed on load(stream, context) with the reference's cuRAND sampler semantics.
    void save(std::ostream &stream) const {
        if (seed_ != 0 && size_ > 2) throw std::invalid_argument("Seed exists but size is not 2.");
        std::vector<std::uint64_t> w((seed_ != 0 ? 1 : size_) * n_ * limbs_);
        data_.download(w.data(), w.size());
        saveHeader(stream, false);
        const std::size_t count = w.size();
        stream.write(reinterpret_cast<const char *>(&count), sizeof(count));
        stream.write(reinterpret_cast<const char *>(w.data()), count * 8);
    }
    void load(std::istream &stream) { loadImpl(stream, nullptr); }
    void load(std::istream &stream, const SEALContextCuda &context) { loadImpl(stream, &context); }
    // saveTerms / loadTerms (:45-77, :101-140): only the listed coefficients of poly 0 travel (coefficient form), poly 1.. whole
    inline void saveTerms(std::ostream &stream, EvaluatorCuda &evaluator, const std::vector<std::size_t> &term_ids) const;
    inline void loadTerms(std::istream &stream, EvaluatorCuda &evaluator, const std::vector<std::size_t> &term_ids);

    // used by the evaluator when an operation changes the limb count
    void adopt(b200::DeviceWords &&data, std::size_t size, std::size_t n, std::size_t limbs) {
        data_ = std::move(data); size_ = size; n_ = n; limbs_ = limbs; seed_ = 0;
    }

private:
    void loadImpl(std::istream &stream, const SEALContextCuda *context) {
        std::uint64_t seed = 0;
        if (loadHeader(stream, &seed)) throw std::invalid_argument("Trying to load a termed ciphertext, but indices is not specified");
        if (seed != 0 && !context) throw std::invalid_argument("seed is not zero.");
        if (seed != 0 && size_ > 2) throw std::invalid_argument("Seed exists but size is not 2.");
        std::size_t count = 0;
        stream.read(reinterpret_cast<char *>(&count), sizeof(count));
        std::vector<std::uint64_t> w(count);
        stream.read(reinterpret_cast<char *>(w.data()), count * 8);
        const std::size_t stored = (seed != 0 ? 1 : size_) * n_ * limbs_;
        if (count < stored) throw std::invalid_argument("ciphertext data is too short");
        data_ = b200::DeviceWords(size_ * n_ * limbs_);
        data_.upload(w.data(), stored);
        if (seed != 0) { // src/ciphertext_cuda.cu:158-190: c1 from the seed, inverted for the coefficient-form schemes (the
                         // reference's loader inverts for BFV only, which leaves its BGV seeded ciphertexts undecryptable)
            const bool key_level = parms_id_ == context->keyParmsID() && context->using_keyswitching();
            const int level = key_level ? -1 : context->levelOf(parms_id_);
            b200::DeviceWords d_seed(1);
            d_seed.upload(&seed, 1);
            std::uint64_t *c1 = data_.get() + n_ * limbs_;
            b200::check(troyb_expand_seeded_poly(context->handle(), level, d_seed.get(), c1, n_ * limbs_, 1, b200::stream()));
            if (context->keyContextData()->parms().scheme() != SchemeType::ckks) {
                if (key_level) b200::check(troyb_ntt_inverse_key(context->handle(), c1, 1, b200::stream()));
                else b200::check(troyb_ntt_inverse(context->handle(), level, c1, 1, b200::stream()));
            }
            b200::synchronize(); // `seed` and `w` are host temporaries of this call
        }
        seed_ = 0; // as the reference's loaders leave it
    }
    void saveHeader(std::ostream &stream, bool terms) const {
        const std::uint64_t seed = seed_;
        stream.write(reinterpret_cast<const char *>(&parms_id_), sizeof(ParmsID));
        stream.write(reinterpret_cast<const char *>(&is_ntt_), sizeof(bool));
        stream.write(reinterpret_cast<const char *>(&size_), sizeof(std::size_t));
        stream.write(reinterpret_cast<const char *>(&n_), sizeof(std::size_t));
        stream.write(reinterpret_cast<const char *>(&limbs_), sizeof(std::size_t));
        stream.write(reinterpret_cast<const char *>(&scale_), sizeof(double));
        stream.write(reinterpret_cast<const char *>(&correction_), sizeof(std::uint64_t));
        stream.write(reinterpret_cast<const char *>(&seed), sizeof(std::uint64_t));
        stream.write(reinterpret_cast<const char *>(&terms), sizeof(bool));
    }
    bool loadHeader(std::istream &stream, std::uint64_t *seed_out = nullptr) { // returns the `terms` flag
        std::uint64_t seed = 0;
        bool terms = false;
        stream.read(reinterpret_cast<char *>(&parms_id_), sizeof(ParmsID));
        stream.read(reinterpret_cast<char *>(&is_ntt_), sizeof(bool));
        stream.read(reinterpret_cast<char *>(&size_), sizeof(std::size_t));
        stream.read(reinterpret_cast<char *>(&n_), sizeof(std::size_t));
        stream.read(reinterpret_cast<char *>(&limbs_), sizeof(std::size_t));
        stream.read(reinterpret_cast<char *>(&scale_), sizeof(double));
        stream.read(reinterpret_cast<char *>(&correction_), sizeof(std::uint64_t));
        stream.read(reinterpret_cast<char *>(&seed), sizeof(std::uint64_t));
        stream.read(reinterpret_cast<char *>(&terms), sizeof(bool));
        if (seed_out) *seed_out = seed;
        else if (seed != 0) throw std::invalid_argument("seed is not zero.");
        return terms;
    }
    void resizeInternal(std::size_t size, std::size_t n, std::size_t limbs) {
        if ((size < 2 && size != 0) || size > 6) throw std::invalid_argument("invalid size");
        if (n == n_ && limbs == limbs_ && size * n * limbs <= data_.size()) { size_ = size; return; }
        b200::DeviceWords fresh(size * n * limbs);
        if (n == n_ && limbs == limbs_) {
            const std::size_t keep = std::min(size_, size) * n * limbs;
            if (keep) b200::check(troyb_copy_on_device(fresh.get(), data_.get(), keep * 8, b200::stream()));
        }
        data_ = std::move(fresh);
        size_ = size; n_ = n; limbs_ = limbs;
    }
    std::size_t size_ = 0, n_ = 0, limbs_ = 0;
    ParmsID parms_id_ = parmsIDZero;
    bool is_ntt_ = false;
    double scale_ = 1.0;
    std::uint64_t correction_ = 1;
    std::uint64_t seed_ = 0;
    b200::DeviceWords data_;
};

// LWECiphertextCuda (src/ciphertext_cuda.cuh:270-330): one coefficient of an RLWE ciphertext as an LWE sample - c1 (a whole
// polynomial, [limbs][N], coefficient form) and c0 (one word per limb).
class LWECiphertextCuda {
public:
    ParmsID &parmsID() noexcept { return parms_id_; }
    const ParmsID &parmsID() const noexcept { return parms_id_; }
    bool isNttForm() const noexcept { return false; }
    double &scale() noexcept { return scale_; }
    double scale() const noexcept { return scale_; }
    std::uint64_t correctionFactor() const noexcept { return correction_; }
    std::uint64_t &correctionFactor() noexcept { return correction_; }
    std::size_t polyModulusDegree() const noexcept { return n_; }
    std::size_t coeffModulusSize() const noexcept { return limbs_; }
    std::uint64_t *c1() { return c1_.get(); }
    const std::uint64_t *c1() const { return c1_.get(); }
    std::uint64_t *c0() { return c0_.get(); }
    const std::uint64_t *c0() const { return c0_.get(); }
    static LWECiphertextCuda fromMembers(ParmsID parms_id, std::size_t n, std::size_t limbs, double scale, std::uint64_t correction,
                                         b200::DeviceWords &&c1, b200::DeviceWords &&c0) {
        LWECiphertextCuda ct;
        ct.parms_id_ = parms_id; ct.n_ = n; ct.limbs_ = limbs; ct.scale_ = scale; ct.correction_ = correction;
        ct.c1_ = std::move(c1); ct.c0_ = std::move(c0);
        return ct;
    }

private:
    ParmsID parms_id_ = parmsIDZero;
    std::size_t n_ = 0, limbs_ = 0;
    double scale_ = 1.0;
    std::uint64_t correction_ = 1;
    b200::DeviceWords c1_, c0_;
};

// keys()[index] = one key-switching key as device words [digit][2][key limbs][N]   (src/kswitchkeys_cuda.cuh:43-56,366)
class KSwitchKeysCuda {
public:
    KSwitchKeysCuda() = default;
    KSwitchKeysCuda(const KSwitchKeys &h) { assign(h); }
    void assign(const KSwitchKeys &h) {
        parms_id_ = h.parmsID();
        keys_.clear();
        keys_.resize(h.data().size());
        for (std::size_t idx = 0; idx < h.data().size(); idx++) {
            const auto &digits = h.data()[idx];
            if (digits.empty()) continue;
            const Ciphertext &first = digits[0].data();
            n_ = first.polyModulusDegree();
            limbs_ = first.coeffModulusSize();
            const std::size_t per = first.size() * n_ * limbs_;
            keys_[idx].allocate(per * digits.size());
            for (std::size_t j = 0; j < digits.size(); j++) keys_[idx].upload(digits[j].data().data(), per, j * per);
        }
    }
    std::size_t size() const { std::size_t c = 0; for (auto &k : keys_) c += k.size() ? 1 : 0; return c; }
    const ParmsID &parmsID() const { return parms_id_; }
    const std::uint64_t *key(std::size_t index) const {
        if (index >= keys_.size() || !keys_[index].size()) throw std::invalid_argument("keyswitching key does not exist");
        return keys_[index].get();
    }
    bool hasIndex(std::size_t index) const { return index < keys_.size() && keys_[index].size() != 0; }
    // KSwitchKeysCuda::save / load  src/kswitchkeys_cuda.cuh:330-353: parms_id, n, then per index m and m key digits, each one
    // a size-2 NTT-form ciphertext at the key level in CiphertextCuda's format
    void save(std::ostream &stream) const {
        stream.write(reinterpret_cast<const char *>(&parms_id_), sizeof(ParmsID));
        const std::size_t n = keys_.size(), per = 2 * n_ * limbs_;
        stream.write(reinterpret_cast<const char *>(&n), sizeof(n));
        std::vector<std::uint64_t> w(per);
        for (const auto &k : keys_) {
            const std::size_t m = per ? k.size() / per : 0;
            stream.write(reinterpret_cast<const char *>(&m), sizeof(m));
            for (std::size_t j = 0; j < m; j++) {
                k.download(w.data(), per, j * per);
                b200::synchronize();
                const bool ntt = true, terms = false;
                const std::size_t size = 2;
                const double scale = 1.0;
                const std::uint64_t one = 1, zero = 0;
                stream.write(reinterpret_cast<const char *>(&parms_id_), sizeof(ParmsID));
                stream.write(reinterpret_cast<const char *>(&ntt), sizeof(bool));
                stream.write(reinterpret_cast<const char *>(&size), sizeof(size));
                stream.write(reinterpret_cast<const char *>(&n_), sizeof(n_));
                stream.write(reinterpret_cast<const char *>(&limbs_), sizeof(limbs_));
                stream.write(reinterpret_cast<const char *>(&scale), sizeof(scale));
                stream.write(reinterpret_cast<const char *>(&one), sizeof(one));
                stream.write(reinterpret_cast<const char *>(&zero), sizeof(zero));
                stream.write(reinterpret_cast<const char *>(&terms), sizeof(bool));
                stream.write(reinterpret_cast<const char *>(&per), sizeof(per));
                stream.write(reinterpret_cast<const char *>(w.data()), per * 8);
            }
        }
    }
    void load(std::istream &stream) {
        stream.read(reinterpret_cast<char *>(&parms_id_), sizeof(ParmsID));
        std::size_t n = 0;
        stream.read(reinterpret_cast<char *>(&n), sizeof(n));
        keys_.clear();
        keys_.resize(n);
        for (std::size_t i = 0; i < n; i++) {
            std::size_t m = 0;
            stream.read(reinterpret_cast<char *>(&m), sizeof(m));
            std::vector<std::uint64_t> all;
            for (std::size_t j = 0; j < m; j++) {
                ParmsID id;
                bool ntt = false, terms = false;
                std::size_t size = 0, count = 0;
                double scale = 1.0;
                std::uint64_t cf = 1, seed = 0;
                stream.read(reinterpret_cast<char *>(&id), sizeof(ParmsID));
                stream.read(reinterpret_cast<char *>(&ntt), sizeof(bool));
                stream.read(reinterpret_cast<char *>(&size), sizeof(size));
                stream.read(reinterpret_cast<char *>(&n_), sizeof(n_));
                stream.read(reinterpret_cast<char *>(&limbs_), sizeof(limbs_));
                stream.read(reinterpret_cast<char *>(&scale), sizeof(scale));
                stream.read(reinterpret_cast<char *>(&cf), sizeof(cf));
                stream.read(reinterpret_cast<char *>(&seed), sizeof(seed));
                stream.read(reinterpret_cast<char *>(&terms), sizeof(bool));
                stream.read(reinterpret_cast<char *>(&count), sizeof(count));
                if (!stream || terms || seed != 0 || size != 2 || count != 2 * n_ * limbs_) throw std::invalid_argument("kswitch_keys data is invalid");
                all.resize(all.size() + count);
                stream.read(reinterpret_cast<char *>(all.data() + all.size() - count), count * 8);
            }
            if (!all.empty()) {
                keys_[i].allocate(all.size());
                keys_[i].upload(all.data(), all.size());
                b200::synchronize(); // `all` is a host temporary
            }
        }
    }

protected:
    ParmsID parms_id_ = parmsIDZero;
    std::size_t n_ = 0, limbs_ = 0;
    std::vector<b200::DeviceWords> keys_;
};

class RelinKeysCuda : public KSwitchKeysCuda {
public:
    RelinKeysCuda() = default;
    RelinKeysCuda(const RelinKeys &h) : KSwitchKeysCuda(h) {}
    static std::size_t getIndex(std::size_t key_power) {
        if (key_power < 2) throw std::invalid_argument("key_power cannot be less than 2");
        return key_power - 2;
    }
    bool hasKey(std::size_t key_power) const { return hasIndex(getIndex(key_power)); }
};

class GaloisKeysCuda : public KSwitchKeysCuda {
public:
    GaloisKeysCuda() = default;
    GaloisKeysCuda(const GaloisKeys &h) : KSwitchKeysCuda(h) {}
    static std::size_t getIndex(std::uint32_t galois_elt) { return util::GaloisTool::GetIndexFromElt(galois_elt); }
    bool hasKey(std::uint32_t galois_elt) const { return hasIndex(getIndex(galois_elt)); }
};

// Key generation runs on the host in the reference as well (src/keygenerator_cuda.cuh:36-60); results are uploaded.
class KeyGeneratorCuda {
public:
    KeyGeneratorCuda(const SEALContextCuda &context) : gen_(context.host()) {}
    KeyGeneratorCuda(const SEALContextCuda &context, const SecretKey &sk) : gen_(context.host(), sk) {}
    SecretKeyCuda secretKey() const { return SecretKeyCuda(gen_.secretKey()); }
    void createPublicKey(PublicKeyCuda &dst) const { PublicKey pk; gen_.createPublicKey(pk); dst = PublicKeyCuda(pk); }
    void createRelinKeys(RelinKeysCuda &dst) { RelinKeys k; gen_.createRelinKeys(k); dst = RelinKeysCuda(k); }
    void createGaloisKeys(const std::vector<std::uint32_t> &elts, GaloisKeysCuda &dst) { GaloisKeys k; gen_.createGaloisKeys(elts, k); dst = GaloisKeysCuda(k); }
    void createGaloisKeys(const std::vector<int> &steps, GaloisKeysCuda &dst) { GaloisKeys k; gen_.createGaloisKeys(steps, k); dst = GaloisKeysCuda(k); }
    void createGaloisKeys(GaloisKeysCuda &dst) { GaloisKeys k; gen_.createGaloisKeys(k); dst = GaloisKeysCuda(k); }
    // keys for the automorphisms X -> X^(2^k + 1) of the LWE packing (src/keygenerator.cpp:350-358)
    GaloisKeysCuda createAutomorphismKeys() { return GaloisKeysCuda(gen_.createAutomorphismKeys()); }
    // key-switching key from another party's secret key to this generator's (src/keygenerator.h:292)
    KSwitchKeysCuda createKeySwitchingKeys(const SecretKey &new_key) { return KSwitchKeysCuda(gen_.createKeySwitchingKeys(new_key)); }
    KeyGenerator &host() { return gen_; }

private:
    KeyGenerator gen_;
};

// EncryptorCuda (src/encryptor_cuda.cuh / .cu): encryption runs on the device and follows the reference's CPU encryptor word
// for word (src/encryptor.cpp:88-253, src/utils/rlwe.cpp:160-356): every call takes a generator from the parameters' factory
// (parms.randomGenerator()->create(), so a factory with a fixed seed reproduces the CPU path's ciphertexts) and hands its
// 64-byte seed to troyb_encrypt_zero_*; the plaintext is then added into c0 on the device. Only the Blake2xb generator (the
// default) is reproduced. (The reference's own CUDA encryptor draws from cuRAND and matches no host generator.)
class EncryptorCuda {
public:
    EncryptorCuda(const SEALContextCuda &context, const PublicKey &pk) : context_(context) { setPublicKey(pk); }
    EncryptorCuda(const SEALContextCuda &context, const SecretKey &sk) : context_(context) { setSecretKey(sk); }
    EncryptorCuda(const SEALContextCuda &context, const PublicKey &pk, const SecretKey &sk) : context_(context) {
        setPublicKey(pk);
        setSecretKey(sk);
    }
    EncryptorCuda(const EncryptorCuda &) = delete;
    EncryptorCuda(EncryptorCuda &&) = delete;
    void setPublicKey(const PublicKey &pk) {
        if (!isMetadataValidFor(pk, context_.host()) || !isBufferValid(pk)) throw std::invalid_argument("public key is not valid for encryption parameters");
        const Ciphertext &k = pk.data();
        pk_.allocate(k.size() * k.coeffModulusSize() * k.polyModulusDegree());
        pk_.upload(k.data(), pk_.size());
    }
    void setSecretKey(const SecretKey &sk) {
        if (!isMetadataValidFor(sk, context_.host()) || !isBufferValid(sk)) throw std::invalid_argument("secret key is not valid for encryption parameters");
        sk_.allocate(sk.data().coeffCount());
        sk_.upload(sk.data().data(), sk_.size());
    }
    // Without a public key but with a secret key, encrypt() encrypts symmetrically: the reference's own CUDA test BFVKeySwitching
    // calls it that way (its CUDA class reads the unset key there; its CPU class throws "public key is not set").
    void encrypt(const Plaintext &plain, CiphertextCuda &dst) const { encryptInternal(plain, pk_.size() != 0 || sk_.size() == 0, dst); }
    void encryptSymmetric(const Plaintext &plain, CiphertextCuda &dst) const { encryptInternal(plain, false, dst); }
    CiphertextCuda encrypt(const Plaintext &plain) const { CiphertextCuda d; encrypt(plain, d); return d; }
    CiphertextCuda encryptSymmetric(const Plaintext &plain) const { CiphertextCuda d; encryptSymmetric(plain, d); return d; }
    void encryptZero(CiphertextCuda &dst) const { encryptZeroInternal(context_.firstParmsID(), true, dst); }
    void encryptZero(ParmsID id, CiphertextCuda &dst) const { encryptZeroInternal(id, true, dst); }
    CiphertextCuda encryptZero() const { CiphertextCuda d; encryptZero(d); return d; }
    CiphertextCuda encryptZero(ParmsID id) const { CiphertextCuda d; encryptZero(id, d); return d; }
    void encryptZeroSymmetric(CiphertextCuda &dst) const { encryptZeroInternal(context_.firstParmsID(), false, dst); }
    void encryptZeroSymmetric(ParmsID id, CiphertextCuda &dst) const { encryptZeroInternal(id, false, dst); }
    CiphertextCuda encryptZeroSymmetric() const { CiphertextCuda d; encryptZeroSymmetric(d); return d; }
    CiphertextCuda encryptZeroSymmetric(ParmsID id) const { CiphertextCuda d; encryptZeroSymmetric(id, d); return d; }
    // Symmetric encryptions in the reference's CUDA path are SEEDED (c1 = cuRAND expansion of destination.seed(), so that save
    // writes c0 + seed; src/utils/rlwe_cuda.cu:290-305) - the default here too. false: c1 is drawn from the Blake2xb ciphertext
    // generator as the CPU path does (src/utils/rlwe.cpp:290-296), word for word.
    void setSeededSymmetric(bool on) { seeded_symmetric_ = on; }

private:
    // Encryptor::encryptZeroInternal  src/encryptor.cpp:88-164
    void encryptZeroInternal(const ParmsID &id, bool asymmetric, CiphertextCuda &dst) const {
        auto cd = context_.getContextData(id);
        if (!cd) throw std::invalid_argument("parms_id is not valid for encryption parameters");
        const b200::DeviceWords &key = asymmetric ? pk_ : sk_;
        if (!key.size()) throw std::logic_error(asymmetric ? "public key is not set" : "secret key is not set");
        const auto &parms = cd->parms();
        auto prng = parms.randomGenerator()->create();
        if (prng->info().type() != PRNGType::blake2xb) throw std::logic_error("only the Blake2xb generator is reproduced on the device");
        const PRNGSeed seed = prng->seed();
        b200::DeviceWords d_seed(seed.size());
        d_seed.upload(seed.data(), seed.size());
        const int level = id == context_.keyParmsID() && context_.using_keyswitching() ? -1 : context_.levelOf(id);
        const std::size_t n = parms.polyModulusDegree(), limbs = parms.coeffModulus().size();
        b200::DeviceWords out(2 * n * limbs);
        if (asymmetric) b200::check(troyb_encrypt_zero_asymmetric(context_.handle(), level, key.get(), d_seed.get(), out.get(), out.size(), 1, b200::stream()));
        else if (!seeded_symmetric_) b200::check(troyb_encrypt_zero_symmetric(context_.handle(), level, key.get(), d_seed.get(), out.get(), out.size(), 1, b200::stream()));
        std::uint64_t ct_seed = 0;
        if (!asymmetric && seeded_symmetric_) {
            b200::DeviceWords d_ct_seed(1);
            b200::check(troyb_encrypt_zero_symmetric_seeded(context_.handle(), level, key.get(), d_seed.get(), out.get(), out.size(), d_ct_seed.get(), 1, b200::stream()));
            d_ct_seed.download(&ct_seed, 1);
            b200::synchronize();
        }
        dst.adopt(std::move(out), 2, n, limbs);
        dst.parmsID() = id;
        dst.isNttForm() = parms.scheme() == SchemeType::ckks;
        dst.scale() = 1.0;
        dst.correctionFactor() = 1;
        dst.seed() = ct_seed;
    }
    // Encryptor::encryptInternal  src/encryptor.cpp:166-253
    void encryptInternal(const Plaintext &plain, bool asymmetric, CiphertextCuda &dst) const {
        if (!isValidFor(plain, context_.host())) throw std::invalid_argument("plain is not valid for encryption parameters");
        const SchemeType scheme = context_.keyContextData()->parms().scheme();
        if (scheme == SchemeType::ckks) {
            if (!plain.isNttForm()) throw std::invalid_argument("plain must be in NTT form");
            if (!context_.getContextData(plain.parmsID())) throw std::invalid_argument("plain is not valid for encryption parameters");
            encryptZeroInternal(plain.parmsID(), asymmetric, dst);
            const std::uint64_t ct_seed = dst.seed(); // adding into c0 leaves c1, and with it the seed, valid
            b200::DeviceWords p(dst.polyCoeffSize());
            p.upload(plain.data(), p.size());
            std::uint64_t *c0 = dst.data();
            b200::check(troyb_add(context_.handle(), context_.levelOf(plain.parmsID()), c0, p.get(), c0, 1, b200::stream()));
            dst.scale() = plain.scale();
            dst.seed() = ct_seed;
            return;
        }
        if (plain.isNttForm()) throw std::invalid_argument("plain cannot be in NTT form");
        encryptZeroInternal(context_.firstParmsID(), asymmetric, dst);
        // BFV: multiplyAddPlainWithScalingVariant; BGV: addPlainWithoutScalingVariant (the fresh correction factor is 1)
        const std::size_t count = plain.coeffCount();
        if (!count) return;
        const std::uint64_t ct_seed = dst.seed();
        b200::DeviceWords p(count);
        p.upload(plain.data(), count);
        b200::check(troyb_add_plain(context_.handle(), 0, dst.data(), 2 * dst.polyCoeffSize(), p.get(), count, 0, 1, 0, 1, b200::stream()));
        dst.seed() = ct_seed;
    }
    const SEALContextCuda &context_;
    b200::DeviceWords pk_, sk_;
    bool seeded_symmetric_ = true;
};

// DecryptorCuda (src/decryptor_cuda.cuh / .cu): the secret key powers live on the device; decrypt() is the dot product
// c_0 + c_1 s + ... on the device followed by decryptScaleAndRound (BFV) / decryptModt (BGV) on the device; the plaintext
// comes back to its host representation. invariantNoiseBudget stays host work (multi-precision arithmetic).
class DecryptorCuda {
public:
    DecryptorCuda(const SEALContextCuda &context, const SecretKey &sk) : context_(context), dec_(context.host(), sk) {
        key_words_ = sk.data().coeffCount();
        powers_.allocate(key_words_);
        powers_.upload(sk.data().data(), key_words_);
        n_powers_ = 1;
    }
    void decrypt(const CiphertextCuda &ct, Plaintext &dst) {
        if (ct.size() < 2) throw std::invalid_argument("encrypted is empty");
        const SchemeType s = context_.firstContextData()->parms().scheme();
        if (s == SchemeType::ckks && !ct.isNttForm()) throw std::invalid_argument("encrypted must be in NTT form");
        if (s != SchemeType::ckks && ct.isNttForm()) throw std::invalid_argument("encrypted cannot be in NTT form");
        const std::size_t need = ct.size() - 1;
        if (need > n_powers_) { // computeSecretKeyArray  src/decryptor_cuda.cu:188-235
            b200::DeviceWords p(need * key_words_);
            b200::check(troyb_secret_key_powers(context_.handle(), powers_.get(), need, p.get(), b200::stream()));
            powers_ = std::move(p);
            n_powers_ = need;
        }
        const int lv = context_.levelOf(ct.parmsID());
        const std::size_t n = ct.polyModulusDegree(), words = ct.polyCoeffSize();
        b200::DeviceWords phase(words);
        b200::check(troyb_decrypt_dot_product(context_.handle(), lv, ct.data(), ct.size(), ct.size() * words, ct.isNttForm() ? 1 : 0,
                                              powers_.get(), phase.get(), 1, b200::stream()));
        dst.parmsID() = parmsIDZero;
        if (s == SchemeType::ckks) { // ckksDecrypt  src/decryptor_cuda.cu:118-141
            dst.resize(words);
            phase.download(dst.data(), words);
            dst.parmsID() = ct.parmsID();
            dst.scale() = ct.scale();
            return;
        }
        b200::DeviceWords out(n);
        if (s == SchemeType::bfv) {
            b200::check(troyb_decrypt_scale_and_round(context_.handle(), lv, phase.get(), out.get(), 1, b200::stream()));
        } else {
            std::uint64_t fix = 1;
            const auto &t = context_.firstContextData()->parms().plainModulus();
            if (ct.correctionFactor() != 1 && !util::tryInvertUintMod(ct.correctionFactor(), t, fix))
                throw std::logic_error("invalid correction factor");
            b200::check(troyb_decrypt_mod_t(context_.handle(), lv, phase.get(), out.get(), fix, 1, b200::stream()));
        }
        // full size, like the reference's CUDA decryptor (src/decryptor_cuda.cu:119-131; its CPU decryptor trims to the significant
        // coefficients, the CUDA one has that step commented out, and decodePolynomial callers index up to the slot count)
        dst.resize(n);
        out.download(dst.data(), n);
    }
    int invariantNoiseBudget(const CiphertextCuda &ct) { return dec_.invariantNoiseBudget(ct.cpu()); }

private:
    SEALContextCuda context_;
    Decryptor dec_;
    b200::DeviceWords powers_;
    std::size_t key_words_ = 0, n_powers_ = 0;
};

class EvaluatorCuda {
public:
    EvaluatorCuda(const SEALContextCuda &context) : context_(context) {}

    // ---- add / sub / negate (src/evaluator_cuda.cu:143-260)
    void negateInplace(CiphertextCuda &ct) const {
        b200::check(troyb_negate(h(), level(ct), ct.data(), ct.data(), ct.size(), b200::stream()));
    }
    void negate(const CiphertextCuda &a, CiphertextCuda &dst) const { dst = a; negateInplace(dst); }
    void addInplace(CiphertextCuda &a, const CiphertextCuda &b) const { addSub(a, b, false); }
    void subInplace(CiphertextCuda &a, const CiphertextCuda &b) const { addSub(a, b, true); }
    void add(const CiphertextCuda &a, const CiphertextCuda &b, CiphertextCuda &dst) const {
        if (&b == &dst) addInplace(dst, a);
        else { dst = a; addInplace(dst, b); }
    }
    void sub(const CiphertextCuda &a, const CiphertextCuda &b, CiphertextCuda &dst) const {
        if (&b == &dst) { subInplace(dst, a); negateInplace(dst); }
        else { dst = a; subInplace(dst, b); }
    }
    void addMany(const std::vector<CiphertextCuda> &cts, CiphertextCuda &dst) const {
        if (cts.empty()) throw std::invalid_argument("encrypteds cannot be empty");
        for (auto &c : cts) if (&c == &dst) throw std::invalid_argument("encrypteds must be different from destination");
        dst = cts[0];
        for (std::size_t i = 1; i < cts.size(); i++) addInplace(dst, cts[i]);
    }

    // ---- multiply / square (src/evaluator_cuda.cu:262-701)
    void multiplyInplace(CiphertextCuda &a, const CiphertextCuda &b) const {
        if (a.parmsID() != b.parmsID()) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        const SchemeType s = scheme();
        if (s == SchemeType::ckks && !(a.isNttForm() && b.isNttForm())) throw std::invalid_argument("encrypted1 or encrypted2 must be in NTT form");
        if (s != SchemeType::ckks && (a.isNttForm() || b.isNttForm())) throw std::invalid_argument("encrypted1 or encrypted2 cannot be in NTT form");
        const int lv = level(a);
        const std::size_t a_size = a.size(), b_size = b.size(), out = a_size + b_size - 1;
        double new_scale = a.scale() * b.scale();
        if (s == SchemeType::ckks && !scaleWithinBounds(new_scale, a.parmsID())) throw std::invalid_argument("scale out of bounds");
        const bool square = &a == &b;
        const std::size_t stride = out * a.polyCoeffSize();
        a.resize(out); // keeps the first a_size polynomials
        const std::uint64_t *pb = square ? a.data() : b.data();
        b200::check(troyb_multiply(h(), lv, a.data(), a_size, stride, pb, b_size, square ? stride : b_size * b.polyCoeffSize(), a.data(),
                                   stride, 1, b200::stream()));
        if (s == SchemeType::ckks) a.scale() = new_scale;
        if (s == SchemeType::bgv) {
            const std::uint64_t t = context_.firstContextData()->parms().plainModulus().value();
            a.correctionFactor() = (std::uint64_t)((unsigned __int128)a.correctionFactor() * b.correctionFactor() % t);
        }
    }
    void multiply(const CiphertextCuda &a, const CiphertextCuda &b, CiphertextCuda &dst) const {
        if (&b == &dst) multiplyInplace(dst, a);
        else { dst = a; multiplyInplace(dst, b); }
    }
    void squareInplace(CiphertextCuda &a) const { multiplyInplace(a, a); }
    void square(const CiphertextCuda &a, CiphertextCuda &dst) const { dst = a; squareInplace(dst); }

    // ---- key switching (src/evaluator_cuda.cu:703-747, 1163-1378)
    void relinearizeInplace(CiphertextCuda &ct, const RelinKeysCuda &rlk) const {
        if (ct.size() == 2) return;
        if (ct.size() != 3) throw std::logic_error("relinearization of size > 3 is outside the ported path");
        checkKeySwitch(ct, rlk);
        if (!rlk.hasKey(2)) throw std::invalid_argument("not enough relinearization keys");
        b200::check(troyb_relinearize(h(), level(ct), ct.data(), 3 * ct.polyCoeffSize(), rlk.key(RelinKeysCuda::getIndex(2)), 1, b200::stream()));
        ct.resize(2);
    }
    void relinearize(const CiphertextCuda &ct, const RelinKeysCuda &rlk, CiphertextCuda &dst) const { dst = ct; relinearizeInplace(dst, rlk); }
    // EvaluatorCuda::multiplyMany  src/evaluator_cuda.cu:1447-1523: product tree, every product relinearised (BFV / BGV)
    void multiplyMany(const std::vector<CiphertextCuda> &cts, const RelinKeysCuda &rlk, CiphertextCuda &dst) const {
        if (cts.empty()) throw std::invalid_argument("encrypteds vector must not be empty");
        for (auto &c : cts) if (&c == &dst) throw std::invalid_argument("encrypteds must be different from destination");
        if (!context_.getContextData(cts[0].parmsID())) throw std::invalid_argument("encrypteds is not valid for encryption parameters");
        if (scheme() != SchemeType::bfv && scheme() != SchemeType::bgv) throw std::logic_error("unsupported scheme");
        if (cts.size() == 1) { dst = cts[0]; return; }
        std::vector<CiphertextCuda> prod;
        for (std::size_t i = 0; i + 1 < cts.size(); i += 2) {
            CiphertextCuda t;
            if (cts[i].data() == cts[i + 1].data()) square(cts[i], t);
            else multiply(cts[i], cts[i + 1], t);
            relinearizeInplace(t, rlk);
            prod.emplace_back(std::move(t));
        }
        if (cts.size() & 1) prod.emplace_back(cts.back());
        for (std::size_t i = 0; i + 1 < prod.size(); i += 2) {
            CiphertextCuda t;
            multiply(prod[i], prod[i + 1], t);
            relinearizeInplace(t, rlk);
            prod.emplace_back(std::move(t));
        }
        dst = prod.back();
    }
    // EvaluatorCuda::exponentiateInplace  src/evaluator_cuda.cu:1525-1553
    void exponentiateInplace(CiphertextCuda &ct, std::uint64_t exponent, const RelinKeysCuda &rlk) const {
        if (!context_.getContextData(ct.parmsID())) throw std::invalid_argument("encrypted is not valid for encryption parameters");
        if (!context_.getContextData(rlk.parmsID())) throw std::invalid_argument("relin_keys is not valid for encryption parameters");
        if (exponent == 0) throw std::invalid_argument("exponent cannot be 0");
        if (exponent == 1) return;
        // copies of one ciphertext: the product tree multiplies distinct buffers where the reference squares aliased ones - the same words
        std::vector<CiphertextCuda> copies(static_cast<std::size_t>(exponent), ct);
        multiplyMany(copies, rlk, ct);
    }
    void exponentiate(const CiphertextCuda &ct, std::uint64_t exponent, const RelinKeysCuda &rlk, CiphertextCuda &dst) const {
        dst = ct;
        exponentiateInplace(dst, exponent, rlk);
    }
    // EvaluatorCuda::applyKeySwitchingInplace  src/evaluator_cuda.cu:1365-1378: c1 becomes the target of a key switch with the
    // single key of `keys`, and starts again from zero
    void applyKeySwitchingInplace(CiphertextCuda &ct, const KSwitchKeysCuda &keys) const {
        if (keys.size() != 1 || !keys.hasIndex(0)) throw std::invalid_argument("kswitch_keys.data().size() != 1");
        if (ct.size() != 2) throw std::invalid_argument("encrypted.size() != 2");
        const std::size_t words = ct.polyCoeffSize();
        b200::DeviceWords target(words);
        b200::check(troyb_copy_on_device(target.get(), ct.data(1), words * 8, b200::stream()));
        b200::check(troyb_memset_zero(ct.data(1), words * 8, b200::stream()));
        b200::check(troyb_switch_key(h(), level(ct), ct.data(), 2 * words, target.get(), words, keys.key(0), 1, b200::stream()));
    }
    void applyKeySwitching(const CiphertextCuda &ct, const KSwitchKeysCuda &keys, CiphertextCuda &dst) const {
        dst = ct;
        applyKeySwitchingInplace(dst, keys);
    }

    // ---- LWE extraction / packing (src/evaluator_cuda.cu:2178-2365): compositions over the negacyclic shift, applyGalois and add
    void negacyclicShift(const CiphertextCuda &ct, std::size_t shift, CiphertextCuda &dst) const {
        CiphertextCuda out = ct; // metadata and size
        b200::check(troyb_negacyclic_shift(h(), level(ct), ct.data(), out.data(), ct.size(), shift, b200::stream()));
        dst = std::move(out);
    }
    void negacyclicShiftInplace(CiphertextCuda &ct, std::size_t shift) const { CiphertextCuda x = ct; negacyclicShift(x, shift, ct); }
    LWECiphertextCuda extractLWE(const CiphertextCuda &ct, std::size_t term) const {
        if (ct.size() != 2) throw std::invalid_argument("Encrypted size must be 2 to be extracted.");
        if (ct.isNttForm()) {
            CiphertextCuda t;
            transformFromNtt(ct, t);
            return extractLWE(t, term);
        }
        const std::size_t n = ct.polyModulusDegree(), limbs = ct.coeffModulusSize();
        b200::DeviceWords c1(n * limbs), c0(limbs);
        b200::check(troyb_negacyclic_shift(h(), level(ct), ct.data(1), c1.get(), 1, term == 0 ? 0 : 2 * n - term, b200::stream()));
        for (std::size_t l = 0; l < limbs; l++) // extractLWECopyC0: coefficient `term` of every limb of c0
            b200::check(troyb_copy_on_device(c0.get() + l, ct.data(0) + l * n + term, 8, b200::stream()));
        return LWECiphertextCuda::fromMembers(ct.parmsID(), n, limbs, ct.scale(), ct.correctionFactor(), std::move(c1), std::move(c0));
    }
    CiphertextCuda assembleLWE(const LWECiphertextCuda &lwe, std::size_t term) const {
        const std::size_t n = lwe.polyModulusDegree(), limbs = lwe.coeffModulusSize(), poly = n * limbs;
        const int lv = context_.levelOf(lwe.parmsID());
        b200::DeviceWords data(2 * poly);
        b200::check(troyb_memset_zero(data.get(), poly * 8, b200::stream()));
        b200::check(troyb_negacyclic_shift(h(), lv, lwe.c1(), data.get() + poly, 1, term, b200::stream()));
        for (std::size_t l = 0; l < limbs; l++) // assembleLWECopyC0
            b200::check(troyb_copy_on_device(data.get() + l * n + term, lwe.c0() + l, 8, b200::stream()));
        CiphertextCuda out;
        out.adopt(std::move(data), 2, n, limbs);
        out.parmsID() = lwe.parmsID();
        out.isNttForm() = false;
        out.scale() = lwe.scale();
        out.correctionFactor() = lwe.correctionFactor();
        return out;
    }
    void fieldTraceInplace(CiphertextCuda &ct, const GaloisKeysCuda &automorphism_keys, std::size_t logn) const {
        std::size_t degree = ct.polyModulusDegree();
        CiphertextCuda t;
        while (degree > (std::size_t(1) << logn)) {
            applyGalois(ct, (std::uint32_t)(degree + 1), automorphism_keys, t);
            addInplace(ct, t);
            degree >>= 1;
        }
    }
    void divideByPolyModulusDegreeInplace(CiphertextCuda &ct, std::uint64_t mul = 1) const {
        const int lv = level(ct);
        std::uint64_t *d = ct.data();
        b200::check(troyb_multiply_inv_degree(h(), lv, d, d, ct.size(), b200::stream()));
        if (mul > 1) b200::check(troyb_multiply_scalar(h(), lv, d, mul, d, ct.size(), b200::stream()));
    }
    CiphertextCuda packLWECiphertexts(const std::vector<LWECiphertextCuda> &lwes, GaloisKeysCuda &automorphism_keys) const {
        if (lwes.empty()) throw std::invalid_argument("LWE ciphertexts must not be empty.");
        for (auto &x : lwes) if (x.parmsID() != lwes[0].parmsID()) throw std::invalid_argument("LWE ciphertexts must have same parmsID.");
        const std::size_t n = lwes[0].polyModulusDegree();
        const bool ckks = scheme() == SchemeType::ckks;
        std::size_t l = 0;
        while ((std::size_t(1) << l) < lwes.size()) l++;
        std::vector<CiphertextCuda> rlwes(std::size_t(1) << l);
        CiphertextCuda zero = assembleLWE(lwes[0], 0);
        b200::check(troyb_memset_zero(zero.data(), 2 * zero.polyCoeffSize() * 8, b200::stream()));
        for (std::size_t i = 0; i < rlwes.size(); i++) {
            std::size_t index = 0;
            for (std::size_t b = 0; b < l; b++) index |= ((i >> b) & 1) << (l - 1 - b); // reverseBits(i, l)
            if (index < lwes.size()) {
                rlwes[i] = assembleLWE(lwes[index], 0);
                divideByPolyModulusDegreeInplace(rlwes[i]);
            } else {
                rlwes[i] = zero;
            }
        }
        CiphertextCuda temp = zero;
        for (std::size_t layer = 0; layer < l; layer++) {
            const std::size_t gap = std::size_t(1) << layer, shift = n >> (layer + 1);
            for (std::size_t offset = 0; offset < rlwes.size(); offset += 2 * gap) {
                CiphertextCuda &even = rlwes[offset], &odd = rlwes[offset + gap];
                negacyclicShift(odd, shift, temp);
                sub(even, temp, odd);
                addInplace(even, temp);
                if (ckks) transformToNttInplace(odd);
                applyGaloisInplace(odd, (std::uint32_t)((std::size_t(1) << (layer + 1)) + 1), automorphism_keys);
                if (ckks) transformFromNttInplace(odd);
                addInplace(even, odd);
            }
        }
        CiphertextCuda ret = std::move(rlwes[0]);
        fieldTraceInplace(ret, automorphism_keys, l);
        if (ckks) transformToNttInplace(ret);
        return ret;
    }

    // ---- Galois / rotations (src/evaluator_cuda.cu:2024-2176)
    void applyGaloisInplace(CiphertextCuda &ct, std::uint32_t elt, const GaloisKeysCuda &gk) const {
        if (gk.parmsID() != context_.keyParmsID()) throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        const std::size_t n = ct.polyModulusDegree();
        if (!(elt & 1) || elt >= 2 * n) throw std::invalid_argument("Galois element is not valid");
        if (ct.size() > 2) throw std::invalid_argument("encrypted size must be 2");
        if (!gk.hasKey(elt)) throw std::invalid_argument("Galois key not present");
        checkKeySwitch(ct, gk);
        b200::check(troyb_apply_galois(h(), level(ct), ct.data(), 2 * ct.polyCoeffSize(), elt, gk.key(GaloisKeysCuda::getIndex(elt)), 1, b200::stream()));
    }
    void applyGalois(const CiphertextCuda &ct, std::uint32_t elt, const GaloisKeysCuda &gk, CiphertextCuda &dst) const { dst = ct; applyGaloisInplace(dst, elt, gk); }
    void rotateRowsInplace(CiphertextCuda &ct, int steps, const GaloisKeysCuda &gk) const {
        if (scheme() != SchemeType::bfv && scheme() != SchemeType::bgv) throw std::logic_error("unsupported scheme");
        rotateInternal(ct, steps, gk);
    }
    void rotateRows(const CiphertextCuda &ct, int steps, const GaloisKeysCuda &gk, CiphertextCuda &dst) const { dst = ct; rotateRowsInplace(dst, steps, gk); }
    void rotateColumnsInplace(CiphertextCuda &ct, const GaloisKeysCuda &gk) const {
        if (scheme() != SchemeType::bfv && scheme() != SchemeType::bgv) throw std::logic_error("unsupported scheme");
        applyGaloisInplace(ct, eltFromStep(0), gk);
    }
    void rotateColumns(const CiphertextCuda &ct, const GaloisKeysCuda &gk, CiphertextCuda &dst) const { dst = ct; rotateColumnsInplace(dst, gk); }
    void rotateVectorInplace(CiphertextCuda &ct, int steps, const GaloisKeysCuda &gk) const {
        if (scheme() != SchemeType::ckks) throw std::logic_error("unsupported scheme");
        rotateInternal(ct, steps, gk);
    }
    void rotateVector(const CiphertextCuda &ct, int steps, const GaloisKeysCuda &gk, CiphertextCuda &dst) const { dst = ct; rotateVectorInplace(dst, steps, gk); }
    void complexConjugateInplace(CiphertextCuda &ct, const GaloisKeysCuda &gk) const {
        if (scheme() != SchemeType::ckks) throw std::logic_error("unsupported scheme");
        applyGaloisInplace(ct, eltFromStep(0), gk);
    }
    void complexConjugate(const CiphertextCuda &ct, const GaloisKeysCuda &gk, CiphertextCuda &dst) const { dst = ct; complexConjugateInplace(dst, gk); }

    // ---- rescale / modulus switching (src/evaluator_cuda.cu:749-1010, 1380)
    void rescaleToNext(const CiphertextCuda &ct, CiphertextCuda &dst) const {
        if (scheme() != SchemeType::ckks) throw std::invalid_argument("unsupported operation for scheme type");
        scaleToNext(ct, dst);
    }
    void rescaleToNextInplace(CiphertextCuda &ct) const { CiphertextCuda out; rescaleToNext(ct, out); ct = std::move(out); }
    void rescaleToInplace(CiphertextCuda &ct, const ParmsID &target) const {
        const int want = context_.levelOf(target);
        if (level(ct) > want) throw std::invalid_argument("cannot switch to higher level modulus");
        while (level(ct) < want) rescaleToNextInplace(ct);
    }
    void modSwitchToNext(const CiphertextCuda &ct, CiphertextCuda &dst) const {
        if (scheme() != SchemeType::ckks) { scaleToNext(ct, dst); return; }
        // CKKS: drop the last limb (modSwitchDropToNext, src/evaluator_cuda.cu:826-880)
        const int lv = level(ct);
        if (lv + 1 >= (int)context_.chainLength()) throw std::invalid_argument("end of modulus switching chain reached");
        const std::size_t n = ct.polyModulusDegree(), L = ct.coeffModulusSize();
        b200::DeviceWords out(ct.size() * (L - 1) * n);
        for (std::size_t p = 0; p < ct.size(); p++)
            b200::check(troyb_copy_on_device(out.get() + p * (L - 1) * n, ct.data(p), (L - 1) * n * 8, b200::stream()));
        CiphertextCuda r = metaOf(ct);
        r.adopt(std::move(out), ct.size(), n, L - 1);
        r.parmsID() = context_.parmsIDOf(lv + 1);
        dst = std::move(r);
    }
    void modSwitchToNextInplace(CiphertextCuda &ct) const { CiphertextCuda out; modSwitchToNext(ct, out); ct = std::move(out); }
    void modSwitchToInplace(CiphertextCuda &ct, const ParmsID &target) const {
        const int want = context_.levelOf(target);
        if (level(ct) > want) throw std::invalid_argument("cannot switch to higher level modulus");
        while (level(ct) < want) modSwitchToNextInplace(ct);
    }

    // ---- NTT transforms (src/evaluator_cuda.cu:1950-2022)
    void transformToNttInplace(CiphertextCuda &ct) const {
        if (ct.isNttForm()) throw std::invalid_argument("encrypted is already in NTT form");
        b200::check(troyb_ntt_forward(h(), level(ct), ct.data(), ct.size(), b200::stream()));
        ct.isNttForm() = true;
    }
    void transformFromNttInplace(CiphertextCuda &ct) const {
        if (!ct.isNttForm()) throw std::invalid_argument("encrypted_ntt is not in NTT form");
        b200::check(troyb_ntt_inverse(h(), level(ct), ct.data(), ct.size(), b200::stream()));
        ct.isNttForm() = false;
    }
    // ---- ciphertext (x) plaintext (src/evaluator_cuda.cu:1555-1935). Plaintexts keep their host representation in this
    // layer (see PlaintextCuda above); each call uploads the plaintext words it needs.
    void addPlainInplace(CiphertextCuda &ct, const PlaintextCuda &plain) const { addSubPlain(ct, plain, false); }
    void subPlainInplace(CiphertextCuda &ct, const PlaintextCuda &plain) const { addSubPlain(ct, plain, true); }
    void addPlain(const CiphertextCuda &ct, const PlaintextCuda &plain, CiphertextCuda &dst) const { dst = ct; addPlainInplace(dst, plain); }
    void subPlain(const CiphertextCuda &ct, const PlaintextCuda &plain, CiphertextCuda &dst) const { dst = ct; subPlainInplace(dst, plain); }
    void multiplyPlainInplace(CiphertextCuda &ct, const PlaintextCuda &plain) const {
        if (ct.isNttForm() != plain.isNttForm()) throw std::invalid_argument("NTT form mismatch");
        const int lv = level(ct);
        const std::size_t stride = ct.size() * ct.polyCoeffSize();
        b200::DeviceWords d(plain.coeffCount());
        d.upload(plain.data(), plain.coeffCount());
        if (ct.isNttForm()) {
            // multiplyPlainNtt  src/evaluator_cuda.cu:1818-1862
            if (ct.parmsID() != plain.parmsID()) throw std::invalid_argument("encrypted_ntt and plain_ntt parameter mismatch");
            b200::check(troyb_multiply_plain_ntt(h(), lv, ct.data(), ct.size(), stride, d.get(), 0, ct.data(), stride, 1, 0, b200::stream()));
            ct.scale() *= plain.scale();
            if (!scaleWithinBounds(ct.scale(), ct.parmsID())) throw std::invalid_argument("scale out of bounds");
        } else {
            // multiplyPlainNormal  src/evaluator_cuda.cu:1757-1816
            b200::check(troyb_multiply_plain_normal(h(), lv, ct.data(), ct.size(), stride, d.get(), plain.coeffCount(), 1, b200::stream()));
            if (scheme() == SchemeType::ckks) {
                ct.scale() *= plain.scale();
                if (!scaleWithinBounds(ct.scale(), ct.parmsID())) throw std::invalid_argument("scale out of bounds");
            }
        }
    }
    void multiplyPlain(const CiphertextCuda &ct, const PlaintextCuda &plain, CiphertextCuda &dst) const { dst = ct; multiplyPlainInplace(dst, plain); }
    // transformToNttInplace(PlaintextCuda &, ParmsID)  src/evaluator_cuda.cu:1865-1935
    void transformToNttInplace(PlaintextCuda &plain, const ParmsID &parms_id) const {
        auto cd = context_.getContextData(parms_id);
        if (!cd) throw std::invalid_argument("parms_id is not valid for the current context");
        if (plain.isNttForm()) throw std::invalid_argument("plain is already in NTT form");
        const std::size_t n = cd->parms().polyModulusDegree(), L = cd->parms().coeffModulus().size(), count = plain.coeffCount();
        b200::DeviceWords in(count), out(L * n);
        in.upload(plain.data(), count);
        b200::check(troyb_plain_lift(h(), context_.levelOf(parms_id), in.get(), count, count, out.get(), 1, 1, b200::stream()));
        plain.resize(L * n);
        out.download(plain.data(), L * n);
        plain.parmsID() = parms_id;
    }
    void transformToNtt(const PlaintextCuda &plain, const ParmsID &parms_id, PlaintextCuda &dst) const {
        dst = plain;
        transformToNttInplace(dst, parms_id);
    }

    void transformToNtt(const CiphertextCuda &ct, CiphertextCuda &dst) const { dst = ct; transformToNttInplace(dst); }
    void transformFromNtt(const CiphertextCuda &ct, CiphertextCuda &dst) const { dst = ct; transformFromNttInplace(dst); }

private:
    troyb_context *h() const { return context_.handle(); }
    int level(const CiphertextCuda &ct) const { return context_.levelOf(ct.parmsID()); }
    SchemeType scheme() const { return context_.firstContextData()->parms().scheme(); }
    std::uint32_t eltFromStep(int step) const { std::uint32_t e = 0; b200::check(troyb_galois_elt_from_step(h(), step, &e)); return e; }
    static CiphertextCuda metaOf(const CiphertextCuda &ct) {
        CiphertextCuda r;
        r.parmsID() = ct.parmsID(); r.isNttForm() = ct.isNttForm(); r.scale() = ct.scale(); r.correctionFactor() = ct.correctionFactor();
        return r;
    }
    // isScaleWithinBounds, src/evaluator_cuda.cu:32-51
    bool scaleWithinBounds(double scale, const ParmsID &id) const {
        auto cd = context_.getContextData(id);
        int bound = scheme() == SchemeType::ckks ? cd->totalCoeffModulusBitCount() : cd->parms().plainModulus().bitCount();
        return !(scale <= 0 || (int)std::log2(scale) >= bound);
    }
    void addSubPlain(CiphertextCuda &ct, const PlaintextCuda &plain, bool sub) const {
        const SchemeType s = scheme();
        if (s == SchemeType::bfv && ct.isNttForm()) throw std::invalid_argument("BFV encrypted cannot be in NTT form");
        if (s == SchemeType::ckks && !ct.isNttForm()) throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (s == SchemeType::bgv && ct.isNttForm()) throw std::invalid_argument("BGV encrypted cannot be in NTT form");
        if (plain.isNttForm() != ct.isNttForm()) throw std::invalid_argument("NTT form mismatch");
        if (ct.isNttForm() && ct.parmsID() != plain.parmsID()) throw std::invalid_argument("encrypted and plain parameter mismatch");
        if (s == SchemeType::ckks && ct.scale() != plain.scale() && std::fabs(ct.scale() - plain.scale()) > 1e-15 * std::fabs(ct.scale()))
            throw std::invalid_argument("scale mismatch");
        const int lv = level(ct);
        b200::DeviceWords d(plain.coeffCount());
        d.upload(plain.data(), plain.coeffCount());
        if (s == SchemeType::ckks) {
            // kAddPolyCoeffmod / kSubPolyCoeffmod on poly 0  src/evaluator_cuda.cu:1607-1612,1680-1685
            if (plain.coeffCount() != ct.polyCoeffSize()) throw std::invalid_argument("plain is not valid for encryption parameters");
            b200::check((sub ? troyb_sub : troyb_add)(h(), lv, ct.data(), d.get(), ct.data(), 1, b200::stream()));
        } else {
            b200::check(troyb_add_plain(h(), lv, ct.data(), ct.size() * ct.polyCoeffSize(), d.get(), plain.coeffCount(), 0,
                                        ct.correctionFactor(), sub ? 1 : 0, 1, b200::stream()));
        }
    }
    void checkKeySwitch(const CiphertextCuda &ct, const KSwitchKeysCuda &keys) const {
        if (!context_.using_keyswitching()) throw std::logic_error("keyswitching is not supported by the context");
        if (keys.parmsID() != context_.keyParmsID()) throw std::invalid_argument("parameter mismatch");
        const SchemeType s = scheme();
        if (s == SchemeType::bfv && ct.isNttForm()) throw std::invalid_argument("BFV encrypted cannot be in NTT form");
        if (s == SchemeType::ckks && !ct.isNttForm()) throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (s == SchemeType::bgv && ct.isNttForm()) throw std::invalid_argument("BGV encrypted cannot be in NTT form");
    }
    void addSub(CiphertextCuda &a, const CiphertextCuda &b, bool sub) const {
        if (a.parmsID() != b.parmsID()) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (a.isNttForm() != b.isNttForm()) throw std::invalid_argument("NTT form mismatch");
        if (scheme() == SchemeType::bgv && a.correctionFactor() != b.correctionFactor()) {
            // balance the correction factors with scalar multiplications first (src/evaluator_cuda.cu:171-190)
            const std::uint64_t t = context_.firstContextData()->parms().plainModulus().value();
            std::uint64_t f, e1, e2;
            balanceCorrectionFactors(a.correctionFactor(), b.correctionFactor(), t, f, e1, e2);
            CiphertextCuda b2 = b;
            b200::check(troyb_multiply_scalar(h(), level(a), a.data(), e1, a.data(), a.size(), b200::stream()));
            b200::check(troyb_multiply_scalar(h(), level(b2), b2.data(), e2, b2.data(), b2.size(), b200::stream()));
            a.correctionFactor() = f;
            b2.correctionFactor() = f;
            addSub(a, b2, sub);
            return;
        }
        if (scheme() == SchemeType::ckks && a.scale() != b.scale() && std::fabs(a.scale() - b.scale()) > 1e-15 * std::fabs(a.scale()))
            throw std::invalid_argument("scale mismatch");
        const int lv = level(a);
        const std::size_t a_old = a.size(), mn = std::min(a.size(), b.size()), mx = std::max(a.size(), b.size());
        a.resize(mx);
        b200::check((sub ? troyb_sub : troyb_add)(h(), lv, a.data(), b.data(), a.data(), mn, b200::stream()));
        if (a_old < b.size()) {
            if (!sub) b200::check(troyb_copy_on_device(a.data(a_old), b.data(a_old), (b.size() - a_old) * a.polyCoeffSize() * 8, b200::stream()));
            else b200::check(troyb_negate(h(), lv, b.data(a_old), a.data(a_old), b.size() - a_old, b200::stream()));
        }
    }
    // balanceCorrectionFactors (src/evaluator_cuda.cu:53-114): the pair (e1, e2) with e1 f1 = e2 f2 (mod t) of least
    // |e1| + |e2| among the remainders of the extended Euclid run on (t, f2 / f1)
    static void balanceCorrectionFactors(std::uint64_t f1, std::uint64_t f2, std::uint64_t t, std::uint64_t &factor, std::uint64_t &e1,
                                         std::uint64_t &e2) {
        typedef unsigned __int128 u128;
        const std::uint64_t half = t / 2;
        auto bal = [&](std::uint64_t x) { return x > half ? (std::int64_t)(t - x) : (std::int64_t)x; }; // |balanced x|
        auto sum_abs = [&](std::uint64_t x, std::uint64_t y) { return bal(x) + bal(y); };
        auto gcd = [](std::uint64_t x, std::uint64_t y) { while (y) { std::uint64_t r = x % y; x = y; y = r; } return x; };
        std::uint64_t inv = 0;
        if (!util::tryInvertUintMod(f1 % t, t, inv)) throw std::logic_error("invalid correction factor1");
        const std::uint64_t ratio = (std::uint64_t)((u128)inv * f2 % t);
        e1 = ratio; e2 = 1;
        std::int64_t best = sum_abs(e1, e2);
        std::int64_t prev_a = (std::int64_t)t, prev_b = 0, a = (std::int64_t)ratio, b = 1;
        while (a != 0) {
            const std::int64_t q = prev_a / a, ra = prev_a % a;
            prev_a = a; a = ra;
            const std::int64_t rb = prev_b - b * q;
            prev_b = b; b = rb;
            const std::uint64_t a_mod = (std::uint64_t)a % t; // a >= 0 throughout
            std::uint64_t b_mod = (std::uint64_t)(b < 0 ? -b : b) % t;
            if (b < 0 && b_mod) b_mod = t - b_mod;
            if (a_mod != 0 && gcd(a_mod, t) == 1) {
                const std::int64_t s = sum_abs(a_mod, b_mod);
                if (s < best) { best = s; e1 = a_mod; e2 = b_mod; }
            }
        }
        factor = (std::uint64_t)((u128)e1 * f1 % t);
    }
    // rotateInternal with the NAF fallback for missing keys (src/evaluator_cuda.cu:2119-2176)
    void rotateInternal(CiphertextCuda &ct, int steps, const GaloisKeysCuda &gk) const {
        if (gk.parmsID() != context_.keyParmsID()) throw std::invalid_argument("galois_keys is not valid for encryption parameters");
        if (steps == 0) return;
        const std::uint32_t elt = eltFromStep(steps);
        if (gk.hasKey(elt)) { applyGaloisInplace(ct, elt, gk); return; }
        std::vector<int> naf = util::naf(steps);
        if (naf.size() == 1) throw std::invalid_argument("Galois key not present");
        const int half = (int)(ct.polyModulusDegree() >> 1);
        for (int s : naf) if (std::abs(s) != half) rotateInternal(ct, s, gk);
    }
    // modSwitchScaleToNext (src/evaluator_cuda.cu:749-825)
    void scaleToNext(const CiphertextCuda &ct, CiphertextCuda &dst) const {
        const SchemeType s = scheme();
        const int lv = level(ct);
        if (lv + 1 >= (int)context_.chainLength()) throw std::invalid_argument("end of modulus switching chain reached");
        if (s == SchemeType::ckks && !ct.isNttForm()) throw std::invalid_argument("CKKS encrypted must be in NTT form");
        if (s != SchemeType::ckks && ct.isNttForm()) throw std::invalid_argument(s == SchemeType::bfv ? "BFV encrypted cannot be in NTT form" : "BGV encrypted cannot be in NTT form");
        const std::size_t n = ct.polyModulusDegree(), L = ct.coeffModulusSize();
        b200::DeviceWords out(ct.size() * (L - 1) * n);
        b200::check(troyb_mod_switch_scale_to_next(h(), lv, ct.data(), ct.size() * L * n, out.get(), ct.size() * (L - 1) * n, ct.size(), 1, b200::stream()));
        const auto &q = context_.getContextData(ct.parmsID())->parms().coeffModulus();
        const std::uint64_t q_last = q.back().value();
        CiphertextCuda r = metaOf(ct);
        r.adopt(std::move(out), ct.size(), n, L - 1);
        r.parmsID() = context_.parmsIDOf(lv + 1);
        if (s == SchemeType::ckks) r.scale() = ct.scale() / static_cast<double>(q_last);
        if (s == SchemeType::bgv) {
            const std::uint64_t t = context_.firstContextData()->parms().plainModulus().value();
            std::uint64_t inv = 0;
            if (!util::tryInvertUintMod(q_last % t, t, inv)) throw std::logic_error("invalid plain modulus");
            r.correctionFactor() = (std::uint64_t)((unsigned __int128)ct.correctionFactor() * inv % t);
        }
        dst = std::move(r);
    }

    SEALContextCuda context_;
};

inline void PublicKeyCuda::save(std::ostream &stream) const { CiphertextCuda(data()).save(stream); }
inline void PublicKeyCuda::load(std::istream &stream) { CiphertextCuda c; c.load(stream); data() = c.cpu(); }

inline void CiphertextCuda::saveTerms(std::ostream &stream, EvaluatorCuda &evaluator, const std::vector<std::size_t> &term_ids) const {
    CiphertextCuda copy = *this;
    if (is_ntt_) evaluator.transformFromNttInplace(copy);
    std::vector<std::uint64_t> w(size_ * n_ * limbs_);
    copy.data_.download(w.data(), w.size());
    // the header records the form of the ORIGINAL ciphertext (src/ciphertext_cuda.cu:46-47)
    saveHeader(stream, true);
    if (seed_ != 0) throw std::invalid_argument("Seed is not zero."); // after the header, as the reference (:64-66)
    for (std::size_t id : term_ids)
        for (std::size_t j = 0; j < limbs_; j++) stream.write(reinterpret_cast<const char *>(&w[j * n_ + id]), 8);
    const std::size_t offset = n_ * limbs_, count = w.size() - offset;
    stream.write(reinterpret_cast<const char *>(&count), sizeof(count));
    stream.write(reinterpret_cast<const char *>(w.data() + offset), count * 8);
}
inline void CiphertextCuda::loadTerms(std::istream &stream, EvaluatorCuda &evaluator, const std::vector<std::size_t> &term_ids) {
    if (!loadHeader(stream)) throw std::invalid_argument("Trying to load a normal ciphertext, but term indices is specified");
    std::vector<std::uint64_t> w(size_ * n_ * limbs_, 0);
    for (std::size_t id : term_ids)
        for (std::size_t j = 0; j < limbs_; j++) stream.read(reinterpret_cast<char *>(&w[j * n_ + id]), 8);
    std::size_t count = 0;
    stream.read(reinterpret_cast<char *>(&count), sizeof(count));
    const std::size_t offset = n_ * limbs_;
    if (offset + count > w.size()) throw std::invalid_argument("ciphertext data is too long");
    stream.read(reinterpret_cast<char *>(w.data() + offset), count * 8);
    data_ = b200::DeviceWords(w.size());
    data_.upload(w.data(), w.size());
    if (is_ntt_) {
        is_ntt_ = false;
        evaluator.transformToNttInplace(*this);
    }
}

} // namespace troy

namespace troyn {
using troy::ParmsID;
using troy::SchemeType;
using troy::SecurityLevel;
using troy::Modulus;
using troy::CoeffModulus;
using troy::PlainModulus;
using EncryptionParameters = troy::EncryptionParametersCuda;
using SEALContext = troy::SEALContextCuda;
using Plaintext = troy::PlaintextCuda;
using Ciphertext = troy::CiphertextCuda;
using LWECiphertext = troy::LWECiphertextCuda;
using Encryptor = troy::EncryptorCuda;
using Decryptor = troy::DecryptorCuda;
using Evaluator = troy::EvaluatorCuda;
using KeyGenerator = troy::KeyGeneratorCuda;
using PublicKey = troy::PublicKeyCuda;
using SecretKey = troy::SecretKeyCuda;
using KSwitchKeys = troy::KSwitchKeysCuda;
using RelinKeys = troy::RelinKeysCuda;
using GaloisKeys = troy::GaloisKeysCuda;
using CKKSEncoder = troy::CKKSEncoderCuda;
using BatchEncoder = troy::BatchEncoderCuda;
using KernelProvider = troy::KernelProvider;
} // namespace troyn
