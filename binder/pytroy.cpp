// pytroy — the reference's Python module (binder/binder.cu: PYBIND11_MODULE(pytroy, m)) for the classes of the RNS polynomial
// path, bound over the drop-in header include/troyn/troy_cuda.cuh (and so over libtroy_b200.so). Python-visible names and
// argument order follow the reference's module, so that scripts written for it (e.g. its binder/test.py) run unchanged;
// the Conv2d / BFV-matmul application helpers of the reference (app/LinearHelper.cuh) are outside the path and not bound.
// Serialisation goes through Python `bytes` in the reference's wire format.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "troy_cuda.cuh"

#include <complex>
#include <sstream>
#include <string>
#include <vector>

namespace py = pybind11;
using namespace troyn;
using cplx = std::complex<double>;

namespace {

template <class T> py::bytes saved(const T &obj) {
    std::ostringstream s;
    obj.save(s);
    return py::bytes(s.str());
}
template <class T> void loaded(T &obj, const py::bytes &b) {
    std::istringstream s(static_cast<std::string>(b));
    obj.load(s);
}

template <class T> std::vector<T> from_array(const py::array_t<T, py::array::c_style | py::array::forcecast> &a) {
    return std::vector<T>(a.data(), a.data() + a.size());
}
template <class T> py::array_t<T> to_array(const std::vector<T> &v) { return py::array_t<T>(v.size(), v.data()); }

// same-named "inplace" / "new object" pairs of the evaluator
template <class F> Ciphertext fresh(F f) {
    Ciphertext out;
    f(out);
    return out;
}

} // namespace

PYBIND11_MODULE(pytroy, m) {
    m.doc() = "troy_b200: the pytroy module of lightbulb128/troy over the B200-native library";
    m.def("initialize_kernel", [](int device) { KernelProvider::initialize(device); }, py::arg("device") = 0);

    py::enum_<SchemeType>(m, "SchemeType").value("none", SchemeType::none).value("bfv", SchemeType::bfv).value("ckks", SchemeType::ckks).value("bgv", SchemeType::bgv);
    py::enum_<SecurityLevel>(m, "SecurityLevel").value("none", SecurityLevel::none).value("tc128", SecurityLevel::tc128).value("tc192", SecurityLevel::tc192).value("tc256", SecurityLevel::tc256);

    py::class_<Modulus>(m, "Modulus").def(py::init<std::uint64_t>()).def("value", &Modulus::value).def("bit_count", &Modulus::bitCount);
    py::class_<CoeffModulus>(m, "CoeffModulus")
        .def_static("create", [](std::size_t n, const std::vector<int> &bits) { return CoeffModulus::Create(n, bits); })
        .def_static("bfv_default", [](std::size_t n, SecurityLevel sec) { return CoeffModulus::BFVDefault(n, sec); }, py::arg("poly_modulus_degree"), py::arg("sec") = SecurityLevel::tc128)
        .def_static("max_bit_count", [](std::size_t n, SecurityLevel sec) { return CoeffModulus::MaxBitCount(n, sec); }, py::arg("poly_modulus_degree"), py::arg("sec") = SecurityLevel::tc128);
    py::class_<PlainModulus>(m, "PlainModulus").def_static("batching", [](std::size_t n, int bits) { return PlainModulus::Batching(n, bits); });

    py::class_<ParmsID>(m, "ParmsID").def("vec", [](const ParmsID &p) { return std::vector<std::uint64_t>(p.begin(), p.end()); }).def("__eq__", [](const ParmsID &a, const ParmsID &b) { return a == b; });

    py::class_<EncryptionParameters>(m, "EncryptionParameters")
        .def(py::init<SchemeType>())
        .def("set_poly_modulus_degree", &EncryptionParameters::setPolyModulusDegree)
        .def("set_coeff_modulus", [](EncryptionParameters &p, const std::vector<Modulus> &q) { p.setCoeffModulus(q); })
        .def("set_plain_modulus", [](EncryptionParameters &p, const Modulus &t) { p.setPlainModulus(t); })
        .def("set_plain_modulus", [](EncryptionParameters &p, std::uint64_t t) { p.setPlainModulus(t); })
        .def("scheme", [](const EncryptionParameters &p) { return p.host().scheme(); })
        .def("poly_modulus_degree", [](const EncryptionParameters &p) { return p.host().polyModulusDegree(); })
        .def("coeff_modulus", [](const EncryptionParameters &p) { return p.host().coeffModulus(); })
        .def("plain_modulus", [](const EncryptionParameters &p) { return p.host().plainModulus(); });

    py::class_<SEALContext>(m, "SEALContext")
        .def(py::init<const EncryptionParameters &, bool, SecurityLevel>(), py::arg("parms"), py::arg("expand_mod_chain") = true, py::arg("sec_level") = SecurityLevel::tc128)
        .def("first_parms_id", [](const SEALContext &c) { return c.firstParmsID(); })
        .def("last_parms_id", [](const SEALContext &c) { return c.lastParmsID(); })
        .def("key_parms_id", [](const SEALContext &c) { return c.keyParmsID(); })
        .def("using_keyswitching", &SEALContext::using_keyswitching)
        .def("parms_id_of_level", [](const SEALContext &c, int level) { return c.parmsIDOf(level); })
        .def("chain_length", &SEALContext::chainLength);

    py::class_<Plaintext>(m, "Plaintext")
        .def(py::init<>())
        .def("coeff_count", [](const Plaintext &p) { return p.coeffCount(); })
        .def("is_ntt_form", [](const Plaintext &p) { return p.isNttForm(); })
        .def("parms_id", [](const Plaintext &p) { return p.parmsID(); })
        .def("scale", [](const Plaintext &p) { return p.scale(); })
        .def("set_scale", [](Plaintext &p, double s) { p.scale() = s; })
        .def("copy", [](const Plaintext &p) { return Plaintext(p); })
        .def("data", [](const Plaintext &p) { return to_array(std::vector<std::uint64_t>(p.data(), p.data() + p.coeffCount())); })
        .def("save", [](const Plaintext &p) { return saved(p); })
        .def("load", [](Plaintext &p, const py::bytes &b) { loaded(p, b); });

    py::class_<Ciphertext>(m, "Ciphertext")
        .def(py::init<>())
        .def("size", &Ciphertext::size)
        .def("coeff_modulus_size", &Ciphertext::coeffModulusSize)
        .def("poly_modulus_degree", &Ciphertext::polyModulusDegree)
        .def("is_ntt_form", [](const Ciphertext &c) { return c.isNttForm(); })
        .def("parms_id", [](const Ciphertext &c) { return c.parmsID(); })
        .def("scale", [](const Ciphertext &c) { return c.scale(); })
        .def("set_scale", [](Ciphertext &c, double s) { c.scale() = s; })
        .def("correction_factor", [](const Ciphertext &c) { return c.correctionFactor(); })
        .def("seed", [](const Ciphertext &c) { return c.seed(); })
        .def("copy", [](const Ciphertext &c) { return Ciphertext(c); })
        .def("data", [](const Ciphertext &c) { // the words as the host sees them, [size * limbs * N]
            troy::Ciphertext h = c.cpu();
            return to_array(std::vector<std::uint64_t>(h.data(), h.data() + c.size() * c.polyCoeffSize()));
        })
        .def("save", [](const Ciphertext &c) { return saved(c); })
        .def("load", [](Ciphertext &c, const py::bytes &b) { loaded(c, b); })
        .def("load", [](Ciphertext &c, const py::bytes &b, const SEALContext &context) {
            std::istringstream s(static_cast<std::string>(b));
            c.load(s, context);
        })
        .def("save_terms", [](const Ciphertext &c, Evaluator &ev, const std::vector<std::size_t> &terms) {
            std::ostringstream s;
            c.saveTerms(s, ev, terms);
            return py::bytes(s.str());
        })
        .def("load_terms", [](Ciphertext &c, const py::bytes &b, Evaluator &ev, const std::vector<std::size_t> &terms) {
            std::istringstream s(static_cast<std::string>(b));
            c.loadTerms(s, ev, terms);
        });

    py::class_<SecretKey>(m, "SecretKey").def(py::init<>()).def("save", [](const SecretKey &k) { return saved(k); }).def("load", [](SecretKey &k, const py::bytes &b) { loaded(k, b); });
    py::class_<PublicKey>(m, "PublicKey").def(py::init<>()).def("save", [](const PublicKey &k) { return saved(k); }).def("load", [](PublicKey &k, const py::bytes &b) { loaded(k, b); });
    py::class_<KSwitchKeys>(m, "KSwitchKeys").def(py::init<>()).def("save", [](const KSwitchKeys &k) { return saved(k); }).def("load", [](KSwitchKeys &k, const py::bytes &b) { loaded(k, b); });
    py::class_<RelinKeys>(m, "RelinKeys").def(py::init<>()).def("save", [](const RelinKeys &k) { return saved(k); }).def("load", [](RelinKeys &k, const py::bytes &b) { loaded(k, b); });
    py::class_<GaloisKeys>(m, "GaloisKeys").def(py::init<>()).def("save", [](const GaloisKeys &k) { return saved(k); }).def("load", [](GaloisKeys &k, const py::bytes &b) { loaded(k, b); });

    py::class_<KeyGenerator>(m, "KeyGenerator")
        .def(py::init<const SEALContext &>())
        .def(py::init<const SEALContext &, const SecretKey &>())
        .def("secret_key", &KeyGenerator::secretKey)
        .def("create_public_key", [](KeyGenerator &g) { PublicKey k; g.createPublicKey(k); return k; })
        .def("create_relin_keys", [](KeyGenerator &g) { RelinKeys k; g.createRelinKeys(k); return k; })
        .def("create_keyswitching_keys", [](KeyGenerator &g, const SecretKey &other) { return g.createKeySwitchingKeys(other); })
        .def("create_galois_keys", [](KeyGenerator &g) { GaloisKeys k; g.createGaloisKeys(k); return k; })
        .def("create_galois_keys", [](KeyGenerator &g, const std::vector<int> &steps) { GaloisKeys k; g.createGaloisKeys(steps, k); return k; })
        .def("create_automorphism_keys", [](KeyGenerator &g) { return g.createAutomorphismKeys(); });

    // binder/binder.cu:298-302
    py::class_<LWECiphertext>(m, "LWECiphertext")
        .def("copy", [](const LWECiphertext &c) { return LWECiphertext(c); })
        .def("poly_modulus_degree", &LWECiphertext::polyModulusDegree)
        .def("coeff_modulus_size", &LWECiphertext::coeffModulusSize);

    py::class_<BatchEncoder>(m, "BatchEncoder")
        .def(py::init<const SEALContext &>())
        .def("slot_count", [](const BatchEncoder &e) { return e.slotCount(); })
        .def("encode", [](BatchEncoder &e, const std::vector<std::uint64_t> &v, Plaintext &p) { e.encode(v, p); })
        .def("encode", [](BatchEncoder &e, const std::vector<std::uint64_t> &v) { Plaintext p; e.encode(v, p); return p; })
        .def("decode", [](BatchEncoder &e, const Plaintext &p) { std::vector<std::uint64_t> v; e.decode(p, v); return v; })
        .def("encode_polynomial", [](BatchEncoder &e, const std::vector<std::uint64_t> &v) { Plaintext p; e.encodePolynomial(v, p); return p; })
        .def("decode_polynomial", [](BatchEncoder &e, const Plaintext &p) { std::vector<std::uint64_t> v; e.decodePolynomial(p, v); return v; });

    py::class_<CKKSEncoder>(m, "CKKSEncoder")
        .def(py::init<const SEALContext &>())
        .def("slot_count", [](const CKKSEncoder &e) { return e.slotCount(); })
        .def("encode", [](CKKSEncoder &e, const std::vector<cplx> &v, double scale, Plaintext &p) { e.encode(v, scale, p); })
        .def("encode", [](CKKSEncoder &e, const std::vector<cplx> &v, const ParmsID &id, double scale, Plaintext &p) { e.encode(v, id, scale, p); })
        .def("encode", [](CKKSEncoder &e, const std::vector<cplx> &v, double scale) { Plaintext p; e.encode(v, scale, p); return p; })
        .def("encode_polynomial", [](CKKSEncoder &e, const std::vector<double> &v, double scale, Plaintext &p) { e.encodePolynomial(v, scale, p); })
        .def("encode_polynomial", [](CKKSEncoder &e, const std::vector<double> &v, const ParmsID &id, double scale, Plaintext &p) { e.encodePolynomial(v, id, scale, p); })
        .def("decode", [](CKKSEncoder &e, const Plaintext &p) { std::vector<cplx> v; e.decode(p, v); return v; })
        .def("decode_polynomial", [](CKKSEncoder &e, const Plaintext &p) { std::vector<double> v; e.decodePolynomial(p, v); return v; });

    py::class_<Encryptor>(m, "Encryptor")
        .def(py::init<const SEALContext &, const PublicKey &>())
        .def(py::init<const SEALContext &, const SecretKey &>())
        .def(py::init<const SEALContext &, const PublicKey &, const SecretKey &>())
        .def("set_public_key", &Encryptor::setPublicKey)
        .def("set_secret_key", &Encryptor::setSecretKey)
        .def("set_seeded_symmetric", &Encryptor::setSeededSymmetric)
        .def("encrypt", [](const Encryptor &e, const Plaintext &p, Ciphertext &c) { e.encrypt(p, c); })
        .def("encrypt", [](const Encryptor &e, const Plaintext &p) { return e.encrypt(p); })
        .def("encrypt_symmetric", [](const Encryptor &e, const Plaintext &p, Ciphertext &c) { e.encryptSymmetric(p, c); })
        .def("encrypt_symmetric", [](const Encryptor &e, const Plaintext &p) { return e.encryptSymmetric(p); })
        .def("encrypt_zero", [](const Encryptor &e, Ciphertext &c) { e.encryptZero(c); })
        .def("encrypt_zero", [](const Encryptor &e, const ParmsID &id, Ciphertext &c) { e.encryptZero(id, c); })
        .def("encrypt_zero_symmetric", [](const Encryptor &e, Ciphertext &c) { e.encryptZeroSymmetric(c); })
        .def("encrypt_zero_symmetric", [](const Encryptor &e, const ParmsID &id, Ciphertext &c) { e.encryptZeroSymmetric(id, c); });

    py::class_<Decryptor>(m, "Decryptor")
        .def(py::init<const SEALContext &, const SecretKey &>())
        .def("decrypt", [](Decryptor &d, const Ciphertext &c, Plaintext &p) { d.decrypt(c, p); })
        .def("decrypt", [](Decryptor &d, const Ciphertext &c) { Plaintext p; d.decrypt(c, p); return p; });

    py::class_<Evaluator>(m, "Evaluator")
        .def(py::init<const SEALContext &>())
        .def("negate_inplace", &Evaluator::negateInplace)
        .def("negate", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.negate(a, o); }); })
        .def("add_inplace", &Evaluator::addInplace)
        .def("add", [](const Evaluator &e, const Ciphertext &a, const Ciphertext &b) { return fresh([&](Ciphertext &o) { e.add(a, b, o); }); })
        .def("add_many", [](const Evaluator &e, const std::vector<Ciphertext> &v) { return fresh([&](Ciphertext &o) { e.addMany(v, o); }); })
        .def("sub_inplace", &Evaluator::subInplace)
        .def("sub", [](const Evaluator &e, const Ciphertext &a, const Ciphertext &b) { return fresh([&](Ciphertext &o) { e.sub(a, b, o); }); })
        .def("multiply_inplace", &Evaluator::multiplyInplace)
        .def("multiply", [](const Evaluator &e, const Ciphertext &a, const Ciphertext &b) { return fresh([&](Ciphertext &o) { e.multiply(a, b, o); }); })
        .def("square_inplace", &Evaluator::squareInplace)
        .def("square", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.square(a, o); }); })
        .def("relinearize_inplace", &Evaluator::relinearizeInplace)
        .def("relinearize", [](const Evaluator &e, const Ciphertext &a, const RelinKeys &k) { return fresh([&](Ciphertext &o) { e.relinearize(a, k, o); }); })
        .def("multiply_many", [](const Evaluator &e, const std::vector<Ciphertext> &v, const RelinKeys &k) { return fresh([&](Ciphertext &o) { e.multiplyMany(v, k, o); }); })
        .def("exponentiate_inplace", &Evaluator::exponentiateInplace)
        .def("exponentiate", [](const Evaluator &e, const Ciphertext &a, std::uint64_t n, const RelinKeys &k) { return fresh([&](Ciphertext &o) { e.exponentiate(a, n, k, o); }); })
        .def("apply_keyswitching_inplace", &Evaluator::applyKeySwitchingInplace)
        .def("apply_keyswitching", [](const Evaluator &e, const Ciphertext &a, const KSwitchKeys &k) { return fresh([&](Ciphertext &o) { e.applyKeySwitching(a, k, o); }); })
        .def("mod_switch_to_next_inplace", [](const Evaluator &e, Ciphertext &a) { e.modSwitchToNextInplace(a); })
        .def("mod_switch_to_next", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.modSwitchToNext(a, o); }); })
        .def("mod_switch_to_inplace", [](const Evaluator &e, Ciphertext &a, const ParmsID &id) { e.modSwitchToInplace(a, id); })
        .def("rescale_to_next_inplace", &Evaluator::rescaleToNextInplace)
        .def("rescale_to_next", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.rescaleToNext(a, o); }); })
        .def("rescale_to_inplace", &Evaluator::rescaleToInplace)
        .def("add_plain_inplace", &Evaluator::addPlainInplace)
        .def("add_plain", [](const Evaluator &e, const Ciphertext &a, const Plaintext &p) { return fresh([&](Ciphertext &o) { e.addPlain(a, p, o); }); })
        .def("sub_plain_inplace", &Evaluator::subPlainInplace)
        .def("sub_plain", [](const Evaluator &e, const Ciphertext &a, const Plaintext &p) { return fresh([&](Ciphertext &o) { e.subPlain(a, p, o); }); })
        .def("multiply_plain_inplace", &Evaluator::multiplyPlainInplace)
        .def("multiply_plain", [](const Evaluator &e, const Ciphertext &a, const Plaintext &p) { return fresh([&](Ciphertext &o) { e.multiplyPlain(a, p, o); }); })
        .def("transform_to_ntt_inplace", [](const Evaluator &e, Ciphertext &a) { e.transformToNttInplace(a); })
        .def("transform_to_ntt_inplace", [](const Evaluator &e, Plaintext &p, const ParmsID &id) { e.transformToNttInplace(p, id); })
        .def("transform_to_ntt", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.transformToNtt(a, o); }); })
        .def("transform_from_ntt_inplace", &Evaluator::transformFromNttInplace)
        .def("transform_from_ntt", [](const Evaluator &e, const Ciphertext &a) { return fresh([&](Ciphertext &o) { e.transformFromNtt(a, o); }); })
        .def("apply_galois_inplace", &Evaluator::applyGaloisInplace)
        .def("apply_galois", [](const Evaluator &e, const Ciphertext &a, std::uint32_t elt, const GaloisKeys &k) { return fresh([&](Ciphertext &o) { e.applyGalois(a, elt, k, o); }); })
        .def("rotate_rows_inplace", &Evaluator::rotateRowsInplace)
        .def("rotate_rows", [](const Evaluator &e, const Ciphertext &a, int steps, const GaloisKeys &k) { return fresh([&](Ciphertext &o) { e.rotateRows(a, steps, k, o); }); })
        .def("rotate_columns_inplace", &Evaluator::rotateColumnsInplace)
        .def("rotate_columns", [](const Evaluator &e, const Ciphertext &a, const GaloisKeys &k) { return fresh([&](Ciphertext &o) { e.rotateColumns(a, k, o); }); })
        .def("rotate_vector_inplace", &Evaluator::rotateVectorInplace)
        .def("rotate_vector", [](const Evaluator &e, const Ciphertext &a, int steps, const GaloisKeys &k) { return fresh([&](Ciphertext &o) { e.rotateVector(a, steps, k, o); }); })
        .def("complex_conjugate_inplace", &Evaluator::complexConjugateInplace)
        .def("complex_conjugate", [](const Evaluator &e, const Ciphertext &a, const GaloisKeys &k) { return fresh([&](Ciphertext &o) { e.complexConjugate(a, k, o); }); })
        // LWE extraction / packing (binder/binder.cu:712-722)
        .def("extract_lwe", &Evaluator::extractLWE)
        .def("assemble_lwe", &Evaluator::assembleLWE)
        .def("field_trace_inplace", &Evaluator::fieldTraceInplace)
        .def("divide_by_poly_modulus_degree_inplace", &Evaluator::divideByPolyModulusDegreeInplace, py::arg("ciphertext"), py::arg("mul") = 1)
        .def("pack_lwe_ciphertexts", &Evaluator::packLWECiphertexts)
        .def("negacyclic_shift", [](const Evaluator &e, const Ciphertext &a, std::size_t shift) { return fresh([&](Ciphertext &o) { e.negacyclicShift(a, shift, o); }); })
        .def("negacyclic_shift_inplace", &Evaluator::negacyclicShiftInplace);
}
