// TEST CODE. Bit-exact parity of the troyn:: drop-in layer (include/troyn/troy_cuda.cuh over libtroy_b200.so) against the
// reference's own troy:: CPU evaluator, on identical seeded keys and host-encrypted inputs: every output word and the
// metadata are compared after each operation. Built by tests/cpp/build_dropin_tests.py next to the reference's own
// CUDA tests (which are compiled from where they lie in the reference tree; nothing is copied into this repository).
#include "troy_cuda.cuh"

#include "gtest/gtest.h"

#include <cstring>
#include <random>
#include <sstream>

namespace {

using namespace troy;

std::shared_ptr<UniformRandomGeneratorFactory> seeded() {
    return std::make_shared<Blake2xbPRNGFactory>(PRNGSeed{1, 2, 3, 4, 5, 6, 7, 8});
}

void expectSame(const troyn::Ciphertext &dev, const Ciphertext &host, const troyn::SEALContext &, const char *what) {
    Ciphertext got = dev.cpu();
    ASSERT_EQ(got.size(), host.size()) << what;
    ASSERT_TRUE(got.parmsID() == host.parmsID()) << what;
    ASSERT_EQ(got.isNttForm(), host.isNttForm()) << what;
    ASSERT_EQ(got.scale(), host.scale()) << what;
    ASSERT_EQ(got.correctionFactor(), host.correctionFactor()) << what;
    const size_t words = host.size() * host.polyModulusDegree() * host.coeffModulusSize();
    ASSERT_EQ(0, std::memcmp(got.data(), host.data(), words * 8)) << what << ": ciphertext words differ";
}

TEST(DropInParity, CKKSMultiplyRelinRescaleRotateConjugate) {
    troyn::KernelProvider::initialize();
    const size_t n = 8192;
    troyn::EncryptionParameters parms(SchemeType::ckks);
    parms.setPolyModulusDegree(n);
    parms.setCoeffModulus(CoeffModulus::Create(n, {60, 40, 40, 60}));
    parms.setRandomGenerator(seeded());
    troyn::SEALContext context(parms, true, SecurityLevel::none);
    troyn::KeyGenerator keygen(context);
    troyn::PublicKey pk;
    keygen.createPublicKey(pk);
    // the same host key objects feed both evaluators
    RelinKeys h_rlk;
    GaloisKeys h_gk;
    keygen.host().createRelinKeys(h_rlk);
    keygen.host().createGaloisKeys(std::vector<int>{1, -2}, h_gk);
    GaloisKeys h_conj;
    troyn::RelinKeys rlk(h_rlk);
    troyn::GaloisKeys gk(h_gk);

    CKKSEncoder encoder(context.host());
    Encryptor encryptor(context.host(), pk);
    Evaluator cpu(context.host());
    troyn::Evaluator gpu(context);

    std::vector<std::complex<double>> v1(n / 2), v2(n / 2);
    for (size_t i = 0; i < n / 2; i++) { v1[i] = double(i % 64); v2[i] = double((3 * i) % 17); }
    Plaintext p1, p2;
    const double scale = std::pow(2.0, 40);
    encoder.encode(v1, context.firstParmsID(), scale, p1);
    encoder.encode(v2, context.firstParmsID(), scale, p2);
    Ciphertext c1, c2;
    encryptor.encrypt(p1, c1);
    encryptor.encrypt(p2, c2);
    troyn::Ciphertext d1(c1), d2(c2);

    cpu.multiplyInplace(c1, c2);
    gpu.multiplyInplace(d1, d2);
    expectSame(d1, c1, context, "multiply");
    cpu.relinearizeInplace(c1, h_rlk);
    gpu.relinearizeInplace(d1, rlk);
    expectSame(d1, c1, context, "relinearize");
    cpu.rescaleToNextInplace(c1);
    gpu.rescaleToNextInplace(d1);
    expectSame(d1, c1, context, "rescale");
    cpu.rotateVectorInplace(c1, 1, h_gk);
    gpu.rotateVectorInplace(d1, 1, gk);
    expectSame(d1, c1, context, "rotate_vector(1)");
    cpu.rotateVectorInplace(c1, -2, h_gk);
    gpu.rotateVectorInplace(d1, -2, gk);
    expectSame(d1, c1, context, "rotate_vector(-2)");
    // NAF fallback: 3 = 4 - 1 needs keys for 4 and -1, which are absent -> both must throw the same kind of error
    EXPECT_THROW(cpu.rotateVectorInplace(c1, 3, h_gk), std::invalid_argument);
    EXPECT_THROW(gpu.rotateVectorInplace(d1, 3, gk), std::invalid_argument);
    cpu.squareInplace(c2);
    gpu.squareInplace(d2);
    expectSame(d2, c2, context, "square");
    cpu.transformFromNttInplace(c2);
    gpu.transformFromNttInplace(d2);
    expectSame(d2, c2, context, "transform_from_ntt");

    // decrypt the GPU result on the host: the slots must be the rotated product
    Decryptor decryptor(context.host(), keygen.secretKey());
    Plaintext out;
    decryptor.decrypt(d1.cpu(), out);
    std::vector<std::complex<double>> slots;
    encoder.decode(out, slots);
    for (size_t i = 0; i < 32; i++) {
        const size_t src = (i + n / 2 - 1) % (n / 2); // rotate by 1 then by -2  ==  by -1
        EXPECT_NEAR(slots[i].real(), v1[src].real() * v2[src].real(), 0.5);
    }
}

template <SchemeType S> void integerSchemeFlow() {
    troyn::KernelProvider::initialize();
    const size_t n = 4096;
    troyn::EncryptionParameters parms(S);
    parms.setPolyModulusDegree(n);
    parms.setCoeffModulus(CoeffModulus::Create(n, {40, 36, 40}));
    parms.setPlainModulus(PlainModulus::Batching(n, 20));
    parms.setRandomGenerator(seeded());
    troyn::SEALContext context(parms, true, SecurityLevel::none);
    troyn::KeyGenerator keygen(context);
    troyn::PublicKey pk;
    keygen.createPublicKey(pk);
    RelinKeys h_rlk;
    GaloisKeys h_gk;
    keygen.host().createRelinKeys(h_rlk);
    keygen.host().createGaloisKeys(h_gk);
    troyn::RelinKeys rlk(h_rlk);
    troyn::GaloisKeys gk(h_gk);
    BatchEncoder encoder(context.host());
    Encryptor encryptor(context.host(), pk);
    Evaluator cpu(context.host());
    troyn::Evaluator gpu(context);
    std::vector<uint64_t> v(n);
    for (size_t i = 0; i < n; i++) v[i] = i % 64;
    Plaintext p;
    encoder.encode(v, p);
    Ciphertext c1, c2;
    encryptor.encrypt(p, c1);
    encryptor.encrypt(p, c2);
    troyn::Ciphertext d1(c1), d2(c2);
    cpu.multiplyInplace(c1, c2);
    gpu.multiplyInplace(d1, d2);
    expectSame(d1, c1, context, "multiply");
    cpu.relinearizeInplace(c1, h_rlk);
    gpu.relinearizeInplace(d1, rlk);
    expectSame(d1, c1, context, "relinearize");
    cpu.rotateRowsInplace(c1, 3, h_gk);
    gpu.rotateRowsInplace(d1, 3, gk);
    expectSame(d1, c1, context, "rotate_rows(3)");
    cpu.rotateColumnsInplace(c1, h_gk);
    gpu.rotateColumnsInplace(d1, gk);
    expectSame(d1, c1, context, "rotate_columns");
    cpu.modSwitchToNextInplace(c1);
    gpu.modSwitchToNextInplace(d1);
    expectSame(d1, c1, context, "mod_switch_to_next");
    cpu.addInplace(c1, c1);
    gpu.addInplace(d1, d1);
    expectSame(d1, c1, context, "add");
    cpu.negateInplace(c1);
    gpu.negateInplace(d1);
    expectSame(d1, c1, context, "negate");
    Decryptor decryptor(context.host(), keygen.secretKey());
    Plaintext out;
    decryptor.decrypt(d1.cpu(), out);
    std::vector<uint64_t> slots;
    encoder.decode(out, slots);
    const uint64_t t = parms.plainModulus().value();
    // row rotation by 3 then column swap: slot i of row 0 came from row 1, index i+3
    const size_t row = n / 2;
    for (size_t i = 0; i < 16; i++) {
        const uint64_t x = v[row + (i + 3) % row];
        EXPECT_EQ(slots[i], (t - (2 * x * x) % t) % t);
    }
}

TEST(DropInParity, BFVMultiplyRelinRotateModSwitch) { integerSchemeFlow<SchemeType::bfv>(); }
TEST(DropInParity, BGVMultiplyRelinRotateModSwitch) { integerSchemeFlow<SchemeType::bgv>(); }

// Encryption on the device through the drop-in EncryptorCuda: with the parameters' generator factory seeded, every word
// equals the reference CPU Encryptor's (asymmetric and symmetric, encrypt-zero at each level and encrypt of a plaintext).
template <SchemeType scheme> void encryptFlow() {
    troyn::KernelProvider::initialize();
    const size_t n = 4096;
    troyn::EncryptionParameters parms(scheme);
    parms.setPolyModulusDegree(n);
    parms.setCoeffModulus(CoeffModulus::Create(n, {60, 40, 40, 60}));
    if (scheme != SchemeType::ckks) parms.setPlainModulus(PlainModulus::Batching(n, 20));
    parms.setRandomGenerator(seeded());
    troyn::SEALContext context(parms, true, SecurityLevel::none);
    troyn::KeyGenerator keygen(context);
    troyn::PublicKey pk;
    keygen.createPublicKey(pk);
    Encryptor cpu(context.host(), pk, keygen.secretKey());
    troyn::Encryptor gpu(context, pk, keygen.secretKey());
    gpu.setSeededSymmetric(false); // the CPU path's symmetric encryption (c1 from the Blake2xb ciphertext generator)
    for (auto cd = context.firstContextData(); cd; cd = cd->nextContextData()) {
        Ciphertext want;
        troyn::Ciphertext got;
        cpu.encryptZero(cd->parmsID(), want);
        gpu.encryptZero(cd->parmsID(), got);
        expectSame(got, want, context, "encryptZero");
        cpu.encryptZeroSymmetric(cd->parmsID(), want);
        gpu.encryptZeroSymmetric(cd->parmsID(), got);
        expectSame(got, want, context, "encryptZeroSymmetric");
    }
    Plaintext plain;
    if (scheme == SchemeType::ckks) {
        CKKSEncoder encoder(context.host());
        std::vector<std::complex<double>> v(n / 2);
        for (size_t i = 0; i < n / 2; i++) v[i] = {double(i % 31), double(i % 7)};
        encoder.encode(v, context.firstParmsID(), std::pow(2.0, 40), plain);
    } else {
        BatchEncoder encoder(context.host());
        std::vector<uint64_t> v(n);
        for (size_t i = 0; i < n; i++) v[i] = (7 * i + 3) % 1000;
        encoder.encode(v, plain);
    }
    Ciphertext want;
    troyn::Ciphertext got;
    cpu.encrypt(plain, want);
    gpu.encrypt(plain, got);
    expectSame(got, want, context, "encrypt");
    cpu.encryptSymmetric(plain, want);
    gpu.encryptSymmetric(plain, got);
    expectSame(got, want, context, "encryptSymmetric");
    // the CUDA path's seeded form: save writes c0 + seed (half the words), load(stream, context) re-expands c1
    gpu.setSeededSymmetric(true);
    troyn::Ciphertext seeded;
    gpu.encryptSymmetric(plain, seeded);
    ASSERT_NE(seeded.seed(), 0u);
    std::stringstream ss, full;
    seeded.save(ss);
    got.save(full);
    ASSERT_LT(ss.str().size(), full.str().size() / 2 + 128);
    troyn::Ciphertext reloaded;
    ASSERT_THROW({ std::stringstream again(ss.str()); reloaded.load(again); }, std::invalid_argument);
    reloaded.load(ss, context);
    {
        Ciphertext a = seeded.cpu(), b = reloaded.cpu();
        ASSERT_EQ(0, std::memcmp(a.data(), b.data(), 2 * a.polyModulusDegree() * a.coeffModulusSize() * 8)) << "seeded round trip";
    }
    got = reloaded;
    // and the device decryptor gives the plaintext back
    troyn::Decryptor dec(context, keygen.secretKey());
    Plaintext back;
    dec.decrypt(got, back);
    if (scheme != SchemeType::ckks) {
        for (size_t i = 0; i < plain.coeffCount(); i++) ASSERT_EQ(i < back.coeffCount() ? back.data()[i] : 0, plain.data()[i]) << "coefficient " << i;
    }
}
TEST(DropInParity, CKKSEncryptOnDevice) { encryptFlow<SchemeType::ckks>(); }
TEST(DropInParity, BFVEncryptOnDevice) { encryptFlow<SchemeType::bfv>(); }
TEST(DropInParity, BGVEncryptOnDevice) { encryptFlow<SchemeType::bgv>(); }

} // namespace
