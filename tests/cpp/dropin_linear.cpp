// TEST INFRASTRUCTURE. The reference's own application header app/LinearHelperCKKS.cuh (MatmulHelper: encode weights,
// encrypt inputs, matmul = multiplyPlain + addInplace over the block grid, add a plaintext, serialise the outputs by terms,
// decrypt, read the outputs back), compiled UNCHANGED (its text minus the one line that includes the reference's own troy_cuda.cuh, assembled in a
// scratch directory at build time) against include/troyn/troy_cuda.cuh — the flow of
// test/app/linear_ckks.cu (testMatmul / testFullMatmul) as gtest cases with a numeric check.
#include "troy_cuda.cuh"

#include "LinearHelperCKKS.cuh" // the reference's app header; see tests/cpp/build_dropin_tests.py for how it is found

#include "gtest/gtest.h"

#include <random>
#include <sstream>

using namespace troyn;

namespace {

struct Fixture {
    size_t n, slots;
    double delta;
    std::unique_ptr<SEALContext> context;
    std::unique_ptr<KeyGenerator> keygen;
    std::unique_ptr<Encryptor> encryptor;
    std::unique_ptr<Decryptor> decryptor;
    std::unique_ptr<Evaluator> evaluator;
    std::unique_ptr<CKKSEncoder> encoder;
    PublicKey pk;
    Fixture(size_t n_, std::vector<int> bits, double delta_) : n(n_), slots(n_ / 2), delta(delta_) {
        KernelProvider::initialize();
        EncryptionParameters parms(SchemeType::ckks);
        parms.setPolyModulusDegree(n);
        parms.setCoeffModulus(CoeffModulus::Create(n, bits));
        context = std::make_unique<SEALContext>(parms);
        keygen = std::make_unique<KeyGenerator>(*context);
        keygen->createPublicKey(pk);
        encoder = std::make_unique<CKKSEncoder>(*context);
        encryptor = std::make_unique<Encryptor>(*context, pk);
        encryptor->setSecretKey(keygen->secretKey());
        decryptor = std::make_unique<Decryptor>(*context, keygen->secretKey());
        evaluator = std::make_unique<Evaluator>(*context);
    }
};

std::vector<double> random_reals(std::mt19937 &rng, size_t count, int bound) {
    std::vector<double> v(count);
    for (auto &x : v) x = (double)(rng() % bound);
    return v;
}

double max_error(const std::vector<double> &got, const std::vector<double> &x, const std::vector<double> &w, size_t B, size_t I, size_t O) {
    double worst = 0;
    for (size_t b = 0; b < B; b++)
        for (size_t k = 0; k < O; k++) {
            double y = 0;
            for (size_t j = 0; j < I; j++) y += x[b * I + j] * w[j * O + k];
            worst = std::max(worst, std::fabs(y - got[b * O + k]));
        }
    return worst;
}

} // namespace

TEST(DropInLinear, MatmulThroughTheReferenceAppHeader) {
    Fixture f(8192, {60, 40, 40, 60}, (double)(1 << 16));
    std::mt19937 rng(7);
    const size_t B = 3, I = 50, O = 30;
    auto weights = random_reals(rng, I * O, 64), x = random_reals(rng, B * I, 64);
    auto last = f.context->lastParmsID();
    LinearHelperCKKS::MatmulHelper helper(B, I, O, f.slots);
    helper.encodeWeights(*f.encoder, last, weights, f.delta);
    auto xEnc = helper.encryptInputs(*f.encryptor, *f.encoder, last, x, f.delta);
    auto yEnc = helper.matmul(*f.evaluator, xEnc);
    auto yDec = helper.decryptOutputs(*f.encoder, *f.decryptor, yEnc);
    EXPECT_LT(max_error(yDec, x, weights, B, I, O), 0.5);
}

TEST(DropInLinear, FullProtocolWithSerialisation) {
    Fixture f(8192, {60, 40, 40, 60}, (double)(1 << 16));
    std::mt19937 rng(9);
    const size_t B = 2, I = 40, O = 20;
    auto weights = random_reals(rng, I * O, 64), xc = random_reals(rng, B * I, 64), xs = random_reals(rng, B * I, 64);
    auto r = random_reals(rng, B * O, 64);
    auto last = f.context->lastParmsID();
    LinearHelperCKKS::MatmulHelper helper(B, I, O, f.slots);
    helper.encodeWeights(*f.encoder, last, weights, f.delta);
    auto xsEncoded = helper.encodeInputs(*f.encoder, last, xs, f.delta);
    auto rEncoded = helper.encodeOutputs(*f.encoder, last, r, f.delta * f.delta);
    auto xEnc = helper.encryptInputs(*f.encryptor, *f.encoder, last, xc, f.delta);
    { // client -> server
        std::ostringstream sout;
        xEnc.save(sout);
        std::istringstream sin(sout.str());
        xEnc = LinearHelperCKKS::Cipher2d();
        xEnc.load(sin, *f.context);
    }
    helper.addPlainInplace(*f.evaluator, xEnc, xsEncoded);
    auto yEnc = helper.matmul(*f.evaluator, xEnc);
    helper.addPlainInplace(*f.evaluator, yEnc, rEncoded);
    { // server -> client, only the coefficients that carry outputs
        std::ostringstream sout;
        helper.serializeOutputs(*f.evaluator, yEnc, sout);
        std::istringstream sin(sout.str());
        yEnc = helper.deserializeOutputs(*f.evaluator, sin);
    }
    auto yDec = helper.decryptOutputs(*f.encoder, *f.decryptor, yEnc);
    for (size_t i = 0; i < B * O; i++) yDec[i] -= r[i];
    std::vector<double> x(B * I);
    for (size_t i = 0; i < B * I; i++) x[i] = xc[i] + xs[i];
    EXPECT_LT(max_error(yDec, x, weights, B, I, O), 0.5);
}
