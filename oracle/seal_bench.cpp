// TEST / BASELINE INFRASTRUCTURE ONLY — never linked into the product (troy_b200/).
//
// SEAL 4.0 (the copy the reference vendors under extern/SEAL) as a second CPU baseline for bench.py, next to the reference's
// own troy:: CPU path: the same workload (CKKS N=16384, L=8 + one special prime {60,40x7,60}, scale 2^40), one independent
// Evaluator loop per host thread, multiply + relinearize_inplace and rotate_vector on real encryptions.
// Follows the pattern of the reference's test/timetest_seal.cpp (TimeTestCKKS: testMultiplyRescale / testRotateVector).
//
// usage: seal_bench <threads> <reps per thread>      -> one JSON line on stdout
#include <seal/seal.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

using namespace seal;

int main(int argc, char **argv) {
    int threads = argc > 1 ? std::atoi(argv[1]) : 1;
    int reps = argc > 2 ? std::atoi(argv[2]) : 8;
    const size_t n = 16384;
    std::vector<int> bits{60};
    for (int i = 0; i < 7; i++) bits.push_back(40);
    bits.push_back(60);

    EncryptionParameters parms(scheme_type::ckks);
    parms.set_poly_modulus_degree(n);
    parms.set_coeff_modulus(CoeffModulus::Create(n, bits));
    SEALContext context(parms, true, sec_level_type::none);
    KeyGenerator keygen(context);
    PublicKey pk;
    keygen.create_public_key(pk);
    RelinKeys rlk;
    keygen.create_relin_keys(rlk);
    GaloisKeys gk;
    keygen.create_galois_keys(std::vector<int>{1}, gk);
    Encryptor encryptor(context, pk);
    CKKSEncoder encoder(context);
    const double scale = std::pow(2.0, 40);

    std::vector<double> slots(encoder.slot_count());
    for (size_t i = 0; i < slots.size(); i++) slots[i] = double(i % 64);
    Plaintext pa, pb;
    encoder.encode(slots, scale, pa);
    encoder.encode(slots, scale, pb);
    Ciphertext a, b;
    encryptor.encrypt(pa, a);
    encryptor.encrypt(pb, b);

    auto run = [&](bool rotate) {
        std::vector<std::thread> pool;
        auto t0 = std::chrono::steady_clock::now();
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, rotate] {
                Evaluator evaluator(context);
                Ciphertext d;
                for (int r = 0; r < reps; r++) {
                    if (rotate) {
                        evaluator.rotate_vector(a, 1, gk, d);
                    } else {
                        evaluator.multiply(a, b, d);
                        evaluator.relinearize_inplace(d, rlk);
                    }
                }
            });
        for (auto &th : pool) th.join();
        std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
        return double(threads) * reps / dt.count();
    };
    run(false);  // warm the memory pools
    double mr = run(false);
    double rot = run(true);
    std::printf("{\"impl\": \"SEAL 4.0.0 (extern/SEAL of the reference, plain g++ -O3, no HEXL)\", \"workload\": \"CKKS N=16384 L=8 "
                "{60,40x7,60}\", \"threads\": %d, \"reps_per_thread\": %d, \"mult_relin_ops_per_s\": %.2f, \"rotate_ops_per_s\": %.2f}\n",
                threads, reps, mr, rot);
    return 0;
}
