// TEST INFRASTRUCTURE ONLY — never linked into the product (troy_b200/).
//
// Drives the UNMODIFIED reference CUDA path (the troyn:: classes of /root/reference/src, compiled for sm_100a into
// oracle/_ref/libtroyref_cuda.so by oracle/Makefile) on the GPU box:
//
//   troyn_ref_bench bench <batch> <iters>
//       CKKS N=16384 {60,40x7,60} (BASELINE configs[2]): multiply+relinearize, rotate_vector, rescale_to_next and the
//       standalone NTT on `batch` synthetic ciphertexts, one reference call per ciphertext (the reference has no batch
//       API), device synchronised before and after each timed region. Prints one JSON line.
//   troyn_ref_bench dump <in> <out>
//       Parity: reads {scheme, N, n_moduli, t, moduli[], two size-2 level-0 ciphertexts} (uint64 words), generates keys with
//       the seeded host KeyGenerator exactly as oracle/ref_shim.cpp does (same PRNG seed, same call order), runs
//       multiply -> relinearize -> (CKKS: rescale -> rotate_vector(1) | BFV/BGV: rotate_rows(1) -> mod_switch_to_next)
//       and writes the ciphertext words after every step.
#include "troy_cuda.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

using namespace troyn;
typedef troy::SEALContext HostContext;
typedef troy::Ciphertext HostCiphertext;
typedef troy::EncryptionParameters HostParms;

static HostParms make_parms(int scheme, size_t n, const std::vector<uint64_t> &moduli, uint64_t t) {
    HostParms parms(static_cast<uint8_t>(scheme));
    parms.setPolyModulusDegree(n);
    std::vector<troy::Modulus> q;
    for (uint64_t m : moduli) q.emplace_back(m);
    parms.setCoeffModulus(q);
    if (scheme != 2) parms.setPlainModulus(t);
    troy::PRNGSeed seed;
    for (size_t i = 0; i < seed.size(); i++) seed[i] = i + 1; // tests/common.py SEED
    parms.setRandomGenerator(std::make_shared<troy::Blake2xbPRNGFactory>(seed));
    return parms;
}

static Ciphertext upload(const HostContext &hc, const uint64_t *words, size_t size, bool ntt, double scale) {
    HostCiphertext h;
    h.resize(hc, hc.firstParmsID(), size);
    h.isNttForm() = ntt;
    h.scale() = scale;
    std::memcpy(h.data(), words, sizeof(uint64_t) * size * h.coeffModulusSize() * h.polyModulusDegree());
    return Ciphertext(h);
}

static void append(std::vector<uint64_t> &out, const Ciphertext &ct) {
    HostCiphertext h = ct.cpu();
    size_t words = h.size() * h.coeffModulusSize() * h.polyModulusDegree();
    out.push_back(words);
    out.insert(out.end(), h.data(), h.data() + words);
}

static int dump(const char *in_path, const char *out_path) {
    std::ifstream in(in_path, std::ios::binary);
    uint64_t hdr[4];
    in.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    const int scheme = (int)hdr[0];
    const size_t n = hdr[1], nmod = hdr[2];
    const uint64_t t = hdr[3];
    std::vector<uint64_t> moduli(nmod);
    in.read(reinterpret_cast<char *>(moduli.data()), nmod * 8);
    const size_t L = nmod - 1, ct_words = 2 * L * n;
    std::vector<uint64_t> a(ct_words), b(ct_words);
    in.read(reinterpret_cast<char *>(a.data()), ct_words * 8);
    in.read(reinterpret_cast<char *>(b.data()), ct_words * 8);

    KernelProvider::initialize();
    HostContext hc(make_parms(scheme, n, moduli, t), true, troy::SecurityLevel::none);
    SEALContext context(hc);
    // key generation order of oracle/ref_shim.cpp: KeyGenerator, public key, relin keys, Galois keys {step 1, conj}
    KeyGenerator keygen(context);
    PublicKey pk;
    keygen.createPublicKey(pk);
    RelinKeys rlk;
    keygen.createRelinKeys(rlk);
    const troy::util::GaloisTool *gt = hc.keyContextData()->galoisTool();
    std::vector<uint32_t> elts{gt->getEltFromStep(1), gt->getEltFromStep(0)};
    GaloisKeys gk;
    keygen.createGaloisKeys(elts, gk);
    Evaluator ev(context);
    const bool ckks = scheme == 2;
    const double sc = ckks ? 1099511627776.0 : 1.0;
    Ciphertext ca = upload(hc, a.data(), 2, ckks, sc), cb = upload(hc, b.data(), 2, ckks, sc);
    std::vector<uint64_t> out;
    ev.multiplyInplace(ca, cb);
    append(out, ca);
    ev.relinearizeInplace(ca, rlk);
    append(out, ca);
    if (ckks) {
        ev.rescaleToNextInplace(ca);
        append(out, ca);
        ev.rotateVectorInplace(ca, 1, gk);
        append(out, ca);
    } else {
        ev.rotateRowsInplace(ca, 1, gk);
        append(out, ca);
        ev.modSwitchToNextInplace(ca);
        append(out, ca);
    }
    std::ofstream o(out_path, std::ios::binary);
    o.write(reinterpret_cast<const char *>(out.data()), out.size() * 8);
    std::printf("dump ok: %zu words\n", out.size());
    return 0;
}

// Wire-format parity: the reference saves a ciphertext and its relinearisation keys (<dir>/ct.bin, <dir>/rlk.bin), and
// loads <dir>/ours_ct.bin + <dir>/ours_rlk.bin (written by troy_b200/serialize.py), squares and relinearises with them and
// writes the result words to <dir>/ours_result.bin.
static int serialize(const char *in_path, const char *dir) {
    std::ifstream in(in_path, std::ios::binary);
    uint64_t hdr[4];
    in.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    const int scheme = (int)hdr[0];
    const size_t n = hdr[1], nmod = hdr[2];
    std::vector<uint64_t> moduli(nmod);
    in.read(reinterpret_cast<char *>(moduli.data()), nmod * 8);
    const size_t L = nmod - 1, ct_words = 2 * L * n;
    std::vector<uint64_t> a(ct_words);
    in.read(reinterpret_cast<char *>(a.data()), ct_words * 8);
    KernelProvider::initialize();
    HostContext hc(make_parms(scheme, n, moduli, hdr[3]), true, troy::SecurityLevel::none);
    SEALContext context(hc);
    KeyGenerator keygen(context);
    PublicKey pk;
    keygen.createPublicKey(pk);
    RelinKeys rlk;
    keygen.createRelinKeys(rlk);
    const bool ckks = scheme == 2;
    const double sc = ckks ? 1048576.0 : 1.0;
    Ciphertext ca = upload(hc, a.data(), 2, ckks, sc);
    const std::string d(dir);
    { std::ofstream f(d + "/ct.bin", std::ios::binary); ca.save(f); }
    { std::ofstream f(d + "/rlk.bin", std::ios::binary); rlk.save(f); }
    Ciphertext c2;
    RelinKeys r2;
    { std::ifstream f(d + "/ours_ct.bin", std::ios::binary); c2.load(f); }
    { std::ifstream f(d + "/ours_rlk.bin", std::ios::binary); r2.load(f); }
    Evaluator ev(context);
    ev.squareInplace(c2);
    ev.relinearizeInplace(c2, r2);
    std::vector<uint64_t> out;
    append(out, c2);
    std::ofstream o(d + "/ours_result.bin", std::ios::binary);
    o.write(reinterpret_cast<const char *>(out.data()), out.size() * 8);
    // Seeded ciphertexts (the transport form of fresh symmetric encryptions): the reference encrypts zero symmetrically, saves
    // (c0 + seed), reloads with the context (c1 re-expanded by its cuRAND sampler) and writes the words; then it loads
    // <dir>/ours_seeded_ct.bin (a seeded ciphertext written by troy_b200) the same way, decrypts it and writes the plaintext.
    {
        Encryptor enc(context, pk, keygen.secretKey());
        Ciphertext z;
        enc.encryptZeroSymmetric(z);
        { std::ofstream f(d + "/seeded_ct.bin", std::ios::binary); z.save(f); }
        Ciphertext z2;
        { std::ifstream f(d + "/seeded_ct.bin", std::ios::binary); z2.load(f, context); }
        std::vector<uint64_t> w;
        append(w, z2);
        std::ofstream o1(d + "/seeded_expanded.bin", std::ios::binary);
        o1.write(reinterpret_cast<const char *>(w.data()), w.size() * 8);
        std::ifstream f(d + "/ours_seeded_ct.bin", std::ios::binary);
        if (f.good()) {
            Ciphertext mine;
            mine.load(f, context);
            Decryptor dec(context, keygen.secretKey());
            Plaintext p;
            dec.decrypt(mine, p);
            troy::Plaintext hp = p.cpu();
            std::ofstream o2(d + "/ours_seeded_plain.bin", std::ios::binary);
            o2.write(reinterpret_cast<const char *>(hp.data()), hp.coeffCount() * 8);
        }
    }
    std::printf("serialize ok\n");
    return 0;
}

template <class F> static double timed(F f, int iters) {
    f();
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) f();
    cudaDeviceSynchronize();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / iters;
}

static int bench(int batch, int iters) {
    const size_t n = 16384;
    std::vector<int> bits{60, 40, 40, 40, 40, 40, 40, 40, 60};
    std::vector<uint64_t> moduli;
    for (auto &m : troy::CoeffModulus::Create(n, bits)) moduli.push_back(m.value());
    const size_t L = moduli.size() - 1;
    KernelProvider::initialize();
    HostContext hc(make_parms(2, n, moduli, 0), true, troy::SecurityLevel::none);
    SEALContext context(hc);
    KeyGenerator keygen(context);
    RelinKeys rlk;
    keygen.createRelinKeys(rlk);
    GaloisKeys gk;
    keygen.createGaloisKeys(std::vector<int>{1}, gk);
    Evaluator ev(context);
    const double sc = 1099511627776.0;
    std::vector<Ciphertext> a, b;
    std::vector<uint64_t> w(2 * L * n);
    for (int i = 0; i < batch; i++) {
        for (int k = 0; k < 2; k++) {
            std::mt19937_64 rng(0x5EED0000 + 2 * i + k);
            for (size_t p = 0; p < 2; p++)
                for (size_t l = 0; l < L; l++)
                    for (size_t c = 0; c < n; c++) w[(p * L + l) * n + c] = rng() % moduli[l];
            (k ? b : a).push_back(upload(hc, w.data(), 2, true, sc));
        }
    }
    std::vector<Ciphertext> work(batch);
    double t_mr = timed([&] {
        for (int i = 0; i < batch; i++) {
            work[i] = a[i];
            ev.multiplyInplace(work[i], b[i]);
            ev.relinearizeInplace(work[i], rlk);
        }
    }, iters);
    double t_rot = timed([&] {
        for (int i = 0; i < batch; i++) {
            work[i] = a[i];
            ev.rotateVectorInplace(work[i], 1, gk);
        }
    }, iters);
    double t_resc = timed([&] {
        for (int i = 0; i < batch; i++) {
            Ciphertext r;
            ev.rescaleToNext(a[i], r);
        }
    }, iters);
    double t_ntt = timed([&] {
        for (int i = 0; i < batch; i++) {
            ev.transformFromNttInplace(a[i]);
            ev.transformToNttInplace(a[i]);
        }
    }, iters);
    std::printf("{\"impl\": \"reference troyn:: CUDA path compiled for sm_100a\", \"workload\": \"CKKS N=16384 L=8 {60,40x7,60}\", "
                "\"batch\": %d, \"mult_relin_ops_per_s\": %.1f, \"rotate_ops_per_s\": %.1f, \"rescale_ops_per_s\": %.1f, "
                "\"limb_ntt_per_s\": %.1f}\n",
                batch, batch / t_mr, batch / t_rot, batch / t_resc, batch * 2.0 * 2 * L / t_ntt);
    return 0;
}

int main(int argc, char **argv) {
    try {
        if (argc >= 4 && !std::strcmp(argv[1], "dump")) return dump(argv[2], argv[3]);
        if (argc >= 4 && !std::strcmp(argv[1], "serialize")) return serialize(argv[2], argv[3]);
        if (argc >= 4 && !std::strcmp(argv[1], "bench")) return bench(std::atoi(argv[2]), std::atoi(argv[3]));
        std::fprintf(stderr, "usage: troyn_ref_bench bench <batch> <iters> | dump <in> <out>\n");
        return 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
