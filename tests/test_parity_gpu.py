"""`-m gpu`: the parity tests proper. Every call goes through the C ABI of the CUDA library; every output word and the
metadata are compared with the reference troy:: CPU library (oracle/_ref) on identical seeded keys and inputs."""
import numpy as np
import pytest

from common import (assert_same, make_pair, random_ct_words, run_decrypt_flow, run_encrypt_flow, run_evaluator_flow,
                    run_matmul_flow, run_plain_flow, run_prng_flow, run_sampler_flow)

pytestmark = pytest.mark.gpu

FLOWS = [
    # (scheme, N, coeff_modulus bits)            config in BASELINE.json
    (1, 8192, [60, 40, 40, 60]),                  # config 1: BFV N=8192 {60,40,40,60}
    (2, 16384, [60, 40, 40, 40, 40, 40, 40, 40, 60]),  # config 3: CKKS N=16384 L=8
    (1, 32768, [60, 50, 50, 50, 60]),             # config 4: BFV N=32768, 4 data primes
    (3, 32768, [60, 50, 50, 50, 60]),             # config 4: BGV
    (2, 65536, [60, 50, 50, 50, 60]),             # config 5 shape at reduced limb count
    (2, 4096, [60, 40, 40, 60]),
    (3, 8192, [60, 40, 40, 60]),
    (2, 64, [40, 30, 30, 40]),                    # the reference's toy sizes
    (1, 128, [40, 30, 30, 40]),
    (3, 32, [40, 30, 40]),
]


@pytest.mark.parametrize("scheme,n,bits", FLOWS)
def test_evaluator_flow(ref, cuda_lib, scheme, n, bits):
    run_evaluator_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(1000 + n + scheme))


def test_batched_flow(ref, cuda_lib):
    run_evaluator_flow(ref, cuda_lib, ref.CKKS, 16384, [60, 40, 40, 40, 60], np.random.default_rng(77), batch=5)
    run_evaluator_flow(ref, cuda_lib, ref.BFV, 8192, [60, 40, 40, 60], np.random.default_rng(78), batch=3)


def test_many_digits(ref, cuda_lib):
    """More than 8 digits: the key inner product runs over digit groups that are added up in the product buffer."""
    run_evaluator_flow(ref, cuda_lib, ref.CKKS, 1024, [40] * 10 + [50], np.random.default_rng(11), batch=2)
    run_evaluator_flow(ref, cuda_lib, ref.BFV, 512, [36] * 9 + [40], np.random.default_rng(12))
    run_evaluator_flow(ref, cuda_lib, ref.CKKS, 8192, [40] * 17 + [50], np.random.default_rng(13))


@pytest.mark.parametrize("scheme,n,bits", FLOWS)
def test_plain_ops(ref, cuda_lib, scheme, n, bits):
    """multiplyPlain / addPlain / subPlain and the fused multiply-accumulate."""
    run_plain_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(3000 + n + scheme))
    if len(bits) > 3:
        run_plain_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(4000 + n + scheme), batch=1, level=1)


@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768, 65536])
def test_ntt_sweep(ref, cuda_lib, n):
    """BASELINE config 2: the standalone batched NTT / INTT sweep, N = 4096 .. 65536 x 1 .. 32 limbs of 60-bit primes.
    Size-independent properties on a batch of 3 polys (round trip = identity, linearity) and direct parity of the first and
    the last limb with the reference's transform."""
    tb = cuda_lib
    from troy_b200.host import DeviceArray, _L, _vp
    rng = np.random.default_rng(n)
    for L in (1, 8, 32):
        q = ref.coeff_modulus_create(n, [60] * (L + 1))
        p = tb.EncryptionParameters(tb.SchemeType.ckks)
        p.set_poly_modulus_degree(n)
        p.set_coeff_modulus(q)
        ctx = tb.SEALContextCuda(p, expand_mod_chain=False)
        assert ctx.level_limbs(0) == L
        a = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in q[:L]]) for _ in range(3)])
        b = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in q[:L]]) for _ in range(3)])
        qv = np.array(q[:L], dtype=np.uint64)[None, :, None]
        s = a + b
        s = np.where(s >= qv, s - qv, s)

        def transform(x, inverse=False):
            d = DeviceArray(x.size)
            d.upload(x.reshape(-1))
            _L().call("troyb_ntt_inverse" if inverse else "troyb_ntt_forward", _vp(ctx.h), 0, _vp(d.ptr), 3, _vp(None))
            return d.download().reshape(x.shape)

        fa, fb, fs = transform(a), transform(b), transform(s)
        assert np.array_equal(transform(fa, inverse=True), a), f"round trip N={n} L={L}"
        lin = fa + fb
        lin = np.where(lin >= qv, lin - qv, lin)
        assert np.array_equal(fs, lin), f"linearity N={n} L={L}"
        for limb in range(L):  # every limb, word for word against the reference's own transform
            assert np.array_equal(fa[1, limb], ref.ntt_prime(n, q[limb], a[1, limb])), f"forward N={n} L={L} limb {limb}"
            assert np.array_equal(a[2, limb], ref.ntt_prime(n, q[limb], fa[2, limb], inverse=True)), f"inverse N={n} limb {limb}"


def test_config5_full_size_key_switch(ref, cuda_lib):
    """BASELINE config 5 at its full size: CKKS N=65536, 40 data primes + special prime. One relinearisation (40 digits,
    1640 mod-up transforms, a 1.7 GB key) word for word against the reference CPU library, on a synthetic key
    (uniform residues: key VALUES are irrelevant to the arithmetic being compared)."""
    tb = cuda_lib
    n, bits = 65536, [60] + [50] * 39 + [60]
    rng = np.random.default_rng(65536)
    rc, ctx = make_pair(ref, tb, ref.CKKS, n, bits)
    kq = rc.level_moduli(-1)
    L = len(kq) - 1
    assert L == 40 and ctx.level_limbs(0) == 40
    key = np.empty((L, 2, L + 1, n), dtype=np.uint64)
    for i, m in enumerate(kq):
        key[:, :, i, :] = rng.integers(0, m, (L, 2, n), dtype=np.uint64)
    rks = rc.synthetic_keys(False, 1)
    rks.set_words(0, key)
    d_rlk = tb.RelinKeysCuda.from_host(ctx, {0: key})
    del key
    ev = tb.EvaluatorCuda(ctx)
    w = random_ct_words(rc, 0, 3, n, rng)
    a = tb.CiphertextCuda.from_host(ctx, w, 0, True, 2.0 ** 50)
    ra = rc.ciphertext(w, 0, True, 2.0 ** 50)
    ev.relinearize_inplace(a, d_rlk)
    rc.eval(ref.OP_RELIN, ra, keys=rks)
    assert_same(a, ra, "config 5 relinearize (N=65536, 40 digits)")
    # two levels down (38 limbs): only digits < 38 and key limbs {0..37, special} are touched
    for _ in range(2):
        ev.rescale_to_next_inplace(a)
        rc.eval(ref.OP_RESCALE_NEXT, ra)
    assert_same(a, ra, "config 5 rescale x2")
    ev.multiply_inplace(a, a.clone())
    rc.eval(ref.OP_MULTIPLY, ra, ra.clone())
    ev.relinearize_inplace(a, d_rlk)
    rc.eval(ref.OP_RELIN, ra, keys=rks)
    assert_same(a, ra, "config 5 multiply + relinearize at level 2")


@pytest.mark.parametrize("scheme,n,bits", FLOWS)
def test_decrypt(ref, cuda_lib, scheme, n, bits):
    """CKKS / BFV / BGV decryption on the device against the reference CPU decryptor."""
    run_decrypt_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(5000 + n + scheme))


def test_prng_stream(ref, cuda_lib):
    """Blake2xbPRNG's byte stream expanded on the device."""
    run_prng_flow(ref, cuda_lib, np.random.default_rng(6000))


@pytest.mark.parametrize("scheme,n,bits", [(2, 16384, [60, 40, 40, 60]), (1, 8192, [60, 40, 40, 60]), (3, 32768, [60, 50, 60]),
                                           (2, 65536, [60, 50, 60]), (1, 256, [40, 30, 30, 40])])
def test_samplers(ref, cuda_lib, scheme, n, bits):
    """samplePolyTernary / samplePolyCbd / samplePolyUniform on the device, real and crafted streams (up to 90 % of the draws
    rejected), words and end positions against the reference's samplers."""
    run_sampler_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(6100 + n + scheme))


@pytest.mark.parametrize("scheme,n,bits", FLOWS)
def test_encrypt(ref, cuda_lib, scheme, n, bits):
    """Encryption on the device with the CPU generator's semantics: same seed, same ciphertext words as the reference CPU
    encryptor, asymmetric and symmetric, at every level; then decryption on the device gives the plaintext back."""
    run_encrypt_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(6200 + n + scheme))


def test_matmul(ref, cuda_lib):
    """The one-launch HE matmul (MatmulHelper::matmul) against the reference's multiplyPlain + addInplace loop."""
    run_matmul_flow(ref, cuda_lib, 16384, [60, 40, 40, 40, 60], np.random.default_rng(51), B=3, I=5, J=3)
    run_matmul_flow(ref, cuda_lib, 65536, [60, 50, 60], np.random.default_rng(52), B=2, I=2, J=4, level=1)
    run_matmul_flow(ref, cuda_lib, 4096, [60, 40, 60], np.random.default_rng(53), B=1, I=70, J=1)
    # long sums on 40-bit limbs (short accumulators renewed every eight terms), random and with every word at p - 1
    run_matmul_flow(ref, cuda_lib, 8192, [40, 40, 36, 60], np.random.default_rng(54), B=3, I=21, J=2)
    run_matmul_flow(ref, cuda_lib, 4096, [60, 40, 40, 60], np.random.default_rng(55), B=2, I=19, J=3, extreme=True)


# ---- committed golden fixtures (generated from the reference by tests/golden/make_golden.py) ----------------------
import os  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["ckks_n256", "bfv_n256", "bgv_n256", "ckks_n1024"])
def test_golden_fixture(cuda_lib, name):
    tb = cuda_lib
    g = np.load(os.path.join(GOLD, name + ".npz"))
    scheme, n, t = int(g["scheme"]), int(g["n"]), int(g["t"])
    p = tb.EncryptionParameters(scheme)
    p.set_poly_modulus_degree(n)
    p.set_coeff_modulus([int(x) for x in g["q"]])
    if t:
        p.set_plain_modulus(t)
    ctx = tb.SEALContextCuda(p)
    ev = tb.EvaluatorCuda(ctx)
    e1, ec = int(g["elt_step1"]), int(g["elt_conj"])
    rlk = tb.RelinKeysCuda.from_host(ctx, {0: g["relin_key"]})
    gk = tb.GaloisKeysCuda.from_host(ctx, {(e1 - 1) >> 1: g["galois_key_step1"], (ec - 1) >> 1: g["galois_key_conj"]})
    assert np.array_equal(ctx.root_powers(0)[:, 0], g["root_powers_q0"])
    ntt = scheme == 2
    sc = float(g["scale"])
    a = tb.CiphertextCuda.from_host(ctx, g["in_a"], 0, ntt, sc)
    b = tb.CiphertextCuda.from_host(ctx, g["in_b"], 0, ntt, sc)
    ev.multiply_inplace(a, b)
    assert np.array_equal(a.to_host()[0], g["multiply"])
    ev.relinearize_inplace(a, rlk)
    assert np.array_equal(a.to_host()[0], g["relinearize"])
    if ntt:
        ev.rescale_to_next_inplace(a)
        assert np.array_equal(a.to_host()[0], g["rescale"])
        ev.rotate_vector_inplace(a, 1, gk)
        assert np.array_equal(a.to_host()[0], g["rotate"])
        ev.complex_conjugate_inplace(a, gk)
        assert np.array_equal(a.to_host()[0], g["conjugate"])
    else:
        ev.rotate_rows_inplace(a, 1, gk)
        assert np.array_equal(a.to_host()[0], g["rotate"])
        ev.mod_switch_to_next_inplace(a)
        assert np.array_equal(a.to_host()[0], g["mod_switch"])
        ev.rotate_columns_inplace(a, gk)
        assert np.array_equal(a.to_host()[0], g["rotate_columns"])
        if scheme == 1:
            assert ctx.bsk(0) == [int(x) for x in g["bsk"]]
    assert a.scale == float(g["final_scale"]) and a.correction_factor == int(g["final_correction_factor"])


# ---- full-size, size-independent properties (BASELINE.json config 2 and 5 shapes) --------------------------------------
def _ckks_ctx(tb, n, n_data_limbs, bits=60):
    from oracle import oracle as orc
    q = orc.coeff_modulus_create(n, [bits] * (n_data_limbs + 1))
    p = tb.EncryptionParameters(tb.SchemeType.ckks)
    p.set_poly_modulus_degree(n)
    p.set_coeff_modulus(q)
    return tb.SEALContextCuda(p, expand_mod_chain=False), q


@pytest.mark.parametrize("n,limbs", [(4096, 1), (8192, 3), (16384, 32), (32768, 8), (65536, 16), (65536, 32)])
def test_ntt_sweep_round_trip_and_spot_check(cuda_lib, n, limbs):
    """NTT sweep of config 2 (60-bit primes): inverse(forward(x)) == x on every word, forward is linear, and one limb
    is checked word-for-word against the oracle."""
    from oracle import oracle as orc
    tb = cuda_lib
    ctx, q = _ckks_ctx(tb, n, limbs)
    ev = tb.EvaluatorCuda(ctx)
    rng = np.random.default_rng(n + limbs)
    polys = 3
    x = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in q[:limbs]]) for _ in range(polys)])
    ct = tb.CiphertextCuda.from_host(ctx, x, 0, False)
    ev.transform_to_ntt_inplace(ct)
    f = ct.to_host()[0]
    assert all((f[:, l, :] < q[l]).all() for l in range(limbs)), "outputs must be canonical residues"
    l = limbs - 1
    assert np.array_equal(f[1, l], orc.ntt_forward(q[l], x[1, l]))
    # linearity: NTT(x0) + NTT(x1) == NTT(x0 + x1) (mod q)
    s = np.stack([(x[0, k].astype(object) + x[1, k].astype(object)) % q[k] for k in range(limbs)]).astype(np.uint64)
    cs = tb.CiphertextCuda.from_host(ctx, np.stack([s, s]), 0, False)
    ev.transform_to_ntt_inplace(cs)
    want = np.stack([(f[0, k].astype(object) + f[1, k].astype(object)) % q[k] for k in range(limbs)]).astype(np.uint64)
    assert np.array_equal(cs.to_host()[0, 0], want)
    ev.transform_from_ntt_inplace(ct)
    assert np.array_equal(ct.to_host()[0], x)


def test_large_keyswitch_against_oracle(cuda_lib):
    """Config 5 shape (CKKS N=65536) at a limb count the C oracle finishes in seconds: relinearize word-for-word."""
    from oracle import oracle as orc
    tb = cuda_lib
    n, L = 65536, 3
    ctx, q = _ckks_ctx(tb, n, L, bits=55)
    ev = tb.EvaluatorCuda(ctx)
    rng = np.random.default_rng(9)
    key = np.stack([np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in q]) for _ in range(2)]) for _ in range(L)])
    ct3 = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in q[:L]]) for _ in range(3)])
    rlk = tb.RelinKeysCuda.from_host(ctx, {0: key})
    a = tb.CiphertextCuda.from_host(ctx, ct3, 0, True, 2.0 ** 40)
    ev.relinearize_inplace(a, rlk)
    want = orc.switch_key(orc.CKKS, q, 0, ct3[:2], ct3[2], key)
    assert np.array_equal(a.to_host()[0], want)


def test_error_behaviour(cuda_lib):
    """Argument errors surface as the reference's exception kinds before any launch."""
    tb = cuda_lib
    from troy_b200.lib import TroyInvalidArgument, TroyLogicError
    ctx, q = _ckks_ctx(tb, 4096, 2)
    ev = tb.EvaluatorCuda(ctx)
    x = np.zeros((2, 2, 4096), dtype=np.uint64)
    ct = tb.CiphertextCuda.from_host(ctx, x, 0, False, 2.0 ** 20)
    with pytest.raises(TroyInvalidArgument):
        ev.multiply_inplace(ct, ct)  # CKKS operands must be in NTT form
    with pytest.raises(TroyInvalidArgument):
        ev.transform_from_ntt_inplace(ct)
    gk = tb.GaloisKeysCuda(ctx)
    ct.is_ntt_form = True
    with pytest.raises(TroyInvalidArgument, match="Galois key not present"):
        ev.rotate_vector_inplace(ct, 1, gk)
    with pytest.raises(TroyInvalidArgument):
        ev.rescale_to_next(tb.CiphertextCuda.from_host(ctx, x, 0, False))
    p = tb.EncryptionParameters(tb.SchemeType.bfv)
    p.set_poly_modulus_degree(4096)
    p.set_coeff_modulus(q[:1])
    p.set_plain_modulus(65537)
    one = tb.SEALContextCuda(p)
    assert not one.using_keyswitching()
    with pytest.raises(TroyLogicError, match="keyswitching is not supported"):
        c3 = tb.CiphertextCuda.from_host(one, np.zeros((3, 1, 4096), dtype=np.uint64), 0, False)
        tb.EvaluatorCuda(one).relinearize_inplace(c3, tb.RelinKeysCuda(one))
    with pytest.raises(TroyInvalidArgument):
        bad = tb.EncryptionParameters(tb.SchemeType.ckks)
        bad.set_poly_modulus_degree(4096)
        bad.set_coeff_modulus([q[0], q[0]])
        tb.SEALContextCuda(bad)


def test_bgv_add_sub_with_different_correction_factors(ref, cuda_lib):
    """BGV add / sub of ciphertexts whose correction factors differ: the factors are balanced with scalar
    multiplications first (src/evaluator_cuda.cu:53-114, 171-190; kMultiplyPolyScalarCoeffmod)."""
    from common import assert_same, make_pair, random_ct_words, upload_keys
    tb = cuda_lib
    n = 4096
    rc, ctx = make_pair(ref, tb, ref.BGV, n, [40, 36, 36, 40])
    rlk, gk, d_rlk, d_gk = upload_keys(tb, ctx, rc, ref)
    ev = tb.EvaluatorCuda(ctx)
    rng = np.random.default_rng(31)
    w = random_ct_words(rc, 0, 2, n, rng)
    a, ra = tb.CiphertextCuda.from_host(ctx, w, 0, False), rc.ciphertext(w, 0, False, 1.0)
    ev.mod_switch_to_next_inplace(a); rc.eval(ref.OP_MODSWITCH_NEXT, ra)      # correction factor becomes q_last^-1 mod t
    sq, rsq = a.clone(), ra.clone()
    ev.square_inplace(sq); rc.eval(ref.OP_SQUARE, rsq)
    ev.relinearize_inplace(sq, d_rlk); rc.eval(ref.OP_RELIN, rsq, keys=rlk)  # correction factor squared
    assert sq.correction_factor != a.correction_factor
    s2, rs2 = sq.clone(), rsq.clone()
    ev.add_inplace(sq, a); rc.eval(ref.OP_ADD, rsq, ra)
    assert_same(sq, rsq, "BGV add with correction-factor balancing")
    ev.sub_inplace(s2, a); rc.eval(ref.OP_SUB, rs2, ra)
    assert_same(s2, rs2, "BGV sub with correction-factor balancing")


@pytest.mark.parametrize("scheme,n,bits", [(2, 16384, [60, 40, 40, 40, 40, 40, 40, 40, 60]), (1, 8192, [60, 40, 40, 60]),
                                           (3, 32768, [60, 50, 50, 50, 60]), (2, 65536, [60, 50, 50, 50, 60]), (2, 4096, [60, 40, 40, 60])])
def test_fused_key_switch(ref, cuda_lib, monkeypatch, scheme, n, bits):
    """TROYB_KS_FUSED=1: mod-up transform, key product and accumulation in ONE kernel (the transformed digits are multiplied by the
    key in the tile and reduce-added into the accumulators at the L2; D is never allocated, ks_inner_product is not launched).
    Same flows, same words as the reference; the launch list of a key switch has no inner-product kernel."""
    import ctypes as C
    import json
    from troy_b200.host import _L
    monkeypatch.setenv("TROYB_KS_FUSED", "1")
    run_evaluator_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(7000 + n + scheme), batch=3)
    _L().call("troyb_profile_enable", 1)
    run_evaluator_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(7100 + n + scheme))
    buf = C.create_string_buffer(1 << 16)
    _L().call("troyb_profile_report", buf, len(buf))
    _L().call("troyb_profile_enable", 0)
    names = {json.loads(x)["kernel"] for x in buf.value.decode().splitlines() if x.strip()}
    assert "ntt_fwd_modup" in names and "ks_inner_product" not in names


@pytest.mark.parametrize("scheme,n,bits", [(1, 8192, [60, 40, 40, 60]), (2, 16384, [60, 40, 40, 40, 60]), (3, 64, [40, 30, 30, 40])])
def test_negacyclic_shift_and_divide_by_degree(ref, cuda_lib, scheme, n, bits):
    """troyb_negacyclic_shift / troyb_multiply_inv_degree (the kernels behind extractLWE / assembleLWE / packLWECiphertexts)
    word for word against the reference's CPU statements."""
    from common import run_lwe_kernels_flow
    run_lwe_kernels_flow(ref, cuda_lib, scheme, n, bits)


@pytest.mark.parametrize("scheme,n,bits", [(2, 16384, [60] + [40] * 7 + [60]), (1, 8192, [40] * 8 + [60]), (3, 4096, [40] * 9), (2, 4096, [30, 40, 36, 50])])
def test_keyswitch_with_extreme_operands(ref, cuda_lib, scheme, n, bits):
    """All key and ciphertext words at p - 1 (and mixed extremes): the bounds of the inner product's accumulators."""
    from common import run_extreme_keyswitch_flow
    run_extreme_keyswitch_flow(ref, cuda_lib, scheme, n, bits, np.random.default_rng(7))
