"""`-m gpu`: the pybind module `pytroy` (binder/pytroy.cpp; the reference's binder/binder.cu for the classes of the path) driven
the way the reference's binder/test.py drives its module: a key owner and an evaluator that only exchange `bytes` in the
reference's wire format. Checks decrypted values (CKKS) and exact slots (BFV / BGV), and that a script-level user sees device
encryption, seeded ciphertexts and the evaluator through the same names as in the reference's module."""
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pytroy():
    if not glob.glob(os.path.join(ROOT, "binder", "pytroy*.so")):
        pytest.skip("binder/pytroy*.so not built (run __graft_entry__.build() where /root/reference exists)")
    sys.path.insert(0, os.path.join(ROOT, "binder"))
    import pytroy as m
    m.initialize_kernel()
    return m


def test_ckks_two_party_flow(pytroy):
    n, bits = 1 << 14, [40, 40, 40, 40, 40, 40]
    scale = float(1 << 40)

    def context():
        parms = pytroy.EncryptionParameters(pytroy.SchemeType.ckks)
        parms.set_poly_modulus_degree(n)
        parms.set_coeff_modulus(pytroy.CoeffModulus.create(n, bits))
        return pytroy.SEALContext(parms)

    # key owner
    ctx_a = context()
    encoder = pytroy.CKKSEncoder(ctx_a)
    keygen = pytroy.KeyGenerator(ctx_a)
    pk = keygen.create_public_key()
    encryptor = pytroy.Encryptor(ctx_a, pk, keygen.secret_key())
    decryptor = pytroy.Decryptor(ctx_a, keygen.secret_key())
    rlk_bytes, gk_bytes = keygen.create_relin_keys().save(), keygen.create_galois_keys([1]).save()
    m1, m2 = [1, 2, 3, 4], [0.5, 0.6, 0.7, 0.8]
    p1, p2 = pytroy.Plaintext(), pytroy.Plaintext()
    encoder.encode(m1, scale, p1)
    encoder.encode(m2, scale, p2)
    c1, c2 = pytroy.Ciphertext(), pytroy.Ciphertext()
    encryptor.encrypt(p1, c1)                # asymmetric, on the device
    encryptor.encrypt_symmetric(p2, c2)      # symmetric: seeded, travels as c0 + seed
    assert c2.seed() != 0
    b1, b2 = c1.save(), c2.save()
    assert len(b2) < len(b1) // 2 + 128

    # evaluator: a second context built from the same parameters, keys and ciphertexts arrive as bytes
    ctx_b = context()
    ev = pytroy.Evaluator(ctx_b)
    rlk, gk = pytroy.RelinKeys(), pytroy.GaloisKeys()
    rlk.load(rlk_bytes)
    gk.load(gk_bytes)
    x, y = pytroy.Ciphertext(), pytroy.Ciphertext()
    x.load(b1)
    with pytest.raises(ValueError):
        y.load(b2)                           # a seeded ciphertext needs the context to re-expand c1
    y.load(b2, ctx_b)
    ev.multiply_inplace(x, y)
    ev.relinearize_inplace(x, rlk)
    ev.rescale_to_next_inplace(x)
    rot = ev.rotate_vector(x, 1, gk)
    out_bytes, rot_bytes = x.save(), rot.save()

    # back at the key owner
    def dec(b):
        c = pytroy.Ciphertext()
        c.load(b)
        return np.array(encoder.decode(decryptor.decrypt(c)))
    got = dec(out_bytes)
    want = np.array(m1) * np.array(m2)
    assert np.allclose(got[:4].real, want, atol=1e-3) and np.allclose(got[4:64], 0, atol=1e-3)
    assert np.allclose(dec(rot_bytes)[:3].real, want[1:], atol=1e-3)


@pytest.mark.parametrize("scheme", ["bfv", "bgv"])
def test_integer_schemes(pytroy, scheme):
    n = 8192
    parms = pytroy.EncryptionParameters(getattr(pytroy.SchemeType, scheme))
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus(pytroy.CoeffModulus.create(n, [60, 40, 40, 60]))
    parms.set_plain_modulus(pytroy.PlainModulus.batching(n, 20))
    t = parms.plain_modulus().value()
    ctx = pytroy.SEALContext(parms)
    encoder = pytroy.BatchEncoder(ctx)
    keygen = pytroy.KeyGenerator(ctx)
    encryptor = pytroy.Encryptor(ctx, keygen.create_public_key(), keygen.secret_key())
    decryptor = pytroy.Decryptor(ctx, keygen.secret_key())
    ev = pytroy.Evaluator(ctx)
    rlk, gk = keygen.create_relin_keys(), keygen.create_galois_keys([1])
    rng = np.random.default_rng(5)
    a, b = rng.integers(0, 1000, n), rng.integers(0, 1000, n)
    ca = encryptor.encrypt(encoder.encode([int(v) for v in a]))
    cb = encryptor.encrypt_symmetric(encoder.encode([int(v) for v in b]))
    cb2 = pytroy.Ciphertext()
    cb2.load(cb.save(), ctx)                 # through the seeded wire form
    prod = ev.multiply(ca, cb2)
    ev.relinearize_inplace(prod, rlk)
    total = ev.add(prod, ca)
    ev.rotate_rows_inplace(total, 1, gk)
    ev.mod_switch_to_next_inplace(total)
    got = np.array(encoder.decode(decryptor.decrypt(total)), dtype=np.int64)
    want = (a * b + a) % t
    half = n // 2
    want = np.concatenate([np.roll(want[:half], -1), np.roll(want[half:], -1)])
    assert np.array_equal(got, want)
    # product tree and powers (multiply_many / exponentiate), and a key switch to another party's secret key
    small = rng.integers(0, 7, n)
    cs = encryptor.encrypt(encoder.encode([int(v) for v in small]))
    cube = ev.exponentiate(cs, 3, rlk)
    assert np.array_equal(np.array(encoder.decode(decryptor.decrypt(cube)), dtype=np.int64), small ** 3 % t)
    tree = ev.multiply_many([cs, ca, cs], rlk)
    assert np.array_equal(np.array(encoder.decode(decryptor.decrypt(tree)), dtype=np.int64), small * a * small % t)
    other = pytroy.KeyGenerator(ctx)
    ksk = keygen.create_keyswitching_keys(other.secret_key())
    theirs = pytroy.Encryptor(ctx, other.secret_key()).encrypt_symmetric(encoder.encode([int(v) for v in a]))
    mine = ev.apply_keyswitching(theirs, ksk)
    assert np.array_equal(np.array(encoder.decode(decryptor.decrypt(mine)), dtype=np.int64), a % t)


@pytest.mark.parametrize("scheme,n", [("bfv", 4096), ("bgv", 2048)])
def test_lwe_extract_assemble_pack(pytroy, scheme, n):
    """extract_lwe / assemble_lwe / field_trace_inplace / pack_lwe_ciphertexts (binder/binder.cu:712-722) at a working size: the
    flow of the reference's BFVAssembleLWETest (test/evaluator_cuda.cu:6340-6440), which runs at N = 32."""
    parms = pytroy.EncryptionParameters(getattr(pytroy.SchemeType, scheme))
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus(pytroy.CoeffModulus.create(n, [50, 40, 50]))
    parms.set_plain_modulus(pytroy.PlainModulus.batching(n, 20))
    t = parms.plain_modulus().value()
    ctx = pytroy.SEALContext(parms, True, pytroy.SecurityLevel.none)
    encoder = pytroy.BatchEncoder(ctx)
    keygen = pytroy.KeyGenerator(ctx)
    encryptor = pytroy.Encryptor(ctx, keygen.create_public_key(), keygen.secret_key())
    decryptor = pytroy.Decryptor(ctx, keygen.secret_key())
    ev = pytroy.Evaluator(ctx)
    auto_keys = keygen.create_automorphism_keys()
    rng = np.random.default_rng(n)
    coeffs = [int(v) for v in rng.integers(0, t, n)]
    ct = encryptor.encrypt(encoder.encode_polynomial(coeffs))

    def dec(c):
        return encoder.decode_polynomial(decryptor.decrypt(c))
    # X^a * X^b = X^(a+b), and the shift is the negacyclic one
    s = ev.negacyclic_shift(ct, 5)
    got = dec(s)
    assert got[5:] == coeffs[:n - 5] and got[:5] == [(t - v) % t for v in coeffs[n - 5:]]
    ev.negacyclic_shift_inplace(s, 2 * n - 5)
    assert dec(s) == coeffs
    # one coefficient out as an LWE sample and back in at another position
    term = n // 2 + 7
    lwe = ev.extract_lwe(ct, term)
    assert lwe.poly_modulus_degree() == n and lwe.coeff_modulus_size() == 2
    assert dec(ev.assemble_lwe(lwe, 0))[0] == coeffs[term]
    assert dec(ev.assemble_lwe(lwe.copy(), 9))[9] == coeffs[term]
    # field trace down to the constant coefficient: N * c_0
    tr = ct.copy()
    ev.field_trace_inplace(tr, auto_keys, 0)
    got = dec(tr)
    assert got[0] == coeffs[0] * n % t and not any(got[1:])
    # 18 LWE samples packed into one ciphertext: coefficient i * N / 32 carries sample i
    count = 18
    cts = [encryptor.encrypt(encoder.encode_polynomial([int(v) for v in rng.integers(0, t, n)])) for _ in range(count)]
    want = [dec(c)[0] for c in cts]
    packed = ev.pack_lwe_ciphertexts([ev.extract_lwe(c, 0) for c in cts], auto_keys)
    got = dec(packed)
    step = n // 32
    for i in range(count):
        assert got[i * step] == want[i], i
    assert not any(got[i] for i in range(n) if i % step)
