"""`-m "not gpu"`: pins the plain-C oracle (oracle/troy_oracle.c) against (1) the known-answer values of the reference's
own tests, (2) committed golden fixtures generated from the reference library, and (3) the reference library itself
(oracle/_ref) on fresh seeded inputs."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not orc.available(), reason="oracle/liboracle.so not built")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_known_answers_of_reference_tests():
    with open(os.path.join(GOLD, "kat_reference_tests.json")) as f:
        k = json.load(f)
    q = k["ntt_modulus"]
    assert list(orc.ntt_tables(2, q)[0]) == k["root_powers_logn1"]
    assert list(orc.ntt_tables(4, q)[0]) == k["root_powers_logn2"]
    assert list(orc.ntt_forward(q, np.array([1, 1], dtype=np.uint64))) == k["ntt_fwd_logn1_of_1_1"]
    assert list(orc.ntt_forward(q, np.array([1, 0], dtype=np.uint64))) == k["ntt_fwd_logn1_of_1_0"]
    for step, elt in k["galois_elt_from_step_n8"].items():
        assert orc.galois_elt_from_step(8, int(step)) == elt
    g = k["apply_galois_n8_q17_elt3"]
    assert list(orc.galois_poly([17], np.array([g["in"]], dtype=np.uint64), 3, False)[0]) == g["out"]
    g = k["apply_galois_ntt_n8_elt3"]
    assert list(orc.galois_poly([17], np.array([g["in"]], dtype=np.uint64), 3, True)[0]) == g["out"]
    for case in k["divide_round_q_last_13_7"]:
        ct = np.array(case["in"], dtype=np.uint64).reshape(1, 2, 2)
        assert list(orc.drop_last(orc.BFV, [13, 7], 0, ct).reshape(-1)) == case["out"]


def test_ntt_round_trip_random():
    rng = np.random.default_rng(0)
    q = 0xffffffffffc0001
    a = rng.integers(0, q, 8, dtype=np.uint64)
    assert np.array_equal(orc.ntt_inverse(q, orc.ntt_forward(q, a)), a)


@pytest.mark.parametrize("name", ["ckks_n256", "bfv_n256", "bgv_n256", "ckks_n1024"])
def test_oracle_matches_golden_fixture(name):
    g = load(name)
    scheme, n, t = int(g["scheme"]), int(g["n"]), int(g["t"])
    q = [int(x) for x in g["q"]]
    data_q = q[:-1]
    assert orc.coeff_modulus_create(n, [int(x).bit_length() for x in q]) == q
    assert np.array_equal(orc.ntt_tables(n, q[0])[0], g["root_powers_q0"])
    a, b = g["in_a"], g["in_b"]
    assert np.array_equal(orc.ntt_poly(data_q, a, inverse=(scheme == orc.CKKS)), g["ntt_of_a"])
    m = orc.multiply(scheme, data_q, t, a, b)
    assert np.array_equal(m, g["multiply"])
    r = orc.switch_key(scheme, q, t, m[:2], m[2], g["relin_key"])
    assert np.array_equal(r, g["relinearize"])
    e1, ec = int(g["elt_step1"]), int(g["elt_conj"])
    assert orc.galois_elt_from_step(n, 1) == e1 and orc.galois_elt_from_step(n, 0) == ec
    if scheme == orc.CKKS:
        r = orc.drop_last(orc.CKKS, data_q, t, r)
        assert np.array_equal(r, g["rescale"])
        r = orc.apply_galois(scheme, q, t, r, e1, g["galois_key_step1"])
        assert np.array_equal(r, g["rotate"])
        r = orc.apply_galois(scheme, q, t, r, ec, g["galois_key_conj"])
        assert np.array_equal(r, g["conjugate"])
    else:
        r = orc.apply_galois(scheme, q, t, r, e1, g["galois_key_step1"])
        assert np.array_equal(r, g["rotate"])
        r = orc.drop_last(scheme, data_q, t, r)
        assert np.array_equal(r, g["mod_switch"])
        r = orc.apply_galois(scheme, q, t, r, ec, g["galois_key_conj"])
        assert np.array_equal(r, g["rotate_columns"])
        if scheme == orc.BFV:
            assert orc.behz_base(n, data_q, t) == [int(x) for x in g["bsk"]]


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_oracle_matches_reference_library(ref, scheme):
    """Fresh inputs, different shape than the fixtures (5 data primes, N=128), straight against oracle/_ref."""
    n, bits = 128, [40, 30, 30, 30, 30, 40]
    rng = np.random.default_rng(4242 + scheme)
    q = ref.coeff_modulus_create(n, bits)
    assert orc.coeff_modulus_create(n, bits) == q
    t = ref.coeff_modulus_create(n, [20])[0] if scheme != ref.CKKS else 0
    rc = ref.RefContext(scheme, n, q, t)
    rc.keygen()
    rlk = rc.relin_keys()
    e1 = rc.galois_elt_from_step(3)
    gk = rc.galois_keys(elts=[e1])
    ntt = scheme == ref.CKKS
    mods = rc.level_moduli(0)
    w1, w2 = (np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(2)]) for _ in range(2))
    sc = 2.0 ** 30 if ntt else 1.0
    a, b = rc.ciphertext(w1, 0, ntt, sc), rc.ciphertext(w2, 0, ntt, sc)
    rc.eval(ref.OP_MULTIPLY, a, b)
    m = orc.multiply(scheme, mods, t, w1, w2)
    assert np.array_equal(m, a.words())
    rc.eval(ref.OP_RELIN, a, keys=rlk)
    r = orc.switch_key(scheme, q, t, m[:2], m[2], rlk.words(0))
    assert np.array_equal(r, a.words())
    # one level down, then rotate by 3 (exercises key-switching below the top level)
    rc.eval(ref.OP_RESCALE_NEXT if ntt else ref.OP_MODSWITCH_NEXT, a)
    r = orc.drop_last(scheme, mods, t, r)
    assert np.array_equal(r, a.words())
    rc.eval(ref.OP_APPLY_GALOIS, a, keys=gk, arg=e1)
    r = orc.apply_galois(scheme, q, t, r, e1, gk.words((e1 - 1) >> 1))
    assert np.array_equal(r, a.words())
