"""Shared helpers for the parity tests: build the reference context and the troy_b200 context on identical parameters,
seeded keys and synthetic ciphertexts (uniform residues), and compare every output word."""
import numpy as np

SEED = (1, 2, 3, 4, 5, 6, 7, 8)


def make_pair(ref, tb, scheme, n, bits, plain_bits=20, moduli=None):
    q = moduli if moduli is not None else ref.coeff_modulus_create(n, bits)
    t = ref.coeff_modulus_create(n, [plain_bits])[0] if scheme != ref.CKKS else 0
    rc = ref.RefContext(scheme, n, q, t, seed=SEED)
    p = tb.EncryptionParameters(scheme)
    p.set_poly_modulus_degree(n)
    p.set_coeff_modulus(q)
    if t:
        p.set_plain_modulus(t)
    return rc, tb.SEALContextCuda(p)


def random_ct_words(rc, level, size, n, rng, batch=None):
    mods = rc.level_moduli(level)

    def one():
        return np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(size)])
    if batch is None:
        return one()
    return np.stack([one() for _ in range(batch)])


def assert_same(mine, refct, what, batch_index=0):
    m = refct.meta()
    assert mine.size == m["size"], what
    assert mine.parms_id == m["level"], what
    assert mine.is_ntt_form == m["is_ntt"], what
    assert mine.correction_factor == m["correction_factor"], what
    assert mine.scale == m["scale"], what
    got, want = mine.to_host()[batch_index], refct.words()
    assert got.shape == want.shape, what
    assert np.array_equal(got, want), f"{what}: {int((got != want).sum())} of {got.size} words differ"


def upload_keys(tb, ctx, rc, ref, steps=(1,), with_conj=True):
    rc.keygen()
    rlk = rc.relin_keys()
    elts = [rc.galois_elt_from_step(s) for s in steps]
    if with_conj:
        elts.append(rc.galois_elt_from_step(0))
    gk = rc.galois_keys(elts=elts)
    d_rlk = tb.RelinKeysCuda.from_host(ctx, {0: rlk.words(0)})
    d_gk = tb.GaloisKeysCuda.from_host(ctx, {(e - 1) >> 1: gk.words((e - 1) >> 1) for e in elts})
    return rlk, gk, d_rlk, d_gk


def run_evaluator_flow(ref, tb, scheme, n, bits, rng, batch=1):
    """multiply → relinearize → (rescale → rotate_vector → conjugate | rotate_rows → mod_switch → rotate_columns)
    → square → add → negate, comparing every word and the metadata after each step."""
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    assert ctx.num_levels == rc.num_levels()
    rlk, gk, d_rlk, d_gk = upload_keys(tb, ctx, rc, ref)
    ev = tb.EvaluatorCuda(ctx)
    ntt = scheme == ref.CKKS
    sc = 2.0 ** 40 if ntt else 1.0
    w1 = random_ct_words(rc, 0, 2, n, rng, batch)
    w2 = random_ct_words(rc, 0, 2, n, rng, batch)
    a = tb.CiphertextCuda.from_host(ctx, w1, 0, ntt, sc)
    b = tb.CiphertextCuda.from_host(ctx, w2, 0, ntt, sc)
    ras = [rc.ciphertext(w1[i], 0, ntt, sc) for i in range(batch)]
    rbs = [rc.ciphertext(w2[i], 0, ntt, sc) for i in range(batch)]

    def both(name, mine_fn, op, **kw):
        mine_fn()
        for i in range(batch):
            rc.eval(op, ras[i], kw.get("b", [None] * batch)[i], keys=kw.get("keys"), arg=kw.get("arg", 0))
            assert_same(a, ras[i], f"{name} [scheme {scheme}, N {n}, batch elt {i}]", i)

    both("multiply", lambda: ev.multiply_inplace(a, b), ref.OP_MULTIPLY, b=rbs)
    both("relinearize", lambda: ev.relinearize_inplace(a, d_rlk), ref.OP_RELIN, keys=rlk)
    if ntt:
        both("rescale", lambda: ev.rescale_to_next_inplace(a), ref.OP_RESCALE_NEXT)
        both("rotate_vector", lambda: ev.rotate_vector_inplace(a, 1, d_gk), ref.OP_ROTATE_VECTOR, keys=gk, arg=1)
        both("conjugate", lambda: ev.complex_conjugate_inplace(a, d_gk), ref.OP_CONJUGATE, keys=gk)
    else:
        both("rotate_rows", lambda: ev.rotate_rows_inplace(a, 1, d_gk), ref.OP_ROTATE_ROWS, keys=gk, arg=1)
        both("mod_switch", lambda: ev.mod_switch_to_next_inplace(a), ref.OP_MODSWITCH_NEXT)
        both("rotate_columns", lambda: ev.rotate_columns_inplace(a, d_gk), ref.OP_ROTATE_COLUMNS, keys=gk)
    both("add", lambda: ev.add_inplace(a, a.clone()), ref.OP_ADD, b=[r.clone() for r in ras])
    both("negate", lambda: ev.negate_inplace(a), ref.OP_NEGATE)
    # square on fresh inputs (keeps the CKKS scale in bounds)
    a2 = tb.CiphertextCuda.from_host(ctx, w2, 0, ntt, sc)
    ev.square_inplace(a2)
    for i in range(batch):
        r2 = rc.ciphertext(w2[i], 0, ntt, sc)
        rc.eval(ref.OP_SQUARE, r2)
        assert_same(a2, r2, f"square [scheme {scheme}, N {n}]", i)
    # NTT round trip through the public transforms
    c = tb.CiphertextCuda.from_host(ctx, w1, 0, ntt, sc)
    rc0 = rc.ciphertext(w1[0], 0, ntt, sc)
    if ntt:
        ev.transform_from_ntt_inplace(c); rc.eval(ref.OP_FROM_NTT, rc0)
        assert_same(c, rc0, "transform_from_ntt")
        ev.transform_to_ntt_inplace(c); rc.eval(ref.OP_TO_NTT, rc0)
    else:
        ev.transform_to_ntt_inplace(c); rc.eval(ref.OP_TO_NTT, rc0)
        assert_same(c, rc0, "transform_to_ntt")
        ev.transform_from_ntt_inplace(c); rc.eval(ref.OP_FROM_NTT, rc0)
    assert_same(c, rc0, "transform round trip")
