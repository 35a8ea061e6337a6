"""Shared helpers for the parity tests: build the reference context and the troy_b200 context on identical parameters,
seeded keys and synthetic ciphertexts (uniform residues), and compare every output word."""
import numpy as np
import pytest

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

    # out-of-place forms write straight into a destination and leave their inputs alone
    dst = tb.CiphertextCuda(ctx, 0, 3, ntt, sc, batch=batch)
    before = a.to_host().copy()
    ev.multiply(a, b, dst)
    assert np.array_equal(a.to_host(), before), "multiply(a, b, dst) modified its input"
    both("multiply", lambda: ev.multiply_inplace(a, b), ref.OP_MULTIPLY, b=rbs)
    assert np.array_equal(dst.to_host(), a.to_host()) and dst.size == a.size and dst.scale == a.scale, "multiply(a, b, dst)"
    both("relinearize", lambda: ev.relinearize_inplace(a, d_rlk), ref.OP_RELIN, keys=rlk)
    if ntt:
        both("rescale", lambda: ev.rescale_to_next_inplace(a), ref.OP_RESCALE_NEXT)
        dst2 = tb.CiphertextCuda(ctx, a.parms_id, 2, ntt, a.scale, batch=batch)
        ev.rotate_vector(a, 1, d_gk, dst2)
        both("rotate_vector", lambda: ev.rotate_vector_inplace(a, 1, d_gk), ref.OP_ROTATE_VECTOR, keys=gk, arg=1)
        assert np.array_equal(dst2.to_host(), a.to_host()), "rotate_vector(a, steps, keys, dst)"
        both("conjugate", lambda: ev.complex_conjugate_inplace(a, d_gk), ref.OP_CONJUGATE, keys=gk)
    else:
        dst2 = tb.CiphertextCuda(ctx, a.parms_id, 2, ntt, a.scale, a.correction_factor, batch=batch)
        ev.rotate_rows(a, 1, d_gk, dst2)
        both("rotate_rows", lambda: ev.rotate_rows_inplace(a, 1, d_gk), ref.OP_ROTATE_ROWS, keys=gk, arg=1)
        assert np.array_equal(dst2.to_host(), a.to_host()), "rotate_rows(a, steps, keys, dst)"
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


def run_plain_flow(ref, tb, scheme, n, bits, rng, batch=2, level=0):
    """The ciphertext (x) plaintext callers next to the path, every word against the reference:
    CKKS: multiply_plain (NTT form), add_plain, sub_plain, fused multiply-accumulate (matmul step);
    BFV / BGV: multiply_plain (coefficient form), add_plain, sub_plain, plaintext transform_to_ntt + multiply_plain (NTT)."""
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    ev = tb.EvaluatorCuda(ctx)
    ckks = scheme == ref.CKKS
    sc = 2.0 ** 30 if ckks else 1.0
    mods = rc.level_moduli(level)
    w = random_ct_words(rc, level, 2, n, rng, batch)
    cf = 1 if scheme != ref.BGV else 3  # a non-trivial BGV correction factor
    a = tb.CiphertextCuda.from_host(ctx, w, level, ckks, sc, cf)
    ras = [rc.ciphertext(w[i], level, ckks, sc, cf) for i in range(batch)]

    def check(name):
        for i in range(batch):
            assert_same(a, ras[i], f"{name} [scheme {scheme}, N {n}, level {level}, batch elt {i}]", i)

    if ckks:
        pw = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
        pt = tb.PlaintextCuda.from_host(ctx, pw, level, sc)
        rpt = rc.plaintext(pw, level, sc)
        ev.add_plain_inplace(a, pt)
        [rc.eval_plain(ref.PLAIN_ADD, r, rpt) for r in ras]
        check("add_plain")
        ev.sub_plain_inplace(a, pt)
        ev.sub_plain_inplace(a, pt)
        [rc.eval_plain(ref.PLAIN_SUB, rc.eval_plain(ref.PLAIN_SUB, r, rpt), rpt) for r in ras]
        check("sub_plain")
        # matmul step: acc = x * w0, acc += x * w1 in one pass vs multiplyPlain + addInplace
        pw1 = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods])
        pt1, rpt1 = tb.PlaintextCuda.from_host(ctx, pw1, level, sc), rc.plaintext(pw1, level, sc)
        x = a.clone()
        rxs = [r.clone() for r in ras]
        ev.multiply_plain_inplace(a, pt)
        [rc.eval_plain(ref.PLAIN_MULTIPLY, r, rpt) for r in ras]
        check("multiply_plain")
        ev.multiply_plain_accumulate(x, pt1, a)
        for r, rx in zip(ras, rxs):
            rc.eval_plain(ref.PLAIN_MULTIPLY, rx, rpt1)
            rc.eval(ref.OP_ADD, r, rx)
        check("multiply_plain_accumulate")
        # one plaintext per batch element
        pwb = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(batch)])
        ptb = tb.PlaintextCuda.from_host(ctx, pwb, level, 1.0)
        ev.multiply_plain_inplace(a, ptb)
        for i in range(batch):
            rc.eval_plain(ref.PLAIN_MULTIPLY, ras[i], rc.plaintext(pwb[i], level, 1.0))
        check("multiply_plain (batched plaintexts)")
        return
    t = rc.plain_modulus
    for count in (n, n // 2 + 2):  # full and short plaintexts (coeff_count < N)
        pw = rng.integers(0, t, count, dtype=np.uint64)
        pt, rpt = tb.PlaintextCuda.from_host(ctx, pw), rc.plaintext(pw)
        ev.add_plain_inplace(a, pt)
        [rc.eval_plain(ref.PLAIN_ADD, r, rpt) for r in ras]
        check(f"add_plain ({count} coefficients)")
        ev.multiply_plain_inplace(a, pt)
        [rc.eval_plain(ref.PLAIN_MULTIPLY, r, rpt) for r in ras]
        check(f"multiply_plain ({count} coefficients)")
        pw2 = rng.integers(0, t, count, dtype=np.uint64)
        pt2, rpt2 = tb.PlaintextCuda.from_host(ctx, pw2), rc.plaintext(pw2)
        ev.sub_plain_inplace(a, pt2)
        [rc.eval_plain(ref.PLAIN_SUB, r, rpt2) for r in ras]
        check(f"sub_plain ({count} coefficients)")
    # NTT-form route: transformToNttInplace on both operands, multiplyPlainNtt, back
    pw = rng.integers(0, t, n, dtype=np.uint64)
    pt, rpt = tb.PlaintextCuda.from_host(ctx, pw), rc.plaintext(pw)
    ev.transform_plain_to_ntt_inplace(pt, level)
    rpt.to_ntt(level)
    assert np.array_equal(pt.to_host()[0], rpt.words()), "plaintext transform_to_ntt"
    ev.transform_to_ntt_inplace(a)
    ev.multiply_plain_inplace(a, pt)
    ev.transform_from_ntt_inplace(a)
    for r in ras:
        rc.eval(ref.OP_TO_NTT, r)
        rc.eval_plain(ref.PLAIN_MULTIPLY, r, rpt)
        rc.eval(ref.OP_FROM_NTT, r)
    check("multiply_plain (NTT form)")


def run_matmul_flow(ref, tb, n, bits, rng, B=3, I=3, J=3, level=0, extreme=False):
    """MatmulHelper::matmul (app/LinearHelperCKKS.cuh:225-249) on random NTT-form operands: the one-launch product against
    the reference's own loop of multiplyPlain + addInplace, every word."""
    rc, ctx = make_pair(ref, tb, ref.CKKS, n, bits)
    ev = tb.EvaluatorCuda(ctx)
    sc = 2.0 ** 20
    mods = rc.level_moduli(level)
    x = random_ct_words(rc, level, 2, n, rng, B * I)
    w = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(I * J)])
    if extreme:  # every operand word at p - 1: the accumulators' bounds
        top = np.array(mods, dtype=np.uint64) - np.uint64(1)
        x = np.ascontiguousarray(np.broadcast_to(top[None, None, :, None], x.shape))
        w = np.ascontiguousarray(np.broadcast_to(top[None, :, None], w.shape))
    helper = tb.MatmulHelper(B, 1, 1, n // 2)
    helper.grid = (I, J)  # operands are synthetic: only the grid shape matters here
    a = tb.Cipher2d(B, I, tb.CiphertextCuda.from_host(ctx, x, level, True, sc))
    wp = tb.Plain2d(I, J, tb.PlaintextCuda.from_host(ctx, w, level, sc))
    y = helper.matmul(ev, a, wp)
    got = y.ct.to_host()
    assert y.ct.scale == sc * sc and y.rows == B and y.cols == J
    for b in range(B):
        for j in range(J):
            acc = None
            for i in range(I):
                prod = rc.ciphertext(x[b * I + i], level, True, sc)
                rc.eval_plain(ref.PLAIN_MULTIPLY, prod, rc.plaintext(w[i * J + j], level, sc))
                acc = prod if acc is None else rc.eval(ref.OP_ADD, acc, prod)
            want = acc.words()
            assert np.array_equal(got[b * J + j], want), f"matmul [{b}][{j}]: {int((got[b * J + j] != want).sum())} words differ"


def run_decrypt_flow(ref, tb, scheme, n, bits, rng, batch=2):
    """Decryptor::decrypt on synthetic ciphertexts of size 2 and 3 at levels 0 and 1, and on a real
    encryption: every plaintext word against the reference CPU decryptor holding the same secret key."""
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    rc.keygen()
    dec = tb.DecryptorCuda(ctx, rc.secret_key_words())
    ckks = scheme == ref.CKKS
    sc = 2.0 ** 30 if ckks else 1.0
    for level in ([0, 1] if ctx.num_levels > 1 else [0]):
        for size in (2, 3):
            w = random_ct_words(rc, level, size, n, rng, batch)
            cf = 1 if scheme != ref.BGV or size == 2 else 5  # BGV: also a non-trivial correction factor
            got = dec.decrypt(tb.CiphertextCuda.from_host(ctx, w, level, ckks, sc, cf)).to_host()
            for i in range(batch):
                want = rc.decrypt_plain(rc.ciphertext(w[i], level, ckks, sc, cf)).words()
                full = np.zeros(got.shape[1], dtype=np.uint64)
                full[: want.size] = want  # the CPU decryptor trims leading zero coefficients of BFV plaintexts
                assert np.array_equal(got[i], full), f"decrypt [scheme {scheme}, N {n}, level {level}, size {size}, elt {i}]"
    if not ckks:
        vals = rng.integers(0, rc.plain_modulus, n, dtype=np.uint64)
        rct = rc.encrypt_batch(vals)
        got = dec.decrypt(tb.CiphertextCuda.from_host(ctx, rct.words(), 0, False, 1.0)).to_host()[0]
        want = rc.decrypt_plain(rct).words()
        assert np.array_equal(got[: want.size], want) and not got[want.size:].any(), "decrypt of a real encryption"


def _expand_stream(tb, seeds, nbytes):
    """UniformRandomGenerator::generate on the device: [batch] streams of nbytes (a multiple of 4096)."""
    from troy_b200.host import DeviceArray, _L, _vp
    sd = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1, 8)
    d_seeds, out = DeviceArray(sd.size), DeviceArray(sd.shape[0] * nbytes // 8)
    d_seeds.upload(sd.reshape(-1))
    _L().call("troyb_prng_fill", _vp(d_seeds.ptr), 8, _vp(out.ptr), nbytes // 8, nbytes, sd.shape[0], _vp(None))
    return d_seeds, out


def run_prng_flow(ref, tb, rng):
    """Blake2xbPRNG (src/randomgen.cpp:164-194): the expanded stream of several generators, byte for byte."""
    seeds = rng.integers(0, 2 ** 63, (3, 8), dtype=np.uint64)
    for nbytes in (4096, 5 * 4096, 66 * 4096):
        _, out = _expand_stream(tb, seeds, nbytes)
        got = out.download().view(np.uint8).reshape(3, nbytes)
        for b in range(3):
            assert np.array_equal(got[b], ref.prng_bytes(seeds[b], nbytes)), f"stream of generator {b}, {nbytes} bytes"


def run_sampler_flow(ref, tb, scheme, n, bits, rng):
    """samplePolyTernary / samplePolyCbd / samplePolyUniform (src/utils/rlwe.cpp:21-158) at the key level and two data
    levels: from the real generator (also with the expanded part cut short, so that the rest is computed on demand), and
    from crafted streams in which 1 %, 30 % and 90 % of the draws are rejected (zeros for the ternary sampler, words above
    the largest multiple of the prime for the uniform one); words and the generator's end position against the reference."""
    import ctypes as C
    from troy_b200.host import DeviceArray, _L, _vp
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    seed = rng.integers(0, 2 ** 63, 8, dtype=np.uint64)
    for level in [-1, 0, 1][: 1 + ctx.num_levels]:
        L = ctx.level_limbs(level)
        o, end = DeviceArray(L * n), DeviceArray(1)

        def call(which, st, nbytes, expanded, d_seed, start):
            h, lv, sp, ss = _vp(ctx.h), level, _vp(st.ptr), nbytes // 8
            if which == 0:
                _L().call("troyb_sample_poly_ternary", h, lv, sp, ss, expanded, d_seed, 8, C.c_uint64(start), _vp(o.ptr), L * n,
                          _vp(end.ptr), 1, _vp(None))
            elif which == 1:
                _L().call("troyb_sample_poly_cbd", h, lv, sp, ss, expanded, d_seed, 8, C.c_uint64(start), _vp(o.ptr), L * n, 1,
                          _vp(None))
            else:
                _L().call("troyb_sample_poly_uniform", h, lv, sp, ss, expanded, d_seed, 8, _vp(o.ptr), L * n, 1, _vp(None))
            return o.download().reshape(L, n)

        nbytes = ((L * n * 8 + 8192 + 4095) // 4096) * 4096
        d_seed, st = _expand_stream(tb, seed, nbytes)
        for which, start in ((0, 0), (0, 4 * n + 8), (1, 64), (1, 4 * n), (1, 4 * n + 6), (2, 0)):
            want, pos = rc.sample_poly(level, which, seed, start)
            assert np.array_equal(call(which, st, nbytes, nbytes, _vp(d_seed.ptr), start), want), f"sampler {which} level {level}"
            if which == 0:
                assert int(end.download()[0]) == pos
                # only the first 64 bytes expanded: every later draw is computed from the seed on demand
                assert np.array_equal(call(which, st, nbytes, 64, _vp(d_seed.ptr), start), want), f"on-demand sampler level {level}"
        nbytes = ((L * n * 16 + 4095) // 4096) * 4096
        st = DeviceArray(nbytes // 8)
        for frac in (0.01, 0.3, 0.9):
            words = rng.integers(0, 2 ** 63, nbytes // 8, dtype=np.uint64) * 2 + rng.integers(0, 2, nbytes // 8, dtype=np.uint64)
            w32 = words.view(np.uint32).copy()
            r32 = rng.random(w32.size) < frac
            r32[w32.size // 2:] = False  # the stream must not run dry
            w32[r32] = 0
            st.upload(w32.view(np.uint64))
            want, pos = rc.sample_poly_replay(level, 0, w32.view(np.uint8), 8)
            assert np.array_equal(call(0, st, nbytes, nbytes, _vp(None), 8), want), f"ternary, {frac} rejected, level {level}"
            assert int(end.download()[0]) == pos
            wu = words.copy()
            rej = rng.random(wu.size) < frac
            rej[L * n + L * n // 2:] = False
            wu[rej] = np.uint64(2 ** 64 - 1) - rng.integers(0, 3, int(rej.sum()), dtype=np.uint64)
            st.upload(wu)
            want, pos = rc.sample_poly_replay(level, 2, wu.view(np.uint8))
            assert pos > L * n * 8
            assert np.array_equal(call(2, st, nbytes, nbytes, _vp(None), 0), want), f"uniform, {frac} rejected, level {level}"


def run_encrypt_flow(ref, tb, scheme, n, bits, rng, batch=2):
    """Encryptor::encryptZero / encryptZeroSymmetric at every level, then encrypt / encryptSymmetric of a
    plaintext followed by decryption on the device: for the seed the reference context was built with, every ciphertext
    word equals the reference CPU encryptor's (src/utils/rlwe.cpp:160-356, src/encryptor.cpp:88-253)."""
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    rc.keygen()
    enc = tb.EncryptorCuda(ctx, rc.public_key_words(), rc.secret_key_words())
    seeds = np.array([SEED] * batch, dtype=np.uint64)
    levels = ([-1] if ctx.using_keyswitching() else []) + list(range(ctx.num_levels))
    for level in levels:
        for asym in (True, False):
            want = rc.encrypt_zero(level, asym)
            got = (enc.encrypt_zero if asym else enc.encrypt_zero_symmetric)(level, batch=batch, seeds=seeds)
            for i in range(batch):
                assert_same(got, want, f"encrypt_zero [scheme {scheme}, N {n}, level {level}, asymmetric {asym}]", i)
    ckks = scheme == ref.CKKS
    dec = tb.DecryptorCuda(ctx, rc.secret_key_words())
    for asym in (True, False):
        if ckks:
            level = min(1, ctx.num_levels - 1)
            mods = rc.level_moduli(level)
            pw = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(batch)])
            pt = tb.PlaintextCuda.from_host(ctx, pw, level, 2.0 ** 30)
            rpts = [rc.plaintext(pw[i], level, 2.0 ** 30) for i in range(batch)]
        else:
            pw = rng.integers(0, rc.plain_modulus, (batch, n), dtype=np.uint64)
            pt = tb.PlaintextCuda.from_host(ctx, pw)
            rpts = [rc.plaintext(pw[i]) for i in range(batch)]
        got = (enc.encrypt if asym else enc.encrypt_symmetric)(pt, seeds=seeds)
        for i in range(batch):
            assert_same(got, rc.encrypt_plain(rpts[i], asym), f"encrypt [scheme {scheme}, N {n}, asymmetric {asym}]", i)
        if not ckks:  # and back: the device decryptor returns the plaintext
            assert np.array_equal(dec.decrypt(got).to_host(), pw), "decrypt(encrypt(plain))"
    # different seeds give different ciphertexts; the default draws fresh seeds
    two = enc.encrypt_zero(0, batch=2, seeds=rng.integers(0, 2 ** 63, (2, 8), dtype=np.uint64)).to_host()
    assert not np.array_equal(two[0], two[1])
    assert not np.array_equal(enc.encrypt_zero(0).to_host(), enc.encrypt_zero(0).to_host())


def run_lwe_kernels_flow(ref, tb, scheme, n, bits):
    """The two kernels behind extractLWE / assembleLWE / packLWECiphertexts (kNegacyclicShiftPolyCoeffmod,
    kMultiplyInvPolyDegreeCoeffmod, src/kernelutils.cu:537-592) against the reference's CPU statements, every word, at
    shifts on both sides of N (including 0, N and 2N - 1) and on a lower level; X^a * X^b = X^(a+b) as the size-free property."""
    from troy_b200.lib import TroyInvalidArgument
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    ev = tb.EvaluatorCuda(ctx)
    rng = np.random.default_rng(900 + n)
    for level in (0, 1):
        mods = rc.level_moduli(level)
        w = random_ct_words(rc, level, 2, n, rng, batch=2)
        w[0, 0, 0, :7] = 0  # zero coefficients stay zero on the negated side (the reference tests poly[i] != 0)
        ct = tb.CiphertextCuda.from_host(ctx, w, level, False)
        for shift in (0, 1, n // 2 + 3, n - 1, n, n + 5, 2 * n - 1):
            got = ev.negacyclic_shift(ct, shift).to_host()
            for b in range(2):
                for k in range(2):
                    for l, q in enumerate(mods):
                        want = ref.negacyclic_shift(n, int(q), shift, w[b, k, l])
                        assert np.array_equal(got[b, k, l], want), (level, shift, b, k, l)
        a, bb = 3, n - 2
        twice = ev.negacyclic_shift(ev.negacyclic_shift(ct, a), bb).to_host()
        assert np.array_equal(twice, ev.negacyclic_shift(ct, a + bb).to_host())
        with pytest.raises(TroyInvalidArgument):
            ev.negacyclic_shift(ct, 2 * n)
        d = ct.clone()
        ev.divide_by_poly_modulus_degree_inplace(d)
        got = d.to_host()
        for b in range(2):
            for k in range(2):
                for l, q in enumerate(mods):
                    assert np.array_equal(got[b, k, l], ref.multiply_inv_degree(n, int(q), w[b, k, l])), (level, b, k, l)


def run_extreme_keyswitch_flow(ref, tb, scheme, n, bits, rng):
    """Relinearisation with every key word and every ciphertext word at its extreme (p - 1 everywhere; then p - 1 against random
    and alternating 0 / p - 1): the inner product's accumulators (ColumnAcc and the 40-bit SmallAcc with its FP64 quotient
    estimate) are sized for exactly these sums. Compared word for word with the reference, like every other flow."""
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    rc.keygen()
    rlk = rc.relin_keys()
    kq = np.array(rc.level_moduli(-1), dtype=np.uint64)
    mods = np.array(rc.level_moduli(0), dtype=np.uint64)
    shape = rlk.words(0).shape  # [digit][2][key limbs][N]
    ev = tb.EvaluatorCuda(ctx)
    ntt = scheme == ref.CKKS
    sc = 2.0 ** 40 if ntt else 1.0
    top = np.broadcast_to((kq - np.uint64(1))[None, None, :, None], shape)
    ct_top = np.broadcast_to((mods - np.uint64(1))[None, :, None], (3, len(mods), n))
    alt = (np.arange(n) & 1).astype(np.uint64)
    cases = [("all p - 1", top, ct_top),
             ("keys p - 1, random ciphertext", top, random_ct_words(rc, 0, 3, n, rng)),
             ("random keys, ciphertext p - 1", None, ct_top),
             ("alternating", top * alt[None, None, None, :], ct_top * (np.uint64(1) - alt)[None, None, :] + ct_top * alt[None, None, :])]
    for what, kw, cw in cases:
        if kw is None:
            kw = np.stack([np.stack([np.stack([rng.integers(0, int(m), n, dtype=np.uint64) for m in kq]) for _ in range(2)])
                           for _ in range(shape[0])])
        kw = np.ascontiguousarray(kw, dtype=np.uint64)
        cw = np.ascontiguousarray(cw, dtype=np.uint64)
        rlk.set_words(0, kw)
        d_rlk = tb.RelinKeysCuda.from_host(ctx, {0: kw})
        a, ra = tb.CiphertextCuda.from_host(ctx, cw, 0, ntt, sc), rc.ciphertext(cw, 0, ntt, sc)
        ev.relinearize_inplace(a, d_rlk)
        rc.eval(ref.OP_RELIN, ra, keys=rlk)
        assert_same(a, ra, f"relinearize, {what} [scheme {scheme}, N {n}]")
