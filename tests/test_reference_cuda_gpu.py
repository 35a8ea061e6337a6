"""`-m gpu`: parity with the reference's OWN CUDA path (troyn::), the second target north_star names. oracle/_ref/troyn_ref_bench
is the unmodified reference CUDA library compiled for sm_100a (oracle/Makefile `refcuda`) plus a small driver
(oracle/ref_cuda_bench.cu). Both sides get the same synthetic ciphertexts and the same seeded keys; every ciphertext word
after multiply, relinearize, rescale / rotate_rows and rotate_vector / mod_switch must agree."""
import os
import subprocess

import numpy as np
import pytest

from common import make_pair, random_ct_words, upload_keys

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "troyn_ref_bench")


@pytest.mark.parametrize("scheme,n,bits", [(2, 16384, [60, 40, 40, 40, 40, 40, 40, 40, 60]), (1, 8192, [60, 40, 40, 60]),
                                           (3, 8192, [60, 40, 40, 60]), (2, 4096, [60, 40, 40, 60]),
                                           (1, 32768, [60, 50, 50, 50, 60]), (3, 32768, [60, 50, 50, 50, 60]),
                                           (2, 65536, [60, 50, 50, 50, 60]), (1, 128, [40, 30, 30, 40])])
def test_same_words_as_the_reference_cuda_path(ref, cuda_lib, tmp_path, scheme, n, bits):
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/troyn_ref_bench not built (make -C oracle refcuda where /root/reference exists)")
    tb = cuda_lib
    rng = np.random.default_rng(9000 + n + scheme)
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    q = rc.level_moduli(-1)
    w1, w2 = random_ct_words(rc, 0, 2, n, rng), random_ct_words(rc, 0, 2, n, rng)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([scheme, n, len(q), rc.plain_modulus], dtype=np.uint64).tofile(f)
        np.array(q, dtype=np.uint64).tofile(f)
        w1.tofile(f)
        w2.tofile(f)
    r = subprocess.run([TOOL, "dump", fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    raw = np.fromfile(fout, dtype=np.uint64)
    steps, pos = [], 0
    while pos < raw.size:
        cnt = int(raw[pos])
        steps.append(raw[pos + 1: pos + 1 + cnt])
        pos += 1 + cnt
    assert len(steps) == 4
    # our side: the same keys (the CPU key generator with the same seed and call order) and the same inputs
    rlk, gk, d_rlk, d_gk = upload_keys(tb, ctx, rc, ref)
    ev = tb.EvaluatorCuda(ctx)
    ckks = scheme == ref.CKKS
    sc = 2.0 ** 40 if ckks else 1.0
    a = tb.CiphertextCuda.from_host(ctx, w1, 0, ckks, sc)
    b = tb.CiphertextCuda.from_host(ctx, w2, 0, ckks, sc)
    names = ["multiply", "relinearize"] + (["rescale", "rotate_vector"] if ckks else ["rotate_rows", "mod_switch"])
    ops = [lambda: ev.multiply_inplace(a, b), lambda: ev.relinearize_inplace(a, d_rlk)]
    if ckks:
        ops += [lambda: ev.rescale_to_next_inplace(a), lambda: ev.rotate_vector_inplace(a, 1, d_gk)]
    else:
        ops += [lambda: ev.rotate_rows_inplace(a, 1, d_gk), lambda: ev.mod_switch_to_next_inplace(a)]
    for name, op, want in zip(names, ops, steps):
        op()
        got = a.to_host()[0].reshape(-1)
        assert got.size == want.size, name
        assert np.array_equal(got, want), f"{name}: {int((got != want).sum())} of {got.size} words differ from the reference CUDA path"


@pytest.mark.parametrize("scheme,n,bits", [(2, 4096, [60, 40, 40, 60]), (1, 4096, [60, 40, 40, 60])])
def test_wire_format_against_the_reference_cuda_classes(ref, cuda_lib, tmp_path, scheme, n, bits):
    """troy_b200/serialize.py against CiphertextCuda::save / load and KSwitchKeysCuda::save / load of the reference itself:
    the reference's bytes equal ours for the same objects, and the reference computes with objects it loaded from our bytes."""
    import io
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/troyn_ref_bench not built")
    tb = cuda_lib
    ser = tb.serialize
    rng = np.random.default_rng(9100 + scheme)
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    q = rc.level_moduli(-1)
    w1 = random_ct_words(rc, 0, 2, n, rng)
    fin = str(tmp_path / "in.bin")
    with open(fin, "wb") as f:
        np.array([scheme, n, len(q), rc.plain_modulus], dtype=np.uint64).tofile(f)
        np.array(q, dtype=np.uint64).tofile(f)
        w1.tofile(f)
    ckks = scheme == ref.CKKS
    sc = 2.0 ** 20 if ckks else 1.0
    rc.keygen()
    rlk = rc.relin_keys()
    d_rlk = tb.RelinKeysCuda.from_host(ctx, {0: rlk.words(0)})
    a = tb.CiphertextCuda.from_host(ctx, w1, 0, ckks, sc)
    with open(tmp_path / "ours_ct.bin", "wb") as f:
        ser.save_ciphertext(a, f)
    with open(tmp_path / "ours_rlk.bin", "wb") as f:
        ser.save_kswitch_keys(d_rlk, f)
    # a seeded ciphertext of ours (fresh symmetric encryption: c0 + 64-bit seed) for the reference to load and decrypt
    enc = tb.EncryptorCuda(ctx, None, rc.secret_key_words())
    dec = tb.DecryptorCuda(ctx, rc.secret_key_words())
    if ckks:
        pw = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in rc.level_moduli(0)])
        pt = tb.PlaintextCuda.from_host(ctx, pw, 0, sc)
    else:
        pt = tb.PlaintextCuda.from_host(ctx, rng.integers(0, rc.plain_modulus, n, dtype=np.uint64))
    mine = enc.encrypt_symmetric(pt, save_seed=True)
    assert mine.seeds is not None and int(mine.seeds[0]) != 0
    with open(tmp_path / "ours_seeded_ct.bin", "wb") as f:
        ser.save_ciphertext(mine, f)
    assert os.path.getsize(tmp_path / "ours_seeded_ct.bin") < os.path.getsize(tmp_path / "ours_ct.bin") // 2 + 128
    my_plain = dec.decrypt(mine).to_host()[0]
    back = ser.load_ciphertext(ctx, io.BytesIO(open(tmp_path / "ours_seeded_ct.bin", "rb").read()))
    assert np.array_equal(back.to_host(), mine.to_host()), "seeded save / load round trip"
    r = subprocess.run([TOOL, "serialize", fin, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert open(tmp_path / "ct.bin", "rb").read() == open(tmp_path / "ours_ct.bin", "rb").read(), "ciphertext bytes differ"
    assert open(tmp_path / "rlk.bin", "rb").read() == open(tmp_path / "ours_rlk.bin", "rb").read(), "relin key bytes differ"
    loaded = ser.load_kswitch_keys(ctx, io.BytesIO(open(tmp_path / "rlk.bin", "rb").read()), tb.RelinKeysCuda)
    assert np.array_equal(loaded.key(2).download(), d_rlk.key(2).download())
    ev = tb.EvaluatorCuda(ctx)
    ev.square_inplace(a)
    ev.relinearize_inplace(a, loaded)
    raw = np.fromfile(tmp_path / "ours_result.bin", dtype=np.uint64)
    assert np.array_equal(raw[1:], a.to_host()[0].reshape(-1)), "the reference computed something else from our bytes"
    # the reference's seeded ciphertext: our loader's expansion of c1 equals the reference loader's, word for word
    theirs = ser.load_ciphertext(ctx, io.BytesIO(open(tmp_path / "seeded_ct.bin", "rb").read()))
    want = np.fromfile(tmp_path / "seeded_expanded.bin", dtype=np.uint64)
    assert np.array_equal(theirs.to_host()[0].reshape(-1), want[1:]), "c1 expanded from the reference's seed differs"
    # and the reference, loading OUR seeded ciphertext, decrypts what our decryptor decrypts
    got = np.fromfile(tmp_path / "ours_seeded_plain.bin", dtype=np.uint64)
    assert np.array_equal(got, my_plain[: got.size]) and not my_plain[got.size:].any(), "the reference decrypted something else"
