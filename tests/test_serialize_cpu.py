"""`-m "not gpu"`: the reference's wire format (troy_b200/serialize.py): parms_id hashes against the reference library at
every level, and save -> load -> save round trips of ciphertexts, plaintexts and key-switching keys (through the emulated
build of the library, which only moves words here)."""
import io

import numpy as np
import pytest

from common import make_pair, random_ct_words


@pytest.mark.parametrize("scheme,n,bits", [(1, 128, [40, 30, 30, 40]), (2, 256, [40, 30, 30, 30, 40]), (3, 64, [40, 30, 40])])
def test_wire_format(ref, emul_lib, scheme, n, bits):
    tb = emul_lib
    ser = tb.serialize
    rng = np.random.default_rng(70 + scheme)
    rc, ctx = make_pair(ref, tb, scheme, n, bits)
    for level in [-1] + list(range(ctx.num_levels)):
        assert ser.parms_id(ctx, level) == rc.parms_id(level), f"parms_id of level {level}"
    ckks = scheme == ref.CKKS
    # ciphertext
    w = random_ct_words(rc, 1, 3, n, rng)
    ct = tb.CiphertextCuda.from_host(ctx, w, 1, ckks, 2.0 ** 20 if ckks else 1.0, 1 if scheme != ref.BGV else 7)
    buf = io.BytesIO()
    ser.save_ciphertext(ct, buf)
    raw = buf.getvalue()
    limbs = ctx.level_limbs(1)
    assert len(raw) == 32 + 1 + 8 * 6 + 1 + 8 + 8 * 3 * limbs * n
    assert raw[:32] == rc.parms_id(1)
    back = ser.load_ciphertext(ctx, io.BytesIO(raw))
    assert (back.parms_id, back.size, back.is_ntt_form, back.scale, back.correction_factor) == \
           (ct.parms_id, ct.size, ct.is_ntt_form, ct.scale, ct.correction_factor)
    assert np.array_equal(back.to_host(), ct.to_host())
    buf2 = io.BytesIO()
    ser.save_ciphertext(back, buf2)
    assert buf2.getvalue() == raw
    with pytest.raises(ValueError):
        ser.load_ciphertext(ctx, io.BytesIO(bytes(32) + raw[32:]))  # unknown parms_id
    with pytest.raises(ValueError):
        ser.load_ciphertext(ctx, io.BytesIO(raw[:200]))             # truncated
    # plaintext (coefficient form for BFV / BGV, NTT form for CKKS)
    if ckks:
        pw = np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in rc.level_moduli(0)])
        pt = tb.PlaintextCuda.from_host(ctx, pw, 0, 2.0 ** 20)
    else:
        pt = tb.PlaintextCuda.from_host(ctx, rng.integers(0, rc.plain_modulus, n - 3, dtype=np.uint64))
    buf = io.BytesIO()
    ser.save_plaintext(pt, buf)
    pback = ser.load_plaintext(ctx, io.BytesIO(buf.getvalue()))
    assert (pback.parms_id, pback.coeff_count, pback.scale) == (pt.parms_id, pt.coeff_count, pt.scale)
    assert np.array_equal(pback.to_host(), pt.to_host())
    # key-switching keys (Galois keys with gaps between the present indices)
    rc.keygen()
    elts = [rc.galois_elt_from_step(1), rc.galois_elt_from_step(0)]
    gk = rc.galois_keys(elts=elts)
    d_gk = tb.GaloisKeysCuda.from_host(ctx, {(e - 1) >> 1: gk.words((e - 1) >> 1) for e in elts})
    buf = io.BytesIO()
    ser.save_kswitch_keys(d_gk, buf)
    kback = ser.load_kswitch_keys(ctx, io.BytesIO(buf.getvalue()), tb.GaloisKeysCuda)
    assert len(kback.keys) == len(d_gk.keys)
    for e in elts:
        assert kback.has_key(e)
        assert np.array_equal(kback.key(e).download(), d_gk.key(e).download())
    buf2 = io.BytesIO()
    ser.save_kswitch_keys(kback, buf2)
    assert buf2.getvalue() == buf.getvalue()
