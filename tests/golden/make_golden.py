"""Generates tests/golden/*.npz from the UNMODIFIED reference library (oracle/_ref, built from /root/reference).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Each fixture holds the parameters, the seeded key-switching keys, synthetic input ciphertexts (uniform residues) and the
reference's output after every evaluator step. Small N keeps the files a few hundred KiB; the committed files let the
oracle restatement and the CUDA path be pinned where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = (1, 2, 3, 4, 5, 6, 7, 8)


def words(rc, level, size, n, rng):
    mods = rc.level_moduli(level)
    return np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(size)])


def make(scheme, n, bits, name):
    rng = np.random.default_rng(20260101 + scheme * 1000 + n)
    q = ref.coeff_modulus_create(n, bits)
    t = ref.coeff_modulus_create(n, [20])[0] if scheme != ref.CKKS else 0
    rc = ref.RefContext(scheme, n, q, t, seed=SEED)
    rc.keygen()
    rlk = rc.relin_keys()
    e1, ec = rc.galois_elt_from_step(1), rc.galois_elt_from_step(0)
    gk = rc.galois_keys(elts=[e1, ec])
    ntt = scheme == ref.CKKS
    sc = 2.0 ** 30 if ntt else 1.0
    w1, w2 = words(rc, 0, 2, n, rng), words(rc, 0, 2, n, rng)
    out = dict(scheme=scheme, n=n, q=np.array(q, dtype=np.uint64), t=t, elt_step1=e1, elt_conj=ec,
               relin_key=rlk.words(0), galois_key_step1=gk.words((e1 - 1) >> 1), galois_key_conj=gk.words((ec - 1) >> 1),
               in_a=w1, in_b=w2, scale=sc)
    a, b = rc.ciphertext(w1, 0, ntt, sc), rc.ciphertext(w2, 0, ntt, sc)
    rc.eval(ref.OP_MULTIPLY, a, b)
    out["multiply"] = a.words()
    rc.eval(ref.OP_RELIN, a, keys=rlk)
    out["relinearize"] = a.words()
    if ntt:
        rc.eval(ref.OP_RESCALE_NEXT, a)
        out["rescale"] = a.words()
        rc.eval(ref.OP_ROTATE_VECTOR, a, keys=gk, arg=1)
        out["rotate"] = a.words()
        rc.eval(ref.OP_CONJUGATE, a, keys=gk)
        out["conjugate"] = a.words()
    else:
        rc.eval(ref.OP_ROTATE_ROWS, a, keys=gk, arg=1)
        out["rotate"] = a.words()
        rc.eval(ref.OP_MODSWITCH_NEXT, a)
        out["mod_switch"] = a.words()
        rc.eval(ref.OP_ROTATE_COLUMNS, a, keys=gk)
        out["rotate_columns"] = a.words()
    c = rc.ciphertext(w1, 0, ntt, sc)
    rc.eval(ref.OP_FROM_NTT if ntt else ref.OP_TO_NTT, c)
    out["ntt_of_a"] = c.words()
    m = a.meta()
    out["final_scale"] = m["scale"]
    out["final_correction_factor"] = m["correction_factor"]
    root, _, _ = rc.ntt_tables(-1, 0)
    out["root_powers_q0"] = root[:, 0]
    if scheme == ref.BFV:
        out["bsk"] = np.array(rc.rns_info(0)["Bsk"], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


def kats():
    """Known-answer values held by the reference's own tests (values only)."""
    k = {
        "source": "reference test/utils/ntt.cpp:52-99, test/utils/galois.cpp:26-110, test/utils/rns.cpp:893-960",
        "ntt_modulus": 0xffffffffffc0001,
        "root_powers_logn1": [1, 288794978602139552],
        "root_powers_logn2": [1, 288794978602139552, 178930308976060547, 748001537669050592],
        "ntt_fwd_logn1_of_1_1": [288794978602139553, 864126526004445282],
        "ntt_fwd_logn1_of_1_0": [1, 1],
        "galois_elt_from_step_n8": {"0": 15, "1": 3, "-3": 3, "2": 9, "-2": 9, "3": 11, "-1": 11},
        "apply_galois_n8_q17_elt3": {"in": list(range(8)), "out": [0, 14, 6, 1, 13, 7, 2, 12]},
        "apply_galois_ntt_n8_elt3": {"in": list(range(8)), "out": [4, 5, 7, 6, 1, 0, 2, 3]},
        "divide_round_q_last_13_7": [{"in": [12, 11, 4, 3], "out": [4, 3]}, {"in": [1, 2, 1, 2], "out": [0, 0]}],
    }
    with open(os.path.join(HERE, "kat_reference_tests.json"), "w") as f:
        json.dump(k, f, indent=1)


if __name__ == "__main__":
    kats()
    make(ref.CKKS, 256, [40, 30, 30, 40], "ckks_n256")
    make(ref.BFV, 256, [40, 30, 30, 40], "bfv_n256")
    make(ref.BGV, 256, [40, 30, 30, 40], "bgv_n256")
    make(ref.CKKS, 1024, [50, 40, 40, 50], "ckks_n1024")
