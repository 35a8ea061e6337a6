"""Quick device-side timings of the hot path (development aid; bench.py is the contract)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402
from troy_b200.host import DeviceArray, _L, _vp  # noqa: E402

PRIMES = {
    # CoeffModulus::Create outputs (generated with the reference, see tests/golden/make_golden.py)
}


def primes_for(n, bits):
    from troy_b200 import host  # noqa
    # descending primes = 1 mod 2n, handed out from the back per bit size, like CoeffModulus::Create
    import collections
    cnt = collections.Counter(bits)
    found = {}
    for b, c in cnt.items():
        out, v = [], ((1 << b) - 1) // (2 * n) * (2 * n) + 1
        while len(out) < c:
            if pow(2, v - 1, v) == 1 and all(v % s for s in (3, 5, 7, 11, 13)) and is_prime(v):
                out.append(v)
            v -= 2 * n
        found[b] = out
    res = []
    for b in bits:
        res.append(found[b].pop())
    return res


def is_prime(v):
    if v < 2:
        return False
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if v % a == 0:
            return v == a
    d, r = v - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, v)
        if x in (1, v - 1):
            continue
        for _ in range(r - 1):
            x = x * x % v
            if x == v - 1:
                break
        else:
            return False
    return True


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    torch.cuda.init()
    tb.KernelProvider.initialize(0)
    hbm = 6532.5e9
    for n, nl, polys, bits in [(4096, 8, 2048, 60), (8192, 8, 1024, 60), (16384, 8, 512, 60), (32768, 8, 256, 60), (65536, 8, 128, 60),
                               (4096, 8, 2048, 40), (8192, 8, 1024, 40), (16384, 8, 512, 40), (32768, 8, 256, 40), (65536, 8, 128, 40)]:
        q = primes_for(n, [bits] * nl + [60])
        p = tb.EncryptionParameters(tb.SchemeType.ckks)
        p.set_poly_modulus_degree(n)
        p.set_coeff_modulus(q)
        ctx = tb.SEALContextCuda(p, expand_mod_chain=False)
        buf = DeviceArray(polys * nl * n)
        _L().call("troyb_memset_zero", _vp(buf.ptr), buf.words * 8, _vp(None))
        tf = timeit(lambda: _L().call("troyb_ntt_forward", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None)))
        ti = timeit(lambda: _L().call("troyb_ntt_inverse", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None)))
        limbs = polys * nl
        print(json.dumps(dict(kind="ntt", n=n, prime_bits=bits, limbs_per_poly=nl, polys=polys, fwd_limb_ntt_per_s=limbs / tf,
                              inv_limb_ntt_per_s=limbs / ti, fwd_gbs=limbs * 16 * n / tf / 1e9, inv_gbs=limbs * 16 * n / ti / 1e9,
                              fwd_frac=limbs * 16 * n / tf / hbm, inv_frac=limbs * 16 * n / ti / hbm)), flush=True)
        del buf, ctx
    # CKKS N=16384 L=8
    n, L, batch = 16384, 8, 64
    q = primes_for(n, [60] + [40] * 7 + [60])
    p = tb.EncryptionParameters(tb.SchemeType.ckks)
    p.set_poly_modulus_degree(n)
    p.set_coeff_modulus(q)
    ctx = tb.SEALContextCuda(p)
    ev = tb.EvaluatorCuda(ctx)
    rng = np.random.default_rng(0)
    mods = ctx.level_moduli(0)
    kmods = ctx.level_moduli(-1)
    key = np.stack([np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in kmods]) for _ in range(2)]) for _ in range(L)])
    rlk = tb.RelinKeysCuda.from_host(ctx, {0: key})
    elt = ctx.galois_elt_from_step(1)
    gk = tb.GaloisKeysCuda.from_host(ctx, {(elt - 1) >> 1: key})
    one = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in mods]) for _ in range(2)])
    w = np.broadcast_to(one, (batch,) + one.shape)
    a = tb.CiphertextCuda.from_host(ctx, w, 0, True, 2.0 ** 40, capacity=3)
    b = tb.CiphertextCuda.from_host(ctx, w, 0, True, 2.0 ** 40)

    def mul():
        a.size = 2
        a.scale = 2.0 ** 40
        ev.multiply_inplace(a, b)

    def mulrelin():
        mul()
        ev.relinearize_inplace(a, rlk)

    def rot():
        a.size = 2
        ev.rotate_vector_inplace(a, 1, gk)

    def resc():
        a.size = 2
        ev.rescale_to_next(a)

    for name, fn, bytes_per in [("multiply", mul, 56 * L * n), ("mult_relin", mulrelin, 56 * L * n + 8 * n * (2 * L * L + 7 * L)),
                                ("rotate", rot, 8 * n * (2 * L * L + 11 * L)), ("rescale", resc, 8 * n * (4 * L - 2))]:
        t = timeit(fn, 2, 5)
        print(json.dumps(dict(kind=name, n=n, L=L, batch=batch, ops_per_s=batch / t, us_per_op=t / batch * 1e6,
                              algo_gbs=bytes_per * batch / t / 1e9, frac=bytes_per * batch / t / hbm)), flush=True)


if __name__ == "__main__":
    main()
