"""Device-side NTT throughput sweep (development aid; bench.py is the contract): forward / inverse limb-NTT/s for
N = 4096..65536 at 8 limbs of 40-bit or 60-bit primes, data > L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402
from troy_b200 import lib as _tlib  # noqa: E402
if os.environ.get("AB_LIB"):  # development: an A/B build of the library (TROYB_BUILD_OUT)
    _tlib._library = _tlib.Library(os.path.join(ROOT, os.environ["AB_LIB"]))
from troy_b200.host import DeviceArray, _L, _vp  # noqa: E402


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
    sizes = [int(x) for x in os.environ.get("NTT_SIZES", "4096,8192,16384,32768,65536").split(",")]
    for bits in (60, 40):
        for n in sizes:
            nl, polys = 8, (1 << 23) // n
            q = tb.CoeffModulus.Create(n, [bits] * nl + [60])
            p = tb.EncryptionParameters(tb.SchemeType.ckks)
            p.set_poly_modulus_degree(n)
            p.set_coeff_modulus(q)
            ctx = tb.SEALContextCuda(p, expand_mod_chain=False)
            buf = DeviceArray(polys * nl * n)
            _L().call("troyb_memset_zero", _vp(buf.ptr), buf.words * 8, _vp(None))
            tf = timeit(lambda: _L().call("troyb_ntt_forward", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None)))
            ti = timeit(lambda: _L().call("troyb_ntt_inverse", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None)))
            limbs = polys * nl
            print(json.dumps(dict(n=n, bits=bits, fwd_Mps=round(limbs / tf / 1e6, 2), inv_Mps=round(limbs / ti / 1e6, 2),
                                  fwd_gbs=round(limbs * 16 * n / tf / 1e9), inv_gbs=round(limbs * 16 * n / ti / 1e9))), flush=True)
            del buf, ctx


if __name__ == "__main__":
    main()
