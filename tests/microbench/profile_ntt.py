"""Launch sequence for ncu captures of the standalone transforms: PROFILE_N coefficients, 8 limbs of PROFILE_BITS-bit primes,
PROFILE_POLYS polys (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402
from troy_b200 import lib as _tlib  # noqa: E402
if os.environ.get("AB_LIB"):  # development: an A/B build of the library (TROYB_BUILD_OUT)
    _tlib._library = _tlib.Library(os.path.join(ROOT, os.environ["AB_LIB"]))
from troy_b200.host import DeviceArray, _L, _vp  # noqa: E402

n, bits, polys = int(os.environ.get("PROFILE_N", "16384")), int(os.environ.get("PROFILE_BITS", "40")), int(os.environ.get("PROFILE_POLYS", "512"))
tb.KernelProvider.initialize(0)
q = tb.CoeffModulus.Create(n, [bits] * 8 + [60])
p = tb.EncryptionParameters(tb.SchemeType.ckks)
p.set_poly_modulus_degree(n)
p.set_coeff_modulus(q)
ctx = tb.SEALContextCuda(p, expand_mod_chain=False)
buf = DeviceArray(polys * 8 * n)
_L().call("troyb_memset_zero", _vp(buf.ptr), buf.words * 8, _vp(None))
for _ in range(3):
    _L().call("troyb_ntt_forward", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None))
    _L().call("troyb_ntt_inverse", _vp(ctx.h), 0, _vp(buf.ptr), polys, _vp(None))
_L().call("troyb_stream_synchronize", _vp(None))
