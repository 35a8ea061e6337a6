"""Standalone NTT launches for ncu captures: PROFILE_N (16384), PROFILE_BITS (40), PROFILE_POLYS (222), 8 limbs."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402
from troy_b200.lib import library  # noqa: E402

N = int(os.environ.get("PROFILE_N", "16384"))
bits = int(os.environ.get("PROFILE_BITS", "40"))
polys = int(os.environ.get("PROFILE_POLYS", "222"))
tb.KernelProvider.initialize(0)
q = tb.CoeffModulus.Create(N, [bits] * 8 + [60])
p = tb.EncryptionParameters(tb.SchemeType.ckks)
p.set_poly_modulus_degree(N)
p.set_coeff_modulus(q)
ctx = tb.SEALContextCuda(p, expand_mod_chain=False)
dev = torch.device("cuda", 0)
buf = torch.empty(polys, 8, N, dtype=torch.int64, device=dev)
for i, m in enumerate(q[:8]):
    buf[:, i, :] = torch.randint(0, m, (polys, N), device=dev, dtype=torch.int64)
for it in range(3):
    library().call("troyb_ntt_forward", C.c_void_p(ctx.h), 0, C.c_void_p(buf.data_ptr()), polys, None)
    library().call("troyb_ntt_inverse", C.c_void_p(ctx.h), 0, C.c_void_p(buf.data_ptr()), polys, None)
torch.cuda.synchronize()
print("done")
