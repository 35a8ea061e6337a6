"""Small launch sequence for ncu captures (one standalone NTT batch + one batched CKKS mult+relin+rotate step)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402
from troy_b200.lib import library  # noqa: E402

N, L, B = 16384, 8, int(os.environ.get("PROFILE_BATCH", "12"))
tb.KernelProvider.initialize(0)
bits = int(os.environ.get("PROFILE_BITS", "0"))
q = tb.CoeffModulus.Create(N, [bits] * 8 + [60] if bits else [60] + [40] * 7 + [60])
p = tb.EncryptionParameters(tb.SchemeType.ckks)
p.set_poly_modulus_degree(N)
p.set_coeff_modulus(q)
ctx = tb.SEALContextCuda(p)
ev = tb.EvaluatorCuda(ctx)
dev = torch.device("cuda", 0)


def uniform(prefix, moduli):
    out = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
    for i, m in enumerate(moduli):
        out[..., i, :] = torch.randint(0, m, (*prefix, N), device=dev, dtype=torch.int64)
    return out


kq = ctx.level_moduli(-1)
key = uniform((L, 2), kq)
rlk = tb.RelinKeysCuda(ctx)
rlk.keys = [tb.host.ExternalArray(key.data_ptr(), key.numel())]
elt = ctx.galois_elt_from_step(1)
gk = tb.GaloisKeysCuda(ctx)
gk.keys = [None] * (((elt - 1) >> 1) + 1)
gk.keys[(elt - 1) >> 1] = rlk.keys[0]
a_in, b_in = uniform((B, 2), q[:L]), uniform((B, 2), q[:L])
a_work = torch.empty(B, 3, L, N, dtype=torch.int64, device=dev)
a = tb.CiphertextCuda.from_device(ctx, a_work.data_ptr(), B, 2, 0, True, 2.0 ** 40, capacity=3)
b = tb.CiphertextCuda.from_device(ctx, b_in.data_ptr(), B, 2, 0, True, 2.0 ** 40)
ntt_buf = uniform((148,), q[:L])
for it in range(2):
    library().call("troyb_ntt_forward", C.c_void_p(ctx.h), 0, C.c_void_p(ntt_buf.data_ptr()), 148, None)
    library().call("troyb_ntt_inverse", C.c_void_p(ctx.h), 0, C.c_void_p(ntt_buf.data_ptr()), 148, None)
    a_work[:, :2].copy_(a_in)
    a.size, a.scale = 2, 2.0 ** 40
    ev.multiply_inplace(a, b)
    ev.relinearize_inplace(a, rlk)
    ev.rotate_vector_inplace(a, 1, gk)
    r = ev.rescale_to_next(a)
torch.cuda.synchronize()
print("done")
