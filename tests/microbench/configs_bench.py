"""Device-side timings of the BASELINE.json configurations that bench.py does not time (bench.py is the contract and
times configs[2]; the others are parity-test cases, see tests/test_parity_gpu.py). Development aid; prints JSON lines.

  config 1  BFV  N=8192  {60,40,40,60}      multiply + relinearize
  config 4  BFV / BGV N=32768 {60,50,50,50,60}  rotate_rows (4-digit key switch)
  config 5  CKKS N=65536 40 x 50-bit + special  relinearize (40-digit key switch, 1.7 GB key), HE matmul
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import troy_b200 as tb  # noqa: E402
from quick_bench import timeit  # noqa: E402

HBM = 6539.9e9
dev = torch.device("cuda", 0)


def uniform(prefix, moduli, n):
    out = torch.empty(*prefix, len(moduli), n, dtype=torch.int64, device=dev)
    for i, m in enumerate(moduli):
        out[..., i, :] = torch.randint(0, m, (*prefix, n), device=dev, dtype=torch.int64)
    return out


def setup(scheme, n, bits, t_bits=20):
    q = tb.CoeffModulus.Create(n, bits)
    p = tb.EncryptionParameters(scheme)
    p.set_poly_modulus_degree(n)
    p.set_coeff_modulus(q)
    if scheme != tb.SchemeType.ckks:
        p.set_plain_modulus(tb.PlainModulus.Batching(n, t_bits))
    ctx = tb.SEALContextCuda(p)
    ev = tb.EvaluatorCuda(ctx)
    L = ctx.level_limbs(0)
    kq = ctx.level_moduli(-1)
    key = uniform((L, 2), kq, n)
    rlk = tb.RelinKeysCuda(ctx)
    rlk.keys = [tb.host.ExternalArray(key.data_ptr(), key.numel())]
    elt = ctx.galois_elt_from_step(1)
    gk = tb.GaloisKeysCuda(ctx)
    gk.keys = [None] * (((elt - 1) >> 1) + 1)
    gk.keys[(elt - 1) >> 1] = rlk.keys[0]
    return ctx, ev, q, L, key, rlk, gk


def report(**kw):
    print(json.dumps(kw), flush=True)


def main():
    torch.cuda.init()
    tb.KernelProvider.initialize(0)
    # ---- config 1
    n, B = 8192, 256
    ctx, ev, q, L, key, rlk, gk = setup(tb.SchemeType.bfv, n, [60, 40, 40, 60])
    a_in, b_in = uniform((B, 2), q[:L], n), uniform((B, 2), q[:L], n)
    work = torch.empty(B, 3, L, n, dtype=torch.int64, device=dev)
    a = tb.CiphertextCuda.from_device(ctx, work.data_ptr(), B, 2, 0, False, 1.0, capacity=3)
    b = tb.CiphertextCuda.from_device(ctx, b_in.data_ptr(), B, 2, 0, False, 1.0)

    def mulrelin():
        work[:, :2].copy_(a_in)
        a.size = 2
        ev.multiply_inplace(a, b)
        ev.relinearize_inplace(a, rlk)

    t = timeit(mulrelin, 3, 10)
    report(config=1, what="BFV N=8192 {60,40,40,60} multiply+relinearize", batch=B, ops_per_s=B / t, us_per_op=t / B * 1e6,
           io_gbs=B * 8 * n * 7 * L / t / 1e9)
    del a, b, work, a_in, b_in, key
    # ---- config 4
    n, B = 32768, 64
    for scheme, name in ((tb.SchemeType.bfv, "BFV"), (tb.SchemeType.bgv, "BGV")):
        ctx, ev, q, L, key, rlk, gk = setup(scheme, n, [60, 50, 50, 50, 60])
        a_in = uniform((B, 2), q[:L], n)
        work = torch.empty(B, 2, L, n, dtype=torch.int64, device=dev)
        a = tb.CiphertextCuda.from_device(ctx, work.data_ptr(), B, 2, 0, False, 1.0)

        def rot():
            work.copy_(a_in)
            ev.rotate_rows_inplace(a, 1, gk)

        t = timeit(rot, 3, 10)
        bytes_per = 8 * n * (2 * L * L + 11 * L)
        report(config=4, what=f"{name} N=32768 {{60,50,50,50,60}} rotate_rows", batch=B, ops_per_s=B / t, us_per_op=t / B * 1e6,
               algo_gbs=bytes_per * B / t / 1e9, frac_of_hbm_roofline=bytes_per * B / t / HBM)
        del a, work, a_in, key
    # ---- config 5
    n, B = 65536, 4
    ctx, ev, q, L, key, rlk, gk = setup(tb.SchemeType.ckks, n, [60] + [50] * 39 + [60])
    a_in = uniform((B, 3), q[:L], n)
    work = torch.empty(B, 3, L, n, dtype=torch.int64, device=dev)
    a = tb.CiphertextCuda.from_device(ctx, work.data_ptr(), B, 3, 0, True, 2.0 ** 50)

    def relin():
        work.copy_(a_in)
        a.size = 3
        ev.relinearize_inplace(a, rlk)

    t = timeit(relin, 2, 5)
    bytes_per = 8 * n * (2 * L * L + 7 * L)
    report(config=5, what="CKKS N=65536 L=40 relinearize (1.7 GB key)", batch=B, ops_per_s=B / t, ms_per_op=t / B * 1e3,
           limb_ntt_per_s=B * (L + 1) * (L + 2) / t, algo_gbs=bytes_per * B / t / 1e9, frac_of_hbm_roofline=bytes_per * B / t / HBM)
    del a, work, a_in
    MB, MI, MJ = 4, 4, 4
    x_buf = uniform((MB * MI, 2), q[:L], n)
    w_buf = uniform((MI * MJ,), q[:L], n)
    x = tb.Cipher2d(MB, MI, tb.CiphertextCuda.from_device(ctx, x_buf.data_ptr(), MB * MI, 2, 0, True, 2.0 ** 20))
    w = tb.Plain2d(MI, MJ, tb.PlaintextCuda.from_device(ctx, w_buf.data_ptr(), L * n, 0, 2.0 ** 20, batch=MI * MJ))
    mm = tb.MatmulHelper(MB, 1, 1, n // 2)
    mm.grid = (MI, MJ)
    keep = []
    t = timeit(lambda: keep.__setitem__(slice(None), [mm.matmul(ev, x, w)]), 2, 5)
    mm_bytes = (MB * MI * 2 + MI * MJ + MB * MJ * 2) * L * n * 8
    report(config=5, what="CKKS N=65536 L=40 HE matmul 4x4 ciphertexts times 4x4 plaintexts (one launch)",
           multiply_plain_accumulate_per_s=MB * MI * MJ / t, ms=t * 1e3, gbs=mm_bytes / t / 1e9, frac_of_hbm_roofline=mm_bytes / t / HBM)


if __name__ == "__main__":
    main()
