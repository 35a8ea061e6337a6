"""Device-side timing of encryption on the device: encrypt-zero, asymmetric and symmetric, at the bench
configuration (CKKS N=16384 L=8) and at BASELINE config 1 (BFV N=8192), batched and batch 1, with the per-kernel breakdown
and the reference CPU encryptor timed beside it. Development aid; prints JSON lines."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import troy_b200 as tb  # noqa: E402
from oracle import ref  # noqa: E402
from quick_bench import timeit  # noqa: E402
from troy_b200.host import _L, _vp  # noqa: E402


def profile(fn):
    _L().call("troyb_profile_enable", 1)
    fn()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    _L().call("troyb_profile_report", buf, len(buf))
    _L().call("troyb_profile_enable", 0)
    return [json.loads(l) for l in buf.value.decode().splitlines() if l.strip()]


def main():
    torch.cuda.init()
    tb.KernelProvider.initialize(0)
    for name, scheme, n, bits in [("ckks_n16384_L8", 2, 16384, [60] + [40] * 7 + [60]), ("bfv_n8192", 1, 8192, [60, 40, 40, 60])]:
        q = ref.coeff_modulus_create(n, bits)
        t = ref.coeff_modulus_create(n, [20])[0] if scheme != 2 else 0
        rc = ref.RefContext(scheme, n, q, t)
        rc.keygen()
        p = tb.EncryptionParameters(scheme)
        p.set_poly_modulus_degree(n)
        p.set_coeff_modulus(q)
        if t:
            p.set_plain_modulus(t)
        ctx = tb.SEALContextCuda(p)
        enc = tb.EncryptorCuda(ctx, rc.public_key_words(), rc.secret_key_words())
        L = ctx.level_limbs(0)
        for batch in (1, 256):
            seeds = np.random.default_rng(1).integers(0, 2 ** 63, (batch, 8), dtype=np.uint64)
            d_seeds = enc._seeds(seeds, batch)
            ct = tb.CiphertextCuda(ctx, 0, 2, scheme == 2, batch=batch)
            for asym in (True, False):
                fname = "troyb_encrypt_zero_asymmetric" if asym else "troyb_encrypt_zero_symmetric"
                key = enc.pk if asym else enc.sk

                def run():
                    _L().call(fname, _vp(ctx.h), 0, _vp(key.ptr), _vp(d_seeds.ptr), _vp(ct.data.ptr), ct.stride, batch, _vp(None))
                sec = timeit(run, warm=3, reps=20)
                rec = {"config": name, "op": "encrypt_zero" + ("" if asym else "_symmetric"), "batch": batch,
                       "ops_per_s": batch / sec, "us_per_call": sec * 1e6,
                       "out_gbs": batch * 2 * L * n * 8 / sec / 1e9}
                if batch == 256:
                    rec["kernels"] = profile(run)
                print(json.dumps(rec), flush=True)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            rc.encrypt_zero(0, True)
        t1 = time.perf_counter()
        for _ in range(reps):
            rc.encrypt_zero(0, False)
        t2 = time.perf_counter()
        print(json.dumps({"config": name, "reference_cpu_1_thread": {"encrypt_zero_per_s": reps / (t1 - t0),
                                                                      "encrypt_zero_symmetric_per_s": reps / (t2 - t1)}}), flush=True)


if __name__ == "__main__":
    main()
