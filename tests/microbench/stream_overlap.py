"""Does the bench step gain from lanes? The 256-ciphertext multiply + relinearize step (bench.py's configuration, inputs
resident in HBM) issued as S independent lanes of 256 / S ciphertexts on S streams, so that one lane's DRAM-bound kernels
(tensor product, key inner product) can run beside another lane's latency-bound transforms. One JSON line per S."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import troy_b200 as tb  # noqa: E402

N, L, B = 16384, 8, int(os.environ.get("OVERLAP_BATCH", "256"))
STEPS = int(os.environ.get("OVERLAP_STEPS", "30"))
SCALE = 2.0 ** 40
tb.KernelProvider.initialize(0)
q = tb.CoeffModulus.Create(N, [60] + [40] * 7 + [60])
p = tb.EncryptionParameters(tb.SchemeType.ckks)
p.set_poly_modulus_degree(N)
p.set_coeff_modulus(q)
ctx = tb.SEALContextCuda(p)
dev = torch.device("cuda", 0)


def uniform(prefix, moduli):
    out = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
    for i, m in enumerate(moduli):
        out[..., i, :] = torch.randint(0, m, (*prefix, N), device=dev, dtype=torch.int64)
    return out


key = uniform((L, 2), ctx.level_moduli(-1))
rlk = tb.RelinKeysCuda(ctx)
rlk.keys = [tb.host.ExternalArray(key.data_ptr(), key.numel())]
a_in, b_in = uniform((B, 2), q[:L]), uniform((B, 2), q[:L])
work = torch.empty(B, 3, L, N, dtype=torch.int64, device=dev)
cur = torch.cuda.current_stream()
ref_out = None
for S in (1, 2, 3, 4, 8):
    nb = (B + S - 1) // S
    lanes = []
    for s in range(S):
        c0, c1 = s * nb, min(B, (s + 1) * nb)
        if c0 >= c1:
            continue
        st = torch.cuda.Stream(device=dev) if S > 1 else cur
        lanes.append(dict(
            st=st, ev=tb.EvaluatorCuda(ctx, stream=st.cuda_stream),
            a=tb.CiphertextCuda.from_device(ctx, a_in[c0:c1].data_ptr(), c1 - c0, 2, 0, True, SCALE),
            b=tb.CiphertextCuda.from_device(ctx, b_in[c0:c1].data_ptr(), c1 - c0, 2, 0, True, SCALE),
            o=tb.CiphertextCuda.from_device(ctx, work[c0:c1].data_ptr(), c1 - c0, 2, 0, True, SCALE, capacity=3)))

    def step():
        for ln in lanes:
            ln["ev"].multiply(ln["a"], ln["b"], ln["o"])
            ln["ev"].relinearize_inplace(ln["o"], rlk)

    def fork():
        for ln in lanes:
            if ln["st"] is not cur:
                ln["st"].wait_stream(cur)

    def join():
        for ln in lanes:
            if ln["st"] is not cur:
                cur.wait_stream(ln["st"])

    for _ in range(3):
        fork(); step(); join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fork()
    for _ in range(STEPS):
        step()
    join()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    out = work[:, :2].clone()
    if ref_out is None:
        ref_out = out
    print(json.dumps({"lanes": len(lanes), "batch": B, "ms_per_step": round(ms, 4), "ops_per_s": round(B / ms * 1e3, 1),
                      "same_words_as_one_lane": bool(torch.equal(out, ref_out))}), flush=True)
