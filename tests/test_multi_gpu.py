"""`-m gpu`: the NCCL path on real GPUs (skipped on a 1-GPU box): tests/multi_gpu_check.py under torchrun, one rank per GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_digit_sharded_key_switch_and_ciphertext_shards_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert f"multi_gpu_check ok on {n} ranks" in r.stdout


def test_two_devices_in_one_process(ref):
    """Evaluators on two GPUs driven from ONE process and one host thread (the reference pins device 0,
    src/kernelprovider.cuh:30): per-device kernel attributes, per-context device selection and allocations on the context's
    device. The same flow runs on both devices, interleaved, and must match the reference CPU library word for word."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import troy_b200 as tb
    from common import make_pair, random_ct_words, upload_keys, assert_same
    n, bits = 16384, [60, 40, 40, 60]
    rng = np.random.default_rng(2)
    per_dev = []
    for dev in (0, 1):
        tb.KernelProvider.initialize(dev)
        rc, ctx = make_pair(ref, tb, ref.CKKS, n, bits)
        assert ctx.device == dev
        rlk, gk, d_rlk, d_gk = upload_keys(tb, ctx, rc, ref)
        per_dev.append((rc, ctx, rlk, gk, d_rlk, d_gk, tb.EvaluatorCuda(ctx)))
    w = [random_ct_words(per_dev[0][0], 0, 2, n, rng) for _ in range(2)]
    cts = []
    for rc, ctx, rlk, gk, d_rlk, d_gk, ev in per_dev:  # allocations for device 0 happen after device 1 was made current
        cts.append([tb.CiphertextCuda.from_host(ctx, x, 0, True, 2.0 ** 30) for x in w])
    for step in range(3):  # interleave the two devices call by call
        for (rc, ctx, rlk, gk, d_rlk, d_gk, ev), (a, b) in zip(per_dev, cts):
            if step == 0:
                ev.multiply_inplace(a, b)
            elif step == 1:
                ev.relinearize_inplace(a, d_rlk)
            else:
                ev.rotate_vector_inplace(a, 1, d_gk)
    for (rc, ctx, rlk, gk, d_rlk, d_gk, ev), (a, b) in zip(per_dev, cts):
        ra, rb = rc.ciphertext(w[0], 0, True, 2.0 ** 30), rc.ciphertext(w[1], 0, True, 2.0 ** 30)
        rc.eval(ref.OP_MULTIPLY, ra, rb)
        rc.eval(ref.OP_RELIN, ra, keys=rlk)
        rc.eval(ref.OP_ROTATE_VECTOR, ra, keys=gk, arg=1)
        assert_same(a, ra, f"device {ctx.device}")
    tb.KernelProvider.initialize(0)
