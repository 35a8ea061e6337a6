"""Small tour of every kernel family for compute-sanitizer (memcheck / racecheck): development aid.
  compute-sanitizer --tool memcheck python tests/microbench/sanitizer_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import troy_b200 as tb  # noqa: E402
from oracle import ref  # noqa: E402
from common import (run_decrypt_flow, run_encrypt_flow, run_evaluator_flow, run_matmul_flow, run_plain_flow,  # noqa: E402
                    run_sampler_flow)

tb.KernelProvider.initialize(0)
rng = np.random.default_rng(1)
for scheme, n, bits in [(2, 16384, [60, 40, 60]), (1, 4096, [60, 40, 40, 60]), (3, 32768, [50, 40, 60]), (2, 256, [40, 30, 30, 40])]:
    run_evaluator_flow(ref, tb, scheme, n, bits, rng, batch=2)
    run_plain_flow(ref, tb, scheme, n, bits, rng)
    run_decrypt_flow(ref, tb, scheme, n, bits, rng)
run_matmul_flow(ref, tb, 4096, [60, 40, 60], rng, B=3, I=3, J=3)
for scheme, n, bits in [(2, 16384, [60, 40, 60]), (1, 4096, [60, 40, 40, 60]), (3, 256, [40, 30, 30, 40])]:
    run_encrypt_flow(ref, tb, scheme, n, bits, rng)
run_sampler_flow(ref, tb, 2, 1024, [60, 40, 40, 60], rng)
# the fused mod-up + key inner product (bulk reduce-adds from the tile), opt-in
os.environ["TROYB_KS_FUSED"] = "1"
for scheme, n, bits in [(2, 16384, [60, 40, 60]), (1, 8192, [60, 40, 40, 60]), (3, 32768, [50, 40, 60])]:
    run_evaluator_flow(ref, tb, scheme, n, bits, rng, batch=2)
os.environ["TROYB_KS_FUSED"] = "0"
print("sanitizer target done")
