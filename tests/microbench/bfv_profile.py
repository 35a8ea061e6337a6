"""Per-kernel device time of one BFV multiply + relinearize step (BASELINE config 1: N=8192 {60,40,40,60}), development aid."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import troy_b200 as tb  # noqa: E402
from troy_b200.lib import library  # noqa: E402
from configs_bench import setup, uniform  # noqa: E402

torch.cuda.init()
tb.KernelProvider.initialize(0)
n, B = 8192, 256
ctx, ev, q, L, key, rlk, gk = setup(tb.SchemeType.bfv, n, [60, 40, 40, 60])
a_in, b_in = uniform((B, 2), q[:L], n), uniform((B, 2), q[:L], n)
work = torch.empty(B, 3, L, n, dtype=torch.int64, device="cuda")
a = tb.CiphertextCuda.from_device(ctx, a_in.data_ptr(), B, 2, 0, False, 1.0)
b = tb.CiphertextCuda.from_device(ctx, b_in.data_ptr(), B, 2, 0, False, 1.0)
d = tb.CiphertextCuda.from_device(ctx, work.data_ptr(), B, 2, 0, False, 1.0, capacity=3)
for it in range(3):
    if it == 2:
        library().call("troyb_profile_enable", 1)
    ev.multiply(a, b, d)
    ev.relinearize_inplace(d, rlk)
    torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
library().call("troyb_profile_report", buf, len(buf))
ks = [json.loads(x) for x in buf.value.decode().strip().splitlines() if x]
tot = sum(k["ms"] for k in ks)
for k in ks:
    print(f'{k["kernel"]:<26}{k["launches"]:>3} launches {k["grid"]:>8} ctas {k["ms"]:8.3f} ms {k["ms"] / tot:6.1%}')
print("total", round(tot, 3), "ms for", B, "->", round(B / tot * 1e3), "ops/s")
