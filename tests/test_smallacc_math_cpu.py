"""The arithmetic of `SmallAcc` (troy_b200/csrc/kernels_elt.cuh: sums of at most nine products modulo a prime of at most 40 bits, 20-bit
halves of x, a remainder from an FP64 quotient estimate) restated step by step with the kernel's word sizes and checked against
big-integer arithmetic: no accumulator overflows, and the estimate is never off by more than one. The kernel itself is checked word
for word against the reference in test_emul_cpu.py (CPU) and test_parity_gpu.py (`test_keyswitch_with_extreme_operands`)."""
import random
from fractions import Fraction

import numpy as np

M64, M32 = (1 << 64) - 1, (1 << 32) - 1


def small_acc(terms, p):
    ca = cb = ca1 = cb1 = 0
    for x, k in terms:
        xa, xb = x & 0xFFFFF, (x >> 20) & M32
        k0, k1 = k & M32, (k >> 32) & M32
        ca, cb = (ca + xa * k0) & M64, (cb + xb * k0) & M64          # mad.wide.u32 with a 64-bit addend
        ca1, cb1 = (ca1 + xa * k1) & M32, (cb1 + xb * k1) & M32      # 32-bit multiply-adds
    sa, sb = (ca + (ca1 << 32)) & M64, (cb + (cb1 << 32)) & M64
    assert sa + (sb << 20) == sum(x * k for x, k in terms), "an accumulator wrapped"
    v = np.float64(float(Fraction(float(sb)) * 1048576 + Fraction(float(sa))))  # fma(sb, 2^20, sa): one rounding
    q = int(v * (np.float64(1.0) / np.float64(p)))                               # __double2ull_rz(v * pinv)
    assert abs(q - sum(x * k for x, k in terms) // p) <= 1
    r = (sa + ((sb << 20) & M64) - q * p) & M64
    if r >> 63:
        r = (r + p) & M64
    return r - p if r >= p else r


def test_small_acc_matches_big_integer_arithmetic():
    rng = random.Random(20261018)
    for it in range(30000):
        bits = rng.choice([17, 20, 30, 36, 39, 40, 40, 40])
        p = rng.randrange((1 << (bits - 1)) + 1, 1 << bits) | 1
        n = rng.choice([1, 2, 8, 9, 9])
        pick = lambda: rng.choice([0, 1, p - 1, p - 2, rng.randrange(p), rng.randrange(p), 1 << (bits - 1)]) % p  # noqa: E731
        terms = [(p - 1, p - 1)] * n if it % 7 == 0 else [(pick(), pick()) for _ in range(n)]
        assert small_acc(terms, p) == sum(x * k for x, k in terms) % p, (p, terms)
