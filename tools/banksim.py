"""Exhaustive shared-memory bank-conflict check for the tile NTT rounds (64-bit accesses: a half-warp of 16 lanes must hit
16 distinct 8-byte bank pairs). Used to choose pad_idx and the per-tile round plans in troy_b200/csrc/ntt.cuh.

    python tools/banksim.py            # checks the plans compiled into ntt.cuh
"""
import itertools
import sys


def conflicts(LOGT, NT, rounds, pad):
    res, s0 = [], 0
    for RR in rounds:
        BLO = LOGT - s0 - RR
        K = 1 << RR
        NV = (1 << LOGT) // K
        G = max(1, NV // NT)
        worst = 0
        for g in range(G):
            for k in range(K):
                for w0 in range(0, min(NT, NV), 16):
                    banks = {}
                    for lane in range(16):
                        v = w0 + lane + g * NT
                        if v >= NV:
                            continue
                        lo, hi = v & ((1 << BLO) - 1), v >> BLO
                        idx = (hi << (BLO + RR)) | (k << BLO) | lo
                        banks.setdefault(pad(idx) % 16, set()).add(pad(idx))
                    worst = max(worst, max(len(s) for s in banks.values()))
        res.append(((BLO, RR), worst))
        s0 += RR
    return res


def pad_default(i):
    return i + (i >> 4)


PLANS = {12: (256, [4, 4, 4]), 13: (256, [3, 3, 4, 3]), 14: (512, [3, 4, 4, 3])}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "search":
        pads = {"i+(i>>4)": pad_default, "i+(i>>4)+(i>>6)": lambda i: i + (i >> 4) + (i >> 6), "i+(i>>3)+(i>>6)": lambda i: i + (i >> 3) + (i >> 6),
                "i+(i>>4)+(i>>7)": lambda i: i + (i >> 4) + (i >> 7), "i+(i>>4)+(i>>8)": lambda i: i + (i >> 4) + (i >> 8),
                "i+(i>>4)+(i>>5)": lambda i: i + (i >> 4) + (i >> 5), "i^((i>>4)&15)": lambda i: i ^ ((i >> 4) & 15) ^ ((i >> 8) & 15) ^ ((i >> 12) & 15)}
        for LOGT, NT in [(14, 1024), (14, 512), (13, 512), (13, 256), (12, 256), (12, 512)]:
            for name, pad in pads.items():
                good = []
                for n in (4, 5):
                    for r in itertools.product([1, 2, 3, 4], repeat=n):
                        if sum(r) != LOGT or max(r) > (3 if NT * 8 >= (1 << LOGT) and NT * 16 > (1 << LOGT) else 4):
                            continue
                        if (1 << LOGT) // (1 << max(r)) < NT:
                            continue
                        c = conflicts(LOGT, NT, list(r), pad)
                        if all(w == 1 for _, w in c):
                            good.append(r)
                print(LOGT, NT, name, good[:12])
    else:
        for LOGT, (NT, r) in PLANS.items():
            print(LOGT, NT, r, conflicts(LOGT, NT, r, pad_default))
