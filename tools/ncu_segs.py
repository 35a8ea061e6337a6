"""SASS segments of one kernel in an ncu report: the executed code cut at barriers / warp syncs and at changes of the execution
count, with each segment's share of samples, instruction mix and leading stall reasons.
    python tools/ncu_segs.py report.ncu-rep [min_share_pct]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
tex = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
stall_keys = [k for k in ix if k.startswith("stall_") and "Not" not in k]


def new(i):
    return dict(start=i, n=0, s=0, e=0, ops=collections.Counter(), st=collections.Counter(), ex0=None)


segs, cur = [], new(0)
for i, r in enumerate(data):
    src = r[ix["Source"]]
    ex = int(r[ix["Instructions Executed"]])
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else "?")
    base = op.split(".")[0]
    if cur["ex0"] is not None and cur["n"] > 8 and (ex > cur["ex0"] * 1.9 or ex * 1.9 < cur["ex0"]) and ex + cur["ex0"] > 0:
        segs.append(cur)
        cur = new(i)
    if cur["ex0"] is None or cur["n"] < 4:
        cur["ex0"] = ex
    cur["n"] += 1
    cur["s"] += int(r[ix["# Samples"]])
    cur["e"] += ex
    key = ".".join(op.split(".")[:2]) if base in ("LDG", "STG", "LDS", "STS", "ST", "LD", "IMAD") else base
    cur["ops"][key] += ex
    for k in stall_keys:
        if r[ix[k]].isdigit():
            cur["st"][k[6:]] += int(r[ix[k]])
    if base in ("BAR", "WARPSYNC", "UCGABAR_ARV", "UCGABAR_WAIT", "EXIT") or op.startswith("BAR") or "UCGABAR" in op:
        segs.append(cur)
        cur = new(i + 1)
segs.append(cur)
print("samples", tot, "warp-instructions", tex, "static", len(data))
for s in segs:
    if s["e"] == 0 or s["s"] < tot * minshare / 100:
        continue
    ops = ", ".join(f"{k} {v / s['e'] * 100:.0f}%" for k, v in s["ops"].most_common(7))
    st = ", ".join(f"{k} {v / max(1, s['s']) * 100:.0f}%" for k, v in s["st"].most_common(4))
    print(f"@{s['start']:6d}+{s['n']:5d}  samples {s['s'] / tot * 100:5.1f}%  exec {s['e'] / tex * 100:5.1f}% (x{s['ex0']})  | {ops} | {st}")
