"""Instruction mix and top stall sites of one kernel from an ncu report's source page.
    python tools/ncu_source.py gpurun_out/prof.ncu-rep <kernel regex> [launch index]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][:2])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
        op = ".".join((m.group(2) if m else "?").split(".")[:2])
        ops[op] += int(r[idx["Instructions Executed"]])
        samp[op] += int(r[idx["# Samples"]])
    te = sum(ops.values()) or 1
    print("static instructions", len(data), "warp instructions executed", te, "samples", tot)
    for op, c in ops.most_common(28):
        print(f"  {op:18s} exec {c / te * 100:5.1f}%  samples {samp[op] / tot * 100:5.1f}%")
    data.sort(key=lambda r: -int(r[idx["# Samples"]]))
    print("top stall sites:")
    for r in data[:30]:
        st = {k[6:]: int(r[i]) for k, i in idx.items() if k.startswith("stall_") and "Not" not in k and r[i].isdigit() and int(r[i])}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"  {int(r[idx['# Samples']]):5d}  {r[idx['Source']].strip()[:64]:64s} {top}")


if __name__ == "__main__":
    main()


def segments(rep, pat, skip="0"):
    """Samples per code segment (split at barriers), in address order: shows which phase of a kernel the time goes to."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) == len(hdr) and r[idx["# Samples"]].isdigit() and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
    seg, segs = dict(n=0, s=0, ex=0, dp=0, ldg=0, first=None, stall=collections.Counter()), []
    for r in data:
        src = r[idx["Source"]].strip()
        seg["n"] += 1
        seg["s"] += int(r[idx["# Samples"]])
        seg["ex"] += int(r[idx["Instructions Executed"]])
        if re.search(r"\b(DFMA|DADD|DMUL)\b", src):
            seg["dp"] += int(r[idx["Instructions Executed"]])
        if "LDG" in src:
            seg["ldg"] += int(r[idx["Instructions Executed"]])
        for k, i in idx.items():
            if k.startswith("stall_") and "Not" not in k and r[i].isdigit():
                seg["stall"][k[6:]] += int(r[i])
        if re.search(r"BAR\.SYNC|UCGABAR_WAIT|EXIT", src):
            segs.append(seg)
            seg = dict(n=0, s=0, ex=0, dp=0, ldg=0, first=None, stall=collections.Counter())
    segs.append(seg)
    print(f"{'seg':>3} {'static':>6} {'exec':>10} {'dp':>10} {'ldg':>8} {'samples%':>8}  top stalls")
    for i, s in enumerate(segs):
        if s["ex"] == 0:
            continue
        print(f"{i:3d} {s['n']:6d} {s['ex']:10d} {s['dp']:10d} {s['ldg']:8d} {s['s'] / tot * 100:8.1f}  {s['stall'].most_common(4)}")


if __name__ == "__main__" and len(sys.argv) > 4 and sys.argv[4] == "segments":
    segments(sys.argv[1], sys.argv[2], sys.argv[3])
