"""Per-source-line view of one kernel in an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py report.ncu-rep [top]
Prints the instruction mix (SASS page), stall-reason totals and the hottest CUDA source lines."""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    # ---- SASS page: opcode mix + stall totals
    rows = page(rep, [])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    ops, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) != len(hdr) or not r[idx["# Samples"]].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
        op = ".".join((m.group(2) if m else "?").split(".")[:2])
        ops[op] += int(r[idx["Instructions Executed"]])
        samp[op] += int(r[idx["# Samples"]])
        for k, i in idx.items():
            if k.startswith("stall_") and "Not" not in k and r[i].isdigit():
                stall[k[6:]] += int(r[i])
    te, ts = sum(ops.values()) or 1, sum(samp.values()) or 1
    print("static SASS", len(rows) - 2, "warp instructions executed", te, "samples", ts)
    for op, c in ops.most_common(32):
        print(f"  {op:18s} exec {c / te * 100:5.1f}%  samples {samp[op] / ts * 100:5.1f}%")
    print("stall totals %:", [(k, round(v / ts * 100, 1)) for k, v in stall.most_common(12)])
    # ---- CUDA lines
    rows = page(rep, ["--print-source", "cuda,sass"])
    f = "?"
    lines = []
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            # two columns are named "Source": metrics sit after the second one
            ci = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        try:
            ns, ne = int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]])
        except ValueError:
            continue
        st = {k[6:]: int(r[i]) for k, i in ci.items() if k.startswith("stall_") and "Not" not in k and r[i].isdigit() and int(r[i])}
        lines.append((ns, ne, f, r[0], r[1], st))
    print("hottest source lines (samples %, executed %):")
    for ns, ne, f, ln, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
        tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"  {ns / ts * 100:5.1f}% {ne / te * 100:5.1f}%  {f}:{ln}: {src.strip()[:80]}  {tops}")


main()
