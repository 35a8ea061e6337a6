"""Summaries from the CSV pages written by tools/r2_ncu_ntt.sh:
    python tools/ncu_csv.py raw  file.raw.csv       # key counters
    python tools/ncu_csv.py segs file.sass.csv [min_share_pct]   # executed SASS cut at barriers / execution-count changes"""
import collections
import csv
import re
import sys

WANT = ["gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    h, v = rows[0], (rows[2] if len(rows) > 2 else rows[1])
    d = dict(zip(h, v))
    for k in WANT:
        if k in d:
            print(f"  {k:78s} {d[k]}")
    for k in sorted(d):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            print(f"  {k[34:-23]:40s} {float(d[k]):.3f}")


def segs(path, minshare=0.7):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    tex = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    stall_keys = [k for k in ix if k.startswith("stall_") and "Not" not in k]
    allops = collections.Counter()

    def new(i):
        return dict(start=i, n=0, s=0, e=0, ops=collections.Counter(), st=collections.Counter(), ex0=None)
    out, cur = [], new(0)
    for i, r in enumerate(data):
        src = r[ix["Source"]]
        ex = int(r[ix["Instructions Executed"]])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else "?"
        base = op.split(".")[0]
        if cur["ex0"] is not None and cur["n"] > 8 and (ex > cur["ex0"] * 1.9 or ex * 1.9 < cur["ex0"]):
            out.append(cur)
            cur = new(i)
        if cur["ex0"] is None or cur["n"] < 4:
            cur["ex0"] = ex
        cur["n"] += 1
        cur["s"] += int(r[ix["# Samples"]])
        cur["e"] += ex
        key = ".".join(op.split(".")[:2]) if base in ("LDG", "STG", "LDS", "STS", "ST", "LD", "IMAD") else base
        cur["ops"][key] += ex
        allops[key] += ex
        for k in stall_keys:
            if r[ix[k]].isdigit():
                cur["st"][k[6:]] += int(r[ix[k]])
        if base in ("BAR", "WARPSYNC", "EXIT", "SYNCS") or "UCGABAR" in op:
            out.append(cur)
            cur = new(i + 1)
    out.append(cur)
    print("samples", tot, "warp-instructions", tex, "executed static", len(data))
    print("mix:", ", ".join(f"{k} {v / tex * 100:.1f}%" for k, v in allops.most_common(16)))
    for s in out:
        if s["e"] == 0 or s["s"] < tot * minshare / 100:
            continue
        ops = ", ".join(f"{k} {v / s['e'] * 100:.0f}%" for k, v in s["ops"].most_common(7))
        st = ", ".join(f"{k} {v / max(1, s['s']) * 100:.0f}%" for k, v in s["st"].most_common(4))
        print(f"@{s['start']:6d}+{s['n']:5d}  samples {s['s'] / tot * 100:5.1f}%  exec {s['e'] / tex * 100:5.1f}% (x{s['ex0']})  | {ops} | {st}")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2])
    else:
        segs(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.7)
