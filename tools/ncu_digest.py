"""Digest of one `ncu --set full` capture for profiles/: the counters that the roofline discussion uses (from `--page raw --csv`)
plus the executed instruction mix and the stall reasons per issue (from the trimmed `--page source --csv`, tools/ncu_trim.py).

    python tools/ncu_digest.py <name>_raw.csv <name>_sass.csv > profiles/<file>.txt
"""
import collections
import csv
import re
import sys

RAW = [
    "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    print(f"kernel: {d['Kernel Name'][0][:110]}")
    for k in RAW:
        if k in d:
            print(f"  {k:<78} {d[k][0]:>16} {d[k][1]}")
    stalls = {}
    for k, (v, _) in d.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
        if m:
            stalls[m.group(1)] = float(v.replace(",", ""))
    if stalls:
        print("  stall reasons, warps per issue: " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]))


def sass(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix, si = hdr.index("Instructions Executed"), hdr.index("Source")
    mix, n = collections.Counter(), 0
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        tok = r[si].strip().split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        parts = op.split(".")
        name = parts[0] + ("." + parts[1] if len(parts) > 1 and parts[0] in ("IMAD", "IADD3", "LDG", "LDS", "STS", "STG", "I2F", "F2I") else "")
        mix[name] += int(r[ix])
        n += int(r[ix])
    print(f"  executed warp instructions {n}; mix: " + ", ".join(f"{k} {v / n:.1%}" for k, v in mix.most_common(16)))


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        sass(sys.argv[2])
