"""Condense ncu CSV output into the small summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv        # per-kernel launch count / total / share
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep             # key counters of a --set full capture
"""
import csv
import collections
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__cycles_active.avg", "smsp__inst_executed.sum", "local_load_requests", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum", "smsp__average_warp_latency_issue_stalled_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def short(name):
    m = re.search(r"k_cta<troyb::([^>]*>?)>?", name)
    return m.group(1).rstrip(">(") if m else name[:60]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"{'kernel':<40} {'grid':>14} {'block':>12} {'launches':>8} {'total_us':>10} {'avg_us':>9} {'share':>6}")
    for (k, g, b), (n, t) in agg.items():
        print(f"{k:<40} {g:>14} {b:>12} {n:>8} {t / 1e3:>10.1f} {t / n / 1e3:>9.2f} {t / total:>6.1%}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", short(d.get("Kernel Name", "")), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in hdr:
            base = k.split(" ")[0]
            if base in KEYS or any(base.startswith(x) for x in ("smsp__average_warps_issue_stalled", "smsp__pcsamp")) and d[k] not in ("0", ""):
                print(f"  {k:<95} {d[k]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
