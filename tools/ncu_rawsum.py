import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum","dram__bytes_read.sum$","dram__bytes_write.sum$","sm__throughput.avg.pct","dram__throughput.avg.pct","smsp__inst_executed.sum$","sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|fp64|lsu|xu|uniform).*sum$","sm__pipe_.*cycles_active.avg.pct","smsp__average_warps_issue_stalled.*ratio","launch__registers","launch__occupancy_limit","sm__warps_active.avg.pct","smsp__issue_active.avg.pct","l1tex__data_bank_conflicts_pipe_lsu.sum$","smsp__warps_eligible.avg.per_cycle","lts__t_bytes.sum$","sm__cycles_elapsed.max","launch__grid_size","l1tex__t_bytes.*lookup_hit.sum$","sm__inst_executed_pipe"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:70], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k,u in zip(hdr, units):
        if any(re.search(p, k) for p in pats) and d[k] not in ("0",""):
            print(f"   {k:<90} {d[k]} {u}")
