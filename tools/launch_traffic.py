"""Condense the launch list of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (CSV, one row
per launch and metric) into a per-kernel table (launches, average time, average DRAM bytes per launch, share of the listed
time of this library's kernels) and profiles/traffic.json (profile name of the kernel -> DRAM bytes per launch at the
bench batch), which bench.py reports as roofline.traffic.

  python tools/launch_traffic.py gpurun_out/<tag>/launches.csv profiles/r1_s6_launches_summary.txt profiles/traffic.json [batch]
"""
import collections
import csv
import json
import sys

# (kernel functor, grid of the bench step's launch at batch B: CKKS N=16384 L=8) -> bench.py profile name. The key inner
# product's grid depends on how many ciphertexts share a staged key slice: it is the slowest launch of its functor.
def bench_step(batch):
    return {
        ("EltTensor", 256 * batch): "ct_tensor_product", ("NttInvTile<13>", 16 * batch): "ntt_inv",
        ("NttFwdTile<13>", 144 * batch): "ntt_fwd_modup", ("NttInvTile<13>", 4 * batch): "ntt_inv_add_half",
        ("NttFwdTile<13>", 32 * batch): "ntt_fwd_moddown",
        # second-generation transforms (round 2)
        ("NttInvTile2<13>", 16 * batch): "ntt_inv", ("NttFwdTile2<13>", 144 * batch): "ntt_fwd_modup",
        ("NttInvTile2<13>", 4 * batch): "ntt_inv_add_half", ("NttFwdTile2<13>", 32 * batch): "ntt_fwd_moddown",
    }


def main(src, out_txt, out_json, batch=256):
    BENCH_STEP = bench_step(int(batch))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": int(r[ix["Grid Size"]].strip("()").split(",")[0])})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[ix["Metric Name"]]] = v * scale
    groups = collections.OrderedDict()
    for d in launches.values():
        if "troyb::" not in d["name"]:
            continue
        short = d["name"].split("troyb::")[-1].split(">(")[0].rstrip(">") if "<troyb::" in d["name"] else d["name"]
        short = d["name"][d["name"].index("<troyb::") + 8:].split(">(")[0]
        groups.setdefault((short, d["grid"]), []).append(d)
    # launches of one functor with one grid can still differ in work (the inner product's grid is the same for 256
    # ciphertexts sharing staged key slices and for the 16-ciphertext chunks of the end-to-end arm): keep the long ones
    for k in list(groups):
        longest = max(x["gpu__time_duration.sum"] for x in groups[k])
        short = [x for x in groups[k] if x["gpu__time_duration.sum"] < 0.5 * longest]
        if short and k[0].startswith("KsInnerProductStaged"):
            groups[k] = [x for x in groups[k] if x["gpu__time_duration.sum"] >= 0.5 * longest]
            groups[(k[0] + " (small batch)", k[1])] = short
    # the bench step's inner product is two launches since the 40-bit rows have their own form: <8, true> (rows with primes of at most
    # 40 bits) and <8, false> (the 60-bit rows); both are tagged and their bytes add up
    ip = [k for k in groups if k[0].startswith("KsInnerProductStaged") and "small batch" not in k[0]]
    for form in ("1", "0"):  # ncu prints the bool template argument as 1 / 0
        sel = [k for k in ip if k[0].replace(" ", "").rstrip(">").endswith("," + form)]
        if sel:
            BENCH_STEP[max(sel, key=lambda k: sum(x["gpu__time_duration.sum"] for x in groups[k]) / len(groups[k]))] = "ks_inner_product"
    if ip and not any(v == "ks_inner_product" for v in BENCH_STEP.values()):
        BENCH_STEP[max(ip, key=lambda k: sum(x["gpu__time_duration.sum"] for x in groups[k]) / len(groups[k]))] = "ks_inner_product"
    step_total = sum(sum(x["gpu__time_duration.sum"] for x in v) / len(v) for k, v in groups.items() if k in BENCH_STEP)
    lines = [f"{'kernel':<28}{'grid':>8}{'launches':>9}{'avg_us':>10}{'dram_rd_MB':>12}{'dram_wr_MB':>12}  bench-step kernel (share of step)"]
    traffic = {}
    for (short, grid), v in groups.items():
        t = sum(x["gpu__time_duration.sum"] for x in v) / len(v)
        rd = sum(x.get("dram__bytes_read.sum", 0) for x in v) / len(v)
        wr = sum(x.get("dram__bytes_write.sum", 0) for x in v) / len(v)
        tag = ""
        if (short, grid) in BENCH_STEP:
            name = BENCH_STEP[(short, grid)]
            tag = f"{name} ({t / step_total:.1%})"
            traffic[name] = traffic.get(name, 0) + int(rd + wr)
        lines.append(f"{short:<28}{grid:>8}{len(v):>9}{t:>10.1f}{rd / 1e6:>12.1f}{wr / 1e6:>12.1f}  {tag}")
    open(out_txt, "w").write("\n".join(lines) + "\n")
    traffic["_source"] = f"{src}: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, bench batch {batch} (python bench.py --steps 2 --warmup 1)"
    json.dump(traffic, open(out_json, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:5])
