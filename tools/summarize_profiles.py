#!/usr/bin/env python
"""Turn the ncu reports brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag>      (reads gpurun_out/<tag>_*.ncu-rep / <tag>_launches.csv)
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]}
        for k in KEEP:
            if k in d:
                rec[k] = d[k] + " " + units[hdr.index(k)]
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) > 0.3:
                        rec["stall_per_issue_" + k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(float(v), 2)
                except ValueError:
                    pass
            if "issue_stalled" in k and k.endswith("per_warp_active.pct"):
                try:
                    if float(v) > 5:
                        rec["stall_" + k.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")] = v + " %"
                except ValueError:
                    pass
        out.append(rec)
    return out


def main():
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    lc = os.path.join(SRC, f"{tag}_launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
        hdr, data = None, []
        for r in rows:
            if r[0] == "ID":
                hdr = r
                continue
            if hdr:
                data.append(dict(zip(hdr, r)))
        agg = defaultdict(list)
        for d in data:
            agg[d["Kernel Name"].split("(")[0]].append(float(d["Metric Value"].replace(",", "")) / 1e3)
        total = sum(sum(v) for v in agg.values())
        summ = [{"kernel": k, "launches": len(v), "avg_us": sum(v) / len(v), "min_us": min(v), "max_us": max(v),
                 "share_of_gpu_time": sum(v) / total} for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]
        json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none (see DESIGN.md 7)", "kernels": summ},
                  open(os.path.join(OUT, f"{tag}_launch_list_summary.json"), "w"), indent=1)
        with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
            f.write(open(lc).read())
    for name in os.listdir(SRC):
        if name.startswith(tag + "_") and name.endswith(".ncu-rep"):
            recs = raw(os.path.join(SRC, name))
            json.dump(recs, open(os.path.join(OUT, name.replace(".ncu-rep", "_summary.json")), "w"), indent=1)
            for r in recs:
                if "k_step" in r["kernel"] or "k_lidar" in r["kernel"] or "k_cost" in r["kernel"]:
                    print(name, r["kernel"][:50], r.get("gpu__time_duration.sum"), r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
