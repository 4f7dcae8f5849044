#!/usr/bin/env python
"""Turn ncu output into the text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv "<command line>" > profiles/rNN_launches_bench_summary.txt
        (csv from: ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ... <command>)
    python tools/ncu_summary.py full gpurun_out/k_render.ncu-rep "<command line>" > profiles/rNN_k_render_full_summary.txt
        (report from: ncu --set full --clock-control none --import-source on -k regex:... -o ... <command>);
        also rewrites profiles/k_render_traffic.json (DRAM bytes per launch) which bench.py reads.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def launches(path, cmd):
    rows = []
    with open(path, newline="") as f:
        text = f.read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        elif r.get("Metric Unit") in ("ms", "msecond"):
            ns *= 1e6
        rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], r["Stream"], r["Grid Size"], ns))
    groups = OrderedDict()
    for _, k, _, grid, ns in rows:
        groups.setdefault((k, grid), []).append(ns)
    total = sum(ns for *_, ns in rows) or 1.0
    print(cmd)
    print(f"(the {len(rows)} launches ncu saw; per-launch times are cold-cache and serialised by the profiler -- compare SHARES)\n")
    print(f"{'kernel / grid':70s} {'launches':>8s} {'total ms':>10s} {'share %':>8s} {'mean us':>9s} {'min us':>9s} {'max us':>9s}")
    for (k, grid), v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        print(f"{(k + ' grid ' + grid)[:70]:70s} {len(v):8d} {sum(v) / 1e6:10.3f} {100 * sum(v) / total:8.2f} "
              f"{sum(v) / len(v) / 1e3:9.2f} {min(v) / 1e3:9.2f} {max(v) / 1e3:9.2f}")
    print("\nfirst 40 launches (id, kernel, stream, grid, ns):")
    for i, k, st, grid, ns in rows[:40]:
        print(f"{i:5d} {k[:40]:40s} stream {st:>3s} grid {grid:>14s} {ns:10.0f}")


def full(path, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out[out.find('"ID"'):])))
    names, units, data = rd[0], rd[1], rd[2:]
    print(cmd)
    traffic, insts = [], []
    for n, row in enumerate(data):
        rec = dict(zip(names, row))
        unit = dict(zip(names, units))
        print(f"\n--- launch {n}: {rec.get('Kernel Name', '?')[:100]}")
        for k in KEEP:
            if k in rec:
                print(f"{k:75s} {rec[k]:>18s} {unit.get(k, '')}")
        stalls = {k[len(STALL):]: float(v.replace(",", "") or 0) for k, v in rec.items()
                  if k.startswith(STALL) and not k.endswith("_not_issued") and v}
        tot = sum(stalls.values()) or 1.0
        top = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]}
        print(f"{'warp stall samples (% of all)':75s} {top}")
        try:
            rd_b = float(rec["dram__bytes_read.sum"].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit["dram__bytes_read.sum"], 1)
            wr_b = float(rec["dram__bytes_write.sum"].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit["dram__bytes_write.sum"], 1)
            traffic.append(int(rd_b + wr_b))
            insts.append(int(float(rec["smsp__inst_executed.sum"].replace(",", ""))))
        except Exception:
            pass
    if traffic:
        with open(os.path.join(ROOT, "profiles", "k_render_traffic.json"), "w") as f:
            json.dump({"kernel": "k_render_persistent", "source": os.path.basename(path) + " (ncu --set full), see the full summary under profiles/",
                       "dram_bytes_per_launch": int(sum(traffic) / len(traffic)), "per_launch": traffic,
                       "warp_instructions_per_launch": int(sum(insts) / len(insts)) if insts else None,
                       "warp_instructions": insts}, f, indent=1)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    (launches if mode == "launches" else full)(path, cmd)
