"""Summarise `ncu --set full` captures for profiles/: per kernel launch the metrics the DESIGN / bench cite, and
profiles/traffic.json (dram read+write bytes per launch, keyed by workload and by the kernel name bench.py reports) so
that `roofline.traffic` comes from a capture instead of a hand-entered constant.

    python tools/ncu_summary.py --out profiles/ncu_summary_r2.json --traffic profiles/traffic.json \
        pics8_8state_b256=gpurun_out/final/prof_decode.ncu-rep mic3_tiles=gpurun_out/final/prof_mic3.ncu-rep ...
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import re
import subprocess

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(full: str) -> str:
    """void micgpu::k_ans_decode_packed<(int)8, (int)1>(...) -> k_ans_decode_packed<8> (the name bench.py prints)"""
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", full)
    if not m:
        return full
    name, targs = m.group(1), m.group(2) or ""
    nums = re.findall(r"\(int\)(\d+)|\b(\d+)\b", targs)
    first = next((a or b for a, b in nums), None)
    if name.startswith("k_ans_decode") and first:
        return f"{name}<{first}>"
    return name


def read_report(path: str):
    """path: a .ncu-rep, or the output of `ncu -i rep --page raw --csv` saved on the GPU box (capture_round.sh)"""
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = {"kernel": r[idx["Kernel Name"]]}
        for w in WANT:
            if w in idx:
                v, u = r[idx[w]], units[idx[w]]
                try:
                    f = float(v.replace(",", ""))
                except ValueError:
                    continue
                if w.startswith("dram__bytes"):
                    f *= UNIT_SCALE.get(u, 1)
                    u = "byte"
                k[w] = f
                k[w + ".unit"] = u
        out.append(k)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic", required=True)
    ap.add_argument("reports", nargs="+", help="workload_key=path.ncu-rep")
    a = ap.parse_args()
    summary, traffic = {}, {}
    for item in a.reports:
        key, path = item.split("=", 1)
        launches = read_report(path)
        summary[key] = {"report": path, "launches": launches}
        t = traffic.setdefault(key, {})
        for k in launches:
            if "dram__bytes_read.sum" in k and "dram__bytes_write.sum" in k:
                n = short_name(k["kernel"])
                # several launches of one kernel in a capture (unit ranges on their own streams): sum = one step's traffic
                t[n] = t.get(n, 0) + int(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"])
    traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per step, summed over the launches of each kernel in one `ncu --set full` "
                        "capture of the workload (tools/capture_round.sh, tools/ncu_summary.py)")
    json.dump(summary, open(a.out, "w"), indent=1)
    json.dump(traffic, open(a.traffic, "w"), indent=1)
    print("wrote", a.out, a.traffic)


if __name__ == "__main__":
    main()
