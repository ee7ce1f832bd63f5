#!/usr/bin/env python
"""One line per kernel from an .ncu-rep (`ncu --set full`): launches, mean duration, DRAM bytes per launch, DRAM / tensor / tc-pipe /
issue utilisation.  usage: python profiles/ncu_table.py rep.ncu-rep [traffic.json]  (the optional JSON receives the per-launch DRAM
bytes that bench.py quotes as `roofline.traffic`)."""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "tc": "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l2hit": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread", "smem": "launch__shared_mem_per_block_dynamic",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "grid": "Grid Size", "block": "Block Size",
     "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "lsu_smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
     "tc_smem": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("espnet::", "")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        a = agg.setdefault(k, {"n": 0, "grid": r[col[M["grid"]]], "block": r[col[M["block"]]]})
        a["n"] += 1
        for key, m in M.items():
            if key in ("grid", "block") or m not in col or r[col[m]] == "":
                continue
            v = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]].split("/")[0], 1.0)
            a[key] = a.get(key, 0.0) + v
    print("# %s   (cold-cache, serialised launches: compare shares and per-launch DRAM bytes, not absolute times)" % rep)
    print("%-58s %3s %9s %9s %9s %6s %6s %6s %6s %6s %6s %6s %5s %7s" % ("kernel", "n", "us/launch", "rd MB", "wr MB", "dram%", "tens%", "tc%", "tcsm%", "fma%", "issue%", "L2hit%", "regs", "smemKB"))
    traffic = {}
    for k, a in agg.items():
        n = a["n"]
        g = lambda key: a.get(key, 0.0) / n
        print("%-58s %3d %9.1f %9.1f %9.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %5d %7.1f" % (k[:58], n, g("dur"), g("rd") / 1e6, g("wr") / 1e6, g("dram"), g("tensor"),
              g("tc"), g("tc_smem"), g("fma"), g("issue"), g("l2hit"), int(g("regs")), g("smem") / 1e3))
        traffic[k] = {"dram_bytes_read": g("rd"), "dram_bytes_write": g("wr"), "launches": n, "us_per_launch": g("dur")}
    if len(sys.argv) > 2:
        json.dump(traffic, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
