"""Summarise an `ncu --set full` report (raw page CSV on stdin) as a markdown table of achieved DRAM bandwidth.

usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_hbm_table.py > profiles/rNN_ncu_hbm_kernels.md
"""
import csv
import json
import os
import re
import sys

peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
HBM = float(peaks.get("hbm_gbs", peaks.get("hbm_gbps", 6540.8)))

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}


def val(r, name):
    return float(r[ix[name]].replace(",", "")) * SCALE.get(units[ix[name]], 1)


print(f"ncu --set full --clock-control none, one launch each at the LAP-3B B=32 shapes (tools/hbm_prof.py); "
      f"peak = {HBM:.1f} GB/s (MEASURED_PEAKS.json)")
print()
print("| kernel | grid | duration (us) | DRAM bytes (MB) | achieved GB/s | % of HBM peak | tensor pipe % | regs |")
print("|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    t = val(r, "gpu__time_duration.sum")
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    print(f"| {name} | {r[ix['Grid Size']]} | {t * 1e6:.1f} | {b / 1e6:.1f} | {b / t / 1e9:.0f} | "
          f"{100 * b / t / 1e9 / HBM:.1f} | {tp:.1f} | {r[ix['launch__registers_per_thread']]} |")
