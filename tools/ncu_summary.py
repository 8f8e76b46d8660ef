"""`ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py` -> one block of the metrics that matter per kernel
(duration, DRAM bytes, tensor pipe, issue utilisation, registers, shared memory, top warp stalls)."""
import csv, re, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
for r in rows[2:]:
    print("-----")
    print(f"{'Kernel Name':70s} {re.sub(r'[(].*', '', r[ix['Kernel Name']])}   grid {r[ix['Grid Size']]}")
    for k in keep:
        if k in ix:
            print(f"{k:70s} {units[ix[k]]:16s} {r[ix[k]]}")
    st = [(h, float(r[ix[h]].replace(',', ''))) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[ix[h]] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1.0
    top = sorted(st, key=lambda t: -t[1])[:5]
    print(f"{'top warp stalls (pc samples)':70s} " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for h, v in top))
