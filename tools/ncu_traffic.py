"""Join an ncu launch list of one train step that carries `dram__bytes_read.sum` / `dram__bytes_write.sum` with the GEMM
calls logged by tools/step_prof.py, and write the GEMM family's DRAM traffic per step into profiles/ncu_traffic.json
(`bench.py` reports it as `roofline.traffic`, per launch, next to the algorithmic bytes of the same launches).

  ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
      --clock-control none --csv --log-file gpurun_out/traffic_step.csv python tools/step_prof.py
  python tools/ncu_traffic.py gpurun_out/traffic_step.csv gpurun_out/gemm_calls.json
"""
import collections, csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csvf, callsf = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(l for l in open(csvf) if not l.startswith("==")))
calls = json.load(open(callsf))
per = collections.OrderedDict()
for r in rows:
    e = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
    e[r["Metric Name"]] = v * scale
kern = list(per.values())
gem = [k for k in kern if "gemm_bf16" in k["name"]]
assert len(gem) == len(calls), (len(gem), len(calls))


def algorithmic_bytes(c):
    bt = c["batch_i"] * c["batch_o"]
    M, N, K, epi = c["M"], c["N"], c["K"], c["epi"]
    nb = 2 * N if epi == 4 else N                       # GeGLU: gate and up rows
    b = bt * 2 * (M * K + nb * K)                       # bf16 operands
    b += bt * M * N * (4 if c["f32"] else 2)            # output
    if c.get("accumulate"):
        b += bt * M * N * 4
    if epi in (2, 3):
        b += bt * M * N * 2                             # residual read
    if epi in (1, 3) and c.get("c2"):
        b += bt * M * N * 2                             # pre-activation / branch output
    if epi == 4 and c.get("c2"):
        b += bt * 2 * M * N * 2                         # saved g, u
    if epi == 6:
        b += bt * 4 * M * N * 2                         # g, u in; dg, du out
    if epi == 7:
        b += bt * M * N * 2
    return b


tot_dram = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in gem)
tot_alg = sum(algorithmic_bytes(c) for c in calls)
tot_ms = sum(k.get("gpu__time_duration.sum", 0) for k in gem)
all_dram = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in kern)
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(out_path)) if os.path.exists(out_path) else {}
d["gemm_family_per_step"] = {"launches": len(gem), "dram_bytes": tot_dram, "algorithmic_bytes": tot_alg,
                             "ratio": tot_dram / tot_alg, "gemm_ms_under_ncu": tot_ms,
                             "source": os.path.basename(csvf)}
d["whole_step"] = {"launches": len(kern), "dram_bytes": all_dram}
by_shape = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, c in zip(gem, calls):
    key = f"{c['M']}x{c['N']}x{c['K']}" + (f"b{c['batch_i'] * c['batch_o']}" if c['batch_i'] * c['batch_o'] > 1 else "") + f"e{c['epi']}"
    e = by_shape[key]
    e[0] += 1
    e[1] += k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)
    e[2] += algorithmic_bytes(c)
d["per_shape"] = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "algorithmic_bytes_per_launch": v[2] / v[0],
                      "ratio": v[1] / v[2]} for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:16]}
json.dump(d, open(out_path, "w"), indent=1)
print(json.dumps(d["gemm_family_per_step"], indent=1))
