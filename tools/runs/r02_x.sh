#!/bin/bash
# round 2, run X: validation of the tree with K10 v2 (full GPU suite, smoke, default bench) + ncu evidence for the new kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=5 > gpurun_out/r02x_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02x_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r02x_smoke.log
timeout 600 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02x_bench.json') if x.startswith('{')][-1]
d=json.loads(l)
print('bench', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], 'infer', d['infer']['value'], d['infer']['device_ms'], d['infer']['roofline']['frac'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoise_loop -c 1 -o gpurun_out/r02x_denoise python tools/denoise_ncu.py > gpurun_out/r02x_ncu_denoise.log 2>&1
ncu -i gpurun_out/r02x_denoise.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02x_ncu_denoise_summary.txt 2>> gpurun_out/r02x_ncu_denoise.log
timeout 600 ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02x_launches_infer.csv python tools/infer_prof.py > gpurun_out/r02x_infer_prof.log 2>&1
tail -3 gpurun_out/r02x_pytest_gpu.log; tail -2 gpurun_out/r02x_smoke.log; cat gpurun_out/r02x_ncu_denoise_summary.txt
