#!/bin/bash
# K10: bias loads hoisted out of the epilogues of the modulation GEMM and of action_out_proj
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fused_denoise or batch1_sampling" 2>&1 | tail -3
timeout 300 python tools/denoise_prof.py full --no-per-op > gpurun_out/r02p3_prof.json 2> gpurun_out/r02p3_prof.err || tail -5 gpurun_out/r02p3_prof.err
timeout 300 python bench.py --mode infer > gpurun_out/r02p3_infer.json 2> gpurun_out/r02p3_infer.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02p3_prof.json'))['us_per_layer_step']
print({k: round(v['median'],1) for k,v in d.items() if k.startswith('prologue') or k.startswith('final') or k.startswith('action_in')})
d=json.load(open('gpurun_out/r02p3_infer.json')); print(d['value'], d['device_ms'])
PY
