#!/bin/bash
# K10 prologue: time-MLP with 8 weight loads in flight per lane (flags bit 5) vs the plain loop; prologue slots of the profile
mkdir -p gpurun_out
for f in 0 32; do
  LAPB_DENOISE_FLAGS=$f timeout 300 python tools/denoise_prof.py full --no-per-op > gpurun_out/r02p2_prof_f$f.json 2> gpurun_out/r02p2_prof_f$f.err || tail -5 gpurun_out/r02p2_prof_f$f.err
  LAPB_DENOISE_FLAGS=$f timeout 300 python bench.py --mode infer > gpurun_out/r02p2_infer_f$f.json 2> gpurun_out/r02p2_infer_f$f.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02p2_prof_f$f.json'))['us_per_layer_step']
print('flags $f', {k: round(v['median'],1) for k,v in d.items() if k.startswith('prologue') or k.startswith('final') or k.startswith('action_in')})
d=json.load(open('gpurun_out/r02p2_infer_f$f.json')); print('flags $f', d['value'], d['device_ms'])
PY
done
