#!/bin/bash
# K10 variants: flags bit0 = no burst prefetch, bit1 = split barrier (arrive before the preloads), bit2 = per-phase prefetch
mkdir -p gpurun_out
for f in 1 3 5 7; do
  LAPB_DENOISE_FLAGS=$f timeout 300 python bench.py --mode infer > gpurun_out/r02t_infer_f$f.json 2> gpurun_out/r02t_infer_f$f.err
  python -c "import json;d=json.load(open('gpurun_out/r02t_infer_f$f.json'));print('flags $f',d['value'],d['device_ms'])"
done
for f in 1 3 7; do
  LAPB_DENOISE_FLAGS=$f timeout 300 python tools/denoise_prof.py full --no-per-op > gpurun_out/r02t_prof_f$f.json 2> gpurun_out/r02t_prof_f$f.err || tail -5 gpurun_out/r02t_prof_f$f.err
done
LAPB_DENOISE_FLAGS=7 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -k "denoise" -q -p no:cacheprovider | tail -2
LAPB_DENOISE_FLAGS=3 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -k "denoise" -q -p no:cacheprovider | tail -2
