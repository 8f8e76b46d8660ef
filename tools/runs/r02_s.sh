#!/bin/bash
mkdir -p gpurun_out
for v in "LAPB_DENOISE_FLAGS=0" "LAPB_DENOISE_FLAGS=1" "LAPB_DENOISE_FLAGS=1 LAPB_DENOISE_CTAS=128"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --mode infer > gpurun_out/r02s_infer_$name.json 2> gpurun_out/r02s_infer_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r02s_infer_$name.json'));print('$v',d['value'],d['device_ms'])"
done
LAPB_DENOISE_FLAGS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -k "fused_denoise" -q -p no:cacheprovider | tail -2
