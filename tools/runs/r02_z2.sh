#!/bin/bash
# K10 v2: sensitivity of the loop time to the L1 left over next to the dynamic shared memory
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -p no:cacheprovider -k"
timeout 300 $T "fused_denoise or batch1_sampling or denoise" 2>&1 | tail -3
for v in "LAPB_DENOISE_PAD_SMEM=0" "LAPB_DENOISE_PAD_SMEM=6144" "LAPB_DENOISE_PAD_SMEM=12288" "LAPB_DENOISE_PAD_SMEM=16384"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --mode infer > gpurun_out/r02z2_infer_$name.json 2> gpurun_out/r02z2_infer_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r02z2_infer_$name.json'));print('$v',d['value'],d['device_ms'])" || tail -3 gpurun_out/r02z2_infer_$name.err
done
