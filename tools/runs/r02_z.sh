#!/bin/bash
# K10 v2 with TMA bulk copies for the gate/up weights
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -p no:cacheprovider -k"
timeout 300 $T "fused_denoise or batch1_sampling or denoise" 2>&1 | tail -15 > gpurun_out/r02z_pytest_v2.log; tail -3 gpurun_out/r02z_pytest_v2.log
for v in "LAPB_DENOISE_CTAS=40" "LAPB_DENOISE_CTAS=128"; do
  env $v timeout 300 $T "fused_denoise and v2" 2>&1 | tail -3 | sed "s/^/$v: /"
done
for v in "LAPB_DENOISE_FLAGS=0" "LAPB_DENOISE_CTAS=128"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --mode infer > gpurun_out/r02z_infer_$name.json 2> gpurun_out/r02z_infer_$name.err
  python -c "import json;d=json.load(open('gpurun_out/r02z_infer_$name.json'));print('$v',d['value'],d['device_ms'])" || tail -3 gpurun_out/r02z_infer_$name.err
done
timeout 300 python tools/denoise_prof.py full --no-per-op > gpurun_out/r02z_prof_v2.json 2> gpurun_out/r02z_prof_v2.err || tail -5 gpurun_out/r02z_prof_v2.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fused_denoise and debug_small and v2 and not strong" > gpurun_out/r02z_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02z_memcheck.log | tail -3
