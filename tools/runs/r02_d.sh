#!/bin/bash
# round 2, run D: image kernels, fused softmax backward, full suite, A/B bench, SigLIP sweep 224/384, GEMM-family DRAM traffic
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "image or softmax_backward_fused" -q -s -p no:cacheprovider > gpurun_out/r02d_pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02d_pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=8 > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02d_bench_fused.json 2> gpurun_out/r02d_bench_fused.err
LAPB_FUSED_SOFTMAX_BWD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02d_bench_unfused_smbwd.json 2> gpurun_out/r02d_bench_unfused_smbwd.err
timeout 900 python tools/siglip_sweep.py > gpurun_out/r02d_siglip_sweep.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02d_traffic_step.csv python tools/step_prof.py > gpurun_out/r02d_step_prof.log 2>&1
tail -3 gpurun_out/r02d_pytest_new.log; tail -3 gpurun_out/r02d_pytest_gpu.log; head -c 300 gpurun_out/r02d_bench_fused.json; echo; head -c 300 gpurun_out/r02d_bench_unfused_smbwd.json; echo; tail -3 gpurun_out/r02d_siglip_sweep.log; tail -2 gpurun_out/r02d_step_prof.log
