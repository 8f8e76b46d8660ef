#!/bin/bash
# round 2, run C: K2 fused SigLIP attention (kernel test first, under a short timeout), then the whole GPU suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "fused_vit" -q -s -p no:cacheprovider -x > gpurun_out/r02c_pytest_k2.log 2>&1
rc=$?; echo "pytest exit $rc" >> gpurun_out/r02c_pytest_k2.log
if [ $rc -ne 0 ]; then export LAPB_FUSED_VIT=0; echo "K2 failed: continuing with LAPB_FUSED_VIT=0" >> gpurun_out/r02c_pytest_k2.log; fi
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=8 > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -4 gpurun_out/r02c_pytest_k2.log; tail -3 gpurun_out/r02c_pytest_gpu.log; head -c 400 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
