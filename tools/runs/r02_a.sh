#!/bin/bash
# round 2, run A: full GPU test suite (incl. full-size + 2-rank parity), K10 fold variant, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
free -g >> gpurun_out/r02a_gpu.txt; nproc >> gpurun_out/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=15 > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_pytest_gpu.log
LAPB_DENOISE_FOLD=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "fused_denoise or batch1_sampling" -q -s -p no:cacheprovider > gpurun_out/r02a_pytest_fold.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_pytest_fold.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
LAPB_DENOISE_FOLD=1 timeout 300 python bench.py --mode infer > gpurun_out/r02a_infer_fold.json 2> gpurun_out/r02a_infer_fold.err
tail -3 gpurun_out/r02a_pytest_gpu.log; tail -2 gpurun_out/r02a_pytest_fold.log; cat gpurun_out/r02a_infer_fold.json | head -c 600
