#!/bin/bash
# round 2, run B: GPU test suite after fixes; K10c (cluster + tile-major packed operands): correctness, phase profile, latency
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=8 > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_pytest_gpu.log
LAPB_DENOISE_MODE=cluster timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -k "fused_denoise or batch1_sampling or full_size_sample" -q -s -p no:cacheprovider > gpurun_out/r02b_pytest_cluster.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_pytest_cluster.log
LAPB_DENOISE_MODE=cluster timeout 300 python tools/denoise_prof.py > gpurun_out/r02b_prof_cluster.json 2> gpurun_out/r02b_prof_cluster.err
timeout 300 python tools/denoise_prof.py > gpurun_out/r02b_prof_grid.json 2> gpurun_out/r02b_prof_grid.err
LAPB_DENOISE_MODE=cluster timeout 300 python bench.py --mode infer > gpurun_out/r02b_infer_cluster.json 2> gpurun_out/r02b_infer_cluster.err
timeout 600 ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_infer.csv python tools/infer_prof.py > gpurun_out/r02b_infer_prof.log 2>&1
tail -3 gpurun_out/r02b_pytest_gpu.log; tail -3 gpurun_out/r02b_pytest_cluster.log; head -c 700 gpurun_out/r02b_infer_cluster.json; grep -A14 phase_us_per_layer_step gpurun_out/r02b_prof_cluster.json
