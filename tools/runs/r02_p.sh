#!/bin/bash
# round 2, run P: full-size gradient parity, release_workspaces, pipelined sgemm; then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -s -p no:cacheprovider --durations=6 > gpurun_out/r02p_pytest_fullsize.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02p_pytest_fullsize.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py > gpurun_out/r02p_pytest_rest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02p_pytest_rest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
grep "parity\]" gpurun_out/r02p_pytest_fullsize.log | tail -8; tail -4 gpurun_out/r02p_pytest_fullsize.log; tail -3 gpurun_out/r02p_pytest_rest.log
python -c "import json;d=json.load(open('gpurun_out/r02p_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['infer']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['cpu_baseline'])"
