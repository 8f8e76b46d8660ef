#!/bin/bash
# round 2, run K: paired K2 kernel: tests, then A/B benches (pair on/off, saved MLP act on/off), inference, sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "fused_vit" -q -s -p no:cacheprovider -x > gpurun_out/r02k_pytest_k2.log 2>&1
rc=$?; echo "pytest exit $rc" >> gpurun_out/r02k_pytest_k2.log
if [ $rc -ne 0 ]; then export LAPB_VIT_PAIR=0; echo "pair kernel failed: LAPB_VIT_PAIR=0 for the rest" >> gpurun_out/r02k_pytest_k2.log; fi
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/r02k_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02k_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
LAPB_VIT_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02k_bench_nopair.json 2> gpurun_out/r02k_bench_nopair.err
LAPB_SAVE_MLP_ACT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02k_bench_noact.json 2> gpurun_out/r02k_bench_noact.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02k_bench2.json 2> gpurun_out/r02k_bench2.err
timeout 300 python bench.py --mode infer > gpurun_out/r02k_infer.json 2> gpurun_out/r02k_infer.err
timeout 900 python tools/siglip_sweep.py > gpurun_out/r02k_siglip_sweep.log 2>&1
tail -3 gpurun_out/r02k_pytest_k2.log; tail -3 gpurun_out/r02k_pytest_gpu.log
for f in r02k_bench r02k_bench_nopair r02k_bench_noact r02k_bench2; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"; done
python -c "import json;d=json.load(open('gpurun_out/r02k_infer.json'));print('infer',d['value'],d['device_ms'])"
grep "'images': 256\|'images': 2," gpurun_out/r02k_siglip_sweep.log | head
