#!/bin/bash
# round 2, run F: K2 v2 + slab split-K + PDL: kernel tests, full suite, inference A/B, train A/B (K2 on/off)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "fused_vit or slab_split or gemm" -q -s -p no:cacheprovider > gpurun_out/r02f_pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02f_pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/r02f_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02f_pytest_gpu.log
LAPB_PDL=1 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02f_pytest_gpu_pdl.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02f_pytest_gpu_pdl.log
timeout 300 python bench.py --mode infer > gpurun_out/r02f_infer.json 2> gpurun_out/r02f_infer.err
LAPB_SMALL_M_SPLIT_K=0 timeout 300 python bench.py --mode infer > gpurun_out/r02f_infer_nosplit.json 2> gpurun_out/r02f_infer_nosplit.err
LAPB_PDL=1 timeout 300 python bench.py --mode infer > gpurun_out/r02f_infer_pdl.json 2> gpurun_out/r02f_infer_pdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
LAPB_FUSED_VIT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02f_bench_novit.json 2> gpurun_out/r02f_bench_novit.err
LAPB_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r02f_bench_pdl.json 2> gpurun_out/r02f_bench_pdl.err
tail -3 gpurun_out/r02f_pytest_new.log; tail -3 gpurun_out/r02f_pytest_gpu.log; tail -3 gpurun_out/r02f_pytest_gpu_pdl.log
for f in r02f_infer r02f_infer_nosplit r02f_infer_pdl; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['device_ms'])"; done
for f in r02f_bench r02f_bench_novit r02f_bench_pdl; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'])"; done
