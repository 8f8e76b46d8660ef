#!/bin/bash
# round 2, run G (2 GPUs): 2-rank engine gradient test over NCCL, data-parallel bench with the phased all-reduce vs coarse
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02g_gpus.txt
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -s -p no:cacheprovider > gpurun_out/r02g_pytest_ddp.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02g_pytest_ddp.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02g_bench_${N}gpu_phased.json 2> gpurun_out/r02g_bench_${N}gpu_phased.err
LAPB_BWD_SEGMENTS=1 LAPB_VIS_SEGMENTS=1 LAPB_COMM_SMS=16 NCCL_MAX_NCHANNELS=16 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02g_bench_${N}gpu_coarse.json 2> gpurun_out/r02g_bench_${N}gpu_coarse.err
tail -3 gpurun_out/r02g_pytest_ddp.log
for f in phased coarse; do python -c "import json;d=json.load(open('gpurun_out/r02g_bench_${N}gpu_$f.json'));print('$f',d['value'],d['ms_per_step'],d['e2e']['value'])"; tail -2 gpurun_out/r02g_bench_${N}gpu_$f.err; done
