#!/bin/bash
# 2-GPU sanity of the final tree: 2-rank engine test over NCCL + the bench line at N = 2 (both arms)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -s -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02n2_bench.json 2> gpurun_out/r02n2_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02n2_bench.json') if x.startswith('{')][-1]
d=json.loads(l); print('N=2', d['value'], d['ms_per_step'], d['n_gpus'], d['clocks']['sm_mhz'])
PY
