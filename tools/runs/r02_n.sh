#!/bin/bash
# round 2, run N (8 GPUs): raw all-reduce time per NCCL setting, then two more overlap variants
mkdir -p gpurun_out
N=${1:-8}
port=29800
probe() {
  port=$((port+1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/nccl_allreduce_probe.py 2>/dev/null | grep '^{' | tee -a gpurun_out/r02n_allreduce_probe.jsonl
}
probe NCCL_MAX_NCHANNELS=16
probe NCCL_MAX_NCHANNELS=24
probe NCCL_MAX_NCHANNELS=32
probe NCCL_MAX_NCHANNELS=16 NCCL_ALGO=Ring
probe NCCL_MAX_NCHANNELS=16 NCCL_ALGO=NVLS
probe NCCL_MAX_NCHANNELS=8 NCCL_ALGO=NVLS
run() {  # name bwdseg visseg comm_sms nchannels extra-env
  port=$((port+1))
  env LAPB_BWD_SEGMENTS=$2 LAPB_VIS_SEGMENTS=$3 LAPB_COMM_SMS=$4 NCCL_MAX_NCHANNELS=$5 $6 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02n_${N}gpu_$1.json 2> gpurun_out/r02n_${N}gpu_$1.err
  python - <<PY
import json
try:
    txt=open("gpurun_out/r02n_${N}gpu_$1.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$1 seg=$2/$3 comm_sms=$4 nch=$5 $6:", round(d["value"],1), "samples/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 failed", e)
PY
}
run coarse_32_32 1 1 32 32 X=1
run coarse_8_8_nvls 1 1 8 8 NCCL_ALGO=NVLS
run phased_8_8_nvls 3 3 8 8 NCCL_ALGO=NVLS
