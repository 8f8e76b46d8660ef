#!/bin/bash
# round 2, run J (8 GPUs): overlap variants at N = 8
mkdir -p gpurun_out
N=${1:-8}
port=29700
run() {  # name bwdseg visseg comm_sms nchannels
  port=$((port+1))
  LAPB_BWD_SEGMENTS=$2 LAPB_VIS_SEGMENTS=$3 LAPB_COMM_SMS=$4 NCCL_MAX_NCHANNELS=$5 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02j_${N}gpu_$1.json 2> gpurun_out/r02j_${N}gpu_$1.err
  python - <<PY
import json
try:
    txt=open("gpurun_out/r02j_${N}gpu_$1.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$1 seg=$2/$3 comm_sms=$4 nch=$5 :", round(d["value"],1), "samples/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 failed", e)
PY
}
run coarse_16_16 1 1 16 16
run phased_16_16 3 3 16 16
run phased_8_8 3 3 8 8
run coarse_24_24 1 1 24 24
