#!/bin/bash
# round 2, run H (N GPUs): data-parallel overlap sweep: backward segments x NCCL channels x SMs left to NCCL
mkdir -p gpurun_out
N=${1:-2}
port=29600
run() {  # name bwdseg visseg comm_sms nchannels
  port=$((port+1))
  LAPB_BWD_SEGMENTS=$2 LAPB_VIS_SEGMENTS=$3 LAPB_COMM_SMS=$4 NCCL_MAX_NCHANNELS=$5 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02h_${N}gpu_$1.json 2> gpurun_out/r02h_${N}gpu_$1.err
  python - <<PY
import json
try:
    txt=open("gpurun_out/r02h_${N}gpu_$1.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$1 seg=$2/$3 comm_sms=$4 nch=$5 :", round(d["value"],1), "samples/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 failed", e)
PY
}
run coarse_16_16 1 1 16 16
run coarse_0_16 1 1 0 16
run coarse_8_8 1 1 8 8
run coarse_4_4 1 1 4 4
run phased_8_8 3 3 8 8
run phased_4_4 3 3 4 4
run phased_0_4 3 3 0 4
run phased2_8_8 2 1 8 8
