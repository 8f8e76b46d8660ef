#!/bin/bash
# round 2, run O (8 GPUs): NVLS with fewer channels (NCCL_NVLS_NCHANNELS) -> smaller SM carve-out
mkdir -p gpurun_out
N=${1:-8}
port=29900
probe() {
  port=$((port+1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/nccl_allreduce_probe.py 2>/dev/null | grep '^{' | sed "s/^/$* /" | tee -a gpurun_out/r02o_allreduce_probe.jsonl
}
probe NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=8
probe NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=4
probe NCCL_MAX_NCHANNELS=8
run() {  # name bwdseg visseg comm_sms extra-env...
  port=$((port+1))
  name=$1; b=$2; v=$3; c=$4; shift 4
  env LAPB_BWD_SEGMENTS=$b LAPB_VIS_SEGMENTS=$v LAPB_COMM_SMS=$c "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02o_${N}gpu_$name.json 2> gpurun_out/r02o_${N}gpu_$name.err
  python - <<PY
import json
try:
    txt=open("gpurun_out/r02o_${N}gpu_$name.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$name seg=$b/$v comm_sms=$c $* :", round(d["value"],1), "samples/s", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e)
PY
}
run coarse_8_nvls8 1 1 8 NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=8
run phased_8_nvls8 3 3 8 NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=8
run coarse_4_nvls4 1 1 4 NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=4
run phased2_8_nvls8 2 1 8 NCCL_MAX_NCHANNELS=16 NCCL_NVLS_NCHANNELS=8
