#!/bin/bash
# round 2, run Q (final tree): ncu launch lists (duration + DRAM bytes) of one train step and one sample_actions
mkdir -p gpurun_out
timeout 900 ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02q_traffic_step.csv python tools/step_prof.py > gpurun_out/r02q_step_prof.log 2>&1
cp gpurun_out/gemm_calls.json gpurun_out/r02q_gemm_calls_step.json
timeout 600 ncu --nvtx --nvtx-include "STEP/" --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02q_launches_infer.csv python tools/infer_prof.py > gpurun_out/r02q_infer_prof.log 2>&1
cp gpurun_out/gemm_calls_infer.json gpurun_out/r02q_gemm_calls_infer.json
tail -2 gpurun_out/r02q_step_prof.log; tail -2 gpurun_out/r02q_infer_prof.log
