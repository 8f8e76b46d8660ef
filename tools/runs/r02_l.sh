#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/k2_time.py > gpurun_out/r02l_k2_pair.json 2> gpurun_out/r02l_k2_pair.err
LAPB_VIT_PAIR=0 timeout 300 python tools/k2_time.py > gpurun_out/r02l_k2_single.json 2> gpurun_out/r02l_k2_single.err
cat gpurun_out/r02l_k2_pair.json; echo; cat gpurun_out/r02l_k2_single.json; tail -3 gpurun_out/r02l_k2_pair.err
