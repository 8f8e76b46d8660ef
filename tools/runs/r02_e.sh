#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gemm_small_m.py > gpurun_out/r02e_gemm_small_m.log 2>&1
tail -8 gpurun_out/r02e_gemm_small_m.log
