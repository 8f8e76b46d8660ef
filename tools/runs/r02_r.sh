#!/bin/bash
# round 2, run R: final validation of the tree (full GPU suite, smoke) + ncu --set full evidence for the GEMM, K1, K2 (pair), K10
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --durations=5 > gpurun_out/r02r_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02r_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02r_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r02r_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 6 -o gpurun_out/r02r_gemm python tools/gemm_prof.py > gpurun_out/r02r_ncu_gemm.log 2>&1
ncu -i gpurun_out/r02r_gemm.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02r_ncu_gemm_summary.txt 2>> gpurun_out/r02r_ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fa_vit|fa_gemma" -c 6 -o gpurun_out/r02r_attn python tools/hbm_prof.py > gpurun_out/r02r_ncu_attn.log 2>&1
ncu -i gpurun_out/r02r_attn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02r_ncu_attn_summary.txt 2>> gpurun_out/r02r_ncu_attn.log
timeout 600 ncu --set full --clock-control none -k regex:denoise_loop -c 1 -o gpurun_out/r02r_denoise python tools/denoise_ncu.py > gpurun_out/r02r_ncu_denoise.log 2>&1
ncu -i gpurun_out/r02r_denoise.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r02r_ncu_denoise_summary.txt 2>> gpurun_out/r02r_ncu_denoise.log
tail -3 gpurun_out/r02r_pytest_gpu.log; tail -2 gpurun_out/r02r_smoke.log; grep -c "Kernel Name" gpurun_out/r02r_ncu_gemm_summary.txt gpurun_out/r02r_ncu_attn_summary.txt gpurun_out/r02r_ncu_denoise_summary.txt
