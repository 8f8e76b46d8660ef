#!/bin/bash
# round 2, run I: compute-sanitizer memcheck on the new kernels; ncu --set full of the HBM-bound kernels + K1 + K2
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "fused_vit or image or slab_split or softmax_backward_fused" > gpurun_out/r02i_sanitizer_kernels.log 2>&1
echo "exit $?" >> gpurun_out/r02i_sanitizer_kernels.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "augmentation or non_224 or batch1_sampling or orbax" > gpurun_out/r02i_sanitizer_parity.log 2>&1
echo "exit $?" >> gpurun_out/r02i_sanitizer_parity.log
LAPB_DENOISE_MODE=cluster timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_denoise and debug_small" > gpurun_out/r02i_sanitizer_cluster.log 2>&1
echo "exit $?" >> gpurun_out/r02i_sanitizer_cluster.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_vit|fa_gemma|geglu_bwd|norm_bwd|norm_fwd_warp|adamw|image_augment|resid_norm" -c 24 -o gpurun_out/r02i_hbm python tools/hbm_prof.py > gpurun_out/r02i_ncu.log 2>&1
ncu -i gpurun_out/r02i_hbm.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_hbm_table.py > gpurun_out/r02i_ncu_hbm_kernels.md 2>> gpurun_out/r02i_ncu.log
tail -3 gpurun_out/r02i_sanitizer_kernels.log; tail -3 gpurun_out/r02i_sanitizer_parity.log; tail -3 gpurun_out/r02i_sanitizer_cluster.log; cat gpurun_out/r02i_ncu_hbm_kernels.md | head -40
