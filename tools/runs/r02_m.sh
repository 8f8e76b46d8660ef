#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "fused_vit" -q -p no:cacheprovider > gpurun_out/r02m_pytest_k2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02m_pytest_k2.log
timeout 300 python tools/k2_time.py > gpurun_out/r02m_k2_pair.json 2> gpurun_out/r02m_k2_pair.err
tail -3 gpurun_out/r02m_pytest_k2.log; cat gpurun_out/r02m_k2_pair.json
