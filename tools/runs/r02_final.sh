#!/bin/bash
# final tree: full GPU suite + smoke (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=3 > gpurun_out/r02final_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/r02final_smoke.log
tail -3 gpurun_out/r02final_pytest_gpu.log; tail -2 gpurun_out/r02final_smoke.log
