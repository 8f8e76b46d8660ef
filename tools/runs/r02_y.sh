#!/bin/bash
# short check of the parametrised K10 test (v2, v2 strong barrier, round-1 layout via flags bit 7)
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fused_denoise or batch1_sampling" 2>&1 | tail -4
