#!/bin/bash
# parity tests of the fused kernels under a forced kernel variant: bash tools/gpu_quick_tests.sh VARIANT
EWB_KERNEL=$1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(seeded or edge_shapes or golden) and auto or sweep-" 2>&1 | tail -2
