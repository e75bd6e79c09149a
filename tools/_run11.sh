mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize_blocks.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/blocks_r02.log
