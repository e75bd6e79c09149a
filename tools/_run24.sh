mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or golden" 2>&1 | tail -2 | tee gpurun_out/rp24_tests.log
bash tools/gpu_bench_variants.sh rp24 "auto auto v1" boxgen100_c3d8_linearelastic
