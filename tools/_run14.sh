mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or golden" 2>&1 | tail -4 | tee gpurun_out/rp14_tests.log
bash tools/gpu_bench_variants.sh rp14 "auto v1" boxgen100_c3d8_linearelastic boxgen200x100x100_c3d8_vonmises boxgen100_c3d8tl_neohookewa
