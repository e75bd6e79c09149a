mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or golden" 2>&1 | tail -2 | tee gpurun_out/rp17_tests.log
bash tools/gpu_bench_variants.sh rp17 "auto rpa4_8_4" boxgen100_c3d8_linearelastic
make -C edelweissfe_b200/csrc EXTRA=-DEWB_VARIANTS timing > /dev/null 2>&1
EWB_KERNEL=rpb4_8_4 timeout 60 python tools/microbench/rp_timing.py le 2>&1 | tee gpurun_out/rp_timing17.log
