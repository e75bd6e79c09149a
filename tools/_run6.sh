mkdir -p gpurun_out
for v in rpa4_8_4 rp4_4_4; do
  echo "== tests with EWB_KERNEL=$v"
  EWB_KERNEL=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seeded or edge_shapes or golden" 2>&1 | tail -6
done 2>&1 | tee gpurun_out/rp6_tests.log
bash tools/gpu_bench_variants.sh rp6 "rpr4_8_4 rpa4_8_4 rpb4_8_4 rpr4_4_4"
for v in rpa4_8_4; do EWB_KERNEL=$v timeout 60 python tools/microbench/rp_timing.py le; done 2>&1 | tee gpurun_out/rp_timing6.log
