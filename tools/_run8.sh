mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -x -q -k "plugin or cuda_backend or pcg or spmv or surface or compute_host or dirichlet" 2>&1 | tail -25 | tee gpurun_out/rp8_tests.log
