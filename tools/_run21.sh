mkdir -p gpurun_out
run() { # N tag args...
  N=$1; TAG=$2; shift; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu "$@" 2> gpurun_out/${TAG}.err | grep '^{' > gpurun_out/${TAG}.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}.json'))
    print('${TAG}', 'N', d['n_gpus'], d['config']['workload'], d['scaling'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1) if 'e2e' in d else None, 'parity', d.get('multi_gpu_parity',{}).get('rel_err'))
except Exception as e:
    print('${TAG} failed', e)
PY
  tail -2 gpurun_out/${TAG}.err | cut -c1-300
}
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3) | tee gpurun_out/multi_r02_8gpu.log
run 8 bench_r02_n8_weak
run 4 bench_r02_n4_weak
run 8 bench_r02_n8_strong_le --scaling strong
run 4 bench_r02_n4_strong_le --scaling strong
run 2 bench_r02_n2_strong_le --scaling strong
run 8 bench_r02_n8_config5_nh200 --workload boxgen200_c3d8tl_neohookewa
run 8 bench_r02_n8_strong_vm --workload boxgen200x100x100_c3d8_vonmises --scaling strong
