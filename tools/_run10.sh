mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --no-extra --no-cpu > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err; tail -2 gpurun_out/bench_r02b_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r02b_n1.json'))
print('N=1 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['e2e']['ms_per_step'],2),'pageable',round(d['e2e']['pageable_numpy_inputs']['value'],1))
"
for N in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_r02b_n$N.json 2> gpurun_out/bench_r02b_n$N.err; tail -3 gpurun_out/bench_r02b_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r02b_n$N.json'))
print('N=$N value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('multi_gpu_parity'))
"
done
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/multi_r02b.log
