mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu 2> gpurun_out/bench_r02_n8_weak_numa.err | grep '^{' > gpurun_out/bench_r02_n8_weak_numa.json
python -c "
import json
d=json.load(open('gpurun_out/bench_r02_n8_weak_numa.json'))
print('N=8 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['e2e']['ms_per_step'],2),'pageable',round(d['e2e']['pageable_numpy_inputs']['value'],1), d['config'].get('host_placement'))
"
tail -2 gpurun_out/bench_r02_n8_weak_numa.err | cut -c1-200
