mkdir -p gpurun_out
(time timeout 900 python bench.py --steps 50 > gpurun_out/bench_r02a_n1.json) 2> gpurun_out/bench_r02a_n1.err
tail -3 gpurun_out/bench_r02a_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02a_n1.json'))
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'fp64',round(d['roofline']['fp64']['frac'],3),'launches',d['gpu_launches'])
print('e2e',d.get('e2e'))
print('extra',json.dumps(d.get('extra_workloads'),indent=1)[:1500])
print('cpu',d.get('cpu_baseline'))
PY
