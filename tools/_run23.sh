mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6) 2>&1 | tee gpurun_out/gputests_r02_full.log
timeout 600 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -2 gpurun_out/bench_r02_n1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r02_n1.json'))
print('N=1 value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
print({k:(round(v.get('value',0),1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in d['extra_workloads'].items()})
"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r02_reference_arm.json
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowPipeKernel -s 3 -c 1 -f -o gpurun_out/prof_r02_le $B > /dev/null 2>&1
ls -la gpurun_out/prof_r02_le.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
