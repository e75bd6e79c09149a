for c in 1 2 3 4 5 6; do
EWB_CHUNKS=$c timeout 120 python bench.py --steps 30 --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('chunks $c', round(d['value'], 1), 'Melem/s', round(d['ms_per_step'], 3), 'ms')
"
done | tee gpurun_out/rp19_chunks.log
