bash tools/gpu_rp.sh rp5 "rp4_4_4 rps4_8_4 rpr4_4_4 rpr4_8_4"
for v in rps4_8_4 rpr4_8_4; do EWB_LIB_PATH=$PWD/tools/microbench/libewb_nopf.so EWB_KERNEL=$v timeout 120 python bench.py --steps 30 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('nopf $v', round(d['value'], 1), 'Melem/s')
"; done
for v in rpr4_8_4; do EWB_KERNEL=$v timeout 60 python tools/microbench/rp_timing.py le; done 2>&1 | tee gpurun_out/rp_timing5.log
