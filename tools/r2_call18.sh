mkdir -p gpurun_out
( GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -4
python tools/diag/coarse_check.py garden-standin
GSB_COARSE_SHIFT=1 python tools/diag/coarse_check.py garden-standin | tail -1
python tools/diag/coarse_check.py truck-standin | tail -2
python tools/diag/coarse_check.py bicycle-standin | tail -2
for lv in 2; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --tile-cull $lv > gpurun_out/r2q_l$lv.json 2> gpurun_out/r2q_l$lv.err || tail -c 400 gpurun_out/r2q_l$lv.err
python -c "
import json;d=json.loads(open('gpurun_out/r2q_l$lv.json').read().strip().splitlines()[-1]);print('level $lv','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
