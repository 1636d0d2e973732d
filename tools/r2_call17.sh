mkdir -p gpurun_out
python tools/diag/coarse_check.py garden-standin
for sh in 1 3; do echo "shift $sh"; GSB_COARSE_SHIFT=$sh python tools/diag/coarse_check.py garden-standin | tail -1; done
python tools/diag/coarse_check.py truck-standin | tail -2
for lv in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --tile-cull $lv > gpurun_out/r2p_l$lv.json 2> gpurun_out/r2p_l$lv.err || tail -c 400 gpurun_out/r2p_l$lv.err
python -c "
import json;d=json.loads(open('gpurun_out/r2p_l$lv.json').read().strip().splitlines()[-1]);print('level $lv','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
