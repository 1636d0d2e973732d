mkdir -p gpurun_out
nvidia-smi -L | head -3
( time GSB_SKIP_HUGE=1 timeout 1200 python -m pytest tests -m gpu -q --durations=8 -x ) 2>&1 | tail -30
for wl in garden-standin; do
  timeout 400 python bench.py --steps 100 --warmup 10 --workload $wl > gpurun_out/r2d_$wl.json 2> gpurun_out/r2d_$wl.err || tail -c 800 gpurun_out/r2d_$wl.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2d_$wl.json').read().strip().splitlines()[-1]);print('$wl','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e'].get('sync_value',0),1),{k:round(v,3) for k,v in d['stage_ms'].items()}, d.get('roofline'), d.get('cpu_baseline'))"
done
