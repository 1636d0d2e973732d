#!/bin/bash
# bench lines of the other single-GPU BASELINE workloads (C2 bicycle stand-in, C4-size truck stand-in, C1) -> gpurun_out/<tag>_bench_<workload>.json
tag=${1:-r2}
mkdir -p gpurun_out
for wl in bicycle-standin truck-standin c1; do
  timeout 400 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err || tail -c 400 gpurun_out/${tag}_bench_$wl.err
  python -c "
import json;d=json.loads(open('gpurun_out/${tag}_bench_$wl.json').read().strip().splitlines()[-1]);print('$wl','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),'M',round(d['config']['instances_M']),'aabb',round(d['config']['instances_aabb']),{k:round(v,3) for k,v in d['stage_ms'].items()})"
done
