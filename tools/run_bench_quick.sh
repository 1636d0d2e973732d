#!/bin/bash
# quick GPU check: parity tests + one bench line (usage: tools/run_bench_quick.sh [tag])
tag=${1:-q}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 300 gpurun_out/bench_$tag.err
python -c "
import json,sys;d=json.load(open('gpurun_out/bench_$tag.json'));print('$tag','M',round(d['config']['instances_M']),round(d['value'],1),round(d['e2e']['value'],1),round(d['e2e'].get('sync_value',0),1),{k:round(v,3) for k,v in d['stage_ms'].items()}, [round(x,3) for x in d['sort_pass_ms_each']])"
