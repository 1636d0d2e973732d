#!/bin/bash
# Round-end validation on one B200: full GPU test suite, smoke, the bench line, the ncu launch list + one --set full capture,
# compute-sanitizer.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) 2>&1 | tail -12
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -c 600 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --sh16 > gpurun_out/${tag}_bench_sh16.json 2> gpurun_out/${tag}_bench_sh16.err || tail -c 300 gpurun_out/${tag}_bench_sh16.err
python -c "
import json
for t in ('bench','bench_sh16'):
    d=json.loads(open('gpurun_out/${tag}_'+t+'.json').read().strip().splitlines()[-1]);print(t,'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, d['roofline'], d.get('cpu_baseline'), d['frame_ms_distribution'])"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
SMI=$!
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
kill $SMI
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_sort_hist|k_project|k_emit|k_blend" -s 80 -c 12 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -1 gpurun_out/${tag}_ncu_full.log
bash tools/run_sanitizer.sh $tag
