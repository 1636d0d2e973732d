#!/bin/bash
# Round profile: bench line + ncu launch list (time per launch) + ncu --set full for one frame's kernels.
tag=${1:-r1}
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 300 gpurun_out/${tag}_bench.err
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
SMI=$!
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
kill $SMI
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_sort_hist|k_project|k_emit|k_blend|k_tile_ranges" -s 80 -c 12 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['roofline'],d.get('cpu_baseline'))"
