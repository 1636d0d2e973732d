mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_cli_and_multi.py -m gpu -q -x -k "in_process" ) 2>&1 | grep -E "Error|error|passed|failed|assert" | head -20
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err || tail -c 600 gpurun_out/r2f_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1]);print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'util',d['blend_lane_utilisation'],'cons',d['blend_records_consumed'],'visits',d['blend_warp_visits'],'M',d['config']['instances_M'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_sort_hist|k_project|k_emit|k_blend" -s 80 -c 12 -o gpurun_out/r2f_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_full.log 2>&1
tail -2 gpurun_out/r2f_ncu_full.log; ls -la gpurun_out/
