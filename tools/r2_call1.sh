mkdir -p gpurun_out
bash tools/probe_vulkan.sh > gpurun_out/r2_vulkan_probe.txt 2>&1
tools/ubench/f32x2 > gpurun_out/r2_ubench_f32x2.txt 2>&1
cat gpurun_out/r2_ubench_f32x2.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -25
for wl in garden-standin bicycle-standin truck-standin; do
  timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --workload $wl > gpurun_out/r2a_$wl.json 2> gpurun_out/r2a_$wl.err || tail -c 400 gpurun_out/r2a_$wl.err
  python -c "
import json;d=json.load(open('gpurun_out/r2a_$wl.json'));print('$wl','M',round(d['config']['instances_M']),'aabb',round(d['config']['instances_aabb']),'vis',round(d['config']['visible']),'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),round(d['e2e'].get('sync_value',0),1),{k:round(v,3) for k,v in d['stage_ms'].items()}, d.get('frame_ms_distribution'))"
done
( time timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload synthetic-50m > gpurun_out/r2a_50m.json 2> gpurun_out/r2a_50m.err ) 2>&1 | tail -3; tail -c 600 gpurun_out/r2a_50m.err
python -c "
import json;d=json.load(open('gpurun_out/r2a_50m.json'));print('50m','M',round(d['config']['instances_M']),'aabb',round(d['config']['instances_aabb']),'vis',round(d['config']['visible']),'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})"
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err ) 2>&1 | tail -3; cat gpurun_out/r2a_ref.json | cut -c1-600
head -40 gpurun_out/r2_vulkan_probe.txt
