mkdir -p gpurun_out
tools/ubench/f32x2 > gpurun_out/r2_ubench_f32x2.txt 2>&1; cat gpurun_out/r2_ubench_f32x2.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 -x ) 2>&1 | tail -30
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --workload ${WL:-garden-standin} > gpurun_out/r2b_$tag.json 2> gpurun_out/r2b_$tag.err || tail -c 400 gpurun_out/r2b_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/r2b_$tag.json'));print('$tag','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e'].get('sync_value',0),1),{k:round(v,3) for k,v in d['stage_ms'].items() if k in ('preprocess_ms','preprocess_sort_ms','sort_depth_ms','sort_tile_ms','render_ms','frame_ms')}, 'med',round(d['frame_ms_distribution']['median'],4))"
}
run v2 A=1
run v1 GSB_BLEND_VARIANT=1
run v2_staged GSB_HOST_DIRECT=0
for v in add0 b128 mb7 mb6 chk16 chk4; do run $v GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_$v.so; done
WL=truck-standin run truck_v2 A=1
WL=truck-standin run truck_v1 GSB_BLEND_VARIANT=1
