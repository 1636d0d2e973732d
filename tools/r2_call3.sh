mkdir -p gpurun_out
( time GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_host_renderer.py tests/test_gpu_cli_and_multi.py -m gpu -q --durations=6 -x ) 2>&1 | tail -40
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --workload ${WL:-garden-standin} > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err || tail -c 600 gpurun_out/r2c_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/r2c_$tag.json'));print('$tag','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e'].get('sync_value',0),1),{k:round(v,3) for k,v in d['stage_ms'].items() if k in ('preprocess_ms','preprocess_sort_ms','sort_depth_ms','sort_tile_ms','render_ms','frame_ms')}, 'med',round(d['frame_ms_distribution']['median'],4), 'visits/s', d.get('blend_warp_visits_per_s'))"
}
run v3 A=1
for v in pred0 add1 mb6 b128 mb6b128 mb5b128 mb10b128; do run $v GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_$v.so; done
WL=truck-standin run truck_v3 A=1
( time GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_sort.py -m gpu -q -x ) 2>&1 | tail -8
