mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_$tag.json 2> gpurun_out/r2o_$tag.err || tail -c 400 gpurun_out/r2o_$tag.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2o_$tag.json').read().strip().splitlines()[-1]);print('$tag','fps',round(d['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items() if k in ('sort_depth_ms','sort_tile_ms','sort_hist_ms','frame_ms')}, [round(x,3) for x in d['sort_pass_ms_each']])"
}
( GSB_SKIP_HUGE=1 timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x ) 2>&1 | tail -2
run small_depth A=1
run ipt16only GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_ipt16only.so
run small_all GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_small3.so
