mkdir -p gpurun_out
( GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x ) 2>&1 | tail -2
run() { tag=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $wl > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err || tail -c 400 gpurun_out/r2u_$tag.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2u_$tag.json').read().strip().splitlines()[-1]);print('$tag','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items() if k in ('render_ms','frame_ms','preprocess_sort_ms')}, d['kernels']['k_blend']['alg_bytes_per_launch'], d['blend_records_staged'])"
}
run tma garden-standin A=1
run notma garden-standin GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_notma.so
run tma_truck truck-standin A=1
run notma_truck truck-standin GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_notma.so
