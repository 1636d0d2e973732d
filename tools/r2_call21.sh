mkdir -p gpurun_out
run() { tag=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $wl > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err || tail -c 400 gpurun_out/r2t_$tag.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2t_$tag.json').read().strip().splitlines()[-1]);print('$tag','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items() if k in ('render_ms','frame_ms','preprocess_sort_ms')})"
}
run tma garden-standin A=1
run notma garden-standin GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_notma.so
run tma2 garden-standin A=1
run notma2 garden-standin GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_notma.so
run tma_truck truck-standin A=1
run notma_truck truck-standin GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_notma.so
