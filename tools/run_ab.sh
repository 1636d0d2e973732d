for lib in "$@"; do
GSB200_LIB=$PWD/3dgs.cpp_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err; tail -c 300 gpurun_out/bench_$lib.err
python -c "
import json,sys;d=json.load(open('gpurun_out/bench_$lib.json'));print('$lib',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, [round(x,3) for x in d['sort_pass_ms_each']])"
done
