timeout 300 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for lib in libgsb200.so libgsb200_t512.so; do
GSB200_LIB=$PWD/3dgs.cpp_b200/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err; tail -c 300 gpurun_out/bench_$lib.err
python -c "
import json,sys;d=json.load(open('gpurun_out/bench_$lib.json'));print('$lib',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, [round(x,3) for x in d['sort_pass_ms_each']])"
done
GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200_t512.so timeout 200 python -m pytest tests/test_gpu_sort.py -m gpu -x -q 2>&1 | tail -2
