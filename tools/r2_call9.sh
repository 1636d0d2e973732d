mkdir -p gpurun_out
tools/ubench/smem_atomic 2>&1 | tee gpurun_out/r2_ubench_smem_atomic.txt
tools/ubench/cub_sort 2>&1 | tee gpurun_out/r2_ubench_cub_sort.txt
( time GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -x ) 2>&1 | tail -5
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err || tail -c 600 gpurun_out/r2i_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1]);print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, [round(x,3) for x in d['sort_pass_ms_each']])"
