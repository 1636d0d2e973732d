mkdir -p gpurun_out
( GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -q -x ) 2>&1 | tail -3
python tools/diag/coarse_check.py garden-standin | tail -1
python tools/diag/coarse_check.py truck-standin | tail -1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2s.json 2> gpurun_out/r2s.err || tail -c 400 gpurun_out/r2s.err
python -c "
import json;d=json.loads(open('gpurun_out/r2s.json').read().strip().splitlines()[-1]);print('fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})"
