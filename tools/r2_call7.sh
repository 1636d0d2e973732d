N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_cli_and_multi.py tests/test_gpu_shard.py -m gpu -q -x --durations=5 ) 2>&1 | grep -E "Error|error|passed|failed|assert|^real|s call" | head -20
GSB_SHARD_GATHER=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2n_g${N}.json 2> gpurun_out/r2n_g${N}.err || tail -c 1500 gpurun_out/r2n_g${N}.err
python -c "
import json;d=json.loads(open('gpurun_out/r2n_g${N}.json').read().strip().splitlines()[-1]);print('gpus',d['n_gpus'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()});print(d.get('per_rank'));print([ (e.get('config',{}).get('workload'), e.get('value'), e.get('e2e'), e.get('error')) for e in d.get('extra_workloads',[])])"
( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --durations=5 ) 2>&1 | grep -E "Error|error|passed|failed|assert|^real|s call" | head -20
