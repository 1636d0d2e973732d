#!/bin/bash
# usage: gpurun --gpus N -- bash tools/run_multi.sh N : multi-GPU tests (N = 2) + bench.py under torchrun on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_gpu_cli_and_multi.py tests/test_gpu_shard.py -m gpu -q -x ) 2>&1 | tail -2; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err || tail -c 1500 gpurun_out/bench_g${N}.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_g${N}.json').read().strip().splitlines()[-1]);print('gpus',d['n_gpus'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})
for r in d.get('per_rank'): print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
for e in d.get('extra_workloads',[]):
    print(e.get('config',{}).get('workload'), 'fps', e.get('value'), 'e2e', e.get('e2e',{}).get('value'), e.get('stage_ms'), e.get('error'))
    for r in e.get('per_rank') or []: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})"
