#!/bin/bash
# usage: tools/run_multi.sh N  -- bench.py under torchrun on N GPUs of one box
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 1500 gpurun_out/bench_g$N.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_g$N.json').read().strip().splitlines()[-1]);print('gpus',d['n_gpus'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['config']['parallelism'],{k:round(v,3) for k,v in d['stage_ms'].items()})"
