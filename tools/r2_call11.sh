N=${1:-4}
mkdir -p gpurun_out
for gather in peer nccl; do
GSB_SHARD_GATHER=$gather timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-extra > gpurun_out/r2k_g${N}_$gather.json 2> gpurun_out/r2k_g${N}_$gather.err || tail -c 1500 gpurun_out/r2k_g${N}_$gather.err
python -c "
import json;d=json.loads(open('gpurun_out/r2k_g${N}_$gather.json').read().strip().splitlines()[-1]);print('$gather gpus',d['n_gpus'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()})
for r in d.get('per_rank'): print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})"
done
