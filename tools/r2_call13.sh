N=${1:-4}
mkdir -p gpurun_out
for cull in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --tile-cull $cull > gpurun_out/r2l_g${N}_$cull.json 2> gpurun_out/r2l_g${N}_$cull.err || tail -c 1500 gpurun_out/r2l_g${N}_$cull.err
python -c "
import json;d=json.loads(open('gpurun_out/r2l_g${N}_$cull.json').read().strip().splitlines()[-1]);print('cull $cull gpus',d['n_gpus'],'fps',round(d['value'],1))
for r in d.get('per_rank'): print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})"
done
