# N GPUs of one box: the real multi-GPU tests, then bench.py under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>&1 | head -14
bash tools/probe_vulkan.sh > gpurun_out/r2_vulkan_probe.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_cli_and_multi.py tests/test_gpu_shard.py -m gpu -q -x --durations=5 ) 2>&1 | tail -15
for gather in peer nccl; do
GSB_SHARD_GATHER=$gather timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2e_g${N}_$gather.json 2> gpurun_out/r2e_g${N}_$gather.err || tail -c 1500 gpurun_out/r2e_g${N}_$gather.err
python -c "
import json;d=json.loads(open('gpurun_out/r2e_g${N}_$gather.json').read().strip().splitlines()[-1]);print('$gather gpus',d['n_gpus'],'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),{k:round(v,3) for k,v in d['stage_ms'].items()});print(d.get('per_rank'));print([ (e.get('workload'), e.get('value'), e.get('error')) for e in d.get('extra_workloads',[])])"
done
