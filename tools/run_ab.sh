#!/bin/bash
# A/B of build variants (tools/build_variant.sh <name> "<-D flags>"): parity + shard tests and one bench line per variant.
# usage: tools/run_ab.sh name1 name2 ...   (libgsb200v_<name>.so); "base" = the default library
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then unset GSB200_LIB; else export GSB200_LIB=$PWD/3dgs.cpp_b200/libgsb200v_$v.so; fi
  t=$( ( GSB_SKIP_HUGE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -1 ) )
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -c 300 gpurun_out/ab_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]);print('$v','fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),{k:round(x,3) for k,x in d['stage_ms'].items() if k in ('preprocess_sort_ms','sort_depth_ms','sort_tile_ms','render_ms','frame_ms')},'| tests:','$t')"
done
