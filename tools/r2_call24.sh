mkdir -p gpurun_out
( GSB_SKIP_HUGE=1 timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -3
timeout 900 python tools/diag/load_times.py 2>&1 | tail -1 | tee gpurun_out/r2_load_times.json
