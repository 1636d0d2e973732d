mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_blend2|k_project" -s 40 -c 6 -o gpurun_out/r2r_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_ncu_full.log 2>&1
tail -2 gpurun_out/r2r_ncu_full.log
