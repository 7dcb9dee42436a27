#!/bin/bash
# round 2, call 12: where configuration 4's tile merge spends its time (per-kernel event timing + one ncu --set full capture of each kernel)
mkdir -p gpurun_out
(SIMKA_B200_TILE_TIMING=1 timeout 900 python bench.py --no-cpu-baseline --config 4 --steps 1 --warmup 1 --e2e-steps 1 --out gpurun_out/r02l_config4.json) > gpurun_out/r02l_config4.log 2>&1
grep "tile merge" gpurun_out/r02l_config4.log | tail -3
tail -1 gpurun_out/r02l_config4.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_merge_tiles|k_group_bins' -c 2 -o gpurun_out/r02l_tiles4 python bench.py --config 4 --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02l_tiles4_run.log 2>&1; echo "ncu rc=$?"
