#!/bin/bash
# Round 2, call 11 (1 GPU): tile merge for 64 < N <= 512 -- parity, A/B against the generic kernel at N = 128, configuration 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "tiles or many_samples or 512_samples or sub_pass_split or merge_variants" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02k_pytest.log)"
run() {  # tag, args, env...
    tag=$1; shift; args=$1; shift
    (env "$@" timeout 900 python bench.py --no-cpu-baseline $args --out gpurun_out/r02k_$tag.json) > gpurun_out/r02k_$tag.log 2>&1
    tail -1 gpurun_out/r02k_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('$tag', 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, 'pairs/s', d['kernels']['k_merge'].get('pair_updates_per_s'))
except Exception as e: print('$tag', 'FAILED', e)
"
}
run n128_tiles "--samples 128 --reads 1000000 --steps 2 --warmup 1 --e2e-steps 1" SIMKA_B200_TILE_TIMING=1
true
run config4 "--config 4 --steps 2 --warmup 1 --e2e-steps 1" SIMKA_B200_TILE_TIMING=1
grep -h "tile merge" gpurun_out/r02k_n128_tiles.log gpurun_out/r02k_config4.log | sort | uniq -c | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_merge_tiles|k_group_bins' -c 2 -o gpurun_out/r02k_tiles python bench.py --samples 128 --reads 200000 --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02k_tiles_run.log 2>&1; echo "ncu rc=$?"
