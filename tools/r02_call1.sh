#!/bin/bash
# Round 2, first GPU call: whole GPU suite (incl. full-size parity), bench of every configuration, stage-1 A/B, ncu evidence.
#   gpurun --timeout 2700 -- 'bash tools/r02_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r02_pytest_gpu.log)"
timeout 600 python bench.py --steps 5 --warmup 3 --verify --out gpurun_out/r02_c2_base.json > gpurun_out/r02_c2_base.log 2>&1; echo "config2 rc=$? $(tail -1 gpurun_out/r02_c2_base.log | cut -c1-400)"
run() {  # tag, env...
    tag=$1; shift
    (env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/r02_ab_$tag.log 2>&1
    tail -1 gpurun_out/r02_ab_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['notes'].get('stage1'), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e:
    print('$tag', 'FAILED', e)
"
}
run superk_flat SIMKA_B200_SUPERK=1
run superk_thread SIMKA_B200_SUPERK=1 SIMKA_B200_SUPERK_EXPAND=thread
run k2minb3 SIMKA_B200_K2_MINB=3
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --out gpurun_out/r02_config4.json > gpurun_out/r02_config4.log 2>&1; echo "config4 rc=$? $(tail -1 gpurun_out/r02_config4.log | cut -c1-600)"
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --out gpurun_out/r02_config5.json > gpurun_out/r02_config5.log 2>&1; echo "config5 rc=$? $(tail -1 gpurun_out/r02_config5.log | cut -c1-600)"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 150 > gpurun_out/r02_refarm.log 2>&1; echo "refarm rc=$? $(tail -1 gpurun_out/r02_refarm.log | cut -c1-300) ... $(tail -1 gpurun_out/r02_refarm.log | grep -o '"sample": "[^"]*"' | head -1)"
timeout 300 python tools/bench_sketch.py > gpurun_out/r02_sketch.log 2>&1; echo "sketch rc=$? $(tail -1 gpurun_out/r02_sketch.log | cut -c1-500)"
timeout 300 python tools/bench_sketch.py --filter > gpurun_out/r02_sketch_filter.log 2>&1; echo "sketch -filter rc=$? $(tail -1 gpurun_out/r02_sketch_filter.log | cut -c1-500)"
timeout 400 python tools/bench_cli.py --samples 32 --reads 1000000 > gpurun_out/r02_cli.log 2>&1; echo "cli rc=$? $(tail -2 gpurun_out/r02_cli.log | cut -c1-600)"
# ncu: launch list of one default step, then full captures of the hot kernels (stage 1 on 2 samples, the merge on the whole step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_run.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extract16|k_partition2|k_bin_count' -c 8 -o gpurun_out/r02_stage1 python bench.py --steps 1 --warmup 0 --samples 2 --no-cpu-baseline > gpurun_out/r02_stage1_run.log 2>&1; echo "ncu stage1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_merge_sm' -c 1 -o gpurun_out/r02_merge python bench.py --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02_merge_run.log 2>&1; echo "ncu merge rc=$?"
ls -la gpurun_out | tail -25
