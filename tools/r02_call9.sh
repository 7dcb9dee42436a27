#!/bin/bash
# Round 2, call 9 (1 GPU): k_merge_sm with 32-bit shared-memory words (A/B against 21.65 ms), parity, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "merge or many_samples or example or nasty or full_config2 or flags" > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02i_pytest.log)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --out gpurun_out/r02i_c2.json > gpurun_out/r02i_c2.log 2>&1; echo "config2 rc=$? $(tail -1 gpurun_out/r02i_c2.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})")"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_merge_sm' -c 1 -o gpurun_out/r02i_merge python bench.py --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02i_merge_run.log 2>&1; echo "ncu merge rc=$?"
