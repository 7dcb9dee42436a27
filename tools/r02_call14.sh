#!/bin/bash
# Round 2, call 14: whole GPU suite on the shipped HEAD, default bench line, configuration 5 (pinned sketches), N = 128 with the shared pair routine
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02n_pytest.log)"
(timeout 600 python bench.py --out gpurun_out/r02n_c2.json) > gpurun_out/r02n_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02n_c2.log | cut -c1-1500
(timeout 600 python bench.py --config 5 --out gpurun_out/r02n_config5.json) > gpurun_out/r02n_config5.log 2>&1; tail -1 gpurun_out/r02n_config5.log | cut -c1-700
(SIMKA_B200_TILE_TIMING=1 timeout 900 python bench.py --no-cpu-baseline --samples 128 --reads 1000000 --steps 2 --warmup 1 --e2e-steps 1 --out gpurun_out/r02n_n128.json) > gpurun_out/r02n_n128.log 2>&1
grep -h "tile merge" gpurun_out/r02n_n128.log | tail -1; tail -1 gpurun_out/r02n_n128.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('n128 ms/step', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
