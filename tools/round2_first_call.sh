#!/bin/bash
# Everything built after round 1's GPU budget ran out, measured in ONE gpurun call (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# 1. the whole -m gpu suite (the files test_zz_* hold what has never run on a B200: super-k-mer front end, DSK known answers,
#    sketch pre-filter);  2. A/B of the stage-1 front ends (tools/ab_superk.sh);  3. the sketch stage with / without its pre-filter.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r02_pytest_gpu.log)"
bash tools/ab_superk.sh
timeout 300 python tools/bench_sketch.py > gpurun_out/r02_sketch.log 2>&1; echo "sketch rc=$? $(tail -1 gpurun_out/r02_sketch.log | cut -c1-600)"
timeout 300 python tools/bench_sketch.py --filter > gpurun_out/r02_sketch_filter.log 2>&1; echo "sketch -filter rc=$? $(tail -1 gpurun_out/r02_sketch_filter.log | cut -c1-600)"
# 4. the `simka` command line end to end (first byte read -> matrices written; its -verbose line has the three timestamps:
#    GPU context ready, all samples parsed + packed, stage 1 done) -- round 1 measured 3.4 s for 32 x 1 M reads before the ingest changes
timeout 600 python tools/bench_cli.py --samples 32 --reads 1000000 > gpurun_out/r02_cli.log 2>&1; echo "cli rc=$? $(tail -2 gpurun_out/r02_cli.log | cut -c1-700)"
