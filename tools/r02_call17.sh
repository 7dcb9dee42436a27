#!/bin/bash
# Round 2, call 17: whole GPU suite on the shipped HEAD, the default bench line, configuration 5 (device buffers kept between calls, pinned
# sketches), the reference arm with the two-point sample sizing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02q_pytest.log)"
(timeout 600 python bench.py --out gpurun_out/r02q_c2.json) > gpurun_out/r02q_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02q_c2.log | cut -c1-1200
(SIMKA_BENCH_VERBOSE=1 timeout 600 python bench.py --config 5 --out gpurun_out/r02q_config5.json) > gpurun_out/r02q_config5.log 2>&1; grep "^step" gpurun_out/r02q_config5.log | tr '\n' ';'; echo; tail -1 gpurun_out/r02q_config5.log | cut -c1-900
(timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 150 --out gpurun_out/r02q_refarm.json) > gpurun_out/r02q_refarm.log 2>&1; tail -1 gpurun_out/r02q_refarm.log | cut -c1-1500
