#!/bin/bash
# A/B of the stage-1 front ends on one B200 (run under gpurun from the repo root):
#   tools/ab_superk.sh            parity tests of the super-k-mer path, then bench lines for default / superk (flat) / superk (thread)
# Writes gpurun_out/ab_*.log; prints one summary line per variant.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_superk_gpu.py -m gpu -x -q > gpurun_out/ab_superk_tests.log 2>&1; echo "superk tests rc=$? $(tail -1 gpurun_out/ab_superk_tests.log)"
run() {  # tag, env...
    tag=$1; shift
    (env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/ab_$tag.log 2>&1
    tail -1 gpurun_out/ab_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['config'].get('stage1'), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['counts'])
except Exception as e:
    print('$tag', 'FAILED', e)
"
}
run default SIMKA_B200_SUPERK=0
run superk_flat SIMKA_B200_SUPERK=1
run superk_thread SIMKA_B200_SUPERK=1 SIMKA_B200_SUPERK_EXPAND=thread
run default_k2minb3 SIMKA_B200_SUPERK=0 SIMKA_B200_K2_MINB=3   # K2 scatter compiled for 3 CTAs / SM (80 registers, no spill) against 4 (64 registers, 40 B spill)
# tools/ab_superk.sh ncu : additionally one full ncu capture of the super-k-mer kernels and of k_bin_count behind them
# (2 samples are enough: every sample launches each kernel once)
if [ "$1" = "ncu" ]; then
    SIMKA_B200_SUPERK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_superk|k_bin_count' -c 8 \
        -o gpurun_out/ab_superk_full python bench.py --steps 1 --warmup 0 --samples 2 --no-cpu-baseline > gpurun_out/ab_superk_ncu.log 2>&1
    echo "ncu rc=$?"
fi
