#!/bin/bash
# Round 2, call 2: the super-k-mer RECORD path of stage 1 (SIMKA_B200_SUPERK=2): parity, timing against the k-mer path, bin-size sweep, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_records_gpu.py -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest records rc=$? $(tail -1 gpurun_out/r02b_pytest.log)"
run() {  # tag, extra args, env...
    tag=$1; shift; extra=$1; shift
    (env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $extra) > gpurun_out/r02b_$tag.log 2>&1
    tail -1 gpurun_out/r02b_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['notes'].get('stage1'), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d.get('verify',{}).get('result'))
except Exception as e:
    print('$tag', 'FAILED', e)
"
}
run records_verify "--verify" SIMKA_B200_SUPERK=2
run records_t1024 "" SIMKA_B200_SUPERK=2 SIMKA_B200_SK2_TARGET=1024
run records_t1536 "" SIMKA_B200_SUPERK=2 SIMKA_B200_SK2_TARGET=1536
run records_t3000 "" SIMKA_B200_SUPERK=2 SIMKA_B200_SK2_TARGET=3000
run kmer_path "" SIMKA_B200_SUPERK=0
SIMKA_B200_SUPERK=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_launches_run.log 2>&1; echo "ncu launches rc=$?"
SIMKA_B200_SUPERK=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_superk_scan|k_sk2_scatter|k_sk2_count|k_compact|k_refine' -c 14 -o gpurun_out/r02b_stage1 python bench.py --steps 1 --warmup 0 --e2e-steps 1 --samples 2 --no-cpu-baseline > gpurun_out/r02b_stage1_run.log 2>&1; echo "ncu stage1 rc=$?"
ls -la gpurun_out | grep r02b
