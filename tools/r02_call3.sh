#!/bin/bash
# Round 2, call 3: record path after the convergent-insert rewrite: parity, timing, build variants, ncu; config 4 with the exact-size arena
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_records_gpu.py -m gpu -q -x > gpurun_out/r02c_pytest.log 2>&1; echo "pytest records rc=$? $(tail -1 gpurun_out/r02c_pytest.log)"
run() {  # tag, extra args, env...
    tag=$1; shift; extra=$1; shift
    (env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $extra) > gpurun_out/r02c_$tag.log 2>&1
    tail -1 gpurun_out/r02c_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d.get('verify',{}).get('result'))
except Exception as e:
    print('$tag', 'FAILED', e)
"
}
run rec_default "--verify" SIMKA_B200_SUPERK=2
SIMKA_B200_SUPERK=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sk2_count' -c 2 -o gpurun_out/r02c_count python bench.py --steps 1 --warmup 0 --e2e-steps 1 --samples 2 --no-cpu-baseline > gpurun_out/r02c_count_run.log 2>&1; echo "ncu count rc=$?"
variant() {  # tag, nvcc extra, env...
    tag=$1; shift; nv=$1; shift
    SIMKA_B200_NVCC_EXTRA="$nv" python -m simka_b200.build --force > gpurun_out/r02c_build_$tag.log 2>&1 || echo "build $tag failed"
    run $tag "" SIMKA_B200_SUPERK=2 "$@"
}
variant rcap512 "-DSIMKA_SK2_RCAP=512"
variant s2048_t1024 "-DSIMKA_SK2_SLOTS=2048" SIMKA_B200_SK2_TARGET=1024
variant s2048_t1024_th128 "-DSIMKA_SK2_SLOTS=2048 -DSIMKA_SK2_THREADS=128 -DSIMKA_SK2_RCAP=256" SIMKA_B200_SK2_TARGET=1024
variant s8192_t4096_rcap512 "-DSIMKA_SK2_SLOTS=8192 -DSIMKA_SK2_RCAP=512" SIMKA_B200_SK2_TARGET=4096
python -m simka_b200.build --force > gpurun_out/r02c_build_restore.log 2>&1
timeout 1200 python bench.py --config 4 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --out gpurun_out/r02_config4.json > gpurun_out/r02c_config4.log 2>&1; echo "config4 rc=$? $(tail -1 gpurun_out/r02c_config4.log | cut -c1-900)"
ls -la gpurun_out | grep r02c
