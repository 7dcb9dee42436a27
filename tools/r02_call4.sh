#!/bin/bash
# Round 2, call 4 (1 GPU): count-kernel occupancy variants of the record path, merge sub-pass split test, config 4
mkdir -p gpurun_out
run() {  # tag, extra args, env...
    tag=$1; shift; extra=$1; shift
    (env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $extra) > gpurun_out/r02d_$tag.log 2>&1
    tail -1 gpurun_out/r02d_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('$tag', 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d.get('verify',{}).get('result'))
except Exception as e:
    print('$tag', 'FAILED', e)
"
}
variant() {  # tag, nvcc extra, env...
    tag=$1; shift; nv=$1; shift
    SIMKA_B200_NVCC_EXTRA="$nv" python -m simka_b200.build --force > gpurun_out/r02d_build_$tag.log 2>&1 || echo "build $tag failed"
    run $tag "" SIMKA_B200_SUPERK=2 "$@"
}
variant s2048_m4_t1024 "-DSIMKA_SK2_SLOTS=2048 -DSIMKA_SK2_MINB=4" SIMKA_B200_SK2_TARGET=1024
variant s2048_m5_t1024 "-DSIMKA_SK2_SLOTS=2048 -DSIMKA_SK2_MINB=5" SIMKA_B200_SK2_TARGET=1024
variant s2048_m4_t768 "-DSIMKA_SK2_SLOTS=2048 -DSIMKA_SK2_MINB=4" SIMKA_B200_SK2_TARGET=768
variant s2048_m5_t768_th128 "-DSIMKA_SK2_SLOTS=2048 -DSIMKA_SK2_MINB=10 -DSIMKA_SK2_THREADS=128 -DSIMKA_SK2_RCAP=128" SIMKA_B200_SK2_TARGET=768
python -m simka_b200.build --force > gpurun_out/r02d_build_restore.log 2>&1
timeout 600 python -m pytest tests/test_zzz_full_size_gpu.py -m gpu -q -x -k "split or 512" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02d_pytest.log)"
timeout 1500 python bench.py --config 4 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --out gpurun_out/r02_config4.json > gpurun_out/r02d_config4.log 2>&1; echo "config4 rc=$? $(tail -1 gpurun_out/r02d_config4.log | cut -c1-1500)"
ls -la gpurun_out | grep r02d
