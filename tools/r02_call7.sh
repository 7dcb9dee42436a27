#!/bin/bash
# Round 2, call 7 (2 GPUs): overlapped round-based exchange: parity, 2-GPU CLI, N=2 bench; SimkaMin kernel with bulk-copy tiles (config 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_nccl_gpu.py -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1; echo "pytest nccl rc=$? $(tail -1 gpurun_out/r02g_pytest.log)"
SIMKA_BENCH_EXCHANGE=native timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --out gpurun_out/r02g_bench2.json > gpurun_out/r02g_bench2.log 2>&1
echo "bench N=2 rc=$? $(tail -1 gpurun_out/r02g_bench2.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['notes'].get('exchange_rank0_ms_per_step'))
except Exception as e: print('FAILED', e)
")"
timeout 600 python -m pytest tests -m gpu -q -x -k "simkamin" > gpurun_out/r02g_pytest_simkamin.log 2>&1; echo "pytest simkamin rc=$? $(tail -1 gpurun_out/r02g_pytest_simkamin.log)"
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --out gpurun_out/r02g_config5_ept6.json > gpurun_out/r02g_config5_ept6.log 2>&1; echo "config5 ept6 rc=$? $(tail -1 gpurun_out/r02g_config5_ept6.log | cut -c1-330)"
SIMKA_B200_NVCC_EXTRA="-DSIMKA_SMT_EPT=4" python -m simka_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline --out gpurun_out/r02g_config5_ept4.json > gpurun_out/r02g_config5_ept4.log 2>&1; echo "config5 ept4 rc=$? $(tail -1 gpurun_out/r02g_config5_ept4.log | cut -c1-330)"
SIMKA_B200_NVCC_EXTRA="-DSIMKA_SMT_EPT=3" python -m simka_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline --out gpurun_out/r02g_config5_ept3.json > gpurun_out/r02g_config5_ept3.log 2>&1; echo "config5 ept3 rc=$? $(tail -1 gpurun_out/r02g_config5_ept3.log | cut -c1-330)"
