#!/bin/bash
# Round 2, call 5 (2 GPUs): NCCL exchange inside the library -- parity (tools/nccl_check.py) and the N=2 bench, against the round-1 torch-level exchange
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py > gpurun_out/r02e_nccl_check.log 2>&1; echo "nccl_check rc=$? $(grep -c 'native exchange ok' gpurun_out/r02e_nccl_check.log) ok lines; $(tail -2 gpurun_out/r02e_nccl_check.log | cut -c1-300)"
for mode in native torch; do
  SIMKA_BENCH_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --out gpurun_out/r02e_bench2_$mode.json > gpurun_out/r02e_bench2_$mode.log 2>&1
  echo "bench N=2 $mode rc=$? $(tail -1 gpurun_out/r02e_bench2_$mode.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['notes'].get('exchange_rank0_ms_per_step'))
except Exception as e: print('FAILED', e)
")"
done
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-budget-s 60 > gpurun_out/r02e_refarm.log 2>&1; echo "refarm rc=$? $(tail -1 gpurun_out/r02e_refarm.log | cut -c1-200)"
