#!/bin/bash
# Round 2, call 13 (8 GPUs): the tile merge under the in-library exchange -- parity at 8 ranks (70 samples included), configuration 3
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py > gpurun_out/r02m_nccl_check8.log 2>&1; echo "nccl_check x8 rc=$? $(grep -c 'native exchange ok' gpurun_out/r02m_nccl_check8.log) ok lines"
b() {  # tag, args, env...
  tag=$1; shift; args=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 $args --out gpurun_out/r02m_$tag.json > gpurun_out/r02m_$tag.log 2>&1
  echo "$tag rc=$? $(tail -1 gpurun_out/r02m_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,1), 'G/s e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['notes'].get('exchange_rank0_ms_per_step'), 'gen_s', d['notes']['input_build_s'])
except Exception as e: print('FAILED', e)
")"
}
b config3 "--config 3 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline" SIMKA_B200_TILE_TIMING=1
grep -h "tile merge" gpurun_out/r02m_config3.log | tail -2
