#!/bin/bash
# Round 2, call 6 (8 GPUs): scale point of configuration 2 with the in-library NCCL exchange, NCCL p2p channel A/B, configuration 3 (128 x 5M reads)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; free -g | head -2
b() {  # tag, args, env...
  tag=$1; shift; args=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 $args --out gpurun_out/r02f_$tag.json > gpurun_out/r02f_$tag.log 2>&1
  echo "$tag rc=$? $(tail -1 gpurun_out/r02f_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']/1e9,1), 'G/s e2e', round(d['e2e']['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['notes'].get('exchange_rank0_ms_per_step'), 'gen_s', d['notes']['input_build_s'])
except Exception as e: print('FAILED', e)
")"
}
b c2_native "--steps 10 --warmup 3"
b c2_native_ch32 "--steps 10 --warmup 3" NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
b c2_torch "--steps 10 --warmup 3" SIMKA_BENCH_EXCHANGE=torch
b config3 "--config 3 --steps 2 --warmup 1 --e2e-steps 1"
tail -5 gpurun_out/r02f_config3.log | cut -c1-400
