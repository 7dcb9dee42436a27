#!/bin/bash
# Round 2, call 20 (2 GPUs): the driver's scaling command on the shipped HEAD, and the two-GPU tests
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --out gpurun_out/r02t_c2_2gpu.json > gpurun_out/r02t_c2_2gpu.log 2>&1; echo "bench x2 rc=$?"; tail -1 gpurun_out/r02t_c2_2gpu.log | cut -c1-400
timeout 400 python -m pytest tests/test_zz_nccl_gpu.py -m gpu -q -x > gpurun_out/r02t_pytest2.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02t_pytest2.log)"
