#!/bin/bash
# Run under gpurun: ncu launch list of the bench command + full captures of the hot kernels.  Outputs in gpurun_out/.
tag=${1:-r01}
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_launch_run.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_bin_count|k_extract|k_partition2|k_compact" -s 10 -c 5 \
    -o gpurun_out/${tag}_stage1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_stage1_run.log 2>&1; echo "stage1 rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_merge" -c 1 \
    -o gpurun_out/${tag}_merge python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_merge_run.log 2>&1; echo "merge rc=$?"
