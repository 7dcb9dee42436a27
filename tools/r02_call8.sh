#!/bin/bash
# Round 2, call 8 (1 GPU): whole GPU suite (k > 31, record path, SimkaMin kernel with bulk-copy tiles, ...), config 5 variants, profiles of the shipped build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r02h_pytest_gpu.log)"
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --out gpurun_out/r02h_config5.json > gpurun_out/r02h_config5.log 2>&1; echo "config5 rc=$? $(tail -1 gpurun_out/r02h_config5.log | cut -c1-330)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_simkamin_tiled' -c 1 -o gpurun_out/r02h_simkamin python bench.py --config 5 --samples 48 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02h_simkamin_run.log 2>&1; echo "ncu simkamin rc=$?"
SIMKA_B200_NVCC_EXTRA="-DSIMKA_SMT_EPT=8" python -m simka_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline --out gpurun_out/r02h_config5_ept8.json > gpurun_out/r02h_config5_ept8.log 2>&1; echo "config5 ept8 rc=$? $(tail -1 gpurun_out/r02h_config5_ept8.log | cut -c1-330)"
SIMKA_B200_NVCC_EXTRA="-DSIMKA_SMT_EPT=4" python -m simka_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline --out gpurun_out/r02h_config5_ept4.json > gpurun_out/r02h_config5_ept4.log 2>&1; echo "config5 ept4 rc=$? $(tail -1 gpurun_out/r02h_config5_ept4.log | cut -c1-330)"
python -m simka_b200.build --force > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --verify --out gpurun_out/r02h_c2.json > gpurun_out/r02h_c2.log 2>&1; echo "config2 rc=$? $(tail -1 gpurun_out/r02h_c2.log | cut -c1-300)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_launches_run.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extract16|k_partition2|k_bin_count|k_compact' -c 8 -o gpurun_out/r02h_stage1 python bench.py --steps 1 --warmup 0 --e2e-steps 1 --samples 2 --no-cpu-baseline > gpurun_out/r02h_stage1_run.log 2>&1; echo "ncu stage1 rc=$?"
timeout 300 python bench.py --kmer 41 --steps 2 --warmup 1 --no-cpu-baseline --out gpurun_out/r02h_k41.json > gpurun_out/r02h_k41.log 2>&1; echo "k=41 rc=$? $(tail -1 gpurun_out/r02h_k41.log | cut -c1-300)"
