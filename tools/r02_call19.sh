#!/bin/bash
# Round 2, call 19: ncu --set full of the two tile-merge kernels in the shape of configuration 4 (512 samples, k = 21; 150 k reads per sample so
# that the capture fits next to ncu's save / restore buffers)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_merge_tiles|k_group_bins' -c 2 -o gpurun_out/r02s_tiles512 python bench.py --samples 512 --reads 150000 --kmer 21 --steps 1 --warmup 0 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02s_tiles512_run.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02s_tiles512_run.log | cut -c1-300
