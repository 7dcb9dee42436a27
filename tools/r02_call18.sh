#!/bin/bash
# Round 2, call 18: the `simka` command line against the reference `simka` on the same FASTA files (first byte read -> matrices written),
# every output matrix compared byte for byte
mkdir -p gpurun_out
timeout 500 python tools/bench_cli.py --samples 32 --reads 1000000 --with-reference > gpurun_out/r02r_cli_1m.json 2> gpurun_out/r02r_cli_1m.err; echo "rc=$?"; cat gpurun_out/r02r_cli_1m.json | cut -c1-1500; tail -3 gpurun_out/r02r_cli_1m.err
