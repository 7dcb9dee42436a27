#!/bin/bash
# usage: tools/bench_short.sh <tag> [extra bench args]   (run under gpurun)
tag=$1; shift
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@") > gpurun_out/bench_$tag.log 2>&1; echo rc=$?
tail -1 gpurun_out/bench_$tag.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'G kmers/s', round(d['value']/1e9,2))
print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
"
