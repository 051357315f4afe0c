#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu mufu.cu && /tmp/mufu) > gpurun_out/r2_ubench_mufu.txt 2>&1; grep -E "warps/SM (4|16)" gpurun_out/r2_ubench_mufu.txt
for cfgb in "C4 8" "C2n4 32"; do set -- $cfgb; layers=2; [ "$1" = "C2n4" ] && layers=4
  timeout 600 python bench.py --config $1 --batch $2 --layers $layers --steps 20 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r2_breakdown_$1.txt > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$1.json')); print('$1', {k:d[k] for k in ('value','ms_per_step','tflops_algorithmic')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; head -8 gpurun_out/r2_breakdown_$1.txt
done
