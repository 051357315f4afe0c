#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matcher_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c5_pytest.log
python tools/bench_matcher.py 2>&1 | grep -E "^C|criterion call|fused   |solve only   " | tee gpurun_out/r2c5_matcher.txt
for k in attn_self attn_cross attn_q; do python tools/run_kernel.py $k 20 2>&1 | tail -1; done | tee gpurun_out/r2c5_attn.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r2c5_breakdown.txt > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c5_bench.json')); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['e2e_bf16_features']['value'], d['roofline']['ms_per_launch'])"
cat gpurun_out/r2c5_breakdown.txt
