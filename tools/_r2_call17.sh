#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r2c17_breakdown.txt > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c17_bench.json')); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['e2e_bf16_features']['value'])"
grep -E "sa_out|ta_out|ca_out|TOTAL|sa_qkv|in_proj" gpurun_out/r2c17_breakdown.txt
touch svol_b200/csrc/gemm_tc.cu; SVOL_EXTRA_NVCC_FLAGS=-DSVOL_GEMM_TRACE bash svol_b200/csrc/build.sh > /dev/null 2>&1; python tools/gemm_trace.py gemm_sa_out | head -8
