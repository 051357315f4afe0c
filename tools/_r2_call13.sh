#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c13_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r2c13_breakdown.txt > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c13_bench.json')); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['e2e_bf16_features']['value'], d['roofline']['ms_per_launch'])"
cat gpurun_out/r2c13_breakdown.txt
timeout 600 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r2c13_bench_train.json 2> gpurun_out/r2c13_bench_train.err; python -c "
import json; d=json.load(open('gpurun_out/r2c13_bench_train.json')); print('train', {k:d[k] for k in ('value','ms_per_step','final_loss')})"
touch svol_b200/csrc/gemm_tc.cu; SVOL_EXTRA_NVCC_FLAGS=-DSVOL_GEMM_TRACE bash svol_b200/csrc/build.sh > /dev/null 2>&1; python tools/gemm_trace.py gemm_sa_out; python tools/gemm_trace.py gemm_qk
