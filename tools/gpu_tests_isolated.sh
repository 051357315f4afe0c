#!/usr/bin/env bash
# Runs the GPU test-suite in isolated processes (a device-side trap poisons the CUDA context of the
# process that hit it), each under its own timeout, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout, pytest args...
  local name=$1 t=$2; shift 2
  timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name: exit $? :: $(tail -1 gpurun_out/$name.log)"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run matcher   300 tests/test_matcher_gpu.py -m gpu
run rowwise   300 tests/test_kernels_gpu.py -m gpu -k "not gemm and not attention"
run gemm_plain 300 tests/test_kernels_gpu.py -m gpu -k "gemm and plain"
run attn_plain 300 tests/test_kernels_gpu.py -m gpu -k "attention and plain"
run gemm_tc   300 tests/test_kernels_gpu.py -m gpu -k "gemm and tcgen05"
run attn_tc   300 tests/test_kernels_gpu.py -m gpu -k "attention and tcgen05"
run model     600 tests/test_model_gpu.py -m gpu -s
