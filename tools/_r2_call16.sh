#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for S in 1 0; do
  touch svol_b200/csrc/attn_tc.cu; SVOL_EXTRA_NVCC_FLAGS=-DSVOL_ATTN_SPEC_MAX=$S bash svol_b200/csrc/build.sh > /dev/null 2>&1
  echo "== SPEC_MAX=$S: $(grep -A2 'attention_tc_kernelILb0' svol_b200/csrc/build/attn_tc.log | grep spill)"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -1
  for k in attn_self attn_cross attn_q; do python tools/run_kernel.py $k 20 2>&1 | tail -1; done
done
