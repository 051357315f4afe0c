#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3
for P in 1 0; do echo "persistent=$P"; for k in attn_self attn_cross attn_q; do SVOL_ATTN_PERSISTENT=$P timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done; done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
