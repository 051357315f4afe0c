#!/usr/bin/env bash
# Runs the training-step GPU tests in isolated processes (a device-side trap poisons the CUDA context of the
# process that hit it), each under its own timeout, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, timeout, pytest args...
  local name=$1 t=$2; shift 2
  timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider -x "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name: exit $? :: $(tail -1 gpurun_out/$name.log)"
}
run tr_rowwise 300 tests/test_train_gpu.py -m gpu -k "layernorm or gelu or transpose or colsum or wgrad_through or heads or ln_linear or batch_sum"
run tr_splitk  300 tests/test_train_gpu.py -m gpu -k "split_k"
run tr_mnmajor 300 tests/test_train_gpu.py -m gpu -k "mn_major"
run tr_gate    200 tests/test_train_gpu.py -m gpu -k "gate"
run tr_attn_a  200 tests/test_train_gpu.py -m gpu -k "attention_backward and 128"
run tr_attn_b  300 tests/test_train_gpu.py -m gpu -k "attention_backward and not 128"
run tr_dropout 300 tests/test_train_gpu.py -m gpu -s -k "dropout"
run tr_head    600 tests/test_train_gpu.py -m gpu -s -k "head_backward_vs"
run tr_step    600 tests/test_train_gpu.py -m gpu -s -k "training_step or overwritten"
grep -h "worst per-tensor" gpurun_out/tr_head.log
