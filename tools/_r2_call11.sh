#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2c11_pytest.log; grep "d/d src_video" gpurun_out/r2c11_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r2c11_bench_train.json 2> gpurun_out/r2c11_bench_train.err; echo "train bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c11_bench_train.json')); print({k:d[k] for k in ('value','ms_per_step','final_loss')}, d['e2e']['value'], d['roofline']['ms_per_launch'])"; tail -3 gpurun_out/r2c11_bench_train.err
