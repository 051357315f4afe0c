#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke exit $? :: $(tail -2 gpurun_out/r2c1_smoke.log)"
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown gpurun_out/r2c1_breakdown.txt > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench exit $?"; cat gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/r2c1_bench_refgpu.json 2> gpurun_out/r2c1_bench_refgpu.err; echo "refgpu exit $?"; cat gpurun_out/r2c1_bench_refgpu.json; tail -3 gpurun_out/r2c1_bench_refgpu.err
timeout 300 python tools/bench_matcher.py > gpurun_out/r2c1_matcher.txt 2>&1; echo "matcher exit $?"; cat gpurun_out/r2c1_matcher.txt
