cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $? :: $(tail -1 $O/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $? :: $(tail -2 $O/smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --steps 100 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-200 $O/bench.json
SVOL_CONFIG=C4 timeout 120 python tools/run_kernel.py attn_self 10 2>&1 | tail -1
SVOL_CONFIG=C4 timeout 120 python tools/run_kernel.py attn_cross 10 2>&1 | tail -1
timeout 200 python tools/bench_matcher.py 2>&1 | tail -6
