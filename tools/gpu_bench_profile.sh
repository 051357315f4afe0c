#!/usr/bin/env bash
# smoke + bench (both arms) + per-kernel breakdown + ncu launch list, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $? :: $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/breakdown.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat gpurun_out/breakdown.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
if [[ "${1:-}" == "ncu" ]]; then
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/ncu.log
fi
