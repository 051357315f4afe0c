#!/usr/bin/env bash
# One `ncu --set full` capture of EVERY kernel of one eager forward + criterion step (C2, B=32); the raw metrics are
# exported as CSV into gpurun_out/ (the .ncu-rep itself is too large to bring back).   bash tools/ncu_step.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-step}
mkdir -p gpurun_out
python tools/fwd_once.py > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --nvtx --nvtx-include "profiled/" -f -o /tmp/${tag} python tools/fwd_once.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${tag}_ncu.log
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2> /dev/null
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}_raw.csv
python tools/ncu_table.py gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_table.txt 2>&1; cat gpurun_out/${tag}_table.txt
