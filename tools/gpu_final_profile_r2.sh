#!/usr/bin/env bash
# Round-2 end-of-round measurement pass (one GPU): full GPU test suite, smoke, the bench with all three arms (ours, the
# reference's CPU path, the reference's eager path on this GPU), training bench, matcher split, ncu launch list of the
# bench command and one `--set full` capture of every kernel of a step.   bash tools/gpu_final_profile_r2.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r02_final}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/ -x -q -m gpu > $O/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $? :: $(tail -1 $O/${tag}_pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${tag}_smoke.log 2>&1; echo "smoke exit $? :: $(tail -2 $O/${tag}_smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_reference.json 2> $O/${tag}_bench_reference.err; echo "ref exit $?"; cut -c1-160 $O/${tag}_bench_reference.json
timeout 600 python bench.py --breakdown $O/${tag}_breakdown.txt > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench exit $?"; cut -c1-220 $O/${tag}_bench.json
timeout 600 python bench.py --impl reference-gpu --steps 5 --warmup 3 > $O/${tag}_bench_reference_gpu.json 2> $O/${tag}_bench_reference_gpu.err; echo "ref-gpu exit $?"; cut -c1-160 $O/${tag}_bench_reference_gpu.json
timeout 600 python bench.py --mode train --steps 30 --warmup 3 > $O/${tag}_bench_train.json 2> $O/${tag}_bench_train.err; echo "train exit $?"; cut -c1-200 $O/${tag}_bench_train.json
timeout 300 python tools/bench_matcher.py > $O/${tag}_matcher_split.txt 2>&1; grep -E "^C|criterion call" $O/${tag}_matcher_split.txt
for k in attn_self attn_cross attn_q ffn_video ffn_query gate_fused; do timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done > $O/${tag}_kernels_alone.txt; cat $O/${tag}_kernels_alone.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${tag}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${tag}_ncu.log 2>&1; echo "ncu launch list $?"
python tools/summarize_launches.py $O/${tag}_ncu_launches.csv > $O/${tag}_ncu_launches_summary.txt; head -30 $O/${tag}_ncu_launches_summary.txt
bash tools/ncu_step.sh ${tag}_step > /dev/null 2>&1; echo "ncu step $?"; cut -c1-200 $O/${tag}_step_table.txt | head -50
