#!/usr/bin/env bash
# End-of-round measurement pass: full GPU test suite (one process, as the driver runs it), smoke, both bench modes with
# their reference arms, kernel breakdowns, ncu launch lists and one --set full capture of the attention backward.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $? :: $(tail -1 $O/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $? :: $(tail -2 $O/smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --steps 100 --warmup 5 --breakdown $O/breakdown.txt > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-200 $O/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?"
timeout 600 python bench.py --mode train --steps 30 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err; echo "train exit $?"; cut -c1-200 $O/bench_train.json
timeout 600 python bench.py --mode train --impl reference --steps 2 --ref-batch 2 > $O/bench_train_ref.json 2> $O/bench_train_ref.err; echo "train ref exit $?"; cut -c1-160 $O/bench_train_ref.json
timeout 300 python tools/bench_train.py --steps 10 --breakdown > $O/train_breakdown.txt 2>&1; grep "train step" $O/train_breakdown.txt
timeout 300 python tools/bench_next_rows.py > $O/next_rows.txt 2>&1; cat $O/next_rows.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu fwd $?"
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/train_launches.csv python bench.py --mode train --steps 2 --warmup 3 > $O/ncu_train.log 2>&1; echo "ncu train $?"
python tools/summarize_launches.py $O/train_launches.csv > $O/train_launches_summary.txt
for k in attn_self attn_cross attn_q ffn_video ffn_query gate_fused gate_split; do timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done > $O/kernels_alone.txt; cat $O/kernels_alone.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o $O/attn_self python tools/run_kernel.py attn_self 1 > $O/ncu_attn.log 2>&1; echo "ncu attn_self $?"
python tools/ncu_summary.py $O/attn_self.ncu-rep > $O/attn_self_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_tc -c 1 -o $O/ffn_video python tools/run_kernel.py ffn_video 1 > $O/ncu_ffn.log 2>&1; echo "ncu ffn_video $?"
python tools/ncu_summary.py $O/ffn_video.ncu-rep > $O/ffn_video_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gate_fused -c 1 -o $O/gate_fused python tools/run_kernel.py gate_fused 1 > $O/ncu_gate.log 2>&1; echo "ncu gate_fused $?"
python tools/ncu_summary.py $O/gate_fused.ncu-rep > $O/gate_fused_summary.txt
