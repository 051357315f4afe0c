cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit $? :: $(tail -1 $O/pytest_gpu.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $? :: $(tail -2 $O/smoke.log | tr '\n' ' ')"
timeout 600 python bench.py --steps 100 --warmup 5 --breakdown $O/breakdown.txt > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-200 $O/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?"
timeout 600 python bench.py --mode train --steps 30 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err; echo "train exit $?"; cut -c1-200 $O/bench_train.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu fwd $?"
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt
for k in attn_self attn_cross attn_q ffn_video ffn_query gate_fused; do timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done > $O/kernels_alone.txt; cat $O/kernels_alone.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -o $O/attn_cross python tools/run_kernel.py attn_cross 1 > $O/ncu_attn.log 2>&1; echo "ncu attn_cross $?"
python tools/ncu_summary.py $O/attn_cross.ncu-rep > $O/attn_cross_summary.txt
