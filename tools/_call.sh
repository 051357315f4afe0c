cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider "$@" > "$O/$name.log" 2>&1; echo "$name: exit $? :: $(tail -1 $O/$name.log)"; }
run attn_tc 300 tests/test_kernels_gpu.py -m gpu -k "attention and tcgen05"
run model 600 tests/test_model_gpu.py -m gpu
for k in attn_self attn_cross attn_q; do timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cut -c1-220 $O/bench.json
