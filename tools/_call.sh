cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider "$@" > "$O/$name.log" 2>&1; echo "$name: exit $? :: $(tail -1 $O/$name.log)"; }
run attn_tc 300 tests/test_kernels_gpu.py -m gpu -k "attention and tcgen05"
run model 600 tests/test_model_gpu.py -m gpu
for k in attn_cross attn_q; do timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done
b() { timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['host_enqueue_ms_per_step'])"; }
b new; b new
cp svol_b200/csrc/libsvol_b200.so /tmp/new.so; cp svol_b200/csrc/libsvol_b200_prev.so svol_b200/csrc/libsvol_b200.so
b prev; b prev
cp /tmp/new.so svol_b200/csrc/libsvol_b200.so
b new
