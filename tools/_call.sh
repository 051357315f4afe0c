cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider "$@" > "$O/$name.log" 2>&1; echo "$name: exit $? :: $(tail -1 $O/$name.log)"; }
run model 600 tests/test_model_gpu.py -m gpu
grep -E "Error|FAILED|assert" $O/model.log | head
b() { timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>$O/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['host_enqueue_ms_per_step'])"; }
b resident; b resident
SVOL_BENCH_NO_RESIDENT_BUFFERS=1 b copied; SVOL_BENCH_NO_RESIDENT_BUFFERS=1 b copied
tail -3 $O/bench.err
