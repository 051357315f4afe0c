cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
run() { local name=$1 t=$2; shift 2; timeout "$t" python -m pytest -q --tb=short -p no:cacheprovider "$@" > "$O/$name.log" 2>&1; echo "$name: exit $? :: $(tail -1 $O/$name.log)"; }
run model 600 tests/test_model_gpu.py -m gpu
grep -E "Error|FAILED|assert" $O/model.log | head -5
b() { timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>$O/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'])"; tail -2 $O/bench.err; }
b stagger; b stagger; b stagger
SVOL_B200_STAGGER=0 b free; SVOL_B200_STAGGER=0 b free; SVOL_B200_STAGGER=0 b free
