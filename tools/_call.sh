cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
b() { timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['host_enqueue_ms_per_step'])"; }
b new; b new
cp svol_b200/csrc/libsvol_b200.so /tmp/new.so; cp svol_b200/csrc/libsvol_b200_prev.so svol_b200/csrc/libsvol_b200.so
b prev; b prev
cp /tmp/new.so svol_b200/csrc/libsvol_b200.so
b new
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown $O/breakdown.txt --no-cpu-baseline > /dev/null 2>&1; head -12 $O/breakdown.txt
