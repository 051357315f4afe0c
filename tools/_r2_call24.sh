#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 300 python tools/attn_p_check.py 20 2>&1 | tail -10
for P in 1 0; do echo "persistent=$P"; for k in attn_self attn_cross attn_q; do SVOL_ATTN_PERSISTENT=$P timeout 120 python tools/run_kernel.py $k 20 2>&1 | tail -1; done; SVOL_CONFIG=C4 SVOL_ATTN_PERSISTENT=$P timeout 120 python tools/run_kernel.py attn_self 10 2>&1 | tail -1; done
SVOL_ATTN_PERSISTENT=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r2c24_breakdown.txt > gpurun_out/r2c24_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2c24_bench.json')); print({k:d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step')}, d['e2e']['value'], d['e2e_bf16_features']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; head -8 gpurun_out/r2c24_breakdown.txt
