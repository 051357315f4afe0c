#!/usr/bin/env bash
# bench.py at N GPUs of this box (both modes).   bash tools/scale_bench.sh N tag
cd "$(dirname "$0")/.."
N=${1:-8}; tag=${2:-r2}
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${tag}_scale_n$N.json 2> gpurun_out/${tag}_scale_n$N.err; echo "fwd_match N=$N exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${tag}_scale_n$N.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','host_enqueue_ms_per_step')}, 'e2e', d['e2e']['value'], 'e2e_bf16', d['e2e_bf16_features']['value'], d['clocks'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 20 --warmup 5 > gpurun_out/${tag}_train_n$N.json 2> gpurun_out/${tag}_train_n$N.err; echo "train N=$N exit $?"
python -c "
import json; d=json.load(open('gpurun_out/${tag}_train_n$N.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','final_loss')}, 'e2e', d['e2e']['value'])"

