cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "n8 exit $?"; cut -c1-260 gpurun_out/scale_n8.json
