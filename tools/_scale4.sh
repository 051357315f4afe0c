cd /root/repo; mkdir -p gpurun_out
for n in 2 4; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "n$n exit $?"; cut -c1-160 gpurun_out/scale_n$n.json; done
