cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
SVOL_FFN_MULTICAST=0 timeout 120 python tools/ffn_trace.py > $O/ftrace_tw.txt 2>&1; echo "trace $?"
