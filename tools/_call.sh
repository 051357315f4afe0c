cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
SVOL_FFN_2SM=0 timeout 120 python tools/ffn_trace.py > $O/ftrace_1sm_noload.txt 2>&1; echo "1sm $?"
SVOL_FFN_2SM=1 timeout 120 python tools/ffn_trace.py > $O/ftrace_2sm_noload.txt 2>&1; echo "2sm $?"
for m in 0 1; do for k in ffn_video; do echo -n "NOLOAD 2SM=$m "; SVOL_FFN_2SM=$m timeout 60 python tools/run_kernel.py $k 20 2>&1 | tail -1; done; done
