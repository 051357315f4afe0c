cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
SVOL_GATE_CL=4 timeout 120 python tools/run_kernel.py gate_fused 50 2>&1 | tail -1
SVOL_GATE_CL=8 timeout 120 python tools/run_kernel.py gate_fused 50 2>&1 | tail -1
timeout 120 python tools/run_kernel.py gate_split 50 2>&1 | tail -1
SVOL_GATE_CL=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gate_fused -c 1 -o $O/gate_fused python tools/run_kernel.py gate_fused 1 > $O/ncu_gate.log 2>&1; echo "ncu $?"
