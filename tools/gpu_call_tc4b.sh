#!/bin/bash
# cycle trace + ncu full capture of the second-generation streaming engine
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/tc3_trace.py 6 2>&1 | tee gpurun_out/tc4_trace_engine6.txt
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_nt_tc4_kernel' -s 4 -c 4 \
    -o gpurun_out/r02_tc4_kernels python tools/prof_target.py edgeconv > gpurun_out/r02_tc4_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
