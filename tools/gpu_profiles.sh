#!/bin/bash
# round-2 final profile pass: launch list of a training step + full captures of the kernels DESIGN.md discusses
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# (1) every launch of 3 eager C2 training steps with its device time and DRAM bytes (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/r02_launches.csv python tools/prof_target.py model 3 > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
# (2) full captures: the streaming row GEMMs (second-generation engine) and the MN weight-gradient GEMM, from an EdgeConv layer fwd + bwd
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_nt_tc4_kernel|gemm_tn_mn_kernel' -s 11 -c 9 \
    -o gpurun_out/r02_final_gemms python tools/prof_target.py edgeconv > gpurun_out/r02_final_gemms.log 2>&1
echo "gemm capture rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
