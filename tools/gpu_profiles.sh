#!/bin/bash
# round-2 profile pass: launch list of a training step + full captures of the kernels DESIGN.md discusses
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# (1) every launch of 3 eager C2 training steps with its device time and DRAM bytes (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/r02_launches.csv python tools/prof_target.py model 3 > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"
# (2) full captures: LSTM kernels, the dominant row GEMM, the weight-gradient GEMM (second step of the run)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:lstm_fwd_kernel|lstm_bwd_kernel|lstm_dw_kernel|gemm_nt_tc3_kernel<3|gemm_tn_tc_kernel' \
    -s 12 -c 9 -o gpurun_out/r02_train_kernels python tools/prof_target.py model 2 > gpurun_out/r02_full_train.log 2>&1
echo "train capture rc=$?"
# (3) full capture of the fused inference EdgeConv kernel
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edgeconv_eval_kernel' -s 2 -c 2 \
    -o gpurun_out/r02_edgeconv_eval python tools/prof_target.py infer > gpurun_out/r02_full_infer.log 2>&1
echo "infer capture rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
# (4) cycle trace of the LSTM kernels
timeout 120 python tools/lstm_trace.py > gpurun_out/r02_lstm_trace.txt 2>&1; tail -8 gpurun_out/r02_lstm_trace.txt
