#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "weight_gradient or edge_conv or edgeconv or mask or gemm_tn or gradient" 2>&1 | tail -3
NT_CENSUS_EACH=gemm_tn timeout 300 python tools/step_kernels.py 2>&1 | sed -n '/every launch/,$p' | tee gpurun_out/tn_each_mn.txt
