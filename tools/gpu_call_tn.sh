#!/bin/bash
# MN-major BF16x3 weight-gradient engine: parity tests, kernel census of one C2 step, C2 bench (both TN precisions)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "weight_gradient or edge_conv or edgeconv or mask or gemm_tn or all_edge_pairs or gradient" > gpurun_out/tn_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tn_pytest.log
tail -15 gpurun_out/tn_pytest.log
timeout 300 python tools/step_kernels.py > gpurun_out/tn_census_mn.txt 2>&1; head -14 gpurun_out/tn_census_mn.txt
NT_TN_PRECISION=tf32x3 timeout 300 python tools/step_kernels.py > gpurun_out/tn_census_tf32.txt 2>&1; head -8 gpurun_out/tn_census_tf32.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/tn_bench_mn.json 2> gpurun_out/tn_bench_mn.err; tail -c 1500 gpurun_out/tn_bench_mn.json
NT_TN_PRECISION=tf32x3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/tn_bench_tf32.json 2> gpurun_out/tn_bench_tf32.err; tail -c 400 gpurun_out/tn_bench_tf32.json
