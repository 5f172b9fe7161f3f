#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/_build/ts_mode_test 2>&1 | tee gpurun_out/ts_mode_test.txt
timeout 200 python tools/engine_bench.py 0 6 7 2>&1 | grep bnrelu | tee gpurun_out/tc4_engine7_bench.txt
