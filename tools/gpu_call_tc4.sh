#!/bin/bash
# first GPU run of the second-generation streaming engine: parity vs the other engines, timing, cycle traces
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x -k "streaming_engine_matches" 2>&1 | tail -15 | tee gpurun_out/tc4_pytest.txt
timeout 200 python tools/engine_bench.py 0 6 2>&1 | tee gpurun_out/tc4_engine_bench.txt
timeout 200 python tools/tc3_trace.py 0 2>&1 | tee gpurun_out/tc4_trace_engine0.txt
timeout 200 python tools/tc3_trace.py 6 2>&1 | tee gpurun_out/tc4_trace_engine6.txt
