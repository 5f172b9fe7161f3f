#!/bin/bash
# streaming engine: parity vs the other engines, timing, cycle trace, then (argument "full" / "quick") the whole GPU suite + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "streaming_engine_matches" 2>&1 | tail -15 | tee gpurun_out/tc4_pytest.txt
grep -q passed gpurun_out/tc4_pytest.txt || exit 1
timeout 200 python tools/engine_bench.py 5 6 2>&1 | tee gpurun_out/tc4_engine_bench.txt
timeout 200 python tools/tc3_trace.py 6 2>&1 | tee gpurun_out/tc4_trace_engine6.txt
if [ "$1" = "full" ]; then bash tools/gpu_call_full.sh; fi
if [ "$1" = "quick" ]; then bash tools/gpu_call_full.sh quick; fi
