#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NT_CENSUS_EACH=gemm timeout 300 python tools/step_kernels.py 2>&1 | tee gpurun_out/census.txt | tail -75
timeout 600 python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
print(d['roofline']['group'], d['roofline']['frac'], d['roofline']['launch_ms'])
print(list(d['kernel_ms_per_step'].items()))
PY
