#!/bin/bash
# full gpu suite + default bench (with extras) after a kernel change
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -14 gpurun_out/full_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline'])
print(d['cpu_baseline'])
print(list(d['kernel_ms_per_step'].items())[:12])
print(json.dumps(d.get('extras'))[:1500])
PY
