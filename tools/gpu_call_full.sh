#!/bin/bash
# full gpu suite + default bench (with extras) after a kernel change;  "quick" as first argument: C2 line only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -14 gpurun_out/full_pytest.log
EXTRA=""; [ "$1" = "quick" ] && EXTRA="--no-extras"
timeout 600 python bench.py --steps 30 --warmup 5 $EXTRA > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/full_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline'])
print(d['cpu_baseline'])
print(list(d['kernel_ms_per_step'].items()))
print(json.dumps(d.get('extras'))[:1500])
PY
