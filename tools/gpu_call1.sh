#!/bin/bash
# round-2 GPU call 1: full gpu suite (with the new BASELINE-shape parity tests), engine-5 experiment, C2 bench baselines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
NT_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -k "streaming_engine_matches" > gpurun_out/c1_engine5.log 2>&1; echo "rc=$?" >> gpurun_out/c1_engine5.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
NT_TC3_TILES=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_engine5.json 2> gpurun_out/c1_bench_engine5.err
tail -5 gpurun_out/c1_pytest.log; tail -3 gpurun_out/c1_engine5.log
python - <<'PY'
import json
for n in ('default','engine5'):
    try:
        d=json.loads(open('gpurun_out/c1_bench_%s.json'%n).read().strip().splitlines()[-1])
        print(n, d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v for k,v in list(d['kernel_ms_per_step'].items())[:8]})
    except Exception as e: print(n, 'ERR', e)
PY
