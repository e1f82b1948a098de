#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_cfg3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_cfg3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['kernel_ms']); print(d['parity']); print(d['result_digest']); print(d['e2e']); print(d['cpu_baseline'])
PY
