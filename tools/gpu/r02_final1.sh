#!/bin/bash
# final single-GPU record of round 2: GPU test suite, bench lines (configs 2-5, FFMA encoder, ragged, reference arm), evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; echo "bench cfg3 rc=$?"
for c in 2 4 5; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg$c.json 2> gpurun_out/r02_bench_cfg$c.err; echo "bench cfg$c rc=$?"; done
timeout 600 python bench.py --encoder ffma --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_ffma.json 2>/dev/null
timeout 600 python bench.py --ragged --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_ragged.json 2>/dev/null
timeout 600 python bench.py --ragged --encoder ffma --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_ragged_ffma.json 2>/dev/null
timeout 600 python bench.py --iters 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_cfg3_iters1.json 2>/dev/null
python - <<'PY'
import json
for f in ['cfg3','cfg2','cfg4','cfg5','cfg3_ffma','ragged','ragged_ffma','cfg3_iters1','reference_arm']:
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3), 'M sites/s', round(d['ms_per_step'],3), 'ms', (d.get('parity') or {}).get('ok'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'), (d.get('result_digest') or {}).get('site_prob_mod_count_sha256','')[:12])
    except Exception as e: print(f, 'ERR', e)
PY
bash tools/gpu/r02_evidence.sh
