#!/bin/bash
# round-2 GPU call 1: tcgen05 bring-up probe, experimental encoder check, dynamic-tile A/B, phase decomposition
mkdir -p gpurun_out
nvidia-smi -L
echo "== probe"; timeout 60 tools/microbench/tcgen05_probe > gpurun_out/r02_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02_probe.log
echo "== quick tc"; timeout 400 python tools/gpu_quick_tc.py gpurun_out/r02_tc.json --time > gpurun_out/r02_tc.log 2>&1; echo "tc rc=$?"; tail -25 gpurun_out/r02_tc.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 3"
echo "== A/B"
$B > gpurun_out/r02_static_1m.json 2>gpurun_out/err1.log
$B --iters 1 > gpurun_out/r02_static_1m_it1.json 2>>gpurun_out/err1.log
$B --sites 125000 > gpurun_out/r02_static_125k.json 2>>gpurun_out/err1.log
$B --ragged > gpurun_out/r02_static_ragged.json 2>>gpurun_out/err1.log
export M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_dyn.so
$B > gpurun_out/r02_dyn_1m.json 2>>gpurun_out/err1.log
$B --sites 125000 > gpurun_out/r02_dyn_125k.json 2>>gpurun_out/err1.log
$B --ragged > gpurun_out/r02_dyn_ragged.json 2>>gpurun_out/err1.log
unset M6A_LIB
for f in gpurun_out/r02_static_*.json gpurun_out/r02_dyn_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), "ms", round(d["roofline"]["kernel_ms"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -5 gpurun_out/err1.log
