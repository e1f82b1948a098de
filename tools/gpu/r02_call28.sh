#!/bin/bash
# stop hand-over fix: short jobs (quick tool + the new pytest sweep), CLI timing, then timing
timeout 600 python tools/gpu_quick_tc2.py gpurun_out/r02_tc2.json --only-tc --time --only-big > gpurun_out/r02_tc2.log 2>&1; echo "tc2 rc=$?"
grep -E "KERNEL|stuck|Error|error" gpurun_out/r02_tc2.log | cut -c1-250 | head -8
grep -cE '^ok' gpurun_out/r02_tc2.log; grep -E '^FAIL' gpurun_out/r02_tc2.log | cut -c1-200
grep '"encoder"' gpurun_out/r02_tc2.log | cut -c1-110
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "short_jobs or edge_cases or sharding" 2>&1 | tail -4
python tools/gpu_cli_timing.py 100000 50 /tmp/m6a_cli > gpurun_out/r02_cli_timing.json 2> gpurun_out/r02_cli_timing.err; echo "cli rc=$?"; cat gpurun_out/r02_cli_timing.json | cut -c1-900; tail -3 gpurun_out/r02_cli_timing.err
