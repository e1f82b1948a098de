#!/bin/bash
# packed FMUL2 products for the two chains of a pooling warp + explicitly un-contracted accumulation in every pooling path
tools/microbench/mc_packed_check; echo "unit rc=$?"
timeout 600 python tools/gpu_quick_tc2.py gpurun_out/r02_tc2.json --time --only-big > gpurun_out/r02_tc2.log 2>&1; echo "tc2 rc=$?"
grep -E "KERNEL|stuck|Error|error" gpurun_out/r02_tc2.log | cut -c1-250 | head -8
grep -cE '^ok' gpurun_out/r02_tc2.log; grep -E '^FAIL' gpurun_out/r02_tc2.log | cut -c1-300
grep '"encoder"' gpurun_out/r02_tc2.log | cut -c1-110
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
