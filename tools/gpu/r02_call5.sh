#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_quick_tc2.py gpurun_out/r02_tc2.json --time --big > gpurun_out/r02_tc2.log 2>&1; echo "tc2 rc=$?"
grep -E "FAIL|KERNEL|Error|error" gpurun_out/r02_tc2.log | cut -c1-250 | head -5
grep '"encoder": "tc"' gpurun_out/r02_tc2.log
export M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_prof.so
python tools/gpu_tc_profile.py 200000 1 > gpurun_out/r02_prof_it1.log 2>&1; cat gpurun_out/r02_prof_it1.log
python tools/gpu_tc_profile.py 200000 1000 > gpurun_out/r02_prof_it1000.log 2>&1; cat gpurun_out/r02_prof_it1000.log
