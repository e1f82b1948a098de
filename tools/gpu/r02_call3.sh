#!/bin/bash
mkdir -p gpurun_out
export M6A_LIB=$PWD/m6anet_b200/libm6anet_b200_prof.so
python tools/gpu_tc_profile.py 200000 1 > gpurun_out/r02_prof_it1.log 2>&1; cat gpurun_out/r02_prof_it1.log
python tools/gpu_tc_profile.py 200000 1000 > gpurun_out/r02_prof_it1000.log 2>&1; cat gpurun_out/r02_prof_it1000.log
