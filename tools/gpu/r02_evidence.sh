#!/bin/bash
# Evidence on the shipped build: launch list of the default bench command, ncu --set full of the three kernels, validate() and
# CLI timings.  Everything lands in gpurun_out/ (summaries are copied to profiles/ afterwards).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-digest > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mil_infer_tc -s 3 -c 1 -o gpurun_out/r02_final_tc -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-digest --no-e2e > /dev/null 2>&1; echo "ncu tc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mil_infer_kernel -s 3 -c 1 -o gpurun_out/r02_final_ffma -f \
    python bench.py --encoder ffma --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-digest --no-e2e > /dev/null 2>&1; echo "ncu ffma rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mil_infer_kernel -s 1 -c 1 -o gpurun_out/r02_final_bags -f \
    python tools/gpu_validate_timing.py 100000 50 5 > /dev/null 2>&1; echo "ncu bags rc=$?"
python tools/gpu_validate_timing.py 100000 50 5 > gpurun_out/r02_validate_timing.json 2> gpurun_out/r02_validate_timing.err; echo "validate rc=$?"; cat gpurun_out/r02_validate_timing.json | cut -c1-600
python tools/gpu_cli_timing.py 100000 50 /tmp/m6a_cli > gpurun_out/r02_cli_timing.json 2> gpurun_out/r02_cli_timing.err; echo "cli rc=$?"; cat gpurun_out/r02_cli_timing.json | cut -c1-900; tail -3 gpurun_out/r02_cli_timing.err
ls -la gpurun_out/*.ncu-rep
