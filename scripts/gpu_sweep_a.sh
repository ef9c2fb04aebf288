#!/bin/bash
# one gpurun call: A/B of the trace-kernel variants + ncu captures of the f64 and mixed kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=blackhole-simulation_b200/variants_tmp
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
SWEEP_TAG=sweep_main python scripts/kernel_sweep.py 16 0,3,1 4 2>&1 | tee gpurun_out/sweep_main.log
GRAVITAS_B200_LIB=$PWD/$V/libr1.so SWEEP_TAG=sweep_r1 python scripts/kernel_sweep.py 16 0,1 4 2>&1 | tee gpurun_out/sweep_r1.log
GRAVITAS_B200_LIB=$PWD/$V/libcw1.so SWEEP_TAG=sweep_cw1 python scripts/kernel_sweep.py 16 0 4 2>&1 | tee gpurun_out/sweep_cw1.log
for rs in 20 50; do GVT_MIXED_RSWITCH=$rs SWEEP_TAG=sweep_mixed_rs$rs python scripts/kernel_sweep.py 16 3 3 2>&1 | tee gpurun_out/sweep_mixed_rs$rs.log; done
ncu --set full --clock-control none --import-source on -k regex:k_trace_tile -s 1 -c 1 -f -o gpurun_out/prof_r2a_f64 python scripts/profile_frame.py 2 0 2 > gpurun_out/ncu_r2a_f64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_tile -s 1 -c 1 -f -o gpurun_out/prof_r2a_mixed python scripts/profile_frame.py 2 3 2 > gpurun_out/ncu_r2a_mixed.log 2>&1
ls -la gpurun_out/*.ncu-rep
