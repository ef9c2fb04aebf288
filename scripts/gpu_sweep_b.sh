#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=blackhole-simulation_b200/variants_tmp
for L in "$@"; do
  GRAVITAS_B200_LIB=$PWD/$V/lib$L.so SWEEP_TAG=sweep_$L python scripts/kernel_sweep.py 16 ${MODES:-0,3,1} 4 2>&1 | tee gpurun_out/sweep_$L.log
done
