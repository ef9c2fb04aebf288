#!/bin/bash
# ncu captures of the trace kernel in its three precision modes, of the TAA resolve, and the launch list of bench.py.
# gpurun brings back at most 64 MiB of gpurun_out/ per call, so the captures are split over two calls: `... a` and `... b`.
# The .ncu-rep files come back in gpurun_out/; scripts/summarize_ncu.py turns them into profiles/r02_*.txt / .json here.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name precision
  ncu --set full --clock-control none --import-source on -k regex:k_trace_tile -s 1 -c 1 -f -o gpurun_out/prof_r02_$1 python scripts/profile_frame.py 2 $2 2 > gpurun_out/ncu_r02_$1.log 2>&1
  tail -1 gpurun_out/ncu_r02_$1.log
}
if [ "$1" = "a" ]; then cap f64 0; cap mixed 3; fi
if [ "$1" = "b" ]; then
  cap f32 1
  ncu --set full --clock-control none --import-source on -k regex:k_taa_resolve -s 2 -c 1 -f -o gpurun_out/prof_r02_taa python scripts/profile_taa.py > gpurun_out/ncu_r02_taa.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_r02.log 2>&1
fi
ls -la gpurun_out/
