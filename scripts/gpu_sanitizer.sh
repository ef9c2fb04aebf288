#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck) over the GPU test-suite: every test except the bench subprocess and the
# multi-GPU launcher (the tools follow one process). Logs go to gpurun_out/ and are copied to profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='not bench and not multi'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_$tool.log
  tail -5 gpurun_out/sanitizer_$tool.log
done
