#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
for sh in interleaved blocks; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --shard $sh > gpurun_out/bench_${N}gpu_r2_$sh.json 2> gpurun_out/bench_${N}gpu_r2_$sh.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${N}gpu_r2_$sh.json").read().strip().splitlines()[-1])
x=d["extra"]
print("$sh N=$N: value %.4e ms/frame %.4f e2e %.4f trace_kernel %.4f gather %.4f | natural %.4f ms | allgather %s | mixed %.4f ms" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], x["trace_kernel_ms"], x["all_gather_ms"], x["natural_termination"]["ms_per_frame"], x["allgather"], x["mixed_precision"]["ms_per_frame"]))
PY
done
