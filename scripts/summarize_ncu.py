"""Turn an .ncu-rep (one k_trace_tile launch, --set full) into the text summary committed under profiles/."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep, out, steps_per_launch = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct"]
lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on, 1 launch)", ""]
for k in keys:
    if k in M:
        lines.append(f"{k:86s} {M[k][0]:>22s} {M[k][1]}")
stall = {h: float(M[h][0]) for h in M if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
lines += ["", "## warp stall reasons (average warps stalled per issue-active cycle)"]
for h, v in sorted(stall.items(), key=lambda kv: -kv[1])[:9]:
    lines.append(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
c = Counter()
tot = 0
for r in data:
    m = re.match(r"(@!?U?P[T\d]+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
    n = int(r[ix["Instructions Executed"]])
    tot += n
    c[m.group(2) if m else "?"] += n
ws = steps_per_launch / 32.0
lines += ["", f"## dynamic SASS mix: warp-instructions per geodesic step (launch = {steps_per_launch:.0f} steps; total {tot / ws:.1f}/step)"]
for k, v in c.most_common(22):
    lines.append(f"{k:10s} {v / ws:8.2f}")
fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")) / ws
fp32 = sum(c[k] for k in ("FFMA", "FMUL", "FADD", "FSETP", "FMNMX")) / ws
lines += ["", f"FP64-pipe instructions/step: {fp64:.1f}   FP32 arithmetic instructions/step: {fp32:.1f}"]
tma = [r[ix["Source"]].strip() for r in data if re.search(r"UBLKCP|UTMALDG|SYNCS", r[ix["Source"]])]
lines += ["", "## TMA / mbarrier SASS present in the kernel:"] + sorted(set(tma))[:12]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
