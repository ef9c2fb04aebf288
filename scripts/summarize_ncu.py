"""Turn an .ncu-rep (one k_trace_tile launch, --set full) into the text summary committed under profiles/."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

import json
rep, out, steps_per_launch = sys.argv[1], sys.argv[2], float(sys.argv[3])
json_out = sys.argv[4] if len(sys.argv) > 4 else None      # machine-readable export (bench.py reads roofline.traffic from it)
build_info = sys.argv[5] if len(sys.argv) > 5 else ""      # gvt_build_info() of the library the capture ran on
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
if json_out:
    def num(k):
        try:
            return float(M[k][0].replace(",", ""))
        except (KeyError, ValueError):
            return None
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    def nbytes(k):
        v = num(k)
        return None if v is None else v * unit_scale.get(M[k][1], 1.0)
    mix = {k: v / ws for k, v in c.most_common(40)}
    exported = {
        "kernel": M["Kernel Name"][0].strip(), "build_info": build_info, "source_report": rep.split("/")[-1],
        "grid": int(num("launch__grid_size")), "block": int(num("launch__block_size")), "registers": int(num("launch__registers_per_thread")),
        "steps_per_launch": steps_per_launch, "time_ms_under_ncu": num("gpu__time_duration.sum"),
        "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
        "fp64_pipe_active_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "fma_pipe_active_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "alu_pipe_active_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "xu_pipe_active_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions_per_step": tot / ws, "sass_mix_per_step": mix,
        "fp64_pipe_instructions_per_step": fp64,
        # executed flops per geodesic step: FMA = 2, MUL = ADD = 1 (compares and MUFU not counted)
        "executed_fp64_flop_per_step": 2 * mix.get("DFMA", 0.0) + mix.get("DMUL", 0.0) + mix.get("DADD", 0.0),
        "executed_fp32_flop_per_step": 2 * mix.get("FFMA", 0.0) + mix.get("FMUL", 0.0) + mix.get("FADD", 0.0),
    }
    exported["dram_bytes_per_launch"] = (exported["dram_bytes_read"] or 0) + (exported["dram_bytes_write"] or 0)
    json.dump(exported, open(json_out, "w"), indent=1)
print("\n".join(lines[:60]))
