"""Static guard on the code generation the design depends on (the trace kernel is bound by issue slots, its step
loops sit at the register cap, and what keeps them fast is visible in the SASS). Runs on the CPU with cuobjdump on the in-tree
library -- no GPU -- so a change that quietly costs the kernel its uniform-register constants, spills it, or drops the TMA
staging fails here, before any GPU time is spent."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "blackhole-simulation_b200", "gravitas_b200", "libgravitas_b200.so")
HEADLINE = "_ZN3gvt12k_trace_tileIdLi2ELb1ELb0ELi512ELb1ELb0EEEvNS_11FrameParamsE"     # <double, symplectic, budget, no debug, 512, WGSL rule, not mixed>

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(SO), reason="needs cuobjdump and the built library")


def _run(*args):
    return subprocess.run(["cuobjdump", *args, SO], capture_output=True, text=True, check=True).stdout


def _innermost_loops(sass):
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", sass):
        t = m.group(2).strip()
        body = re.sub(r"^@!?U?P\d+\s+", "", t)
        ins.append((int(m.group(1), 16), body.split()[0].split(".")[0] if body else "?", t))
    loops = []
    for addr, op, t in ins:
        if op != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", t)
        if not m or int(m.group(1), 16) >= addr:
            continue
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= addr]
        inner = sum(1 for x in body[:-1] if x[1] == "BRA" and re.search(r"0x([0-9a-f]+)", x[2]) and tgt <= int(re.search(r"0x([0-9a-f]+)", x[2]).group(1), 16) < x[0])
        if inner == 0:
            loops.append(body)
    return loops


def test_resource_usage_of_every_trace_instantiation():
    usage = _run("-res-usage")
    seen = 0
    for m in re.finditer(r"Function (\S*k_trace_tile\S*):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage):
        name, reg, stack, shared, local = m.group(1), *map(int, m.groups()[1:])
        seen += 1
        assert reg <= 128, (name, reg)                        # 512 threads x 128 registers = the whole register file of an SM
        assert local == 0, (name, local)
    assert seen >= 20                                         # f64 / f32 / mixed x steppers x budget x debug x step rule
    m = re.search(re.escape(HEADLINE) + r":\s*\n\s*REG:(\d+) STACK:(\d+)", usage)
    assert m and int(m.group(1)) <= 120 and int(m.group(2)) == 0


def test_headline_kernel_sass():
    sass = _run("-sass", "-fun", HEADLINE)
    assert "UBLKCP" in sass and "SYNCS" in sass              # TMA bulk copies of the frame block / spectral LUT + mbarrier waits
    assert not re.search(r"\b(LDL|STL)\b", sass)             # no local memory anywhere in the kernel
    loops = [b for b in _innermost_loops(sass) if sum(1 for x in b if x[1] in ("DFMA", "DMUL", "DADD")) > 100]
    assert len(loops) == 5                                    # generic+polar (twice: march and replay), generic, zone 1, zone 3
    for body in loops:
        ops = collections.Counter(x[1] for x in body)
        assert ops["CALL"] == 0 and ops["LDL"] == 0 and ops["STL"] == 0
        # the loops' FP64 constants come from uniform registers, not from per-step constant loads:
        # a handful of LDC / LDCU per step is the rare paths' (renormalisation, disk shading); dozens mean the constants fell out
        assert ops["LDC"] + ops["LDCU"] <= 12, (len(body), ops["LDC"], ops["LDCU"])
        assert sum(1 for x in body if x[1] == "DFMA" and "UR" in x[2]) >= 15
    zone3 = min(loops, key=len)                              # zone 3: rotated trigonometry, no equatorial-crossing test
    assert len(zone3) <= 240, len(zone3)
    fp64 = sum(1 for x in zone3 if x[1] in ("DFMA", "DMUL", "DADD", "DSETP"))
    assert fp64 <= 175, fp64                                  # 131 on the hot path + the Sigma-free renormalisation every tenth step

    def three_reg(t):
        ops_ = re.sub(r"^@!?U?P\d+\s+", "", t).split(None, 1)[1].split(",")[1:]
        regs = {re.sub(r"[-|]|\.reuse", "", o).strip() for o in ops_ if re.match(r"\s*-?\|?R\d+", o)}
        return len(regs) >= 3 and not any("UR" in o or "c[" in o for o in ops_)
    n3 = sum(1 for x in zone3 if x[1] == "DFMA" and three_reg(x[2]))
    assert n3 <= 45, n3                                       # 34 on the hot path (3 issue cycles each instead of 2) + the rare path's
