#!/usr/bin/env python
"""Static SASS mix of a kernel's hot loop (no GPU needed): cuobjdump -sass, find the backward branch whose body holds
the most FP64/FP32 arithmetic, histogram its opcodes. Usage: sass_loop_mix.py <lib.so|.o> <kernel-name-regex> [unroll]
A quick check before spending GPU time; the dynamic mix of record comes from ncu (scripts/summarize_ncu.py)."""
import collections
import re
import subprocess
import sys


def main():
    so, pat = sys.argv[1], re.compile(sys.argv[2])
    unroll = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        if not pat.search(dem):
            continue
        ins = []   # (addr, opcode, text)
        for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f):
            t = m.group(2).strip()
            body = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = body.split()[0] if body else "?"
            ins.append((int(m.group(1), 16), op.split(".")[0], t))
        best = None
        for i, (addr, op, t) in enumerate(ins):
            if op == "BRA":
                m = re.search(r"0x([0-9a-f]+)", t)
                if m and int(m.group(1), 16) < addr:
                    tgt = int(m.group(1), 16)
                    body = [x for x in ins if tgt <= x[0] <= addr]
                    score = sum(1 for x in body if x[1] in ("DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD"))
                    inner = sum(1 for x in body[:-1] if x[1] == "BRA" and re.search(r"0x([0-9a-f]+)", x[2]) and
                                tgt <= int(re.search(r"0x([0-9a-f]+)", x[2]).group(1), 16) < x[0])
                    if inner == 0 and (best is None or score > best[0]):   # innermost loop with the most arithmetic
                        best = (score, tgt, addr, body, inner)
        print(f"== {dem}")
        if not best:
            print("   no loop found"); continue
        _, tgt, addr, body, inner = best
        h = collections.Counter(x[1] for x in body)
        fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        fp32 = sum(h[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
        print(f"   loop 0x{tgt:x}..0x{addr:x}: {len(body)} instr ({len(body)/unroll:.1f}/iter at unroll {unroll:g}), "
              f"FP64-pipe {fp64/unroll:.1f}/iter, FP32 arith {fp32/unroll:.1f}/iter, nested backward branches {inner}")
        print("   " + "  ".join(f"{k} {v/unroll:.1f}" for k, v in h.most_common(24)))


if __name__ == "__main__":
    main()
