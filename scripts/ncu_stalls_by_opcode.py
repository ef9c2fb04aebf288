"""Aggregate an ncu source page (ncu -i x.ncu-rep --page source --csv --print-source sass) by opcode: executed warp
instructions, samples, and the stall reasons attributed to the instruction a warp was waiting to issue."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and not h.endswith("(Not Issued)")]
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P[T\d]+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    a = agg[op]
    a["inst"] += int(r[ix["Instructions Executed"]]); a["samples"] += int(r[ix["# Samples"]])
    for s in stalls:
        v = int(r[ix[s]] or 0); a[s] += v; tot[s] += v
    tot["samples"] += int(r[ix["# Samples"]]); tot["inst"] += int(r[ix["Instructions Executed"]])
print("total samples", tot["samples"], "inst", tot["inst"])
print("stall totals:", {s.replace("stall_", ""): tot[s] for s in stalls if tot[s]})
print(f"{'op':10s} {'inst%':>7s} {'samples%':>9s}  top stall reasons (share of this opcode's samples)")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:18]:
    top = sorted(((a[s], s) for s in stalls if a[s]), reverse=True)[:4]
    print(f"{op:10s} {100*a['inst']/tot['inst']:7.2f} {100*a['samples']/tot['samples']:9.2f}  " +
          ", ".join(f"{s.replace('stall_','')} {100*v/max(a['samples'],1):.0f}%" for v, s in top))
