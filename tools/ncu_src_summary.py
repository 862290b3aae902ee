"""Summarise an `ncu --page source --csv` export: instruction mix (executed), stall reasons, by opcode class."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); stall = collections.Counter()
tot_exec = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit(): continue
    src = r[ix["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else "?"
    base = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDL", "STL", "LDG", "STG", "LDS", "STS")) else op.split(".")[0]
    ex = int(r[ix["Instructions Executed"]] or 0)
    ops[base] += ex; tot_exec += ex
    samples[base] += int(r[ix["# Samples"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = r[ix[h]]
            if v: stall[h] += int(v)
print("total executed warp-instructions:", tot_exec)
for k, v in ops.most_common(22):
    print(f"  {k:18s} {v:14d} {100*v/tot_exec:6.2f}%   samples {100*samples[k]/max(1,sum(samples.values())):6.2f}%")
ts = sum(stall.values())
print("stall reasons (all samples):")
for k, v in stall.most_common(10):
    print(f"  {k:24s} {100*v/ts:6.2f}%")
