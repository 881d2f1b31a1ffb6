#!/usr/bin/env python
"""Aggregate the warp-stall samples of an `ncu --page source --csv --print-source sass` export per SASS instruction.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv; tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
k = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(k[:-1], k[1:]):
    hdr, body = rows[a + 1], [r for r in rows[a + 2:b] if len(r) > 5]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    print("==", rows[a][1], "instructions", len(body), "samples", tot)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
    print("  ", [(s, round(100 * v / max(tot, 1), 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
    ops = {}
    for r in body:
        op = r[ci["Source"]].split()
        op = op[1] if op[0].startswith("@") else op[0]
        ops[op] = ops.get(op, 0) + int(r[ci["Instructions Executed"]])
    tot_i = sum(ops.values())
    print("  opcode mix:", [(o, round(100 * v / tot_i, 1)) for o, v in sorted(ops.items(), key=lambda x: -x[1])[:12]])
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][ci["# Samples"]]))[:top_n]:
        st = sorted([(int(r[ci[s]]), s) for s in stalls], reverse=True)[:2]
        print(f"  {idx:4d} {int(r[ci['# Samples']]):6d} {r[ci['Source']].strip()[:64]:64s} {st}")
