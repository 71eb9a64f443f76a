#!/usr/bin/env python
"""Aggregate the warp-state samples of an ncu source page (ncu -i X.ncu-rep --page source --csv) by code region.
Regions are found by marker instructions: DP loop = the SHFL.UP clusters, MMA issue = UTCHMMA, ...
usage: ncu_stalls.py source.csv [--dump-region N]"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]

def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0

tot = sum(f(r, '# Samples') for r in data)
print(f"kernel: {rows[0][1][:90]}  total samples {int(tot)}")

def clusters(marker, gap=80, pad=40):
    hit = [i for i, r in enumerate(data) if marker in r[ix['Source']]]
    out = []
    if not hit: return out
    s = p = hit[0]
    for i in hit[1:]:
        if i - p > gap: out.append((max(0, s - pad), p + pad)); s = i
        p = i
    out.append((max(0, s - pad), p + pad))
    return out

def report(name, a, b):
    reg = data[a:b]
    ns = sum(f(r, '# Samples') for r in reg); ne = sum(f(r, 'Instructions Executed') for r in reg)
    agg = {k: sum(f(r, k) for r in reg) for k in stalls}
    top = ", ".join(f"{k[6:]} {100 * v / max(ns, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7] if v > 0)
    print(f"{name:28s} sass[{a}:{b}] samples {int(ns):7d} ({100 * ns / tot:5.1f}%)  warp-instrs {int(ne):9d}  | {top}")

for n, (a, b) in enumerate(clusters('SHFL.UP')): report(f"DP loop {n}", a, b)
for n, (a, b) in enumerate(clusters('UTCHMMA', pad=20)): report(f"MMA issue {n}", a, b)
for n, (a, b) in enumerate(clusters('LDTM', pad=30)): report(f"epilogue (LDTM) {n}", a, b)
for n, (a, b) in enumerate(clusters('STTM', pad=30)): report(f"prologue (STTM) {n}", a, b)
report("whole kernel", 0, len(data))
if len(sys.argv) > 3 and sys.argv[2] == '--dump':
    a, b = map(int, sys.argv[3].split(':'))
    for i in range(a, b):
        r = data[i]
        ns = f(r, '# Samples')
        top = " ".join(f"{k[6:]}:{int(f(r, k))}" for k in stalls if f(r, k) > 0)
        print(f"{i:5d} {int(ns):5d} {int(f(r, 'Instructions Executed')):8d}  {r[ix['Source']].strip()[:70]:70s} {top}")
