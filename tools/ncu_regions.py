#!/usr/bin/env python
"""Stall / instruction breakdown of an ncu report by source-line regions of one file.
usage: tools/ncu_regions.py report.ncu-rep file name:lo:hi [name:lo:hi ...]"""
import csv, subprocess, sys
csv.field_size_limit(10**9)
rep, fname = sys.argv[1], sys.argv[2]
regions = [(a.split(':')[0], int(a.split(':')[1]), int(a.split(':')[2])) for a in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit(): agg.append((cur, int(r[0]), dict(zip(hdr, r))))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def gi(d, k):
    try: return int(d.get(k) or 0)
    except ValueError: return 0
tot = sum(gi(d, '# Samples') for f, l, d in agg); toti = sum(gi(d, 'Instructions Executed') for f, l, d in agg)
print("| region | samples | instructions | top stall reasons (share of all samples) |\n|---|---:|---:|---|")
def line(name, sel):
    s = sum(gi(d, '# Samples') for d in sel); ins = sum(gi(d, 'Instructions Executed') for d in sel)
    st = sorted(((k[6:], sum(gi(d, k) for d in sel)) for k in stalls), key=lambda x: -x[1])[:4]
    print(f"| {name} | {100*s/tot:.1f}% | {100*ins/toti:.1f}% | " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in st) + " |")
for name, a, b in regions: line(name, [d for f, l, d in agg if f == fname and a <= l < b])
line("other files (spatial algebra, sincos, shuffles)", [d for f, l, d in agg if f != fname])
