#!/usr/bin/env python
"""Per-source-line summary of an ncu report (instructions executed + stall samples).
usage: tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
csv.field_size_limit(10**9)
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        def gi(k):
            try: return int(d.get(k) or 0)
            except ValueError: return 0
        agg.append((cur.split('/')[-1], int(r[0]), gi('# Samples'), gi('Instructions Executed'), gi('stall_no_inst'), r[1][:80]))
tot = sum(a[2] for a in agg) or 1; toti = sum(a[3] for a in agg) or 1
print("samples", tot, "instructions", toti)
byf = collections.defaultdict(lambda: [0, 0])
for a in agg: byf[a[0]][0] += a[2]; byf[a[0]][1] += a[3]
for f, v in byf.items(): print(f, "samples %.1f%%" % (100 * v[0] / tot), "instr %.1f%%" % (100 * v[1] / toti))
print("--- by samples")
for a in sorted(agg, key=lambda a: -a[2])[:top]:
    print("%s:%d  smp %.1f%%  ins %.1f%%  noinst %.1f%%  %s" % (a[0], a[1], 100 * a[2] / tot, 100 * a[3] / toti, 100 * a[4] / tot, a[5]))
if len(sys.argv) > 3:
    print("--- by line of", sys.argv[3])
    ntile = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    for a in sorted([a for a in agg if a[0] == sys.argv[3]], key=lambda a: a[1]):
        if a[3] or a[2]:
            print("%4d  ins/tile %7.1f  smp %.1f%%  %s" % (a[1], a[3] / ntile, 100 * a[2] / tot, a[5]))
