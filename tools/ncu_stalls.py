#!/usr/bin/env python3
"""Per-source-line stall breakdown of an ncu report: ncu_stalls.py report.ncu-rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; lines = []
for r in rows:
    if len(r) > 6 and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0].strip().isdigit(): lines.append(r)
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
tot = sum(g(r, '# Samples') for r in lines)
for k in ['stall_wait', 'stall_long_sb', 'stall_no_inst', 'stall_short_sb', 'stall_math', 'stall_branch_resolving', 'stall_dispatch']:
    t = sum(g(r, k) for r in lines)
    print('== %s: %.1f%% of samples' % (k, 100 * t / max(tot, 1)))
    for r in sorted(lines, key=lambda r: -g(r, k))[:n]:
        print('  %5.1f%% line %s: %s' % (100 * g(r, k) / max(t, 1), r[0], r[1].strip()[:110]))
