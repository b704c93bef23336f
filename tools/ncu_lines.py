#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: share of executed warp instructions, of stall
samples, and shared-memory wavefronts (needs -lineinfo builds and --import-source on).
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; lines = []
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0].strip().isdigit():
        lines.append(r)
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
ti = sum(g(r, "Instructions Executed") for r in lines) or 1; ts = sum(g(r, "# Samples") for r in lines) or 1
print("total warp-instructions %.4g, stall samples %d" % (ti, ts))
for r in lines:
    pi, ps = 100 * g(r, "Instructions Executed") / ti, 100 * g(r, "# Samples") / ts
    if pi >= minpct or ps >= minpct:
        print("%4s inst %5.1f%% samples %5.1f%% thr/inst %4.1f smem-wf %.3g/%.3g | %s" % (r[0], pi, ps, g(r, "Avg. Threads Executed"),
              g(r, "L1 Wavefronts Shared"), g(r, "L1 Wavefronts Shared Ideal"), r[1].strip()[:105]))
