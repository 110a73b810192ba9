#!/usr/bin/env python
"""Hot SASS instructions in address order with their source line (needs -lineinfo).
  python tools/ncu_hot_sass.py rep regex:kernel [min_share]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, cur = None, None, None
sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0].isdigit(): cur = (fname, int(r[0])); continue
    if r[0] == "" and r[2].startswith("0x"):
        d = {h: v for h, v in zip(hdr, r)}
        stalls = sorted(((int(v or 0), h[6:]) for h, v in d.items() if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:2]
        sass.append((int(r[2], 16), cur, int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), r[3].strip(), stalls))
sass.sort()
tot = sum(s[2] for s in sass)
base = sass[0][0]
for a, cur, smp, ins, txt, st in sass:
    if smp > tot * minshare:
        print(f"{a-base:6x} {cur[0][:14]:14s}:{cur[1]:4d} samp {100*smp/tot:5.1f}% inst {ins:9d} [{' '.join(f'{n}:{v}' for v,n in st if v)}] {txt[:70]}")
