#!/usr/bin/env python
"""Attribute stall samples of inlined helper lines (e.g. the mbarrier wait loop) to their call
sites: walks the SASS in address order and reports, per contiguous run of `file:line`, the samples
and the nearest preceding/following other source lines.
  python tools/ncu_waits.py rep regex:kernel umma.cuh 24"""
import csv, io, subprocess, sys
rep, kern, tfile, tline = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
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
    if r[0].isdigit(): cur = (fname, int(r[0]), r[1].strip()); continue
    if r[0] == "" and r[2].startswith("0x"):
        iS = hdr.index("# Samples")
        sass.append((int(r[2], 16), cur, int(r[iS] or 0), r[3].strip()))
sass.sort()
tot = sum(s[2] for s in sass)
i = 0
while i < len(sass):
    if sass[i][1][0] == tfile and sass[i][1][1] == tline:
        j = i; samp = 0
        while j < len(sass) and sass[j][1][0] == tfile and sass[j][1][1] in (tline, tline + 1, tline - 1):
            samp += sass[j][2]; j += 1
        prev = next((sass[k][1] for k in range(i - 1, -1, -1) if sass[k][1][0] != tfile), None)
        nxt = next((sass[k][1] for k in range(j, len(sass)) if sass[k][1][0] != tfile), None)
        if samp > tot * 0.003:
            print(f"{100*samp/tot:5.1f}%  after {prev[0]}:{prev[1]} `{prev[2][:60]}`  before {nxt[0] if nxt else ''}:{nxt[1] if nxt else ''} `{(nxt[2] if nxt else '')[:50]}`")
        i = j
    else:
        i += 1
