#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep regex:blend_fwd_tc [min_share]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, recs = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) == len(hdr):
        d = {h: v for h, v in zip(hdr, r) if h != "Source"}
        if d["Instructions Executed"].isdigit():
            recs.append((fname, int(r[0]), r[1], d))
tot_i = sum(int(d["Instructions Executed"]) for *_ , d in recs)
tot_s = sum(int(d["# Samples"]) for *_ , d in recs)
print(f"total warp-instr {tot_i}, samples {tot_s}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for f, ln, src, d in recs:
    i, s = int(d["Instructions Executed"]), int(d["# Samples"])
    if i > tot_i * minshare or s > tot_s * minshare:
        top = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        tops = " ".join(f"{n}:{v}" for v, n in top if v)
        print(f"{f[:16]:16s}{ln:5d} inst {100*i/tot_i:5.1f}% samp {100*s/tot_s:5.1f}%  [{tops}]  {src.strip()[:90]}")
