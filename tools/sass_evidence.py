#!/usr/bin/env python
"""Instruction-count evidence per kernel from the built library's SASS (no GPU needed).
  python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gags_b200", "csrc", "libgags_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)),
                       capture_output=True, text=True).stdout.splitlines()
WANT = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMASTG", "UTMALDG", "UBLKRED",
        "SYNCS", "REDG", "RED.", "ATOMG", "LDGMC", "BAR.ARV", "BAR.SYNC", "BAR.RED")
print("# SASS evidence (cuobjdump -sass gags_b200/csrc/libgags_b200.so, sm_100a): instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit -> mbarrier, LDTM = tcgen05.ld (TMEM), UTCATOMSWS = TMEM alloc,")
print("# UBLKCP = cp.async.bulk (1-D TMA), UTMASTG = cp.async.bulk.tensor store (TMA tensor store, the forward's epilogue),")
print("# SYNCS = mbarrier ops, BAR.ARV / BAR.SYNC = named hardware barriers, REDG = red.global.add, LDGMC = multimem.ld_reduce")
blocks = out.split("Function : ")[1:]
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for line in blk.splitlines():
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WANT:
            if op.startswith(w.rstrip(".")) and (w != "RED." or op.startswith("RED.")):
                cnt[w.rstrip(".")] += 1
                break
    if cnt:
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", name)
        short = re.split(r"\((?![^<]*>)", short)[0][:58]     # cut the argument list, keep <template args>
        print(f"{short:60s} " + " ".join(f"{k}={v}" for k, v in sorted(cnt.items())))
