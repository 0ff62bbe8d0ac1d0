#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source sass --csv` export by device function (split at RET):
executed warp instructions and stall samples per function."""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
segs, cur = [], []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith("0x"):
        continue
    cur.append(r)
    if "RET." in r[1] or " EXIT" in r[1]:
        if "RET." in r[1]:
            segs.append(cur)
            cur = []
if cur:
    segs.append(cur)
ci = {n: i for i, n in enumerate(hdr)}
ix, isamp = ci["Instructions Executed"], ci["# Samples"]
tot_i = sum(int(r[ix] or 0) for s in segs for r in s)
tot_s = sum(int(r[isamp] or 0) for s in segs for r in s)
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for k, s in enumerate(segs):
    ni = sum(int(r[ix] or 0) for r in s)
    ns = sum(int(r[isamp] or 0) for r in s)
    txt = " ".join(r[1] for r in s)
    tags = []
    for key, name in (("0xd2511f53", "philox"), ("DFMA", "fp64"), ("MUFU.RSQ ", "rsq32"), ("BAR.SYNC", "bar"), ("MUFU.SIN", "sin"),
                      ("MUFU.LG2", "log"), ("STL", "spill"), ("DSETP", "dsetp")):
        c = txt.lower().count(key.lower()) if key.startswith("0x") else txt.count(key)
        if c:
            tags.append("%s:%d" % (name, c))
    print("seg %2d  %5d instrs  exec %5.1f%%  samples %5.1f%%  %s" % (k, len(s), 100.0 * ni / max(tot_i, 1), 100.0 * ns / max(tot_s, 1), " ".join(tags)))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    top = sorted(segs[k], key=lambda r: -int(r[isamp] or 0))[: int(sys.argv[3]) if len(sys.argv) > 3 else 25]
    for r in top:
        print(r[0][-6:], r[isamp], r[ix], r[1].strip()[:90])
