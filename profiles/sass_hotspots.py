#!/usr/bin/env python
"""Aggregate ncu's SASS-level warp-stall samples of one kernel by opcode and
list the hottest instructions.  usage: sass_hotspots.py rep kernel-regex [launch-skip]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
s_i, e_i = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[s_i]), int(r[e_i]), r[1].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
byop = collections.Counter()
cnt = collections.Counter()
for s, e, txt in data:
    t = txt.split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    op = ".".join(op.split(".")[:2])
    byop[op] += s
    cnt[op] += e
print(rows[0][1][:120] if rows[0] else "")
print(f"total samples {tot}, instructions executed {sum(d[1] for d in data)}")
for op, s in byop.most_common(14):
    print(f"  {100*s/tot:5.1f}%  {op:16s} executed {cnt[op]}")
print("hottest instructions:")
for i, (s, e, txt) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:16]
for i in order:
    s, e, txt = data[i]
    print(f"  {100*s/tot:5.1f}%  #{i:<5d} exec {e:>9d}  {txt[:100]}")
