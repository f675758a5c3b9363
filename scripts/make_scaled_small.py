#!/usr/bin/env python
"""BASELINE config 5: the 14-relation `small` schema scaled by a factor (rows AND key domains),
written as contest relation files, plus the small.work queries with their constants scaled.

    python scripts/make_scaled_small.py <factor> <outdir>

Every column is resampled from the empirical distribution of the shipped column: value * factor + uniform
jitter in [0, factor); column 0 (the sorted unique key of each relation) stays sorted and unique.  Needs the
shipped relations under oracle/_ref/small (data, not source).  Deterministic (seeded).
"""
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
factor, out = int(sys.argv[1]), Path(sys.argv[2])
out.mkdir(parents=True, exist_ok=True)
small = ROOT / "oracle" / "_ref" / "small"
rng = np.random.default_rng(2018)
names = []
for i in range(14):
    raw = np.fromfile(small / f"r{i}", dtype=np.uint64)
    n, c = int(raw[0]), int(raw[1])
    cols = [raw[2 + j * n: 2 + (j + 1) * n] for j in range(c)]
    big = []
    for j, col in enumerate(cols):
        if j == 0:
            # sorted unique key column: every shipped key becomes `factor` consecutive keys
            v = (np.repeat(col, factor) * np.uint64(factor) + np.tile(np.arange(factor, dtype=np.uint64), n))
        else:
            pick = rng.integers(0, n, n * factor)
            v = col[pick] * np.uint64(factor) + rng.integers(0, factor, n * factor, dtype=np.uint64)
        big.append(v)
    with open(out / f"r{i}", "wb") as f:
        np.array([n * factor, c], np.uint64).tofile(f)
        for v in big:
            v.tofile(f)
    names.append(f"r{i}")
(out / "scaled.init").write_text("\n".join(names) + "\n")
work = (small / "small.work").read_text().splitlines()
scaled = []
for line in work:
    if "|" not in line:
        scaled.append(line)
        continue
    rels, preds, views = line.split("|")
    # filter constants scale with the key domains; keep them below 2^31 (structs.h:146)
    preds = re.sub(r"([<>=])(\d+)(?![\d.])", lambda m: m.group(1) + str(min(int(m.group(2)) * factor, 2**31 - 1)), preds)
    scaled.append("|".join([rels, preds, views]))
(out / "scaled.work").write_text("\n".join(scaled) + "\n")
print(f"wrote 14 relations x{factor} and scaled.work to {out}")
