#!/usr/bin/env python
"""Where config 5's time goes: host/b200_engine -w 1 with B200_TIMING=2 (per-query times and kernel timers) on the
small schema scaled x F, under a few library knobs.   python scripts/exp_config5.py [factor] [knob=value ...]"""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in (sys.argv[2:] or ["", "B200_TAG64=0", "B200_FUSE_FILTERS=0"])]
work = Path(f"/tmp/scaled_small_{factor}")
if not (work / "scaled.work").exists():
    subprocess.run([sys.executable, str(ROOT / "scripts" / "make_scaled_small.py"), str(factor), str(work)], check=True,
                   stdout=sys.stderr)
stdin = "\n".join((work / "scaled.init").read_text().split()) + "\nDone\n" + (work / "scaled.work").read_text()
queries = [l for l in (work / "scaled.work").read_text().splitlines() if l.strip() and l.strip() != "F"]
ref = None
for env in variants:
    out = subprocess.run([str(ROOT / "host" / "b200_engine"), "-w", "1"], input=stdin, capture_output=True, text=True, cwd=work,
                         env=dict(os.environ, B200_TIMING="2", **env), timeout=3000)
    rows = re.findall(r"query (\d+): ([0-9.]+) ms \|(.*?) \| (.*)", out.stderr)
    total = sum(float(r[1]) for r in rows)
    lines = out.stdout.splitlines()
    ref = ref or lines
    print(f"=== {env or 'default'}: rc {out.returncode}, {len(rows)} queries, {total:.1f} ms in queries, output equals the first variant's: {lines == ref}")
    order = sorted(range(len(rows)), key=lambda i: -float(rows[i][1]))
    for i in order[:8]:
        print(f"   #{i:2d} {float(rows[i][1]):9.2f} ms  {queries[i] if i < len(queries) else ''}\n        {rows[i][2].strip()}")
    for l in out.stderr.splitlines():
        if "slow:" in l:
            print("   " + l)
    if out.returncode:
        print(out.stderr[-1500:])
