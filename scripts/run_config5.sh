#!/bin/bash
# BASELINE config 5: the small.work batch on the small schema scaled by <factor>, reference (CPU) against
# host/b200_engine with 1, 4 and 8 concurrent worker threads (one CUDA stream each).  Lines on which the two
# differ are re-evaluated with the oracle executor (tests/orc.py): the reference's threaded partition pass
# loses tuples of tiny relations (tests/test_oracle_vs_reference.py::test_reference_loses_tuples_of_tiny_relations).
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD
F=${1:-10}; DIR=${2:-/tmp/scaled_small_$F}
python scripts/make_scaled_small.py $F $DIR
cd $DIR
(cat scaled.init; echo Done; cat scaled.work) > in.txt
now() { python -c "import time; print(time.time())"; }
t0=$(now); $ROOT/oracle/_ref/radixhash < in.txt > ref.out; t1=$(now)
python -c "print(f'reference radixhash (CPU, whole process): {$t1 - $t0:.2f} s')"
for w in 1 4 8; do
  t0=$(now); B200_TIMING=1 B200_WORKERS=$w $ROOT/host/b200_engine < in.txt > gpu_$w.out 2> gpu_$w.err; t1=$(now)
  python -c "import re,sys; t=open('gpu_$w.err').read(); b=[float(x) for x in re.findall(r'workers: ([0-9.]+) s', t)]; print('  ' + t.splitlines()[0]); print(f'  query time over {len(b)} batches: {sum(b):.3f} s')"
  python -c "print(f'b200_engine workers=$w (whole process incl. CUDA start-up and upload): {$t1 - $t0:.2f} s')"
  cmp -s gpu_1.out gpu_$w.out || echo "  workers=$w output differs from workers=1 !"
done
python - <<PY
import sys
sys.path.insert(0, "$ROOT/tests")
import numpy as np, orc
from pathlib import Path
d = Path("$DIR")
ref, gpu = (d / "ref.out").read_text().splitlines(), (d / "gpu_1.out").read_text().splitlines()
qs = [l for l in (d / "scaled.work").read_text().splitlines() if "|" in l]
diff = [i for i, (a, b) in enumerate(zip(ref, gpu)) if a != b]
print(f"{len(qs)} queries, {len(diff)} line(s) differ between the reference and the GPU engine")
if diff:
    rels = []
    for i in range(14):
        raw = np.fromfile(d / f"r{i}", dtype=np.uint64)
        n, c = int(raw[0]), int(raw[1])
        rels.append([raw[2 + j * n: 2 + (j + 1) * n] for j in range(c)])
    for i in diff:
        o = orc.execute_query(qs[i], rels)
        print(f"  query {i + 1}: reference '{ref[i]}'  gpu '{gpu[i]}'  oracle '{o}'  ->",
              "gpu == oracle" if o == gpu[i] else "reference == oracle" if o == ref[i] else "NEITHER")
PY
