"""Generates tests/golden/reference_vectors.json from the UNMODIFIED reference
(oracle/_ref/libref_ops.so and oracle/_ref/ref_driver, built from
/root/reference by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

Inputs are regenerated from (kind, n, domain, seed) with include/b200_synth.h's
integer generator, so only the reference's OUTPUTS are stored: counts, u64
checksums and SHA-256 digests of the exact output order.
small.init / small.work / small.result beside this file are the reference's
own golden workload files (submission/workloads/small), copied verbatim.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import orc      # noqa: E402
import refbind  # noqa: E402

UNI = 3  # B200_SYNTH_UNIFORM


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, np.uint64).tobytes())
    return h.hexdigest()


def col(n, domain, seed):
    # spread the small case seeds over 64 bits: the generator hashes seed + i
    return orc.synth_column(n, UNI, domain, (seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)


JOIN_CASES = [  # nr, ns, domain, seed
    (1000, 3000, 500, 11), (3000, 1000, 500, 12), (50, 50, 1 << 60, 13), (4000, 4000, 4000, 14),
    (1, 1000, 1, 15), (20000, 30000, 64, 16), (16, 16, 16, 17), (100000, 400000, 1 << 20, 18),
    (100, 100, 1 << 33, 19),
]
QUERY_CASES = [  # sizes, ncols, domain, seed, query
    ([600, 900, 300], 3, 64, 21, "0 1|0.0=1.0|0.1 1.1"),
    ([600, 900, 300], 3, 64, 22, "0 1|0.1=1.1&0.2<40|0.0 1.2 0.1"),
    ([600, 900, 300], 3, 64, 23, "0 1 2|0.0=1.0&1.1=2.1&0.2>10|0.1 1.2 2.0"),
    ([600, 900, 300], 3, 64, 24, "0 0|0.0=1.1|0.2 1.2"),
    ([600, 900, 300], 3, 64, 25, "0 1|0.0=1.0&0.1>1000000|0.1"),
    ([600, 900, 300], 3, 64, 26, "0 1 2|0.0=1.0&1.1=2.1&0.1=33&0.2<60|0.0 2.2"),
    ([600, 900, 300], 3, 64, 27, "0 1|0.0=1.0&0.1=1.1|1.2 0.2"),
    ([5000, 20000, 1000, 300], 3, 1000, 28, "0 1 2 3|1.0=0.0&1.1=2.0&2.1=3.0&1.2>100&1.2<900|1.1 0.2 3.1"),
    ([100, 100], 2, 1 << 40, 29, "0 1|0.0=1.0|0.1 1.1"),   # (almost surely) empty last join -> zeros
]


STATS_CASES = [  # rows, [(domain, offset) per column], seed
    (5000, [(64, 0), (1 << 20, 7), (1, 123)], 31),
    (200000, [(1 << 40, 0), (60_000_000, 1000), (49_999_999, 5)], 32),
    (1, [(10, 3)], 33),
]


def query_relations(sizes, ncols, domain, seed):
    return [[col(n, domain, seed * 1000 + r * 10 + c) for c in range(ncols)] for r, n in enumerate(sizes)]


def main():
    out = {"n_lsb": refbind.N_LSB,
           "next_prime": {str(n): refbind.find_next_prime(n) for n in list(range(0, 64)) + [121, 169, 1000, 4096,
                                                                                             65537, 1 << 20]},
           "joins": [], "reorders": [], "queries": []}
    for nr, ns, domain, seed in JOIN_CASES:
        kr, ks = col(nr, domain, seed), col(ns, domain, seed + 500)
        r, s = refbind.radix_hash_join(kr, ks)
        pay_r, pay_s = col(nr, 1 << 24, seed + 900), col(ns, 1 << 24, seed + 901)
        out["joins"].append({"nr": nr, "ns": ns, "domain": domain, "seed": seed, "m": len(r),
                             "pairs_sha256": sha(r, s),
                             "sum_r": int(pay_r[r.astype(np.int64)].sum(dtype=np.uint64)),
                             "sum_s": int(pay_s[s.astype(np.int64)].sum(dtype=np.uint64))})
        (rk, rr, rh, rp), _ = refbind.reorder(kr, ks)
        out["reorders"].append({"n": nr, "domain": domain, "seed": seed, "hist": [int(x) for x in rh],
                                "psum": [int(x) for x in rp], "tuples_sha256": sha(rk, rr)})
    for sizes, ncols, domain, seed, q in QUERY_CASES:
        rels = query_relations(sizes, ncols, domain, seed)
        line = refbind.run_driver("ref_driver", rels, [q])[0]
        out["queries"].append({"sizes": sizes, "ncols": ncols, "domain": domain, "seed": seed, "query": q,
                               "line": line})
    # column statistics of the reference's loader (relation_map.c:53-83) on relation files written from the generator:
    # small ranges (direct marker array), a range beyond 50 000 000 (the modulo branch) and a constant column
    import tempfile
    out["stats"] = []
    for n, domains, seed in STATS_CASES:
        cols = [col(n, d, seed * 100 + j) + np.uint64(off) for j, (d, off) in enumerate(domains)]
        with tempfile.TemporaryDirectory() as tmp:
            path = Path(tmp) / "rel"
            refbind.write_relation_file(path, cols)
            st = refbind.init_relation_map([path])[0]
        out["stats"].append({"n": n, "domains": domains, "seed": seed, "stats": st})
    (HERE / "reference_vectors.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "reference_vectors.json")


if __name__ == "__main__":
    main()
