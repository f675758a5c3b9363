"""Shared by the CPU and GPU suites: regenerates the inputs of
tests/golden/reference_vectors.json (see golden/make_golden.py)."""
import hashlib
import json
from pathlib import Path

import numpy as np

import orc

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "reference_vectors.json").read_text())
UNI = 3


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, np.uint64).tobytes())
    return h.hexdigest()


def col(n, domain, seed):
    return orc.synth_column(n, UNI, domain, (seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)


def join_inputs(case):
    kr, ks = col(case["nr"], case["domain"], case["seed"]), col(case["ns"], case["domain"], case["seed"] + 500)
    pay_r, pay_s = col(case["nr"], 1 << 24, case["seed"] + 900), col(case["ns"], 1 << 24, case["seed"] + 901)
    return kr, ks, pay_r, pay_s


def query_relations(case):
    return [[col(n, case["domain"], case["seed"] * 1000 + r * 10 + c) for c in range(case["ncols"])]
            for r, n in enumerate(case["sizes"])]


def load_small():
    """The shipped `small` relations (data under oracle/_ref/small), or None."""
    small = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "small"
    if not (small / "r0").exists():
        return None
    rels = []
    for i in range(14):
        raw = np.fromfile(small / f"r{i}", dtype=np.uint64)
        n, c = int(raw[0]), int(raw[1])
        rels.append([raw[2 + j * n: 2 + (j + 1) * n] for j in range(c)])
    return rels


def small_queries():
    g = Path(__file__).resolve().parent / "golden"
    golden = (g / "small.result").read_text().splitlines()
    queries = [l for l in (g / "small.work").read_text().splitlines() if l.strip() and l.strip() != "F"]
    return queries, golden
