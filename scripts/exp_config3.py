#!/usr/bin/env python
"""Where the time of the config-3 query goes: every operator call of host.py's execute_query timed on the host with a
stream synchronise behind it (so launch gaps, allocations and host round trips show), fused and eager filters.
    python scripts/exp_config3.py [scale_bits]"""
import ctypes as C
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 0


class A:
    gpus, scale_bits = 1, scale


env = bench.Env(A)
torch, b200, L = env.torch, env.b200, env.L
h = b200.host
shape = bench.config3_shape(scale)
kinds = {"iota": b200.SYNTH_IOTA, "uni": b200.SYNTH_UNIFORM, "pay": b200.SYNTH_PAYLOAD}
rels = [[env.synth(rows, 0, kinds[k], kk, seed) for k, kk, seed in cols] for rows, cols in shape]
rm = b200.DeviceRelationMap([[(c.data_ptr(), c.numel(), int(c.max().item())) for c in rel] for rel in rels])
q = h.parse_query(bench.CONFIG3_QUERY)


def timed(name, fn, log):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    L.b200_synchronize()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    log.append((name, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
    return r


def run(log):
    node = h._make_batch_node(q)
    relsb = node.relations
    inter = C.POINTER(h.CInterRes)()
    L.InitInterResults(C.byref(inter), len(q.relations))
    for b, c, op, k in q.filters:
        fp = h.CFilterPred(b, c, k, op.encode())
        res = timed(f"Filter {b}.{c}{op}{k}", lambda: L.Filter(inter, C.byref(fp), rm.array, relsb), log)
        timed("InsertSingleRowIds", lambda: L.InsertSingleRowIdsToInterResult(C.byref(inter), b, res), log)
        L.FreeResult(res)
    for b1, c1, b2, c2 in q.joins:
        r1 = timed(f"GetRelation {b1}.{c1}", lambda: L.GetRelation(b1, c1, inter, rm.array, relsb), log)
        r2 = timed(f"GetRelation {b2}.{c2}", lambda: L.GetRelation(b2, c2, inter, rm.array, relsb), log)
        res = timed("RadixHashJoin", lambda: L.RadixHashJoin(r1, r2, None), log)
        L.FreeRelation(r1)
        L.FreeRelation(r2)
        timed("InsertJoin", lambda: L.InsertJoinToInterResults(inter, b1, b2, res), log)
        L.FreeResult(res)
    sums = (C.c_uint64 * len(q.views))()
    rows = C.c_uint64(0)
    timed("calculate_sums", lambda: L.b200_calculate_sums(inter, rm.array, C.byref(node), sums, C.byref(rows)), log)
    L.FreeInterResults(inter)
    return [int(s) for s in sums]


for fuse in (1, 0):
    L.b200_set_fuse_filters(fuse)
    for rep in range(3):
        log = []
        t0 = time.perf_counter()
        sums = run(log)
        total = (time.perf_counter() - t0) * 1e3
    print(f"--- fuse_filters={fuse}: {total:.2f} ms (timed calls sum {sum(x[2] for x in log):.2f}), sums {sums}")
    for name, call_ms, done_ms in log:
        print(f"   {name:28s} call {call_ms:8.3f} ms   done {done_ms:8.3f} ms")
    # untimed-inside run: the whole query, as bench.py times it
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        b200.execute_query(bench.CONFIG3_QUERY, rm)
    torch.cuda.synchronize()
    print(f"   execute_query: {(time.perf_counter() - t0) * 200:.2f} ms per query")
    L.b200_set_profiling(1)
    b200.execute_query(bench.CONFIG3_QUERY, rm)
    torch.cuda.synchronize()
    print("   last kernel ms:", {n: round(b200.last_kernel_ms(n), 3) for n in
                                 ("filter", "filter_fused", "hist_b", "hist_p", "scatter_b", "scatter_p", "join", "join_write")
                                 if b200.last_kernel_ms(n) >= 0})
    L.b200_set_profiling(0)
