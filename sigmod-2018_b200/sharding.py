"""Host-side multi-GPU plan of the join, one process per GPU.

Broadcast plan (small build side): every relation starts position-sharded;
the build side's columns are all-gathered, every rank joins the full build
side with its probe shard, and the k u64 checksums plus the match count are
all-reduced.  u64 sums mod 2^64 are reduced as wrapping int64 sums (NCCL and
gloo have no uint64 SUM).  Works on any torch.distributed backend; the CPU
tests run it under gloo with the oracle as the local join.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[first, first + count) of `n` rows for `rank`; the last rank takes the remainder."""
    per = n // world
    first = rank * per
    count = per if rank < world - 1 else n - first
    return first, count


def u64_to_i64(values) -> list[int]:
    return [int(np.uint64(v).astype(np.int64)) for v in values]


def i64_to_u64(values) -> list[int]:
    return [int(np.int64(v).astype(np.uint64)) for v in values]


def allreduce_checksums(sums, matches, dist=None, device=None):
    """Sum the per-rank checksums (mod 2^64) and match counts over all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(s) % (1 << 64) for s in sums], int(matches)
    import torch
    t = torch.tensor(u64_to_i64(sums) + [int(matches)], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    out = i64_to_u64(t.cpu().tolist())
    return out[:-1], out[-1]


def allgather_column(shard, total_rows: int, dist, out=None):
    """All-gather a position-sharded int64 column (shard_bounds: the last rank may be longer)."""
    import torch
    world = dist.get_world_size()
    per = total_rows // world
    if per * world == total_rows:
        if out is None:
            out = torch.empty(total_rows, dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(out, shard)
        return out
    # ragged: pad every shard to the longest one, gather, trim
    longest = total_rows - per * (world - 1)
    padded = torch.zeros(longest, dtype=shard.dtype, device=shard.device)
    padded[: shard.numel()] = shard
    buf = torch.empty(world * longest, dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(buf, padded)
    parts = [buf[r * longest: r * longest + (per if r < world - 1 else longest)] for r in range(world)]
    return torch.cat(parts)


kMaxResult = 10   # matches + up to 8 sums + overflow count


class BroadcastScatterJoin:
    """Config-2-shaped join over `world` GPUs, one process each (broadcast plan).

    Relations start position-sharded.  Per step and rank:
      1. histogram of the local build shard and of the local probe shard (radix_hist_kernel);
      2. the build histograms are all-gathered (a few KB, NCCL) and turned into this rank's
         scatter cursors inside the GLOBAL partition layout (tiny torch ops on the device);
      3. the build shard is partitioned ONCE, locally, and every partition segment is copied to
         its place in the global partition layout of every rank's build buffer — peer buffers are
         CUDA-IPC mappings, the copy kernel's 256-byte stores travel over NVLink (no all-gather of
         raw columns, no second partition pass on the receivers); up to two build-side SUM
         columns travel with the tuples (early materialisation);
      4. the probe shard is partitioned locally (histogram-free regions + overflow) on a SIDE
         stream, underneath steps 1-3; a stream-ordered all-reduce is the barrier that tells every
         rank the peers' stores have landed;
      5. per-partition build + probe + SUM on local buffers, then the k checksums and the match
         count are all-reduced.
    The probe side never moves.  `dist` may be None (world = 1: the same phases, no exchange),
    which is how the single-GPU test drives the staged C-ABI.
    """

    def __init__(self, b200, torch, dist, rank, world, n_build_total, n_build_local, n_probe_local, n_pay, device,
                 carry32=False, rank_major=False):
        """carry32: the single build-side SUM column (n_pay == 1) holds 32-bit values and travels in the row-id
        slot of the build tuples: no payload buffers, half the bytes in the NVLink broadcast."""
        import ctypes as C
        self.b, self.torch, self.dist, self.rank, self.world = b200, torch, dist, rank, world
        self.C = C
        L = b200.lib()
        self.L = L
        self.n_build_total, self.n_build_local, self.n_probe_local = n_build_total, n_build_local, n_probe_local
        self.n_pay = n_pay
        self.carry32 = bool(carry32) and n_pay == 1
        # rank_major: region r of every build buffer holds rank r's shard in partition order; the broadcast is
        # then ONE contiguous copy per peer on the copy engines (no SM kernel), and the join reads a partition
        # as `world` runs (b200_stage_join_sum_seg).  Needs equal build shards.
        self.rank_major = bool(rank_major) and n_build_local * world == n_build_total
        self.copy_streams = []
        self.bits = int(L.b200_radix_bits_for(n_build_total))
        self.P = 1 << self.bits
        self.device = device
        # build-partition buffers live in cudaMalloc memory so they can be exported over CUDA IPC
        self.tup_b = b200.DeviceColumn(max(n_build_total, 1))
        self.pay_b = [] if self.carry32 else [b200.DeviceColumn(max(n_build_total, 1)) for _ in range(n_pay)]
        # histogram-free probe side: P fixed regions + an overflow array (host.py / DESIGN.md §4)
        self.opt_cap = int(L.b200_opt_region_cap(n_probe_local, self.bits)) if n_probe_local >= (1 << 20) else 0
        self.tup_p = b200.DeviceColumn(max(self.opt_cap * self.P if self.opt_cap else n_probe_local, 1))
        self.ov_p = b200.DeviceColumn(max(n_probe_local, 1)) if self.opt_cap else None
        self.ovcnt = torch.zeros(1, dtype=torch.int32, device=device)
        import os
        on_gpu = device is not None and getattr(device, "type", "") == "cuda"
        self.side = torch.cuda.Stream(device=device) if on_gpu and not os.environ.get("B200_PLAN_SERIAL") else None
        if self.rank_major and on_gpu:
            self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(min(4, max(world - 1, 1)))]
        self.debug = bool(os.environ.get("B200_PLAN_DEBUG"))
        self.marks = []
        self.hist = torch.zeros((2, self.P), dtype=torch.int32, device=device)      # [build, probe] local
        self.hist_all = torch.zeros((world, self.P), dtype=torch.int32, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        self.total_b = torch.zeros(self.P, dtype=torch.int32, device=device)
        self.cur_b = torch.zeros(self.P, dtype=torch.int32, device=device)
        self.result = torch.zeros(kMaxResult, dtype=torch.int64, device=device)
        self.peer_tup = [self.tup_b.ptr] * 1
        self.peer_pay = [[p.ptr] for p in self.pay_b]
        self._imported = []
        if world > 1:
            handles = [self._export(self.tup_b.ptr)] + [self._export(p.ptr) for p in self.pay_b]
            gathered = [None] * world
            dist.all_gather_object(gathered, handles)
            self.peer_tup, self.peer_pay = [], [[] for _ in self.pay_b]
            for r in range(world):
                if r == rank:
                    self.peer_tup.append(self.tup_b.ptr)
                    for k in range(len(self.pay_b)):
                        self.peer_pay[k].append(self.pay_b[k].ptr)
                else:
                    self.peer_tup.append(self._import(gathered[r][0]))
                    for k in range(len(self.pay_b)):
                        self.peer_pay[k].append(self._import(gathered[r][1 + k]))

    def _export(self, ptr):
        buf = self.C.create_string_buffer(64)
        if self.L.b200_ipc_export(ptr, buf) != 0:
            raise RuntimeError("b200_ipc_export failed")
        return bytes(buf.raw)

    def _import(self, handle):
        p = self.L.b200_ipc_import(handle)
        if not p:
            raise RuntimeError("b200_ipc_import failed: " + (self.L.b200_last_error() or b"").decode())
        self._imported.append(p)
        return p

    def step(self, build_keys_ptr, build_pay_ptrs, probe_keys_ptr, proj_cols, proj_side, finish=True):
        """One join: enqueue() then finish().  proj_cols[k]: device pointer of projection k — for a build-side projection (side 0) it must be
        one of build_pay_ptrs (it is read through the early-materialised copy); a probe-side projection
        (side 1) is this rank's local column, indexed by the local probe row id."""
        C, L, torch = self.C, self.L, self.torch
        P, bits, world, rank = self.P, self.bits, self.world, self.rank
        h_b, h_p = self.hist[0], self.hist[1]
        main = torch.cuda.current_stream()
        side = self.side if self.side is not None else main

        def mark(name, stream=None):
            if self.debug and not torch.cuda.is_current_stream_capturing():
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream or main)
                self.marks.append((name, ev))

        self.marks = []
        mark("start")
        # ---- build side: histogram, exchange of the counts, local partition pass ----
        assert L.b200_stage_hist(build_keys_ptr, self.n_build_local, bits, h_b.data_ptr()) == 0
        if world > 1:
            self.dist.all_gather_into_tensor(self.hist_all.view(-1), h_b)
        else:
            self.hist_all[0].copy_(h_b)
        npay = self.n_pay
        pay_cols = (C.c_void_p * max(npay, 1))(*build_pay_ptrs[:npay])
        total_b, cur_b = self.total_b, self.cur_b
        ndst = len(self.peer_tup)
        region = 8 * rank * self.n_build_local          # byte offset of this rank's region (8-byte tuples / values)
        if self.rank_major:
            # partition the local shard straight into this rank's region of its own build buffer
            outs = None if self.carry32 else (C.c_void_p * max(npay, 1))(*[pb.ptr + region for pb in self.pay_b])
            assert L.b200_stage_scatter_build_local(build_keys_ptr, self.n_build_local, rank * self.n_build_local,
                                                    bits, h_b.data_ptr(), self.tup_b.ptr + region, npay, pay_cols,
                                                    outs) == 0
        else:
            assert L.b200_stage_build_cursors(self.hist_all.data_ptr(), world, rank, bits, total_b.data_ptr(),
                                              cur_b.data_ptr()) == 0
            tup_dst = (C.c_void_p * ndst)(*self.peer_tup)
            flat = [self.peer_pay[k][d] for k in range(len(self.pay_b)) for d in range(ndst)]
            pay_dst = None if self.carry32 else (C.c_void_p * max(len(flat), 1))(*flat)
            # the local partition pass of the build shard and the probe-side scatter both want a whole SM's shared
            # memory per CTA, so they run back to back; the NVLink-bound broadcast copy (tiny CTAs) then runs
            # UNDER the probe-side scatter, which is issued on the side stream in between
            assert L.b200_stage_scatter_build(build_keys_ptr, self.n_build_local, rank * self.n_build_local, bits,
                                              h_b.data_ptr(), cur_b.data_ptr(), ndst, tup_dst, npay, pay_cols,
                                              pay_dst, 1) == 0
        mark("build partitioned")
        # ---- probe side: local, independent of the exchange -> side stream ----
        side.wait_stream(main)
        L.b200_set_stream(side.cuda_stream)
        if self.opt_cap:
            assert L.b200_stage_scatter_probe_opt(probe_keys_ptr, self.n_probe_local, bits, self.opt_cap,
                                                  h_p.data_ptr(), self.tup_p.ptr, self.ov_p.ptr,
                                                  self.ovcnt.data_ptr()) == 0
        else:
            assert L.b200_stage_hist(probe_keys_ptr, self.n_probe_local, bits, h_p.data_ptr()) == 0
            with torch.cuda.stream(side):
                cur_p = (torch.cumsum(h_p, 0, dtype=torch.int32) - h_p).contiguous()
            assert L.b200_stage_scatter_probe(probe_keys_ptr, self.n_probe_local, bits, cur_p.data_ptr(),
                                              self.tup_p.ptr) == 0
            cur_p.record_stream(side)
        mark("probe scattered (side)", side)
        L.b200_set_stream(main.cuda_stream)
        # ---- broadcast of the partitioned build shard, concurrent with the probe scatter ----
        if self.rank_major:
            # one contiguous peer copy per destination (copy engines, NVLink), every rank starting at a different peer
            nbytes = 8 * self.n_build_local
            for j in range(1, world):
                dpeer = (rank + j) % world
                cs = self.copy_streams[(j - 1) % len(self.copy_streams)] if self.copy_streams else main
                cs.wait_stream(main)
                L.b200_set_stream(cs.cuda_stream)
                assert L.b200_copy_device_async(self.peer_tup[dpeer] + region, self.tup_b.ptr + region, nbytes) == 0
                for k in range(len(self.pay_b)):
                    assert L.b200_copy_device_async(self.peer_pay[k][dpeer] + region, self.pay_b[k].ptr + region,
                                                    nbytes) == 0
            L.b200_set_stream(main.cuda_stream)
            for cs in self.copy_streams:
                main.wait_stream(cs)
        else:
            assert L.b200_stage_scatter_build(build_keys_ptr, self.n_build_local, rank * self.n_build_local, bits,
                                              h_b.data_ptr(), cur_b.data_ptr(), ndst, tup_dst, npay, pay_cols,
                                              pay_dst, 2) == 0
        mark("broadcast done")
        if world > 1:
            self.dist.all_reduce(self.token)      # stream-ordered barrier: every peer's broadcast has landed
        mark("barrier done")
        main.wait_stream(side)
        k = len(proj_cols)
        cols = (C.c_void_p * max(k, 1))(*proj_cols)
        sides = (C.c_int * max(k, 1))(*proj_side)
        part = []
        for col, side_k in zip(proj_cols, proj_side):
            if side_k == 0:
                part.append(1 if self.carry32 else self.pay_b[build_pay_ptrs.index(col)].ptr)
            else:
                part.append(None)
        part_vals = (C.c_void_p * max(k, 1))(*part)
        if self.rank_major:
            args = (self.tup_b.ptr, self.hist_all.data_ptr(), world, self.n_build_local, self.tup_p.ptr, h_p.data_ptr(),
                    bits, k, cols, sides, part_vals, self.opt_cap, self.ov_p.ptr if self.opt_cap else None,
                    self.ovcnt.data_ptr())
        else:
            args = (self.tup_b.ptr, total_b.data_ptr(), self.tup_p.ptr, h_p.data_ptr(), bits, k, cols, sides,
                    part_vals, self.opt_cap, self.ov_p.ptr if self.opt_cap else None, self.ovcnt.data_ptr())
        # asynchronous join: {matches, sums, overflow count} stay on the device and are all-reduced in place
        # (u64 sums mod 2^64 == wrapping int64 sums); that all-reduce also ends the step on every rank, so no
        # peer can start overwriting this rank's build buffers before its join has finished
        res = self.result[: k + 2]
        if self.rank_major:
            assert L.b200_stage_join_sum_seg(*args, res.data_ptr(), None, None) == 0
        else:
            assert L.b200_stage_join_sum_async(*args, res.data_ptr()) == 0
        if world > 1:
            self.dist.all_reduce(res)
        mark("join + all-reduce done")
        self._pending = (args, k)
        if not finish:
            return None
        return self.finish()

    def enqueue(self, *a):
        """Everything of a step that runs on the device, without any host read-back: capturable in a CUDA
        graph (bench.py does, for N > 1, to take the per-step launch and collective set-up latency out)."""
        return self.step(*a, finish=False)

    def finish(self):
        """Read {matches, sums, overflow count} back; run the exact overflow pass when it is needed."""
        C, L, world, rank = self.C, self.L, self.world, self.rank
        args, k = self._pending
        res = self.result[: k + 2]
        host = res.cpu().tolist()
        if self.debug and rank == 0 and self.marks:
            t0 = self.marks[0][1]
            print("plan timeline (ms): " + ", ".join(f"{n} {t0.elapsed_time(e):.3f}" for n, e in self.marks[1:]),
                  file=__import__("sys").stderr)
        if host[k + 1] != 0:
            # some rank's histogram-free scatter overflowed (skewed keys): redo the join part of this step
            # synchronously, overflow pass included, and reduce again
            sums = (C.c_uint64 * max(k, 1))()
            m = C.c_uint64(0)
            if self.rank_major:
                assert L.b200_stage_join_sum_seg(*args, None, sums, C.byref(m)) == 0
            else:
                assert L.b200_stage_join_sum(*args, sums, C.byref(m)) == 0
            return allreduce_checksums([int(x) for x in sums[:k]], int(m.value), self.dist if world > 1 else None,
                                       self.device)
        return i64_to_u64(host[1: k + 1]), int(host[0])

    def close(self):
        for p in self._imported:
            self.L.b200_ipc_close(p)
        self._imported = []
