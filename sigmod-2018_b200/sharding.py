"""Host-side multi-GPU plan of the join, one process per GPU.

Broadcast plan (small build side): every relation starts position-sharded;
the build side's columns are all-gathered, every rank joins the full build
side with its probe shard, and the k u64 checksums plus the match count are
all-reduced.  u64 sums mod 2^64 are reduced as wrapping int64 sums (NCCL and
gloo have no uint64 SUM).  Works on any torch.distributed backend; the CPU
tests run it under gloo with the oracle as the local join.

Exchange plan (large build side, SURVEY §8e "all-to-all"): both relations are
radix-partitioned locally and every partition travels to its owner rank
(`partition_owner`); `exchange_layout` is the host restatement of the layout
the device computes (exchange_cursors_kernel), `ShardedExchangeJoin` the plan.
"""
from __future__ import annotations

import numpy as np


def _ck(L, rc):
    """Raise when a C-ABI call failed (never `assert call(...) == 0`: python -O strips asserts and with them the call)."""
    if rc != 0:
        raise RuntimeError("libb200join: " + (L.b200_last_error() or b"error").decode())


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
                 carry32=False, rank_major=False, carry_probe=False):
        """carry32: the single build-side SUM column (n_pay == 1) holds 32-bit values and travels in the row-id
        slot of the build tuples: no payload buffers, half the bytes in the NVLink broadcast.
        carry_probe: the single probe-side SUM column holds 32-bit values and is streamed into the row-id slot of
        the probe tuples by the (histogram-free) scatter instead of being gathered per match by the join; worth
        it when matches are not rare (see run_join in engine.cu)."""
        import ctypes as C
        self.b, self.torch, self.dist, self.rank, self.world = b200, torch, dist, rank, world
        self.C = C
        L = b200.lib()
        self.L = L
        self.n_build_total, self.n_build_local, self.n_probe_local = n_build_total, n_build_local, n_probe_local
        self.n_pay = n_pay
        self.carry32 = bool(carry32) and n_pay == 1
        self.carry_probe = bool(carry_probe)
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
        _ck(L, L.b200_stage_hist(build_keys_ptr, self.n_build_local, bits, h_b.data_ptr()))
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
            _ck(L, L.b200_stage_scatter_build_local(build_keys_ptr, self.n_build_local, rank * self.n_build_local,
                                                    bits, h_b.data_ptr(), self.tup_b.ptr + region, npay, pay_cols,
                                                    outs))
        else:
            _ck(L, L.b200_stage_build_cursors(self.hist_all.data_ptr(), world, rank, bits, total_b.data_ptr(),
                                              cur_b.data_ptr()))
            tup_dst = (C.c_void_p * ndst)(*self.peer_tup)
            flat = [self.peer_pay[k][d] for k in range(len(self.pay_b)) for d in range(ndst)]
            pay_dst = None if self.carry32 else (C.c_void_p * max(len(flat), 1))(*flat)
            # the local partition pass of the build shard and the probe-side scatter both want a whole SM's shared
            # memory per CTA, so they run back to back; the NVLink-bound broadcast copy (tiny CTAs) then runs
            # UNDER the probe-side scatter, which is issued on the side stream in between
            _ck(L, L.b200_stage_scatter_build(build_keys_ptr, self.n_build_local, rank * self.n_build_local, bits,
                                              h_b.data_ptr(), cur_b.data_ptr(), ndst, tup_dst, npay, pay_cols,
                                              pay_dst, 1))
        mark("build partitioned")
        # ---- probe side: local, independent of the exchange -> side stream ----
        side.wait_stream(main)
        L.b200_set_stream(side.cuda_stream)
        probe_projs = [kk for kk, sd in enumerate(proj_side) if sd == 1]
        carried = probe_projs[0] if (self.carry_probe and self.opt_cap and len(probe_projs) == 1) else -1
        if carried >= 0:
            _ck(L, L.b200_stage_scatter_probe_opt_carry(probe_keys_ptr, self.n_probe_local, bits, self.opt_cap,
                                                        h_p.data_ptr(), self.tup_p.ptr, self.ov_p.ptr,
                                                        self.ovcnt.data_ptr(), proj_cols[carried]))
        elif self.opt_cap:
            _ck(L, L.b200_stage_scatter_probe_opt(probe_keys_ptr, self.n_probe_local, bits, self.opt_cap,
                                                  h_p.data_ptr(), self.tup_p.ptr, self.ov_p.ptr,
                                                  self.ovcnt.data_ptr()))
        else:
            _ck(L, L.b200_stage_hist(probe_keys_ptr, self.n_probe_local, bits, h_p.data_ptr()))
            with torch.cuda.stream(side):
                cur_p = (torch.cumsum(h_p, 0, dtype=torch.int32) - h_p).contiguous()
            _ck(L, L.b200_stage_scatter_probe(probe_keys_ptr, self.n_probe_local, bits, cur_p.data_ptr(),
                                              self.tup_p.ptr))
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
                _ck(L, L.b200_copy_device_async(self.peer_tup[dpeer] + region, self.tup_b.ptr + region, nbytes))
                for k in range(len(self.pay_b)):
                    _ck(L, L.b200_copy_device_async(self.peer_pay[k][dpeer] + region, self.pay_b[k].ptr + region,
                                                    nbytes))
            L.b200_set_stream(main.cuda_stream)
            for cs in self.copy_streams:
                main.wait_stream(cs)
        else:
            _ck(L, L.b200_stage_scatter_build(build_keys_ptr, self.n_build_local, rank * self.n_build_local, bits,
                                              h_b.data_ptr(), cur_b.data_ptr(), ndst, tup_dst, npay, pay_cols,
                                              pay_dst, 2))
        mark("broadcast done")
        if world > 1:
            self.dist.all_reduce(self.token)      # stream-ordered barrier: every peer's broadcast has landed
        mark("barrier done")
        main.wait_stream(side)
        k = len(proj_cols)
        cols = (C.c_void_p * max(k, 1))(*proj_cols)
        sides = (C.c_int * max(k, 1))(*proj_side)
        part = []
        for kk, (col, side_k) in enumerate(zip(proj_cols, proj_side)):
            if side_k == 0:
                part.append(1 if self.carry32 else self.pay_b[build_pay_ptrs.index(col)].ptr)
            else:
                part.append(1 if kk == carried else None)
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
            _ck(L, L.b200_stage_join_sum_seg(*args, res.data_ptr(), None, None))
        else:
            _ck(L, L.b200_stage_join_sum_async(*args, res.data_ptr()))
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
                _ck(L, L.b200_stage_join_sum_seg(*args, None, sums, C.byref(m)))
            else:
                _ck(L, L.b200_stage_join_sum(*args, sums, C.byref(m)))
            return allreduce_checksums([int(x) for x in sums[:k]], int(m.value), self.dist if world > 1 else None,
                                       self.device)
        return i64_to_u64(host[1: k + 1]), int(host[0])

    def close(self):
        for p in self._imported:
            self.L.b200_ipc_close(p)
        self._imported = []


def partition_owner(p, world: int, bits: int):
    """Rank that owns radix partition p (scalar or numpy array): the top bits of the partition id, so every
    rank owns a contiguous range of the 2^bits partitions."""
    return (np.asarray(p, dtype=np.uint64) * np.uint64(world)) >> np.uint64(bits)


def exchange_layout(hist_all, rank: int, bits: int):
    """Layout of the radix-sharded exchange from the all-gathered histograms hist_all[world][2^bits]:
    (src_off[P+1], dst_start[P], own_total[P], rows_received) for `rank` — the host restatement of
    exchange_cursors_kernel.  An owner's receive buffer is partition-major, source-rank-minor."""
    h = np.asarray(hist_all, dtype=np.uint64)
    world, P = h.shape
    assert P == 1 << bits and world <= P
    src_off = np.concatenate([[0], np.cumsum(h[rank])]).astype(np.uint64)
    total = h.sum(axis=0)
    prefix = np.concatenate([[0], np.cumsum(total)]).astype(np.uint64)       # global exclusive prefix, P + 1
    owner = partition_owner(np.arange(P), world, bits).astype(np.int64)
    first_of_owner = np.searchsorted(owner, np.arange(world), side="left")   # owner ranges are contiguous
    base = prefix[first_of_owner][owner]
    before = h[:rank].sum(axis=0) if rank else np.zeros(P, dtype=np.uint64)
    dst_start = prefix[:-1] - base + before
    own_total = np.where(owner == rank, total, 0).astype(np.uint64)
    return src_off, dst_start.astype(np.uint64), own_total, int(own_total.sum())


class ShardedExchangeJoin:
    """Radix-sharded join over `world` GPUs, one process each (all-to-all plan; config 4's shape).

    Relations start position-sharded.  Per step and rank:
      1. histograms of the local build and probe shards over 2^bits radix partitions; ONE all-gather of both;
      2. exchange_cursors_kernel turns them into the exchange layout (exchange_layout above);
      3. both shards are partitioned locally into staging buffers (the probe side on a side stream, under the
         all-gather), SUM columns travelling with the tuples: one column per side whose values fit 32 bits
         rides in the row-id slot (8 bytes per row on the wire); otherwise up to two 8-byte payload columns
         per side are exchanged alongside and the probe tuples' row ids are rewritten to receive positions;
      4. segment_exchange_kernel stores every staged row into the receive buffer of the partition's owner —
         peers' buffers are CUDA-IPC mappings, a warp stores 256 contiguous bytes over NVLink; a
         stream-ordered all-reduce of a token is the barrier that says every peer's rows have landed;
      5. the owner joins its partitions (masked histograms) and the checksums are all-reduced.
    Receive buffers are sized by `recv_rows_*`; `measure_capacity` returns what a given input needs (skewed
    probe keys make owners uneven).  A step that would overflow them is detected on the device and raises.
    """

    def __init__(self, b200, torch, dist, rank, world, n_build_total, n_build_local, n_probe_local,
                 n_pay_build, n_pay_probe, device, recv_rows_build=None, recv_rows_probe=None, size_from=None,
                 carry_build=False, carry_probe=False, bits=None):
        """recv_rows_*: capacity of this rank's receive buffers in rows; None = measure it from the key columns
        `size_from = (build_keys_ptr, probe_keys_ptr)` (one histogram pass + all-gather at construction)."""
        import ctypes as C
        import os
        self.b, self.torch, self.dist, self.rank, self.world, self.C = b200, torch, dist, rank, world, C
        L = self.L = b200.lib()
        self.n_build_local, self.n_probe_local = n_build_local, n_probe_local
        self.npb, self.npp = n_pay_build, n_pay_probe
        assert 0 <= n_pay_build <= 2 and 0 <= n_pay_probe <= 2
        self.carry_b = bool(carry_build) and n_pay_build == 1
        self.carry_p = bool(carry_probe) and n_pay_probe == 1
        # partitions are sized for the per-partition table exactly as on one GPU: by the GLOBAL build side
        # (and at least four partitions per rank, so that owners of a tiny build side stay balanced)
        self.bits = int(bits) if bits else max(int(L.b200_radix_bits_for(n_build_total)), (world - 1).bit_length() + 2)
        self.P = 1 << self.bits
        assert self.P >= world
        self.device = device
        i32 = dict(dtype=torch.int32, device=device)
        P = self.P
        self.hist = torch.zeros((2, P), **i32)                # [build, probe] local
        self.hist_all = torch.zeros((world, 2, P), **i32)
        if recv_rows_build is None or recv_rows_probe is None:
            assert size_from is not None, "give recv_rows_build / recv_rows_probe or size_from"
            recv_rows_build, recv_rows_probe = self.measure_capacity(*size_from)
        self.cap_b, self.cap_p = max(int(recv_rows_build), 1), max(int(recv_rows_probe), 1)
        col = b200.DeviceColumn
        self.stage_b, self.stage_p = col(max(n_build_local, 1)), col(max(n_probe_local, 1))
        self.stage_pay_b = [] if self.carry_b else [col(max(n_build_local, 1)) for _ in range(n_pay_build)]
        self.stage_pay_p = [] if self.carry_p else [col(max(n_probe_local, 1)) for _ in range(n_pay_probe)]
        self.recv_b, self.recv_p = col(self.cap_b), col(self.cap_p)
        self.recv_pay_b = [] if self.carry_b else [col(self.cap_b) for _ in range(n_pay_build)]
        self.recv_pay_p = [] if self.carry_p else [col(self.cap_p) for _ in range(n_pay_probe)]
        on_gpu = device is not None and getattr(device, "type", "") == "cuda"
        self.side = torch.cuda.Stream(device=device) if on_gpu and not os.environ.get("B200_PLAN_SERIAL") else None
        self.debug = bool(os.environ.get("B200_PLAN_DEBUG"))
        self.marks = []
        self.src_off = torch.zeros((2, P + 1), **i32)
        self.dst_start = torch.zeros((2, P), **i32)
        self.own_total = torch.zeros((2, P), **i32)
        self.need = torch.zeros((2, 2), **i32)                # [side][rows received, overflow flag]
        self.token = torch.zeros(1, **i32)
        self.result = torch.zeros(kMaxResult, dtype=torch.int64, device=device)
        self._imported = []
        mine = [self.recv_b.ptr, self.recv_p.ptr] + [c.ptr for c in self.recv_pay_b + self.recv_pay_p]
        self.peers = [mine]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, [self._export(p) for p in mine])
            self.peers = [mine if r == rank else [self._import(h) for h in gathered[r]] for r in range(world)]

    _export = BroadcastScatterJoin._export
    _import = BroadcastScatterJoin._import

    def measure_capacity(self, build_keys_ptr, probe_keys_ptr):
        """(rows, rows) the receive buffers of the most loaded rank must hold for these inputs."""
        L, torch = self.L, self.torch
        _ck(L, L.b200_stage_hist(build_keys_ptr, self.n_build_local, self.bits, self.hist[0].data_ptr()))
        _ck(L, L.b200_stage_hist(probe_keys_ptr, self.n_probe_local, self.bits, self.hist[1].data_ptr()))
        return self._capacity_from_hist()

    def _capacity_from_hist(self):
        torch, world = self.torch, self.world
        if world > 1:
            self.dist.all_gather_into_tensor(self.hist_all.view(-1), self.hist.view(-1))
        else:
            self.hist_all[0].copy_(self.hist)
        h = self.hist_all.cpu().numpy().astype(np.uint64)
        need = [max(exchange_layout(h[:, s, :], r, self.bits)[3] for r in range(world)) for s in (0, 1)]
        return need[0], need[1]

    def step(self, build_keys_ptr, build_pay_ptrs, probe_keys_ptr, probe_pay_ptrs, finish=True):
        """One join; returns ([Σ build_pay..., Σ probe_pay...], matches) over all ranks."""
        C, L, torch = self.C, self.L, self.torch
        P, bits, world, rank = self.P, self.bits, self.world, self.rank
        main = torch.cuda.current_stream()
        side = self.side if self.side is not None else main
        vp = C.c_void_p

        def mark(name, stream=None):
            if self.debug:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream or main)
                self.marks.append((name, ev))

        def arr(ptrs):
            return (vp * max(len(ptrs), 1))(*ptrs)

        self.marks = []
        mark("start")
        h_b, h_p = self.hist[0], self.hist[1]
        _ck(L, L.b200_stage_hist(build_keys_ptr, self.n_build_local, bits, h_b.data_ptr()))
        _ck(L, L.b200_stage_hist(probe_keys_ptr, self.n_probe_local, bits, h_p.data_ptr()))
        mark("histograms")
        # ---- probe side: the local partition pass needs only the local histogram -> side stream ----
        side.wait_stream(main)
        L.b200_set_stream(side.cuda_stream)
        outs = None if self.carry_p else arr([c.ptr for c in self.stage_pay_p])
        _ck(L, L.b200_stage_scatter_build_local(probe_keys_ptr, self.n_probe_local, 0, bits, h_p.data_ptr(),
                                                self.stage_p.ptr, self.npp, arr(probe_pay_ptrs[:self.npp]),
                                                outs))
        mark("probe partitioned (side)", side)
        L.b200_set_stream(main.cuda_stream)
        # ---- exchange layout ----
        if world > 1:
            self.dist.all_gather_into_tensor(self.hist_all.view(-1), self.hist.view(-1))
        else:
            self.hist_all[0].copy_(self.hist)
        by_side = self.hist_all.permute(1, 0, 2).contiguous()          # [side][world][P]
        for s, cap in ((0, self.cap_b), (1, self.cap_p)):
            _ck(L, L.b200_stage_exchange_cursors(by_side[s].data_ptr(), world, rank, bits, cap,
                                                 self.src_off[s].data_ptr(), self.dst_start[s].data_ptr(),
                                                 self.own_total[s].data_ptr(), self.need[s].data_ptr()))
        # ---- build side: partition, exchange ----
        outs = None if self.carry_b else arr([c.ptr for c in self.stage_pay_b])
        _ck(L, L.b200_stage_scatter_build_local(build_keys_ptr, self.n_build_local, 0, bits, h_b.data_ptr(),
                                                self.stage_b.ptr, self.npb, arr(build_pay_ptrs[:self.npb]),
                                                outs))
        nb, npp = len(self.recv_pay_b), len(self.recv_pay_p)
        tup_dst_b = arr([self.peers[d][0] for d in range(world)])
        tup_dst_p = arr([self.peers[d][1] for d in range(world)])
        pay_dst_b = arr([self.peers[d][2 + k] for k in range(nb) for d in range(world)])
        pay_dst_p = arr([self.peers[d][2 + nb + k] for k in range(npp) for d in range(world)])
        _ck(L, L.b200_stage_exchange_segments(self.stage_b.ptr, nb, arr([c.ptr for c in self.stage_pay_b]),
                                              self.n_build_local, bits, world, self.src_off[0].data_ptr(),
                                              self.dst_start[0].data_ptr(), self.cap_b, 0, tup_dst_b,
                                              pay_dst_b))
        mark("build exchanged")
        main.wait_stream(side)
        _ck(L, L.b200_stage_exchange_segments(self.stage_p.ptr, npp, arr([c.ptr for c in self.stage_pay_p]),
                                              self.n_probe_local, bits, world, self.src_off[1].data_ptr(),
                                              self.dst_start[1].data_ptr(), self.cap_p, 0 if self.carry_p else 1,
                                              tup_dst_p, pay_dst_p))
        mark("probe exchanged")
        if world > 1:
            self.dist.all_reduce(self.token)      # stream-ordered barrier: every peer's rows have landed
        mark("barrier done")
        # ---- local join of the owned partitions ----
        k = self.npb + self.npp
        in_rid = 1
        cols, sides, part = [], [], []
        for j in range(self.npb):
            cols.append(build_pay_ptrs[j]); sides.append(0)
            part.append(in_rid if self.carry_b else self.recv_pay_b[j].ptr)
        for j in range(self.npp):
            # carried: the probe tuple's row-id slot is the value; else the row id is the receive position
            cols.append(probe_pay_ptrs[j] if self.carry_p else self.recv_pay_p[j].ptr); sides.append(1)
            part.append(in_rid if self.carry_p else None)
        res = self.result[: k + 2]
        _ck(L, L.b200_stage_join_sum_async(self.recv_b.ptr, self.own_total[0].data_ptr(), self.recv_p.ptr,
                                           self.own_total[1].data_ptr(), bits, k, arr(cols),
                                           (C.c_int * max(k, 1))(*sides), arr(part), 0, None, None,
                                           res.data_ptr()))
        res[k + 1: k + 2].add_(self.need[:, 1].sum())       # receive-buffer overflow flags of this rank
        if world > 1:
            self.dist.all_reduce(res)     # also ends the step: no peer overwrites buffers still being joined
        mark("join + all-reduce done")
        self._k = k
        return self.finish() if finish else None

    def enqueue(self, *a):
        return self.step(*a, finish=False)

    def finish(self):
        k = self._k
        host = self.result[: k + 2].cpu().tolist()
        if self.debug and self.rank == 0 and self.marks:
            t0 = self.marks[0][1]
            print("exchange plan timeline (ms): " + ", ".join(f"{n} {t0.elapsed_time(e):.3f}"
                                                               for n, e in self.marks[1:]),
                  file=__import__("sys").stderr)
        if host[k + 1] != 0:
            need = self.need.cpu().tolist()
            raise RuntimeError(f"exchange receive buffers too small on {host[k + 1]} (rank, side) pairs; this rank "
                               f"needs build {need[0][0]} / probe {need[1][0]} rows, has {self.cap_b} / {self.cap_p}")
        return i64_to_u64(host[1: k + 1]), int(host[0])

    def close(self):
        for p in self._imported:
            self.L.b200_ipc_close(p)
        self._imported = []


class MultiJoin:
    """Thin binding over the C-ABI multi-GPU plans (csrc/multi.cu: b200_multi_*): no torch in the step, no NCCL.

    One instance per rank.  `dist` (any torch.distributed backend, may be None for world = 1) is used ONCE, at
    construction, to swap the CUDA-IPC handles of the ranks' shared regions and to agree on the row totals; a step
    is `enqueue` (device work only: partition, copy-engine broadcast or NVLink exchange, join, result exchange
    through flags in peer memory) + `finish` (the one host synchronisation).  Keys and SUM values must be below
    2^32 (8-byte tuples); wider columns take the general classes above.
    """

    def __init__(self, b200, dist, rank, world, device_index, plan, n_build_local, n_probe_local, has_build_sum=True,
                 has_probe_sum=True, radix_bits=0, chunks=0, recv_rows_build=0, recv_rows_probe=0, peers_in_process=None,
                 hot_keys=True):
        import ctypes as C
        self.b, self.L, self.C, self.rank, self.world = b200, b200.lib(), C, rank, world
        counts = [(n_build_local, n_probe_local)]
        if world > 1 and peers_in_process is None:
            counts = [None] * world
            dist.all_gather_object(counts, (n_build_local, n_probe_local))
        elif peers_in_process is not None:
            counts = peers_in_process["counts"]
        cfg = b200.CMultiConfig(plan=plan, rank=rank, world=world, device=device_index,
                                n_build_total=sum(c[0] for c in counts), n_probe_total=sum(c[1] for c in counts),
                                n_build_local=n_build_local, n_probe_local=n_probe_local,
                                n_build_local_max=max(c[0] for c in counts), n_probe_local_max=max(c[1] for c in counts),
                                has_build_sum=int(has_build_sum), has_probe_sum=int(has_probe_sum),
                                radix_bits=radix_bits, chunks=chunks, recv_rows_build=recv_rows_build,
                                recv_rows_probe=recv_rows_probe, hot_keys=0 if hot_keys else -1)
        self.nproj = int(has_build_sum) + int(has_probe_sum)
        self.plan = self.L.b200_multi_create(C.byref(cfg))
        if not self.plan:
            raise RuntimeError("b200_multi_create: " + (self.L.b200_last_error() or b"").decode())
        if world > 1 and peers_in_process is None:
            buf = C.create_string_buffer(64)
            _ck(self.L, self.L.b200_multi_export(self.plan, buf))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(buf.raw))
            for r in range(world):
                if r != rank:
                    _ck(self.L, self.L.b200_multi_connect_ipc(self.plan, r, handles[r]))

    def connect_in_process(self, others, devices):
        """Several ranks inside ONE process (tests: ranks emulated on one GPU; b200_join_sum_multi does the same
        with one thread per GPU): peers are plain pointers."""
        for r, o in enumerate(others):
            if r != self.rank:
                _ck(self.L, self.L.b200_multi_connect_ptr(self.plan, r, self.L.b200_multi_shared_ptr(o.plan), devices[r]))

    @property
    def bits(self):
        return int(self.L.b200_multi_radix_bits(self.plan))

    def enqueue(self, build_keys_ptr, build_sum_ptr, probe_keys_ptr, probe_sum_ptr, phases=0):
        _ck(self.L, self.L.b200_multi_enqueue(self.plan, build_keys_ptr, build_sum_ptr, probe_keys_ptr, probe_sum_ptr,
                                              phases))

    def finish(self):
        C = self.C
        sums = (C.c_uint64 * 2)()
        m = C.c_uint64(0)
        _ck(self.L, self.L.b200_multi_finish(self.plan, sums, C.byref(m)))
        return [int(x) for x in sums[: self.nproj]], int(m.value)

    def step(self, build_keys_ptr, build_sum_ptr, probe_keys_ptr, probe_sum_ptr):
        self.enqueue(build_keys_ptr, build_sum_ptr, probe_keys_ptr, probe_sum_ptr)
        return self.finish()

    def received(self):
        out = (self.C.c_uint64 * 2)()
        _ck(self.L, self.L.b200_multi_received(self.plan, out))
        return int(out[0]), int(out[1])

    def close(self):
        if self.plan:
            self.L.b200_multi_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
