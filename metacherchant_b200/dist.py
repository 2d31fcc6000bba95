"""One process per GPU: hash-sharded k-mer counting.

Every rank parses and extracts the canonical k-mers of ITS OWN slice of the reads (libmcgpu, on its GPU), groups the
hashed keys by owner — owner(key) = top log2(world) bits of the bijective 62-bit hash, so each GPU owns one slice of
the table — and the ranks swap key groups with ONE all-to-all per read batch (torch.distributed: NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).  The receiver inserts what it was sent into its table shard.
Counts are exact sums, so the union of the shards equals the single-GPU table whatever the world size or batching.

The exchange itself (`exchange_keys`) is device-agnostic and is what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

import math
import os
import sys
import time

import torch
import torch.distributed as dist


def shard_bits_for(world: int) -> int:
    if world < 1 or world & (world - 1):
        raise ValueError(f"world size must be a power of two (got {world}): owners are the top bits of the k-mer hash")
    return int(math.log2(world))


def owner_of(hashed_key: int, shard_bits: int) -> int:
    """Rank that owns a hashed key (62-bit): its top shard_bits bits."""
    return 0 if shard_bits == 0 else (hashed_key >> (62 - shard_bits)) & ((1 << shard_bits) - 1)


def plan_batches(nbytes: int, batch_bytes: int, record_bytes: int | None):
    """Record-aligned (offset, length) slices of a local read buffer.  With fixed-size records (the synthetic
    workload) slices are multiples of the record size; otherwise the caller must cut at record boundaries itself."""
    if nbytes == 0:
        return []
    if record_bytes:
        per = max(record_bytes, batch_bytes // record_bytes * record_bytes)
    else:
        per = max(1, batch_bytes)
    return [(off, min(per, nbytes - off)) for off in range(0, nbytes, per)]


def exchange_keys(send: torch.Tensor, send_counts, group=None):
    """All-to-all of owner-grouped hashed keys.  `send` (int64, 1-D) holds the keys grouped by destination rank,
    send_counts[r] of them for rank r.  Returns (recv tensor, recv_counts)."""
    world = dist.get_world_size(group)
    assert len(send_counts) == world
    dev = send.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send[: sum(send_counts)], output_split_sizes=recv_counts, input_split_sizes=[int(x) for x in send_counts], group=group)
    return recv, recv_counts


def gather_pairs(keys: torch.Tensor, counts: torch.Tensor, n: int, group=None):
    """All-gather of every rank's (key, count16) pairs (the solid set after the local coverage compaction).
    `keys` (int64) / `counts` (int16) hold this rank's `n` pairs in their first n elements.  Returns a list with one
    (keys, counts) pair of tensors per rank, on the tensors' device.  Device-agnostic: NCCL on GPUs, gloo in the CPU tests."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [(keys[:n], counts[:n])]
    sizes = torch.tensor([n], dtype=torch.int64, device=keys.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [int(x.item()) for x in all_sizes]
    nmax = max(max(all_sizes), 1)
    kbuf = torch.zeros(nmax, dtype=torch.int64, device=keys.device)
    cbuf = torch.zeros(nmax, dtype=torch.int16, device=keys.device)
    kbuf[:n] = keys[:n]
    cbuf[:n] = counts[:n]
    kall = [torch.empty_like(kbuf) for _ in range(world)]
    call = [torch.empty(2 * nmax, dtype=torch.uint8, device=keys.device) for _ in range(world)]
    dist.all_gather(kall, kbuf, group=group)
    dist.all_gather(call, cbuf.view(torch.uint8), group=group)   # neither NCCL nor gloo has an int16 type: the counts travel as bytes
    return [(kall[r][: all_sizes[r]], call[r].view(torch.int16)[: all_sizes[r]]) for r in range(world)]


def target_block(n_targets: int, world: int, rank: int):
    """Targets are independent units (the reference runs one OneSequenceCalculator per target on a pool,
    src/tools/EnvironmentFinderMain.java:188,218-225): rank r takes the contiguous block [lo, hi)."""
    return rank * n_targets // world, (rank + 1) * n_targets // world


class ShardedCounter:
    """Rank-local driver of the sharded counting step (GPU only; there is no CPU fallback)."""

    def __init__(self, ctx, k: int, world: int, rank: int, capacity_hint: int = 0, exchange: str = "peer"):
        """exchange = "peer": the scatter kernel writes every owner's keys straight into that owner's receive region over
        NVLink (CUDA-IPC mapped peer memory, csrc/mc_xchg.cu; ranks must share one node); "nccl": all_to_all_single;
        "merge": count locally, then ship the distinct (key, count) pairs to their owners (also peer memory) -- the choice
        when a rank's own k-mers fit its GPU: several times fewer bytes cross NVLink."""
        self.ctx, self.k, self.world, self.rank = ctx, k, world, rank
        self.shard_bits = shard_bits_for(world)
        self.capacity_hint = capacity_hint
        self.exchange = exchange if (world > 1 or exchange == "merge") else "local"   # "merge" also runs on one rank (profiling)
        self._xchg = None
        self._xchg_keys = 0
        self.table = None
        self.bytes_exchanged = 0
        # wall-clock seconds per phase of the last step (host view, each phase ends with a synchronisation)
        self.phase_s = {"extract": 0.0, "exchange": 0.0, "insert": 0.0, "sweep_filter": 0.0}
        self.verify = False
        self._local = None
        # ship finished hash ranges to their owners DURING the local table's last flush (mc_xchg_flush_and_dump).  Correct (parity
        # worker) but no faster at N = 2 (203.1 vs 203.3 ms per step: the range dumps take from the flush kernels the SM time they
        # use), so it is off unless asked for
        self.fuse_dump = os.environ.get("MC_FUSE_DUMP", "0") != "0"
        self.fused_last = None
        # NCCL sets a collective's channels up on its first call (0.7 s for the first all_gather at N = 2): done here, at construction,
        # so that the solid-set gather after the first counting step is the transfer and nothing else
        if world > 1 and dist.is_initialized() and dist.get_backend() == "nccl":
            tiny = torch.zeros(8, dtype=torch.uint8, device="cuda")
            dist.all_gather([torch.empty_like(tiny) for _ in range(world)], tiny)
            torch.cuda.synchronize()

    def new_table(self):
        if self.table is not None:
            self.table.free()
        self.table = self.ctx.new_table(self.k, capacity_hint=self.capacity_hint, shard_bits=self.shard_bits, shard_id=self.rank)
        return self.table

    def _buffers(self, ub: int):
        """Two send and two receive buffers, allocated once per size class (a fresh torch.empty of a few GB per batch costs
        a cudaMalloc each time the size changes, measured ~10 ms per batch)."""
        cap = getattr(self, "_buf_cap", 0)
        if cap < ub:
            self._send = [torch.empty(ub, dtype=torch.int64, device="cuda") for _ in range(2)]
            self._recv = [torch.empty(ub + ub // 4 + 1024, dtype=torch.int64, device="cuda") for _ in range(2)]
            self._buf_cap = ub
        return self._send, self._recv

    def _start_exchange(self, slot: int, counts):
        """Counts first (tiny all-to-all), then the keys, asynchronously on the communication stream."""
        send, recv = self._send[slot], self._recv[slot]
        with torch.cuda.stream(self._comm_stream):
            sc = torch.tensor(list(counts), dtype=torch.int64, device="cuda")
            rc = torch.empty(self.world, dtype=torch.int64, device="cuda")
            dist.all_to_all_single(rc, sc)
            recv_counts = [int(x) for x in rc.tolist()]
            total = sum(recv_counts)
            if total > recv.numel():   # unbalanced owners: make room (rare)
                self._recv[slot] = recv = torch.empty(total + total // 8, dtype=torch.int64, device="cuda")
            dist.all_to_all_single(recv[:total], send[: sum(counts)], output_split_sizes=recv_counts, input_split_sizes=[int(x) for x in counts])
            done = torch.cuda.Event()
            done.record(self._comm_stream)
        self.bytes_exchanged += 8 * (sum(counts) - counts[self.rank])
        return done, total

    # ---- peer-memory exchange ------------------------------------------------------------------------
    def _peer_setup(self, region_keys: int):
        import ctypes as C
        from ._ffi import check
        lib = self.ctx.lib
        if self._xchg is not None and self._xchg_keys >= region_keys:
            return
        self.close_exchange()
        h = C.c_void_p()
        check(lib.mc_xchg_create(self.ctx.h, self.world, self.rank, region_keys, C.byref(h)))
        buf = C.create_string_buffer(64)
        check(lib.mc_xchg_handle(h, buf))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf.raw))
            for r, raw in enumerate(handles):
                if r != self.rank:
                    check(lib.mc_xchg_open(h, r, C.create_string_buffer(raw, 64)))
        self._xchg, self._xchg_keys = h, region_keys
        if self.world > 1:
            dist.barrier()

    def close_exchange(self):
        if self._xchg is not None:
            if self.world > 1:
                dist.barrier()                   # nobody may still be writing into my regions
            self.ctx.lib.mc_xchg_destroy(self._xchg)
            self._xchg = None

    def _count_device_peer(self, d_reads: int, nbytes: int, fmt: int, batch_bytes: int, record_bytes):
        import ctypes as C
        from ._ffi import check
        lib = self.ctx.lib
        # regions are sized for the worst case of a batch, 2 slots x world regions of them: keep that near 16 GiB
        batch_bytes = min(batch_bytes, (2 << 30) // self.world)
        batches = plan_batches(nbytes, batch_bytes, record_bytes)
        nb = torch.tensor([len(batches), max([n for _, n in batches] + [1])], dtype=torch.int64, device="cuda")
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        n_batches, max_bytes = int(nb[0].item()), int(nb[1].item())
        self._peer_setup(int(lib.mc_max_kmers(max_bytes, fmt, self.k)))
        got = C.c_uint64()
        sync = torch.zeros(1, dtype=torch.int32, device="cuda")
        for b in range(n_batches):
            off, n = batches[b] if b < len(batches) else (0, 0)
            slot = b & 1
            t0 = time.perf_counter()
            check(lib.mc_xchg_scatter_device(self._xchg, C.c_void_p(d_reads + off), n, fmt, self.k, slot, self.table.h))
            t1 = time.perf_counter()
            dist.all_reduce(sync)                # every rank's stores into slot `slot` have completed
            torch.cuda.current_stream().synchronize()
            t2 = time.perf_counter()
            check(lib.mc_xchg_insert(self._xchg, slot, self.table.h, C.byref(got)))
            if b == 0:   # what the first batch brought, times the batches: sizes the shard once (no capacity hint needed)
                self.table.expect_keys(got.value * n_batches)
            t3 = time.perf_counter()
            self.phase_s["extract"] += t1 - t0
            self.phase_s["exchange"] += t2 - t1
            self.phase_s["insert"] += t3 - t2
            self.bytes_exchanged += 8 * got.value
        self.ctx.sync()

    # ---- "merge" exchange: count locally (any number of batches, device- or host-resident), then ship distinct pairs ----------
    def local_begin(self):
        """Start this rank's complete LOCAL table (the single-GPU pipeline); batches are added with local_add_*."""
        self._t_begin = time.perf_counter()
        self._local = self.ctx.new_table(self.k, capacity_hint=self.capacity_hint)
        if self.fuse_dump and self._xchg is not None:
            self.ctx.set_tuning("defer_flush", 1)   # the last batch of bins stays pending: local_finish flushes AND dumps it
        return self._local

    def local_add_device(self, d_reads: int, nbytes: int, fmt: int):
        if nbytes:
            self._local.count_device(d_reads, nbytes, fmt)

    def local_add_host(self, host_ptr: int, nbytes: int, fmt: int):
        """The same from HOST memory (mc_count_bytes: chunked uploads overlapped with the counting): the e2e leg at N > 1."""
        if nbytes:
            self._local.count_host_ptr(host_ptr, nbytes, fmt)

    def local_finish(self):
        """Count locally, then merge (include/mc_b200.h, mc_xchg_dump_owners): the table scan stores every DISTINCT
        (key, count) pair into its owner's region over NVLink, and each rank adds what it received into its shard.
        One barrier, no collective moves data.  With self.verify set, the digests of the local table (before) and of the
        shard (after) are kept: summed over the ranks they must agree (bench.py --verify, tests/dist_parity_worker.py)."""
        import ctypes as C
        from ._ffi import check
        lib = self.ctx.lib
        local = self._local
        self._local = None
        got = C.c_uint64()
        fused = C.c_int(0)
        t1 = time.perf_counter()
        dumped = False
        if self.fuse_dump and self._xchg is not None:
            # The dump rides on the local table's last flush (mc_xchg_flush_and_dump): regions keep the size of the previous step;
            # a rank whose table outgrew them reports -1 and every rank then falls back to the sized, unfused dump below.
            self.ctx.set_tuning("defer_flush", 0)
            check(lib.mc_xchg_flush_and_dump(self._xchg, local.h, 0, C.byref(got), C.byref(fused)))
            bad = torch.tensor([1 if fused.value < 0 else 0], dtype=torch.int32, device="cuda")
            if self.world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            dumped = int(bad.item()) == 0
            self.fused_last = fused.value
            if dumped and getattr(self, "verify", False):
                self.local_digest = local.checksum(0)   # (the local table is complete once its last flush has run)
        self.ctx.set_tuning("defer_flush", 0)
        if not dumped:
            n_local = local.size()
            t1 = time.perf_counter()
            if getattr(self, "verify", False):
                self.local_digest = local.checksum(0)
            mx = torch.tensor([n_local], dtype=torch.int64, device="cuda")
            if self.world > 1:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            # an owner's share of a table is binomial around size / world: 25 % + a constant is far outside its spread
            pairs = int(mx.item()) // self.world
            pairs = pairs + pairs // 4 + (1 << 16)
            self._peer_setup(2 * pairs)
            check(lib.mc_xchg_dump_owners(self._xchg, local.h, 0, C.byref(got)))
        t2 = time.perf_counter()
        if self.world > 1:
            sync = torch.zeros(1, dtype=torch.int32, device="cuda")
            dist.all_reduce(sync)                # every rank's stores into my regions have completed
            torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        check(lib.mc_xchg_merge(self._xchg, 0, self.table.h, C.byref(got)))
        check(lib.mc_table_copy_stats(self.table.h, local.h))
        local.free()      # after the merge: the block goes back to the context's arena for the next step
        self.ctx.sync()
        if self.world > 1:
            dist.barrier()                       # nobody reuses its regions (next step's dump) before every merge has read them
        t4 = time.perf_counter()
        if getattr(self, "verify", False):
            self.shard_digest = self.table.checksum(0)
        self.bytes_exchanged += 10 * got.value * (self.world - 1) // self.world
        self.phase_s["extract"] += t1 - self._t_begin
        self.phase_s["dump"] = self.phase_s.get("dump", 0.0) + (t2 - t1)
        self.phase_s["barrier"] = self.phase_s.get("barrier", 0.0) + (t3 - t2)
        self.phase_s["exchange"] += t3 - t1
        self.phase_s["insert"] += t4 - t3

    def _count_device_merge(self, d_reads: int, nbytes: int, fmt: int):
        self.local_begin()
        self.local_add_device(d_reads, nbytes, fmt)
        self.local_finish()

    def verify_exchange(self):
        """Sum over the ranks of the local tables' (sum of counts, digest) == sum over the ranks of the shards': nothing was
        lost, duplicated or re-keyed on the way to the owners (valid while no count saturates: max_count < 32767 is checked).
        Returns a dict for the bench line; raises on a mismatch."""
        a, b = self.local_digest, self.shard_digest
        import numpy as np
        v = torch.from_numpy(np.array([a["sum_counts"], a["digest"], b["sum_counts"], b["digest"], b["entries"]], dtype=np.uint64).view(np.int64)).cuda()
        mxc = torch.tensor([b["max_count"]], dtype=torch.int64, device="cuda")
        if self.world > 1:
            dist.all_reduce(v)                   # int64 wrap-around = arithmetic mod 2^64, like the digest itself
            dist.all_reduce(mxc, op=dist.ReduceOp.MAX)
        v = [int(x) & 0xFFFFFFFFFFFFFFFF for x in v.cpu().tolist()]
        saturated = int(mxc.item()) >= 32767
        ok = v[0] == v[2] and v[1] == v[3]
        res = {"ok": bool(ok or saturated), "sum_counts": v[2], "digest": f"{v[3]:016x}", "distinct": v[4], "max_count": int(mxc.item()),
               "what": "sum over ranks of the local tables' (sum of counts, key digest) == sum over ranks of the owner shards'"
                       + (" -- NOT CHECKED: a count saturated at 32767, sums are no longer additive" if saturated else "")}
        if not res["ok"]:
            raise AssertionError(f"multi-GPU exchange lost or duplicated k-mers: local {v[0]} / {v[1]:016x} vs shards {v[2]} / {v[3]:016x}")
        return res

    def resolve_format(self, d_reads: int, nbytes: int, fmt: int) -> int:
        """FASTQ with an unknown quality offset: the reference decides Phred+33 / +64 ONCE per file from its first 1000
        records (ITMO!/io/ReadersUtils.java:63-77).  Rank 0 holds the head of the input: it sniffs and broadcasts, and every
        batch of every rank is then parsed with the explicit format (a later batch sniffed on its own could decide otherwise)."""
        if fmt != 1:   # MC_FMT_FASTQ (auto)
            return fmt
        import numpy as np
        off = torch.zeros(1, dtype=torch.int64)
        if self.rank == 0:
            n = min(nbytes, 1 << 20)
            while True:
                head = np.empty(n, dtype=np.uint8)
                if n:
                    self.ctx.d2h(head, d_reads)
                # -1: fewer than 1000 records in the sample and more input exists
                r = int(self.ctx.lib.mc_sniff_quality_offset(head.ctypes.data if n else None, n)) if n == nbytes else self._sniff_partial(head)
                if r > 0 or n == nbytes:
                    break
                n = min(nbytes, n * 4)
            off[0] = r if r > 0 else 64
        if self.world > 1:
            dev = off.cuda() if dist.get_backend() == "nccl" else off
            dist.broadcast(dev, src=0)
            off = dev.cpu()
        return 33 if int(off[0]) == 33 else 64

    def _sniff_partial(self, head) -> int:
        """A head that may end inside a record is only trusted when it holds well over 1000 records (the sniffer stops
        after 1000 and never reaches the cut); otherwise -1 = ask for a larger sample."""
        if int((head == 10).sum()) < 8008:   # 1000 records of 4 lines, twice over for blank lines
            return -1
        return int(self.ctx.lib.mc_sniff_quality_offset(head.ctypes.data, head.size))

    def count_device(self, d_reads: int, nbytes: int, fmt: int, batch_bytes: int, record_bytes: int | None = None):
        """Counts this rank's device-resident reads into the distributed table (collective: every rank must call it
        with the same number of batches; ranks with fewer batches send empty groups).
        Software pipeline: while batch b's keys cross NVLink (communication stream), batch b+1 is parsed and extracted and
        batch b-1's received keys are binned (library stream)."""
        fmt = self.resolve_format(d_reads, nbytes, fmt)
        if self.exchange == "peer":
            return self._count_device_peer(d_reads, nbytes, fmt, batch_bytes, record_bytes)
        if self.exchange == "merge":
            return self._count_device_merge(d_reads, nbytes, fmt)
        lib = self.ctx.lib
        batches = plan_batches(nbytes, batch_bytes, record_bytes)
        n_batches = torch.tensor([len(batches)], dtype=torch.int64, device="cuda")
        if self.world > 1:
            dist.all_reduce(n_batches, op=dist.ReduceOp.MAX)
        n_batches = int(n_batches.item())
        ub = max([int(lib.mc_max_kmers(n, fmt, self.k)) for _, n in batches] + [1])
        send, recv = self._buffers(ub)
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream()
        pending = None   # (event, slot, n_received) of the exchange in flight
        self._expect_set = False

        def finish(p):
            ev, slot, total = p
            t0 = time.perf_counter()
            ev.synchronize()
            t1 = time.perf_counter()
            if total:
                if not self._expect_set:
                    self.table.expect_keys(total * n_batches)
                    self._expect_set = True
                self.table.insert_hashed_device(self._recv[slot].data_ptr(), total)
            self.phase_s["exchange"] += t1 - t0
            self.phase_s["insert"] += time.perf_counter() - t1

        for b in range(n_batches):
            off, n = batches[b] if b < len(batches) else (0, 0)
            slot = b & 1
            t0 = time.perf_counter()
            counts = self.table.extract_partition_device(d_reads + off, n, fmt, self.shard_bits, send[slot].data_ptr()) if n else [0] * self.world
            self.phase_s["extract"] += time.perf_counter() - t0
            if self.world == 1:
                if counts[0]:
                    t0 = time.perf_counter()
                    self.table.insert_hashed_device(send[slot].data_ptr(), counts[0])
                    self.phase_s["insert"] += time.perf_counter() - t0
                continue
            if pending is not None:
                finish(pending)
            t0 = time.perf_counter()
            ev, total = self._start_exchange(slot, counts)
            self.phase_s["exchange"] += time.perf_counter() - t0
            pending = (ev, slot, total)
        if pending is not None:
            finish(pending)
        self.ctx.sync()

    def gather_solid(self, solid_shard):
        """E3: every rank ends up with the COMPLETE coverage-filtered table.  The local compaction already happened
        (solid_shard = this rank's slice); the (key, count16) pairs are dumped into device buffers, all-gathered over
        NVLink and inserted into a fresh unsharded table.  Returns (table, pairs per rank)."""
        n = solid_shard.dump_device(1)
        # The gather's buffers come from torch's allocator, the counting step's blocks (bins, staging: tens of GB) sit in the
        # context's arena cache: when the driver has little left, torch's cudaMalloc fails first, empties its own cache and retries
        # (0.7 s at N = 2).  Counting is over here, so the arena hands its cache back when room is short.
        prof = os.environ.get("MC_DIST_PROFILE")
        t0 = time.perf_counter()
        need = (2 * self.world + 4) * 10 * max(n, 1) * 2
        free_before = torch.cuda.mem_get_info()[0]
        if free_before < need:
            self.ctx.trim()
        keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
        counts = torch.empty(max(n, 1), dtype=torch.int16, device="cuda")
        if n:
            solid_shard.dump_device(1, keys.data_ptr(), counts.data_ptr(), n)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        parts = gather_pairs(keys, counts, n)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        total = sum(int(k.numel()) for k, _ in parts)
        full = self.ctx.new_table(self.k, capacity_hint=2 * total)
        for k_r, c_r in parts:
            if k_r.numel():
                kk, cc = k_r.contiguous(), c_r.contiguous()
                full.add_pairs_device(kk.data_ptr(), cc.data_ptr(), int(kk.numel()))
        if prof:
            self.ctx.sync()
            t3 = time.perf_counter()
            print(f"[dist profile] rank {self.rank} gather_solid: free {free_before / 2**30:.1f} GiB (need {need / 2**30:.1f}), dump {1e3 * (t1 - t0):.1f} ms, "
                  f"all-gather {1e3 * (t2 - t1):.1f} ms, insert {1e3 * (t3 - t2):.1f} ms", file=sys.stderr)
        self.bytes_exchanged += 10 * (total - n)
        return full, [int(k.numel()) for k, _ in parts]

    def find_environments(self, solid_full, seqs, names, out_dir, **kw):
        """E4: the targets are block-distributed over the ranks (no communication: every rank holds the complete solid
        table); each rank writes the output directories of ITS targets under the shared out_dir.  Returns found[] for all
        targets (gathered)."""
        lo, hi = target_block(len(seqs), self.world, self.rank)
        found = self.ctx.find_environments(solid_full, list(seqs[lo:hi]), list(names[lo:hi]), out_dir=out_dir, **kw) if hi > lo else []
        if self.world == 1:
            return list(found)
        parts = [None] * self.world
        dist.all_gather_object(parts, [int(x) for x in found])
        return [x for p in parts for x in p]

    def step_begin(self):
        self.ctx.stage_reset()
        for name in self.phase_s:
            self.phase_s[name] = 0.0
        self.new_table()

    def step_finish(self, min_cov: int, keep_solid: bool = False):
        """Local coverage compaction of the shard + the global solid count (all-reduce).  Returns (global solid k-mers,
        stage times) (+ the solid shard itself when keep_solid: the caller gathers it for the environment stage and frees it)."""
        t0 = time.perf_counter()
        solid = self.table.filter(min_cov)
        n = torch.tensor([solid.size()], dtype=torch.int64, device="cuda")
        self.phase_s["sweep_filter"] = time.perf_counter() - t0
        if self.world > 1:
            dist.all_reduce(n)
        stage = self.ctx.stage_ms()
        if os.environ.get("MC_DIST_PROFILE"):
            print(f"[rank {self.rank}] phases (s): {({k: round(v, 4) for k, v in self.phase_s.items()})} stages (ms): {({k: round(v, 1) for k, v in stage.items()})}", flush=True)
        if keep_solid:
            return int(n.item()), stage, solid
        solid.free()
        return int(n.item()), stage

    def count_device_step(self, d_reads: int, nbytes: int, fmt: int, min_cov: int, batch_bytes: int, record_bytes: int | None = None, keep_solid: bool = False):
        """One bench step: fresh shard, count, coverage compaction."""
        self.step_begin()
        self.count_device(d_reads, nbytes, fmt, batch_bytes, record_bytes)
        return self.step_finish(min_cov, keep_solid)
