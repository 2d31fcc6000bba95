"""Python host layer over libmcgpu: handles + the reference-shaped entry points.

Names follow the reference's domain (reads, k-mers, tables, environments), not ML vocabulary:
  Context.load_reads(files, k)        ~ IOUtils.loadReads(File[] files, int k, ...)        src/io/IOUtils.java:217
  KmerTable.get(keys)                 ~ BigLong2ShortHashMap.get(long)                     Long2ShortHashMap.java:159
  KmerTable.size()                    ~ BigLong2ShortHashMap.size()                        "Hashtable size: N kmers"
  Context.find_environments(...)      ~ new OneSequenceCalculator(...).run() per target    EnvironmentFinderMain.java:218-225
  Context.multi_environment(...)      ~ new MultiSequenceCalculator(...).run()             EnvironmentFinderMultiMain.java:81-91
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import McError, check

FMT_FASTA, FMT_FASTQ, FMT_FASTQ33, FMT_FASTQ64 = 0, 1, 33, 64


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


class Context:
    """One per GPU (mc_ctx): device, streams, staging buffers."""

    def __init__(self, device: int = 0, chunk_bytes: int | None = None):
        self.lib = _ffi.load_library()
        h = C.c_void_p()
        check(self.lib.mc_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if chunk_bytes:
            check(self.lib.mc_ctx_set_chunk_bytes(self.h, chunk_bytes))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunk_bytes(self, n):
        check(self.lib.mc_ctx_set_chunk_bytes(self.h, n))

    @property
    def launches(self) -> int:
        return int(self.lib.mc_ctx_launches(self.h))

    def trim(self):
        """Return the context's cached device blocks to the driver (mc_ctx_trim)."""
        check(self.lib.mc_ctx_trim(self.h))

    def arena_stats(self):
        """(bytes cached for reuse, driver allocations made so far)."""
        a, b = C.c_uint64(), C.c_uint64()
        check(self.lib.mc_ctx_arena_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sync(self):
        check(self.lib.mc_ctx_sync(self.h))

    STAGES = ("parse", "extract", "grow", "scatter", "count", "filter")

    def stage_ms(self):
        """Device milliseconds per stage since the start of the last count call (CUDA events on the library's stream)."""
        v = (C.c_float * 8)()
        check(self.lib.mc_ctx_stage_ms_all(self.h, v))
        d = {name: v[i] for i, name in enumerate(self.STAGES)}
        d["expand"] = v[7]   # super-k-mer records -> hashed keys (mc_sk.cu)
        return d

    def host_ms(self):
        """Host wall milliseconds inside the same phases since the last reset (launch + synchronisation time included)."""
        v = (C.c_double * 8)()
        check(self.lib.mc_ctx_host_ms(self.h, v))
        d = {name: v[i] for i, name in enumerate(self.STAGES)}
        d["finish"] = v[6]
        return d

    def stage_reset(self):
        check(self.lib.mc_ctx_stage_reset(self.h))

    def set_tuning(self, name: str, value: int):
        """Knobs of the binned update path: part_bits, batch_keys, segment_keys, direct (see include/mc_b200.h)."""
        check(self.lib.mc_ctx_set_tuning(self.h, name.encode(), value))

    def event_record(self, slot):
        check(self.lib.mc_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b) -> float:
        ms = C.c_float()
        check(self.lib.mc_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    # ---- Seam 1 --------------------------------------------------------------------------
    def new_table(self, k, capacity_hint=0, shard_bits=0, shard_id=0) -> "KmerTable":
        h = C.c_void_p()
        check(self.lib.mc_table_create(self.h, k, capacity_hint, shard_bits, shard_id, C.byref(h)))
        return KmerTable(self, h, k)

    def load_reads(self, files, k) -> "KmerTable":
        """IOUtils.loadReads: counts every canonical k-mer of the given FASTA/FASTQ files."""
        arr = (C.c_char_p * len(files))(*[os.fspath(f).encode() for f in files])
        h = C.c_void_p()
        check(self.lib.mc_count_files(self.h, arr, len(files), k, C.byref(h)))
        return KmerTable(self, h, k)

    # ---- device memory (bench / multi-GPU plumbing) ----------------------------------------
    def device_alloc(self, nbytes) -> int:
        p = C.c_void_p()
        check(self.lib.mc_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr):
        check(self.lib.mc_device_free(self.h, C.c_void_p(ptr)))

    def h2d(self, d_ptr, a: np.ndarray):
        check(self.lib.mc_memcpy_h2d(self.h, C.c_void_p(d_ptr), a.ctypes.data, a.nbytes))

    def d2h(self, a: np.ndarray, d_ptr):
        check(self.lib.mc_memcpy_d2h(self.h, a.ctypes.data, C.c_void_p(d_ptr), a.nbytes))

    def synth_generate_device(self, params: bytes, first, count, d_out):
        check(self.lib.mc_synth_generate_device(self.h, params, len(params), first, count, C.c_void_p(d_out)))

    # ---- Seam 3 / 4 ------------------------------------------------------------------------
    def find_environments(self, table: "KmerTable", seqs, names, min_cov=1, max_kmers=None, max_radius=None, both_dirs=False,
                          chunk_len=1, out_dir="."):
        n = len(seqs)
        a_seqs = (C.c_char_p * n)(*[s.encode() for s in seqs])
        a_names = (C.c_char_p * n)(*[s.encode() for s in names])
        found = (C.c_int * n)()
        check(self.lib.mc_env_find(self.h, table.h, a_seqs, a_names, n, table.k, min_cov, -1 if max_kmers is None else max_kmers,
                                   -1 if max_radius is None else max_radius, int(both_dirs), chunk_len, os.fspath(out_dir).encode(), found))
        return [bool(x) for x in found]

    def extend_counts(self):
        """"Extending endings by N kmers" of every environment of the last find_environments call (-1: none was built)."""
        n = C.c_int()
        check(self.lib.mc_env_extend_counts(self.h, None, 0, C.byref(n)))
        out = (C.c_longlong * max(n.value, 1))()
        check(self.lib.mc_env_extend_counts(self.h, out, n.value, C.byref(n)))
        return [int(out[i]) for i in range(n.value)]

    def find_environments_ex(self, table: "KmerTable", seqs, names=None, hic_seqs=(), min_cov=1, max_kmers=None, max_radius=None,
                             both_dirs=False, chunk_len=1, merge=False, trim=False, out_dir="."):
        """environment-finder with --merge / --hicseq / --trim (mc_env_find_ex)."""
        n, nh = len(seqs), len(hic_seqs)
        a_seqs = (C.c_char_p * n)(*[s.encode() for s in seqs])
        a_names = (C.c_char_p * n)(*[s.encode() for s in names]) if names is not None else None
        a_hic = (C.c_char_p * max(nh, 1))(*[s.encode() for s in hic_seqs]) if nh else None
        found = (C.c_int * max(n, 1))()
        check(self.lib.mc_env_find_ex(self.h, table.h, a_seqs, a_names, n, a_hic, nh, table.k, min_cov, -1 if max_kmers is None else max_kmers,
                                      -1 if max_radius is None else max_radius, int(both_dirs), chunk_len, int(merge), int(trim),
                                      os.fspath(out_dir).encode(), found))
        return [bool(found[i]) for i in range(1 if merge else n)]

    # ---- "next" row f-4b: reads-classifier -------------------------------------------------------------
    def classify_reads(self, graph: "KmerTable", data, fmt=FMT_FASTQ, z=1.0, found_threshold=0.9, correction=False) -> np.ndarray:
        """found[i] of every read of a FASTQ buffer (ReadsFinderInGraph.findRead / findReadWithCorrection,
        src/algo/ReadsFinderInGraph.java:37-123) against the k-mer table `graph`."""
        a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        n = C.c_size_t()
        check(self.lib.mc_classify_bytes(self.h, graph.h, _ptr(a), a.size, fmt, z, found_threshold, int(correction), None, 0, C.byref(n)))
        found = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            check(self.lib.mc_classify_bytes(self.h, graph.h, _ptr(a), a.size, fmt, z, found_threshold, int(correction), _ptr(found), found.size, C.byref(n)))
        return found

    def env_from_graph_txt(self, graph_txt, target_seq, k, chunk_len, out_dir):
        check(self.lib.mc_env_from_graph_txt(self.h, os.fspath(graph_txt).encode(), target_seq.encode(), k, chunk_len, os.fspath(out_dir).encode()))

    def load_targets(self, fasta_path):
        """RichFastaReader: [(name, sequence)] of a target FASTA."""
        h = C.c_void_p()
        check(self.lib.mc_targets_load(os.fspath(fasta_path).encode(), C.byref(h)))
        try:
            n = self.lib.mc_targets_count(h)
            out = []
            for i in range(n):
                name = self.lib.mc_targets_name(h, i)
                if name is None:
                    raise McError(-4, f"target {i} has no header entry (IndexOutOfBoundsException in the reference)")
                out.append((name.decode(), self.lib.mc_targets_seq(h, i).decode()))
            return out
        finally:
            self.lib.mc_targets_free(h)

    def multi_environment(self, env_paths, seq_fasta, gene_id, out_dir):
        arr = (C.c_char_p * len(env_paths))(*[os.fspath(p).encode() for p in env_paths])
        check(self.lib.mc_env_multi(self.h, arr, len(env_paths), os.fspath(seq_fasta).encode(), gene_id, os.fspath(out_dir).encode()))


class KmerTable:
    """The k-mer count map in HBM (mc_table) — stands where BigLong2ShortHashMap stands in the reference."""

    def __init__(self, ctx: Context, handle, k):
        self.ctx, self.h, self.k = ctx, handle, k
        self.lib = ctx.lib

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.lib.mc_table_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def count_bytes(self, data, fmt) -> "KmerTable":
        """Counts a HOST buffer of whole records (bytes / numpy uint8 / anything with the buffer protocol)."""
        a = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        check(self.lib.mc_count_bytes(self.ctx.h, _ptr(a), a.size, fmt, self.h))
        return self

    def count_host_ptr(self, ptr, nbytes, fmt) -> "KmerTable":
        check(self.lib.mc_count_bytes(self.ctx.h, C.c_void_p(ptr), nbytes, fmt, self.h))
        return self

    def count_device(self, d_ptr, nbytes, fmt) -> "KmerTable":
        check(self.lib.mc_count_device(self.ctx.h, C.c_void_p(d_ptr), nbytes, fmt, self.h))
        return self

    def size(self) -> int:
        return int(self.lib.mc_table_size(self.h))

    def stats(self):
        v = [C.c_uint64() for _ in range(4)]
        check(self.lib.mc_table_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("reads", "runs", "bases", "kmers"), [x.value for x in v]))

    def checksum(self, min_count=0):
        """Device-side digest of the entries with count >= min_count: dict(entries, sum_counts, digest, max_count)."""
        out = (C.c_uint64 * 4)()
        check(self.lib.mc_table_checksum(self.h, min_count, out))
        return dict(zip(("entries", "sum_counts", "digest", "max_count"), [int(x) for x in out]))

    def expect_keys(self, n: int):
        """About n more k-mer instances will be added (sizes the table for a stream of batches); 0 = unknown."""
        check(self.lib.mc_table_expect_keys(self.h, n))

    def add_pairs(self, keys, counts):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.int16)
        check(self.lib.mc_table_add_pairs(self.h, _ptr(keys), _ptr(counts), keys.size))

    def get(self, keys) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(keys.size, dtype=np.int16)
        check(self.lib.mc_table_get(self.h, _ptr(keys), keys.size, _ptr(out)))
        return out

    def dump(self, min_count=0):
        n = C.c_size_t()
        check(self.lib.mc_table_dump(self.h, min_count, None, None, 0, C.byref(n)))
        keys = np.empty(n.value, dtype=np.uint64)
        counts = np.empty(n.value, dtype=np.int16)
        if n.value:
            check(self.lib.mc_table_dump(self.h, min_count, keys.ctypes.data, counts.ctypes.data, n.value, C.byref(n)))
        return keys, counts

    def dump_device(self, min_count, d_keys=0, d_counts=0, cap=0) -> int:
        """Entries with count >= min_count into DEVICE buffers (unsorted); cap = 0 only counts them."""
        n = C.c_size_t()
        check(self.lib.mc_table_dump_device(self.h, min_count, C.c_void_p(d_keys or None), C.c_void_p(d_counts or None), cap, C.byref(n)))
        return n.value

    def add_pairs_device(self, d_keys, d_counts, n):
        check(self.lib.mc_table_add_pairs_device(self.h, C.c_void_p(d_keys), C.c_void_p(d_counts), n))

    def filter(self, min_count) -> "KmerTable":
        h = C.c_void_p()
        check(self.lib.mc_table_filter(self.ctx.h, self.h, min_count, C.byref(h)))
        return KmerTable(self.ctx, h, self.k)

    def bfs(self, seq, min_cov, max_kmers, max_radius, direction, cap=1 << 20):
        kmers = np.empty(cap, dtype=np.uint64)
        dist = np.empty(cap, dtype=np.int32)
        counts = np.empty(cap, dtype=np.int16)
        n = C.c_size_t()
        check(self.lib.mc_bfs(self.ctx.h, self.h, seq.encode(), self.k, min_cov, -1 if max_kmers is None else max_kmers,
                              -1 if max_radius is None else max_radius, direction, kmers.ctypes.data, dist.ctypes.data,
                              counts.ctypes.data, cap, C.byref(n)))
        m = n.value
        return kmers[:m].copy(), dist[:m].copy(), counts[:m].copy()

    def env_kmers(self, seq, min_cov, max_kmers, max_radius, both_dirs, cap=1 << 21):
        kmers = np.empty(cap, dtype=np.uint64)
        counts = np.empty(cap, dtype=np.int32)
        n = C.c_size_t()
        check(self.lib.mc_env_kmers(self.ctx.h, self.h, seq.encode(), self.k, min_cov, -1 if max_kmers is None else max_kmers,
                                    -1 if max_radius is None else max_radius, int(both_dirs), kmers.ctypes.data, counts.ctypes.data,
                                    cap, C.byref(n)))
        return kmers[:n.value].copy(), counts[:n.value].copy()

    def print_kmers(self, threshold, out_file, stat_file=None) -> int:
        good = C.c_uint64()
        check(self.lib.mc_table_print_kmers(self.h, threshold, os.fspath(out_file).encode(),
                                            os.fspath(stat_file).encode() if stat_file else None, C.byref(good)))
        return good.value

    def load_kmers(self, path, freq_threshold=0):
        check(self.lib.mc_table_load_kmers(self.h, os.fspath(path).encode(), freq_threshold))

    # multi-GPU building blocks
    def extract_partition_device(self, d_ptr, nbytes, fmt, shard_bits, d_keys):
        counts = (C.c_uint64 * (1 << shard_bits))()
        check(self.lib.mc_extract_partition_device(self.ctx.h, C.c_void_p(d_ptr), nbytes, fmt, self.k, shard_bits, C.c_void_p(d_keys),
                                                   counts, self.h))
        return [int(x) for x in counts]

    def insert_hashed_device(self, d_hkeys, n):
        check(self.lib.mc_table_insert_hashed_device(self.h, C.c_void_p(d_hkeys), n))
