"""ctypes binding of oracle/libmc_oracle.so — TEST INFRASTRUCTURE (the checker, never the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libmc_oracle.so")

FMT_FASTA, FMT_FASTQ, FMT_FASTQ33, FMT_FASTQ64 = 0, 1, 33, 64

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i16p = C.POINTER(C.c_int16)
_i32p = C.POINTER(C.c_int32)


class OracleError(RuntimeError):
    pass


class SynthParams(C.Structure):
    # opaque blob sized at load time
    pass


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_detect_format.argtypes = [C.c_char_p]
        L.orc_table_new.restype = C.c_void_p
        L.orc_table_new.argtypes = [C.c_int, C.c_int]
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_count_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_count_file.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_table_add_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_table_size.restype = C.c_uint64
        L.orc_table_size.argtypes = [C.c_void_p]
        L.orc_table_get.restype = C.c_int16
        L.orc_table_get.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_table_get_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_table_dump.restype = C.c_uint64
        L.orc_table_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, _u64p]
        L.orc_table_checksum.argtypes = [C.c_void_p, C.c_int, _u64p]
        L.orc_table_stats.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_canonical.restype = C.c_uint64
        L.orc_canonical.argtypes = [C.c_uint64, C.c_int]
        L.orc_encode.restype = C.c_uint64
        L.orc_encode.argtypes = [C.c_char_p, C.c_int]
        L.orc_decode.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.orc_java_string_hash.restype = C.c_int32
        L.orc_java_string_hash.argtypes = [C.c_char_p]
        L.orc_java_hashmap_order.restype = None
        L.orc_java_hashmap_order.argtypes = [C.POINTER(C.c_char_p), C.c_size_t, C.c_void_p]
        L.orc_bfs.restype = C.c_int64
        L.orc_bfs.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64]
        L.orc_env_find.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_char_p]
        L.orc_env_find_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.orc_env_from_graph_txt.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.orc_env_multi.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int, C.c_char_p]
        L.orc_print_kmers.restype = C.c_int64
        L.orc_print_kmers.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
        L.orc_load_kmers.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_bench_count.restype = C.c_double
        L.orc_bench_count.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, _u64p, _u64p]
        L.orc_classify_bytes.restype = C.c_long
        L.orc_classify_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_reads_classifier.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_int, C.c_char_p, C.POINTER(C.c_long)]
        L.orc_env_extend_counts.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int]
        L.orc_bench_env.restype = C.c_double
        L.orc_bench_env.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                    C.POINTER(C.c_int)]
        L.orc_synth_params_size.restype = C.c_size_t
        L.orc_synth_init.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_synth_record_bytes.restype = C.c_uint32
        L.orc_synth_record_bytes.argtypes = [C.c_void_p]
        L.orc_synth_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        L.orc_synth_target.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p]
        L.orc_synth_n_plants.restype = C.c_uint32
        L.orc_synth_n_plants.argtypes = [C.c_void_p]

    def err(self):
        return (self.lib.orc_last_error() or b"").decode("utf-8", "replace")

    def check(self, rc):
        if rc < 0:
            raise OracleError(self.err())
        return rc

    # ---- table -------------------------------------------------------------------------
    def table(self, k, threads=1):
        return OracleTable(self, k, threads)

    def revcomp(self, kmer, k):
        return self.lib.orc_revcomp(kmer, k)

    def canonical(self, kmer, k):
        return self.lib.orc_canonical(kmer, k)

    def encode(self, s):
        return self.lib.orc_encode(s.encode(), len(s))

    def decode(self, kmer, k):
        buf = C.create_string_buffer(k + 1)
        self.lib.orc_decode(kmer, k, buf)
        return buf.value.decode()

    def java_hash(self, s):
        return self.lib.orc_java_string_hash(s.encode())

    def java_hashmap_order(self, keys):
        """Iteration order (as indices into `keys`, distinct strings) of a java.util.HashMap filled in that order."""
        import numpy as np
        arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        out = np.empty(len(keys), dtype=np.uint64)
        self.lib.orc_java_hashmap_order(arr, len(keys), out.ctypes.data)
        return [int(x) for x in out]

    def detect_format(self, path):
        return self.check(self.lib.orc_detect_format(path.encode()))

    def env_from_graph_txt(self, graph_txt, target_seq, k, chunk_len, out_dir):
        return self.check(self.lib.orc_env_from_graph_txt(graph_txt.encode(), target_seq.encode(), k, chunk_len, out_dir.encode()))

    def env_multi(self, env_paths, seq_fasta, gene_id, out_dir):
        arr = (C.c_char_p * len(env_paths))(*[p.encode() for p in env_paths])
        return self.check(self.lib.orc_env_multi(arr, len(env_paths), seq_fasta.encode(), gene_id, out_dir.encode()))

    def bench_count(self, buf: np.ndarray, fmt, k, threads):
        bases, distinct = C.c_uint64(), C.c_uint64()
        secs = self.lib.orc_bench_count(buf.ctypes.data, buf.size, fmt, k, threads, C.byref(bases), C.byref(distinct))
        if secs < 0:
            raise OracleError(self.err())
        return secs, bases.value, distinct.value

    # ---- synthetic workload ------------------------------------------------------------
    def bench_env(self, table, seq_fasta, k, min_cov, max_kmers, max_radius, both_dirs, chunk_len, threads, out_dir):
        """Wall seconds of the environment stage on `threads` pool workers (bench.py's cpu_baseline.env leg) and the targets found."""
        done = C.c_int(0)
        secs = self.lib.orc_bench_env(table.h, seq_fasta.encode(), k, min_cov, -1 if max_kmers is None else max_kmers,
                                      -1 if max_radius is None else max_radius, int(both_dirs), chunk_len, threads, out_dir.encode(), C.byref(done))
        if secs < 0:
            raise OracleError(self.err())
        return secs, done.value

    def synth(self, **kw):
        return Synth(self, **kw)


class OracleTable:
    def __init__(self, o: Oracle, k, threads):
        self.o, self.k = o, k
        self.h = o.lib.orc_table_new(k, threads)
        if not self.h:
            raise OracleError(o.err())

    def __del__(self):
        if getattr(self, "h", None):
            self.o.lib.orc_table_free(self.h)
            self.h = None

    def count_bytes(self, data, fmt):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        self.o.check(self.o.lib.orc_count_bytes(self.h, buf.ctypes.data if buf.size else None, buf.size, fmt))
        return self

    def count_file(self, path):
        self.o.check(self.o.lib.orc_count_file(self.h, path.encode()))
        return self

    def add_pairs(self, keys, counts):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.int16)
        self.o.lib.orc_table_add_pairs(self.h, keys.ctypes.data, counts.ctypes.data, keys.size)

    def size(self):
        return self.o.lib.orc_table_size(self.h)

    def get(self, key):
        return self.o.lib.orc_table_get(self.h, key)

    def get_many(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(keys.size, dtype=np.int16)
        self.o.lib.orc_table_get_many(self.h, keys.ctypes.data, keys.size, out.ctypes.data)
        return out

    def dump(self, min_count=-32768):
        n = self.size()
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.int16)
        total = C.c_uint64()
        m = self.o.lib.orc_table_dump(self.h, min_count, keys.ctypes.data, counts.ctypes.data, n, C.byref(total))
        return keys[:m].copy(), counts[:m].copy()

    def checksum(self, min_count=0):
        """dict(entries, sum_counts, digest, max_count): the digest of the library's mc_table_checksum, computed by the oracle."""
        out = (C.c_uint64 * 4)()
        self.o.lib.orc_table_checksum(self.h, min_count, out)
        return dict(zip(("entries", "sum_counts", "digest", "max_count"), [int(x) for x in out]))

    def stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self.o.lib.orc_table_stats(self.h, *[C.byref(x) for x in v])
        return dict(zip(("reads", "runs", "bases", "kmers"), [x.value for x in v]))

    def bfs(self, seq, min_cov, max_kmers, max_radius, direction, cap=1 << 20):
        k = self.k
        out = C.create_string_buffer(cap * k)
        dist = np.empty(cap, dtype=np.int32)
        n = self.o.lib.orc_bfs(self.h, seq.encode(), k, min_cov, -1 if max_kmers is None else max_kmers,
                               -1 if max_radius is None else max_radius, direction, out, dist.ctypes.data, cap)
        if n == -1:
            return None
        if n < 0:
            raise OracleError(f"orc_bfs -> {n}: {self.o.err()}")
        raw = out.raw[: n * k].decode()
        return [raw[i * k:(i + 1) * k] for i in range(n)], dist[:n].copy()

    def env_find(self, seq_fasta, min_cov, max_kmers, max_radius, both_dirs, chunk_len, out_dir):
        return self.o.check(self.o.lib.orc_env_find(self.h, seq_fasta.encode(), self.k, min_cov,
                                                    -1 if max_kmers is None else max_kmers,
                                                    -1 if max_radius is None else max_radius,
                                                    int(both_dirs), chunk_len, out_dir.encode()))

    def env_find_ex(self, seq_fasta, hic_fasta, min_cov, max_kmers, max_radius, both_dirs, chunk_len, merge, trim, out_dir):
        return self.o.check(self.o.lib.orc_env_find_ex(self.h, seq_fasta.encode(), hic_fasta.encode() if hic_fasta else None, self.k, min_cov,
                                                       -1 if max_kmers is None else max_kmers, -1 if max_radius is None else max_radius,
                                                       int(both_dirs), chunk_len, int(merge), int(trim), out_dir.encode()))

    def classify(self, data, fmt=1, z=1.0, found_threshold=0.9, correction=False):
        """(found, margins) of every read of a FASTQ buffer: the oracle's findRead / findReadWithCorrection; margins[i] = distance of
        the decisive f64 comparison from its boundary (a class is only comparable across math libraries when it is not ~0)."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        cap = int((buf == 10).sum()) // 2 + 2
        found = np.zeros(cap, dtype=np.uint8)
        margins = np.zeros(cap, dtype=np.float64)
        n = self.o.lib.orc_classify_bytes(self.h, buf.ctypes.data if buf.size else None, buf.size, fmt, z, found_threshold, int(correction),
                                          found.ctypes.data, margins.ctypes.data, cap)
        if n < 0:
            raise OracleError(self.o.err())
        return found[:n], margins[:n]

    def reads_classifier(self, reads1, reads2, out_dir, z=1.0, found_threshold=0.9, correction=False):
        counts = (C.c_long * 4)()
        self.o.check(self.o.lib.orc_reads_classifier(self.h, reads1.encode(), reads2.encode() if reads2 else None, z, found_threshold, int(correction),
                                                     out_dir.encode(), counts))
        return list(counts)

    def extend_counts(self, seq_fasta, min_cov, max_kmers, max_radius, both_dirs, cap=1024):
        out = (C.c_longlong * cap)()
        n = self.o.lib.orc_env_extend_counts(self.h, seq_fasta.encode(), self.k, min_cov, -1 if max_kmers is None else max_kmers,
                                             -1 if max_radius is None else max_radius, int(both_dirs), out, cap)
        if n < 0:
            raise OracleError(self.o.err())
        return [int(out[i]) for i in range(min(n, cap))]

    def print_kmers(self, threshold, out_file, stat_file=None):
        return self.o.check(self.o.lib.orc_print_kmers(self.h, threshold, out_file.encode(), stat_file.encode() if stat_file else None))

    def load_kmers(self, path, threshold=0):
        return self.o.check(self.o.lib.orc_load_kmers(self.h, path.encode(), threshold))


class Synth:
    """CPU twin of the synthetic workload generator (include/mc_synth.h)."""

    def __init__(self, o: Oracle, seed=0x6D657461, n_genomes=10, genome_unit=100_000, read_len=150, n_targets=1, target_len=900,
                 repeat_len=0, err_ppm=5000, n_ppm=100, bangq_ppm=100, fasta=False):
        self.o = o
        self.blob = (C.c_uint8 * o.lib.orc_synth_params_size())()
        rc = o.lib.orc_synth_init(self.blob, seed, n_genomes, genome_unit, read_len, n_targets, target_len, repeat_len, err_ppm, n_ppm,
                                  bangq_ppm, int(fasta))
        if rc != 0:
            raise OracleError("mc_synth_init failed (plants do not fit)")
        self.record_bytes = o.lib.orc_synth_record_bytes(self.blob)
        self.read_len, self.n_targets, self.target_len, self.fasta = read_len, n_targets, target_len, fasta

    def params_bytes(self):
        return bytes(self.blob)

    def generate(self, first, count, threads=8):
        out = np.empty(count * self.record_bytes, dtype=np.uint8)
        self.o.lib.orc_synth_generate(self.blob, first, count, out.ctypes.data, threads)
        return out

    def target(self, t):
        buf = C.create_string_buffer(self.target_len)
        self.o.lib.orc_synth_target(self.blob, t, buf)
        return buf.raw.decode()

    def targets_fasta(self):
        return "".join(f">gene{t}\n{self.target(t)}\n" for t in range(self.n_targets))


_cached = None


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)


def load() -> Oracle:
    global _cached
    if _cached is None:
        if not os.path.exists(LIB):
            build()
        _cached = Oracle(C.CDLL(LIB))
    return _cached
