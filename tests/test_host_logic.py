"""CPU: host logic of the product library that needs no device -- the java.util.HashMap iteration order on 2-bit k-mer
codes (KmerMap in csrc/mc_env.cu: it fixes the line order of graph.txt / env.txt and every node id), pinned on the
reference's committed graph.txt and cross-checked against the pure-Python transcription."""
import ctypes as C
import lzma
import os
import random

import numpy as np
import pytest

from metacherchant_b200 import _ffi
from oracle import pyref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hic_example1")


def lib_order(codes, k, peak=0):
    lib = _ffi.load_library()
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    order = np.empty(codes.size, dtype=np.uint32)
    n = C.c_size_t()
    rc = lib.mc_kmer_hashmap_order(codes.ctypes.data, codes.size, k, peak, order.ctypes.data, C.byref(n))
    assert rc == 0, lib.mc_last_error()
    return order[: n.value]


@pytest.mark.parametrize("k", [1, 5, 12, 23, 31])
@pytest.mark.parametrize("n,peak", [(1, 0), (12, 0), (13, 0), (500, 0), (500, 4000), (20000, 0)])
def test_hashmap_order_matches_pyref(k, n, peak):
    rng = random.Random(k * 1000 + n + peak)
    kmers, seen = [], set()
    while len(kmers) < min(n, 4 ** k):
        s = "".join(rng.choice("ACGT") for _ in range(k))
        if s not in seen:
            seen.add(s)
            kmers.append(s)
    codes = [pyref.encode(s) for s in kmers]
    want = pyref.java_order(kmers, peak or None)
    got = lib_order(codes, k, peak)
    assert [kmers[i] for i in got] == want


def test_duplicates_keep_their_first_position():
    kmers = ["ACGTA", "TTTTT", "ACGTA", "GGGCC", "TTTTT", "CATCA"]
    got = lib_order([pyref.encode(s) for s in kmers], 5)
    distinct = list(dict.fromkeys(kmers))
    assert [kmers[i] for i in got] == pyref.java_order(distinct)


def test_reference_graph_txt_is_a_fixed_point():
    """The 93 572 lines of Hi-C_pipline/example_work_dir/output/1/merged/graph.txt are in HashMap iteration order: putting
    them into the map in file order and iterating gives the file back (the same pin as the oracle's, tests/test_oracle_golden.py)."""
    with lzma.open(os.path.join(GOLD, "graph.txt.xz"), "rt") as f:
        kmers = [line.split(" ")[0] for line in f if line.strip()]
    assert len(kmers) == 93572 and all(len(s) == 31 for s in kmers)
    got = lib_order([pyref.encode(s) for s in kmers], 31)
    assert np.array_equal(got, np.arange(len(kmers), dtype=np.uint32))
    # and from a shuffled insertion order the buckets come out the same, insertion order only decides ties inside a bucket
    rng = random.Random(5)
    perm = list(range(len(kmers)))
    rng.shuffle(perm)
    shuffled = [kmers[i] for i in perm]
    got2 = lib_order([pyref.encode(s) for s in shuffled], 31)
    assert [shuffled[i] for i in got2] == pyref.java_order(shuffled)


def test_bad_arguments():
    lib = _ffi.load_library()
    n = C.c_size_t()
    assert lib.mc_kmer_hashmap_order(None, 0, 32, 0, None, C.byref(n)) != 0
    assert lib.mc_kmer_hashmap_order(None, 0, 5, 0, None, C.byref(n)) == 0 and n.value == 0
