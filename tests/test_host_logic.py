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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built_library():
    """The product library is built in-tree (nvcc cross-compiles without a GPU); the host-only entry points need no device."""
    import subprocess
    if not os.path.exists(_ffi.lib_path()):
        subprocess.run(["make", "-C", os.path.join(ROOT, "metacherchant_b200", "csrc"), "-j", "8"], check=True)
    return _ffi.load_library()


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


def crowded_kmers(rng, k, n_total, hot_bins, hot_each, mask, accept=lambda s: True):
    """k-mers of which `hot_each` share each of `hot_bins` residues of the spread String.hashCode modulo mask + 1: those bins are
    treeified by a real java.util.HashMap once the table has mask + 1 buckets or fewer (and >= 64)."""
    def spread(s):
        h = pyref.java_hash(s)
        return h ^ (h >> 16)
    hot = {rng.randrange(mask + 1): [] for _ in range(hot_bins)}
    seen = set()
    while any(len(v) < hot_each for v in hot.values()):
        s = "".join(rng.choice("ACGT") for _ in range(k))
        b = spread(s) & mask
        if s not in seen and accept(s) and b in hot and len(hot[b]) < hot_each:
            hot[b].append(s)
            seen.add(s)
    kmers = [x for v in hot.values() for x in v]
    while len(kmers) < n_total:
        s = "".join(rng.choice("ACGT") for _ in range(k))
        if s not in seen and accept(s):
            seen.add(s)
            kmers.append(s)
    rng.shuffle(kmers)
    return kmers


@pytest.mark.parametrize("seed", range(12))
def test_hashmap_order_with_treeified_bins(seed):
    """Bins of 9 and more nodes (java.util.HashMap.treeifyBin): root moved to the front, new nodes linked after their tree parent,
    resize splits that re-treeify or untreeify at 6 nodes, tables under 64 buckets resized instead.  The library's replay
    (JavaHashReplay, C++) against the model in oracle/pyref.py (JavaHashMapModel): two implementations written separately from the
    JDK's algorithm; the model's own invariants (red-black properties, list/tree consistency) are checked in tests/test_oracle_vs_pyref.py."""
    rng = random.Random(seed)
    k = rng.choice([9, 13, 23, 31])
    n = rng.choice([60, 200, 1500, 3000])                 # final capacities 128 .. 4096: equal residues mod 4096 share a bin throughout
    kmers = crowded_kmers(rng, k, n, rng.randint(1, 3), rng.randint(9, 40), 4095)
    assert pyref.java_order_needs_replay(kmers)
    want = pyref.java_order(kmers)
    got = lib_order([pyref.encode(s) for s in kmers], k)
    assert [kmers[i] for i in got] == want
    # the history matters: the same keys in sorted insertion order give another iteration order, and both sides follow it
    kmers2 = sorted(kmers)
    got2 = lib_order([pyref.encode(s) for s in kmers2], k)
    assert [kmers2[i] for i in got2] == pyref.java_order(kmers2)


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


# ---- target loading: RichFastaReader.readAll (src/io/RichFastaReader.java:38-76) + new DnaQ(str, 0).toString() ---------------
def rich_fasta_reader(text: str):
    """Literal transcription of readAll: returns (dnas, comments)."""
    dnas, comments = [], []
    last_comment, cur_comment, cur_dna = True, "", ""
    for line in pyref.java_lines(text.encode()):
        if line.startswith(">") or line.startswith(";"):
            if not last_comment:
                dnas.append(cur_dna)
                cur_dna, cur_comment = "", ""
            cur_comment += line[1:]
            last_comment = True
        else:
            if last_comment:
                comments.append(cur_comment)
                cur_dna, cur_comment = "", ""
            cur_dna += line
            last_comment = False
    if cur_comment:
        comments.append(cur_comment)
    if cur_dna:
        dnas.append(cur_dna)
    # DnaQ(str, 0): N, n, '.' -> nucleotide 0 ('A'), lower case accepted, toString() upper case (ITMO!/dna/DnaQ.java:21-30)
    tr = {"N": "A", "n": "A", ".": "A"}
    return ["".join(tr.get(c, c.upper()) for c in d) for d in dnas], comments


def lib_targets(path):
    lib = _ffi.load_library()
    h = C.c_void_p()
    rc = lib.mc_targets_load(str(path).encode(), C.byref(h))
    assert rc == 0, lib.mc_last_error()
    try:
        n = lib.mc_targets_count(h)
        seqs = [lib.mc_targets_seq(h, i).decode() for i in range(n)]
        names = []
        i = 0
        while True:
            name = lib.mc_targets_name(h, i)
            if name is None:
                break
            names.append(name.decode())
            i += 1
        return seqs, names
    finally:
        lib.mc_targets_free(h)


TARGET_FILES = {
    "plain": ">gene one\nACGTACGT\nTTGA\n>two\nGGGCCC\n",
    "headerless (the Hi-C example's seq.fasta)": "ACGTTGCAACGT\n",
    "multi-line header": ">part1\n;part2\nACGT\nAC\n",
    "crlf, lower case, N and dot": ">g\r\nacgtNn.\r\nTT\r\n",
    "header only at the end": ">a\nACGT\n>dangling\n",
    "empty header": ">\nACGT\n",
    "no final newline": ">x\nACGT",
    "semicolon header": ";old style\nACGTAC\n",
}


@pytest.mark.parametrize("name", sorted(TARGET_FILES))
def test_target_loading_matches_rich_fasta_reader(tmp_path, name):
    p = tmp_path / "t.fasta"
    p.write_text(TARGET_FILES[name], newline="")
    seqs, names = lib_targets(p)
    want_seqs, want_names = rich_fasta_reader(TARGET_FILES[name])
    assert seqs == want_seqs
    assert names == want_names


def test_target_loading_errors(tmp_path):
    lib = _ffi.load_library()
    h = C.c_void_p()
    assert lib.mc_targets_load(str(tmp_path / "missing.fasta").encode(), C.byref(h)) != 0
    assert b"Could not load sequences from" in (lib.mc_last_error() or b"")
    p = tmp_path / "bad.fasta"
    p.write_text(">g\nACGTXACGT\n")
    assert lib.mc_targets_load(str(p).encode(), C.byref(h)) != 0
    assert b"Incorrect nucleotide char" in (lib.mc_last_error() or b"")


# ---- reader set-up: format by suffix and quality-offset sniffing (ITMO!/io/ReadersUtils.java:27-54,63-77) ---------------------
@pytest.mark.parametrize("name", ["a.fastq", "A.FQ", "x/y.fasta", "b.fa", "c.fn", "d.fna", "r1.fastq.gz", "R.FA.GZ", "reads.txt", "a.fastq.bz2",
                                  "b.binq", "noext", "a.fastq.gz.gz"])
def test_format_detection_matches_the_oracle(oracle, name):
    import oracle_ffi
    lib = _ffi.load_library()
    gz = C.c_int(-1)
    got = lib.mc_detect_file_format(name.encode(), C.byref(gz))
    low, packed = name.lower(), 0           # ReadersUtils strips ".gz", then ".bz2" (ITMO!/io/ReadersUtils.java:36-43)
    plain = name
    if low.endswith(".gz"):
        plain, packed = plain[:-3], 1
    if plain.lower().endswith(".bz2"):
        plain, packed = plain[:-4], 2
    try:
        want = oracle.detect_format(plain)           # the oracle's suffix rules on the name without the compression suffix
    except oracle_ffi.OracleError:
        want = None
    if want is None:
        assert got < 0 and b"Failed to read from file" in (lib.mc_last_error() or b"")
    else:
        assert got == want and gz.value == packed


def test_quality_offset_sniffing_matches_pyref():
    from helpers import make_fastq
    lib = _ffi.load_library()
    rng = random.Random(21)
    cases = [make_fastq(rng, 50, offset=64), make_fastq(rng, 50, offset=33), make_fastq(rng, 1500, offset=64, max_len=30),
             make_fastq(rng, 3, offset=64, crlf=True), b"", b"@r\nACGT\n+\nIIII\n", b"@r\nACGT\n+\nhhhh\n"]
    # a low quality character only after the first 1000 records does not change the verdict (and crashes the Java reader later)
    late = make_fastq(rng, 1001, offset=64, p_n=0, max_len=20, min_len=5, final_newline=True) + b"@late\nACGT\n+\n!!!!\n"
    cases.append(late)
    for data in cases:
        got = lib.mc_sniff_quality_offset(data, len(data))
        assert got == pyref.sniff_offset(data), data[:60]


# ---- String.format("%6.2f ", float): the cells of Jacard_sym.txt / Jacard_alt.txt -------------------------------------------
def lib_fmt(v):
    lib = _ffi.load_library()
    buf = C.create_string_buffer(64)
    assert lib.mc_java_format_6_2f(C.c_float(v), buf) == 0
    return buf.value.decode()


# Hand-derived from java.util.Formatter's rules (EnvironmentFinderMultiMain.java:141-142,163-164): the float is widened to
# double, Double.toString's digits are rounded HALF_UP to two decimals, the result is right-aligned in 6 columns.
#   0.125f = 0.125 exactly            -> tie, HALF_UP -> 0.13   (C's printf gives 0.12: half-to-even on the binary value)
#   0.625f = 0.625 exactly            -> 0.63               (C: 0.62)
#   0.005f = 0.004999999888241291     -> 0.00
#   2.675f = 2.674999952316284        -> 2.67
#   1e-3f  = 0.0010000000474974513    -> 0.00
#   0.995f = 0.9950000047683716       -> 1.00
#   0.015f = 0.014999999664723873     -> 0.01,   0.025f = 0.02500000037252903 -> 0.03
JAVA_FORMAT_VECTORS = [
    (0.125, "  0.13 "), (0.375, "  0.38 "), (0.625, "  0.63 "), (0.875, "  0.88 "), (0.005, "  0.00 "), (2.675, "  2.67 "), (1e-3, "  0.00 "),
    (0.995, "  1.00 "), (0.015, "  0.01 "), (0.025, "  0.03 "), (0.0, "  0.00 "), (1.0, "  1.00 "), (0.5, "  0.50 "), (-0.5, " -0.50 "),
    (12.345, " 12.35 "), (100.0, "100.00 "), (1234.5, "1234.50 "), (float("nan"), "   NaN "), (float("inf"), "Infinity "), (float("-inf"), "-Infinity "),
    (-0.001, " -0.00 "), (3.0e9, "3000000000.00 "), (0.33333334, "  0.33 "), (0.6666667, "  0.67 "),
]


@pytest.mark.parametrize("v,want", JAVA_FORMAT_VECTORS)
def test_java_format_on_hand_derived_vectors(v, want):
    assert lib_fmt(v) == want


def test_java_format_matches_an_independent_model(oracle):
    """Python model of the same rule (repr = shortest round-trip digits, decimal HALF_UP) and the oracle's own formatter over
    random float32 values, Jaccard-like ratios and every exact tie k/8: three implementations, three algorithms."""
    from decimal import ROUND_HALF_UP, Decimal
    olib = oracle.lib
    has_oracle_fmt = hasattr(olib, "orc_java_format_6_2f")
    if has_oracle_fmt:
        olib.orc_java_format_6_2f.argtypes = [C.c_float, C.c_char_p]
    rng = np.random.default_rng(5)
    vals = list(rng.random(3000, dtype=np.float32)) + [np.float32(a) / np.float32(b) for a, b in rng.integers(1, 5000, size=(2000, 2))]
    vals += [np.float32(i / 8) for i in range(0, 80)] + [np.float32(x) for x in rng.normal(0, 1e4, 500)] + [np.float32(i / 200) for i in range(400)]
    for v in vals:
        d = Decimal(repr(float(v))).quantize(Decimal("0.01"), rounding=ROUND_HALF_UP)
        txt = f"{d:.2f}"
        if float(v) < 0 and not txt.startswith("-"):
            txt = "-" + txt
        want = txt.rjust(6) + " "
        got = lib_fmt(float(v))
        assert got == want, (float(v), got, want)
        if has_oracle_fmt:
            buf = C.create_string_buffer(64)
            olib.orc_java_format_6_2f(C.c_float(float(v)), buf)
            assert buf.value.decode() == want, (float(v), buf.value, want)
