"""GPU: reads-classifier ("next" row f-4b) -- the per-read class of the CUDA path equals the oracle's restatement of
ReadsFinderInGraph.findRead / findReadWithCorrection (src/algo/ReadsFinderInGraph.java:37-123), and the CLI tool writes the six
FASTQ files of ReadsClassifier.runImpl (src/tools/ReadsClassifier.java:204-217).

The sums are integers and must agree exactly; the final decision is f64 arithmetic with exp and sqrt, where CUDA's and the host's math
libraries may differ in the last ulp.  The comparison is therefore on CLASSES, and a read is exempt only when the oracle reports that
its decisive comparison sits within 1e-9 of the boundary (none does in these inputs; the test counts them)."""
import os
import random
import subprocess

import numpy as np
import pytest

from helpers import make_fastq
from oracle_ffi import FMT_FASTQ, FMT_FASTQ33

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "metacherchant_b200", "csrc", "metacherchant")
OUT_FILES = ("found_1.fastq", "found_2.fastq", "not_found_1.fastq", "not_found_2.fastq", "found_s.fastq", "not_found_s.fastq")


def same_classes(got, want, margins):
    assert got.size == want.size
    diff = np.nonzero(got != want)[0]
    close = int((margins < 1e-9).sum())
    assert all(margins[i] < 1e-9 for i in diff), f"{diff.size} classes differ from the oracle, {close} reads sit on the interval's edge"
    return close


@pytest.mark.parametrize("k", [15, 31])
@pytest.mark.parametrize("correction", [False, True])
@pytest.mark.parametrize("z,thr", [(1.0, 0.9), (1.96, 0.5)])
def test_classes_match_the_oracle(oracle, k, correction, z, thr):
    from metacherchant_b200 import Context
    s = oracle.synth(genome_unit=4_000)                       # 220 kbp community
    other = oracle.synth(genome_unit=4_000, seed=99)          # unrelated sequences: mostly "not found"
    graph_reads = s.generate(0, 30_000)                       # ~20x: a mix of well and thinly covered genomes
    probe = np.concatenate([s.generate(50_000, 3_000), other.generate(0, 1_000)])
    c = Context(0)
    try:
        g = c.new_table(k).count_bytes(graph_reads, FMT_FASTQ)
        o = oracle.table(k, threads=8).count_bytes(graph_reads, FMT_FASTQ)
        want, margins = o.classify(probe, FMT_FASTQ33, z, thr, correction)
        got = c.classify_reads(g, probe, FMT_FASTQ33, z, thr, correction)
        same_classes(got, want, margins)
        assert 0 < int(want.sum()) < want.size, "the probe set should exercise both classes"
    finally:
        c.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_ragged_reads(oracle, seed):
    """Short reads (< k: never found), N / n / '.', lower case, blank lines, CRLF, low qualities (the correction path: exactly one base
    with phred < 10), chunk boundaries inside records."""
    from metacherchant_b200 import Context
    rng = random.Random(seed)
    k = 9
    genome = "".join(rng.choice("ACGT") for _ in range(600))
    lines = []
    for i in range(1500):
        L = rng.randint(0, 60)
        st = rng.randrange(0, len(genome) - 60)
        seq = list(genome[st:st + L])
        qual = [chr(33 + rng.randint(10, 40)) for _ in range(L)]
        if L and rng.random() < 0.5:                      # one low-quality base, sometimes also a wrong nucleotide there
            p = rng.randrange(L)
            qual[p] = chr(33 + rng.randint(0, 9))
            if rng.random() < 0.7:
                seq[p] = rng.choice("ACGT")
        if L and rng.random() < 0.15:
            p = rng.randrange(L)
            seq[p] = rng.choice("Nn.")
        if rng.random() < 0.3:
            seq = [ch.lower() for ch in seq]
        if rng.random() < 0.1:
            lines.append("")
        lines += [f"@q{i}", "".join(seq), "+", "".join(qual)]
    probe = np.frombuffer(("\r\n" if seed == 2 else "\n").join(lines).encode() + b"\n", dtype=np.uint8)
    graph = make_fastq(rng, 10, max_len=10)                   # something to parse; the real graph comes from the genome below
    fq = "".join(f"@g{i}\n{genome[s:s + 80]}\n+\n{'I' * len(genome[s:s + 80])}\n" for i, s in enumerate(range(0, 520, 7))).encode()
    c = Context(0, chunk_bytes=8192)                          # many chunks: records straddle them
    try:
        g = c.new_table(k).count_bytes(fq, FMT_FASTQ33)
        o = oracle.table(k, threads=2).count_bytes(fq, FMT_FASTQ33)
        for correction in (False, True):
            want, margins = o.classify(probe, FMT_FASTQ33, 1.0, 0.9, correction)
            got = c.classify_reads(g, probe, FMT_FASTQ33, 1.0, 0.9, correction)
            same_classes(got, want, margins)
        # the correction path must matter: some read is found only with it
        a, _ = o.classify(probe, FMT_FASTQ33, 1.0, 0.9, False)
        b, _ = o.classify(probe, FMT_FASTQ33, 1.0, 0.9, True)
        assert int((a != b).sum()) > 0
        del graph
    finally:
        c.close()


def test_errors_follow_the_reader(oracle):
    from metacherchant_b200 import Context, McError
    c = Context(0)
    try:
        g = c.new_table(5).count_bytes(b"@a\nACGTACGTAC\n+\nIIIIIIIIII\n", FMT_FASTQ33)
        with pytest.raises(McError, match="Incorrect nucleotide char"):
            c.classify_reads(g, b"@a\nACGTXCGTAC\n+\nIIIIIIIIII\n", FMT_FASTQ33)
        with pytest.raises(McError, match="length of chars and quality"):
            c.classify_reads(g, b"@a\nACGTACGTAC\n+\nIIII\n", FMT_FASTQ33)
        with pytest.raises(McError, match="Invalid quality code char"):
            c.classify_reads(g, b"@a\nACGTACGTAC\n+\nIIII IIIII\n", FMT_FASTQ33)
        with pytest.raises(McError, match="FASTQ input only"):
            c.classify_reads(g, b">a\nACGT\n", 0)
        assert c.classify_reads(g, b"", FMT_FASTQ33).size == 0
    finally:
        c.close()


@pytest.mark.parametrize("paired", [True, False])
def test_reads_classifier_cli(oracle, tmp_path, paired):
    k = 21
    s = oracle.synth(genome_unit=3_000)
    other = oracle.synth(genome_unit=3_000, seed=7)
    (tmp_path / "graph.fastq").write_bytes(s.generate(0, 20_000).tobytes())
    r1 = np.concatenate([s.generate(30_000, 600), other.generate(0, 200), s.generate(31_000, 200)])
    r2 = np.concatenate([s.generate(40_000, 600), s.generate(41_000, 200), other.generate(500, 200)])
    (tmp_path / "r_1.fastq").write_bytes(r1.tobytes())
    (tmp_path / "r_2.fastq").write_bytes(r2.tobytes())
    reads = [str(tmp_path / "r_1.fastq")] + ([str(tmp_path / "r_2.fastq")] if paired else [])
    cmd = [CLI, "--tool", "reads-classifier", "-k", str(k), "-i", str(tmp_path / "graph.fastq"), "-r", *reads, "-o", str(tmp_path / "gpu"),
           "--work-dir", str(tmp_path / "wd"), "--found-threshold", "80", "--interval95"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Hashtable size:" in r.stderr and "Total:" in r.stderr and "Reads have been written." in r.stderr
    o = oracle.table(k, threads=8)
    o.count_file(str(tmp_path / "graph.fastq"))
    counts = o.reads_classifier(reads[0], reads[1] if paired else None, str(tmp_path / "orc"), z=1.96, found_threshold=0.8)
    assert f"Total: {2 * sum(counts)} reads" in r.stderr
    for f in OUT_FILES:
        a, b = (tmp_path / "gpu" / f).read_bytes(), (tmp_path / "orc" / f).read_bytes()
        assert a == b, f"{f} differs ({len(a)} vs {len(b)} bytes)"
    if paired:
        assert all(x > 0 for x in counts), counts          # every class occurs
    else:
        assert counts[0] == counts[3] == 0 and counts[1] > 0 and counts[2] > 0
    # the same graph from a .kmers.bin file (ReadsClassifier.loadGraph, :90-92)
    subprocess.run([CLI, "--tool", "kmer-counter", "-k", str(k), "-i", str(tmp_path / "graph.fastq"), "-o", str(tmp_path / "kc"), "--work-dir", str(tmp_path / "wd")],
                   check=True, capture_output=True)
    r = subprocess.run([CLI, "--tool", "reads-classifier", "-k", str(k), "-i", str(tmp_path / "kc" / "graph.kmers.bin"), "-r", *reads, "-o", str(tmp_path / "gpu2"),
                        "--work-dir", str(tmp_path / "wd"), "--found-threshold", "80", "--interval95"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for f in OUT_FILES:
        assert (tmp_path / "gpu2" / f).read_bytes() == (tmp_path / "orc" / f).read_bytes()
