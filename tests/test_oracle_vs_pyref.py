"""Stages A-C have no reference fixture (parity unpinned), so the C++ oracle is cross-checked
against an independent pure-Python transcription of the same Java sources (oracle/pyref.py) on small inputs,
and against brute-force properties."""
import random

import numpy as np
import pytest

from oracle import pyref
from oracle_ffi import FMT_FASTA, FMT_FASTQ, FMT_FASTQ33, FMT_FASTQ64, OracleError


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def make_fastq(rng, n_reads, max_len=60, p_n=0.03, p_bang=0.03, crlf=False, blanks=False, lower=False, offset=33):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i in range(n_reads):
        L = rng.randint(0, max_len)
        seq = "".join(("N" if rng.random() < p_n else rng.choice("ACGTacgt" if lower else "ACGT")) for _ in range(L))
        if rng.random() < 0.1 and L:
            p = rng.randrange(L)
            seq = seq[:p] + rng.choice(".n") + seq[p + 1:]
        zero = chr(offset)
        qual = "".join((zero if rng.random() < p_bang else chr(rng.randint(offset + 1, min(126, offset + 41)))) for _ in range(L))
        if blanks and rng.random() < 0.2:
            out.append("")
        out.append(f"@r{i} some comment")
        out.append(seq)
        if blanks and rng.random() < 0.2:
            out.append("")
        out.append("+" + (f"r{i}" if rng.random() < 0.3 else ""))
        out.append(qual)
    text = nl.join(out) + (nl if rng.random() < 0.8 else "")
    return text.encode()


def table_dict(t):
    k, c = t.dump()
    return dict(zip(k.tolist(), c.tolist()))


@pytest.mark.parametrize("k", [3, 5, 11, 21, 31])
@pytest.mark.parametrize("variant", ["plain", "crlf", "blanks", "lower"])
def test_fastq_count_matches_pyref(oracle, k, variant):
    rng = random.Random(1000 * k + len(variant))
    data = make_fastq(rng, 120, max_len=90, crlf=variant == "crlf", blanks=variant == "blanks", lower=variant == "lower")
    t = oracle.table(k, threads=1).count_bytes(data, FMT_FASTQ33)
    want = pyref.count_runs(pyref.fastq_runs(data, 33), k)
    assert table_dict(t) == want
    assert t.size() == len(want)


def test_fastq_sniffing(oracle):
    rng = random.Random(7)
    # all quality chars >= 64 -> sniffed as Phred+64: '@' (64) becomes the split char, 'A'.. are fine
    data = make_fastq(rng, 50, offset=64)
    assert pyref.sniff_offset(data) == 64
    a = table_dict(oracle.table(7).count_bytes(data, FMT_FASTQ))
    assert a == pyref.count_runs(pyref.fastq_runs(data, 64), 7)
    assert a == table_dict(oracle.table(7).count_bytes(data, FMT_FASTQ64))
    # one low char in the first 1000 records -> Phred+33
    data33 = make_fastq(rng, 50, offset=33)
    assert pyref.sniff_offset(data33) == 33
    assert table_dict(oracle.table(7).count_bytes(data33, FMT_FASTQ)) == pyref.count_runs(pyref.fastq_runs(data33, 33), 7)


def test_fastq_phred_wraps_at_64(oracle):
    """Stored phred is 6 bits (DnaQ.java:147-150): Sanger char 97 ('a', phred 64) wraps to 0 and splits the read."""
    seq = "ACGTACGTAC"
    qual = "IIIIaIIIII"
    data = f"@r\n{seq}\n+\n{qual}\n".encode()
    got = table_dict(oracle.table(3).count_bytes(data, FMT_FASTQ33))
    want = pyref.count_runs([[pyref.CODE[c] for c in "ACGT"], [pyref.CODE[c] for c in "CGTAC"]], 3)
    assert got == want


def test_fastq_errors(oracle):
    def bad(data, fmt=FMT_FASTQ33):
        with pytest.raises(OracleError):
            oracle.table(5).count_bytes(data, fmt)
    bad(b"@r\nACGT\n+\nIII\n")            # length mismatch (FastqReader.java:67-70)
    bad(b"@r\nACGT\n+\n")                 # unexpected EOF (:105-108)
    bad(b"@r\nACGT\n")                    # sequence without quality (:61-64)
    bad(b"r\nACGT\n+\nIIII\n")            # header must start with @ or + (:99-103)
    bad(b"@r\nACXT\n+\nIIII\n")           # illegal nucleotide (DnaTools.java:61-62)
    bad(b"@r\nACRT\n+\nIIII\n")           # IUPAC: random in the reference, rejected here
    bad(b"@r\nACGT\n+\nII\x1fI\n")        # quality char < 33 (Sanger.java:7-12)
    bad(b"@r\nACGT\n+\nII!I\n", FMT_FASTQ64)   # forced Phred+64 with a low char (Illumina.java:7-12)
    # N positions skip the quality check (FastqReader.java:75-76)
    oracle.table(5).count_bytes(b"@r\nACNT\n+\nII\x1fI\n", FMT_FASTQ33)
    # empty input and blank-only input are fine
    assert oracle.table(5).count_bytes(b"", FMT_FASTQ33).size() == 0
    assert oracle.table(5).count_bytes(b"\n\n\n", FMT_FASTQ33).size() == 0


@pytest.mark.parametrize("k", [4, 15, 31])
def test_fasta_count_matches_pyref(oracle, k):
    rng = random.Random(k)
    lines = []
    for i in range(80):
        if rng.random() < 0.9:
            lines.append((">" if rng.random() < 0.8 else ";") + f"r{i}")
        if rng.random() < 0.05:
            lines.append(">double header")
        for _ in range(rng.randint(0, 3)):
            L = rng.randint(0, 70)
            s = rand_seq(rng, L, "ACGTacgt")
            if rng.random() < 0.08 and L:
                p = rng.randrange(L)
                s = s[:p] + rng.choice("Nn") + s[p + 1:]
            lines.append(s)
            if rng.random() < 0.1:
                lines.append("")
    data = ("\n".join(lines) + "\n").encode()
    got = table_dict(oracle.table(k).count_bytes(data, FMT_FASTA))
    assert got == pyref.count_runs(pyref.fasta_runs(data), k)
    with pytest.raises(OracleError):
        oracle.table(k).count_bytes(b">r\nAC.T\n", FMT_FASTA)   # '.' is not tolerated by new Dna(s)


def test_saturation_and_absent(oracle):
    t = oracle.table(3)
    data = (">r\n" + "A" * 40000 + "\n").encode()
    t.count_bytes(data, FMT_FASTA)
    assert t.get(0) == 32767          # AAA saturates (NumUtils.java:21-26); key 0 is the FREE-key side slot
    assert t.get(5) == -1             # absent -> -1 (Long2ShortHashMap.java:171-174)
    assert t.size() == 1


def test_threads_do_not_change_the_table(oracle):
    rng = random.Random(3)
    data = make_fastq(rng, 3000, max_len=100)
    a = oracle.table(9, threads=1).count_bytes(data, FMT_FASTQ33)
    b = oracle.table(9, threads=8).count_bytes(data, FMT_FASTQ33)
    ka, ca = a.dump()
    kb, cb = b.dump()
    assert np.array_equal(ka, kb) and np.array_equal(ca, cb)
    assert a.stats() == b.stats()


@pytest.mark.parametrize("k", range(1, 32))
def test_bit_revcomp_equals_string_revcomp(oracle, k):
    rng = random.Random(k)
    for _ in range(50):
        s = rand_seq(rng, k)
        code = oracle.encode(s)
        assert code == pyref.encode(s)
        assert oracle.decode(code, k) == s
        assert oracle.decode(oracle.revcomp(code, k), k) == pyref.rc_str(s)
        assert oracle.canonical(code, k) == min(code, pyref.revcomp_code(code, k))


def small_graph(rng, k, genome_len=400, n_reads=300, read_len=40):
    genome = rand_seq(rng, genome_len)
    # a repeat so the graph branches
    genome = genome[:150] + genome[20:60] + genome[150:]
    reads = []
    for i in range(n_reads):
        s = rng.randrange(0, len(genome) - read_len)
        r = genome[s:s + read_len]
        if rng.random() < 0.5:
            r = pyref.rc_str(r)
        if rng.random() < 0.2:
            p = rng.randrange(read_len)
            r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        reads.append(f">r{i}\n{r}\n")
    return genome, "".join(reads).encode()


@pytest.mark.parametrize("k", [5, 7, 12, 31])
@pytest.mark.parametrize("direction", [-1, 0, 1])
@pytest.mark.parametrize("limits", [(None, 3), (10, None), (40, 6), (100000, None), (1, None), (None, 0)])
def test_bfs_matches_pyref(oracle, k, direction, limits):
    rng = random.Random(k * 100 + direction)
    genome, data = small_graph(rng, k)
    t = oracle.table(k).count_bytes(data, FMT_FASTA)
    table = table_dict(t)
    target = genome[100:100 + max(k + 5, 30)]
    target = target + target[:k]   # duplicate seed k-mers
    max_kmers, max_radius = limits
    for min_cov in (1, 2):
        want = pyref.bfs(table, target, k, min_cov, max_kmers, max_radius, direction)
        got = t.bfs(target, min_cov, max_kmers, max_radius, direction)
        if want is None:
            assert got is None
            continue
        order, dist = got
        assert order == list(want.keys())          # same set AND same discovery order
        assert dist.tolist() == list(want.values())


@pytest.mark.parametrize("both", [False, True])
def test_environment_matches_pyref(oracle, both, tmp_path):
    k = 9
    rng = random.Random(99)
    genome, data = small_graph(rng, k, genome_len=900, n_reads=900)
    t = oracle.table(k).count_bytes(data, FMT_FASTA)
    target = genome[300:360]
    fa = tmp_path / "seq.fasta"
    fa.write_text(f">gene one\n{target[:30]}\n{target[30:]}\n")
    n = t.env_find(str(fa), 2, 150, None, both, 1, str(tmp_path / "out"))
    assert n == 1
    want = pyref.environment(table_dict(t), target, k, 2, 150, None, both)
    lines = (tmp_path / "out" / "gene one" / "graph.txt").read_text().splitlines()
    assert lines == [f"{kmer} {c}" for kmer, c in want]
    assert (tmp_path / "out" / "gene one" / "env.txt").read_text().splitlines() == lines
    for f in ("graph.gfa", "seqs.fasta", "tsvs/nodes.tsv", "tsvs/edges.tsv"):
        assert (tmp_path / "out" / "gene one" / f).exists()


@pytest.mark.parametrize("both", [False, True])
@pytest.mark.parametrize("merge,trim,hic", [(True, False, False), (True, False, True), (False, True, False), (True, True, True)])
def test_merge_hicseq_trim_match_pyref(oracle, both, merge, trim, hic, tmp_path):
    """--merge (+ --hicseq) seeds one environment from every sequence; --trim keeps the k-mers that reach a
    lastKmers member walking against the pass direction.  C++ oracle vs the independent Python transcription."""
    k = 9
    rng = random.Random(1234 + both)
    genome, data = small_graph(rng, k, genome_len=1200, n_reads=1200)
    t = oracle.table(k).count_bytes(data, FMT_FASTA)
    targets = [genome[200:250], genome[235:290], genome[700:745]]      # two overlap: duplicated seed k-mers
    hics = [genome[400:440], "ACGT" * 8]
    fa = tmp_path / "seq.fasta"
    fa.write_text("".join(f">g{i}\n{s}\n" for i, s in enumerate(targets)))
    hf = tmp_path / "hic.fasta"
    hf.write_text("".join(f">h{i}\n{s}\n" for i, s in enumerate(hics)))
    out = tmp_path / "out"
    n = t.env_find_ex(str(fa), str(hf) if hic else None, 2, 120, None, both, 1, merge, trim, str(out))
    if merge:
        assert n == 1
        seeds = targets + (hics if hic else [])
        want = pyref.environment(table_dict(t), seeds, k, 2, 120, None, both, trim)
        lines = (out / "merged" / "graph.txt").read_text().splitlines()
        assert lines == [f"{kmer} {c}" for kmer, c in want]
    else:
        assert n == len(targets)
        names = [f"h{i}" for i in range(len(hics))] if hic else [f"g{i}" for i in range(len(targets))]
        for i, target in enumerate(targets):
            want = pyref.environment(table_dict(t), target, k, 2, 120, None, both, trim)
            lines = (out / names[i] / "graph.txt").read_text().splitlines()
            assert lines == [f"{kmer} {c}" for kmer, c in want]


def test_level_synchronous_rule_equals_fifo_rule():
    """ordering a batch's candidates by (parent queue index, neighbour slot) and keeping the first
    occurrence gives the same set and discovery order as the sequential FIFO loop, also when --maxkmers is hit.
    This is the rule the GPU BFS implements; here it is simulated in Python against pyref.bfs."""
    def batched(table, seq, k, min_cov, max_kmers, max_radius, direction, batch):
        def occ(kmer):
            c = pyref.encode(kmer)
            return table.get(min(c, pyref.revcomp_code(c, k)), -1)
        queue, dist = [], {}
        for i in range(len(seq) - k + 1):
            kmer = seq[i:i + k]
            if occ(kmer) >= min_cov and kmer not in dist:
                queue.append(kmer)
                dist[kmer] = 0
        if not queue:
            return None
        head, stopped = 0, False
        while head < len(queue) and not stopped:
            parents = queue[head:head + batch]
            head += len(parents)
            cands = []
            for p in parents:
                left = [n + p[:-1] for n in pyref.NUC]
                right = [p[1:] + n for n in pyref.NUC]
                nbs = left if direction == -1 else right if direction == 1 else [x for pr in zip(left, right) for x in pr]
                for nb in nbs:
                    if occ(nb) >= min_cov and (max_radius is None or dist[p] + 1 <= max_radius):
                        cands.append((nb, dist[p] + 1))
            seen = set()
            for nb, d in cands:                       # first occurrence in candidate order
                if nb in dist or nb in seen:
                    continue
                seen.add(nb)
                if max_kmers is not None and len(dist) >= max_kmers:
                    stopped = True                    # budget exhausted: nothing can ever be added again
                    break
                queue.append(nb)
                dist[nb] = d
        return dist
    rng = random.Random(5)
    for trial in range(60):
        k = 7
        genome, data = small_graph(rng, k, genome_len=300, n_reads=200, read_len=30)
        table = pyref.count_runs(pyref.fasta_runs(data), k)
        target = genome[50:90] + genome[50:60]
        for direction in (-1, 0, 1):
            for mk, mr in ((None, 4), (25, None), (60, 8)):
                for batch in (1, 3, 64):
                    a = pyref.bfs(table, target, k, 1, mk, mr, direction)
                    b = batched(table, target, k, 1, mk, mr, direction, batch)
                    assert (a is None) == (b is None)
                    if a is not None:
                        assert list(a.items()) == list(b.items())


def test_gz_input_equals_plain_input(oracle, tmp_path):
    """`.fastq.gz` / `.fa.gz` are the plain readers over a GZIPInputStream (ReadersUtils.java:36-39,
    FastqGZReader.java:26-34), concatenated members included; `.bz2` likewise (FastqBZ2Reader.java:26-34)."""
    import gzip
    import random
    from helpers import make_fasta, make_fastq
    rng = random.Random(11)
    fq, fa = make_fastq(rng, 400, max_len=90), make_fasta(rng, 200)
    (tmp_path / "a.fastq").write_bytes(fq)
    (tmp_path / "a.FQ.gz").write_bytes(gzip.compress(fq[:7001]) + gzip.compress(fq[7001:]))     # two members, upper-case suffix
    (tmp_path / "b.fa").write_bytes(fa)
    (tmp_path / "b.fa.gz").write_bytes(gzip.compress(fa))
    for plain, packed, k in (("a.fastq", "a.FQ.gz", 13), ("b.fa", "b.fa.gz", 9)):
        p = oracle.table(k)
        p.count_file(str(tmp_path / plain))
        g = oracle.table(k)
        g.count_file(str(tmp_path / packed))
        assert p.size() > 0 and p.stats() == g.stats()
        pk, pc = p.dump()
        gk, gc = g.dump()
        assert (pk == gk).all() and (pc == gc).all()
    # .bz2: the same readers over Hadoop's BZip2Codec stream, which also reads concatenated streams
    import bz2
    import oracle_ffi
    (tmp_path / "a.fq.bz2").write_bytes(bz2.compress(fq[:9000]) + bz2.compress(fq[9000:]))
    (tmp_path / "b.fasta.bz2").write_bytes(bz2.compress(fa))
    for plain, packed, k in (("a.fastq", "a.fq.bz2", 13), ("b.fa", "b.fasta.bz2", 9)):
        p = oracle.table(k)
        p.count_file(str(tmp_path / plain))
        g = oracle.table(k)
        g.count_file(str(tmp_path / packed))
        assert p.stats() == g.stats()
        assert all((a == b).all() for a, b in zip(p.dump(), g.dump()))
    (tmp_path / "c.fastq.bz2").write_bytes(b"BZh9" + b"garbage" * 10)
    with pytest.raises(oracle_ffi.OracleError):
        oracle.table(13).count_file(str(tmp_path / "c.fastq.bz2"))


# ---- f-4b: the reads-classifier restatement against a literal Python transcription ------------------------------------------
def py_find_read(get, nuc, k, z, thr):
    """ReadsFinderInGraph.findRead (src/algo/ReadsFinderInGraph.java:37-49) + getCoverage (:51-74) + delta (:81-87), line by line."""
    import math
    if len(nuc) < k:
        return False
    cov = []
    for i in range(len(nuc) - k + 1):
        fw = 0
        for b in nuc[i:i + k]:
            fw = (fw << 2) | b
        rc = 0
        for b in reversed(nuc[i:i + k]):
            rc = (rc << 2) | (3 - b)
        cov.append(max(get(min(fw, rc)), 0))
    cov_mean = (sum(cov) + cov[-1] * (k - 1)) / len(nuc)
    width = (sum(1 for c in cov if c > 0) + (1 if cov[-1] > 0 else 0) * (k - 1)) / len(nuc)
    theory = 1.0 - math.exp(-cov_mean)
    std = z * math.sqrt(math.exp(-cov_mean) * (1 - math.exp(-cov_mean)) / len(nuc))
    return (not (width < thr)) and (width == 1 or (width != 0 and -std <= width - theory <= std))


def test_classifier_restatement_matches_a_literal_transcription(oracle):
    import random
    from helpers import rand_seq
    rng = random.Random(11)
    k = 7
    genome = rand_seq(rng, 300)
    fq = "".join(f"@g{i}\n{genome[s:s + 50]}\n+\n{'I' * len(genome[s:s + 50])}\n" for i, s in enumerate(range(0, 260, 3))).encode()
    t = oracle.table(k, threads=2).count_bytes(fq, 33)
    code = {"A": 0, "G": 1, "C": 2, "T": 3}
    reads, lines = [], []
    for i in range(400):
        L = rng.randint(0, 40)
        st = rng.randrange(0, 250)
        seq = list(genome[st:st + L])
        for _ in range(rng.randint(0, 2)):
            if seq:
                seq[rng.randrange(len(seq))] = rng.choice("ACGTN")
        qual = [chr(33 + rng.randint(0, 40)) for _ in seq]
        reads.append((seq, qual))
        lines += [f"@r{i}", "".join(seq), "+", "".join(qual)]
    buf = ("\n".join(lines) + "\n").encode()
    for z, thr in ((1.0, 0.9), (1.96, 0.4)):
        found, margins = t.classify(buf, 33, z, thr, False)
        for i, (seq, qual) in enumerate(reads):
            nuc = [code.get(ch, 0) for ch in seq]
            if margins[i] > 1e-9:
                assert bool(found[i]) == py_find_read(lambda key: int(t.get(key)), nuc, k, z, thr), (i, "".join(seq))
        # correction: a read with exactly one phred < 10 base is found iff one of the four substitutions passes at threshold 0.9
        found_c, margins_c = t.classify(buf, 33, z, thr, True)
        for i, (seq, qual) in enumerate(reads):
            nuc = [code.get(ch, 0) for ch in seq]
            phred = [0 if ch == "N" else (ord(q) - 33) & 63 for ch, q in zip(seq, qual)]
            bad = [j for j, p in enumerate(phred) if p < 10]
            if len(nuc) < k or margins_c[i] <= 1e-9:
                continue
            if len(bad) == 1:
                want = any(py_find_read(lambda key: int(t.get(key)), nuc[:bad[0]] + [x] + nuc[bad[0] + 1:], k, z, 0.9) for x in range(4))
            else:
                want = py_find_read(lambda key: int(t.get(key)), nuc, k, z, thr)
            assert bool(found_c[i]) == want, (i, "".join(seq), bad)


# ---- java.util.HashMap with treeified bins (TREEIFY_THRESHOLD): the history decides the iteration order ----------------------------
def _crowded_keys(rng, k, n_total, hot_bins, hot_each, mask=4095):
    """strings of which `hot_each` share each of `hot_bins` residues of the spread String.hashCode modulo mask + 1"""
    def spread(s):
        h = pyref.java_hash(s)
        return h ^ (h >> 16)
    hot = {rng.randrange(mask + 1): [] for _ in range(hot_bins)}
    seen = set()
    while any(len(v) < hot_each for v in hot.values()):
        s = "".join(rng.choice("ACGT") for _ in range(k))
        b = spread(s) & mask
        if s not in seen and b in hot and len(hot[b]) < hot_each:
            hot[b].append(s)
            seen.add(s)
    keys = [x for v in hot.values() for x in v]
    while len(keys) < n_total:
        s = "".join(rng.choice("ACGT") for _ in range(k))
        if s not in seen:
            seen.add(s)
            keys.append(s)
    rng.shuffle(keys)
    return keys


def _check_model(m):
    """Structural invariants of the model after any number of puts: every key once, every node in the bin of its hash, tree bins are
    valid red-black trees over exactly the nodes of the bin's list, ordered by (hash as int, key), with the root first in the list."""
    seen = 0
    for idx, e in enumerate(m.tab):
        if e == -1:
            continue
        lst, x = [], e
        while x != -1:
            assert m.h[x] & (len(m.tab) - 1) == idx
            lst.append(x)
            x = m.nxt[x]
        seen += len(lst)
        if not m.tree[e]:
            assert not any(m.tree[x] for x in lst) and len(lst) <= 8
            continue
        assert all(m.tree[x] for x in lst) and len(lst) >= 7
        assert all(m.prv[b] == a for a, b in zip(lst, lst[1:]))
        assert m.prv[e] == -1 and m.par[e] == -1 and not m.red[e]
        nodes = []

        def walk(x, lo, hi):
            if x == -1:
                return 1
            nodes.append(x)
            key = (m._sgn(m.h[x]), m.key[x])
            assert (lo is None or lo < key) and (hi is None or key < hi)
            for c in (m.left[x], m.right[x]):
                assert c == -1 or (m.par[c] == x and not (m.red[x] and m.red[c]))
            lh, rh = walk(m.left[x], lo, key), walk(m.right[x], key, hi)
            assert lh == rh
            return lh + (0 if m.red[x] else 1)

        walk(e, None, None)
        assert sorted(nodes) == sorted(lst)
    assert seen == len(m.key)


@pytest.mark.parametrize("seed", range(10))
def test_hashmap_model_with_treeified_bins(oracle, seed):
    """pyref.JavaHashMapModel keeps its invariants through treeify / putTreeVal / resize splits, equals the plain model when no bin
    reaches 9 nodes, and the oracle's literal C++ transcription of the JDK methods (JdkHashMap) visits the keys in the same order."""
    rng = random.Random(100 + seed)
    k = rng.choice([9, 13, 31])
    n = rng.choice([40, 200, 1500, 3000])
    keys = _crowded_keys(rng, k, n, rng.randint(1, 3), rng.randint(7, 40))
    m = pyref.JavaHashMapModel()
    for i, key in enumerate(keys):
        m.put(key)
        if i % 61 == 0:
            _check_model(m)
    _check_model(m)
    assert sorted(m.keys()) == sorted(keys)
    want = pyref.java_order(keys)
    if pyref.java_order_needs_replay(keys):
        assert want == m.keys()
    else:
        assert m.keys() == want            # no bin of 9: the replay and the counting sort are the same thing
    got = oracle.java_hashmap_order(keys)
    assert [keys[i] for i in got] == want


def test_hashmap_model_equals_plain_order_on_random_keys(oracle):
    rng = random.Random(3)
    for n in (1, 12, 13, 97, 5000):
        keys = list({"".join(rng.choice("ACGT") for _ in range(17)) for _ in range(n)})
        rng.shuffle(keys)
        m = pyref.JavaHashMapModel()
        for key in keys:
            m.put(key)
        assert not pyref.java_order_needs_replay(keys)
        assert m.keys() == pyref.java_order(keys)
        assert [keys[i] for i in oracle.java_hashmap_order(keys)] == m.keys()
