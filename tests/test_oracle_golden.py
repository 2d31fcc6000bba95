"""Pins the oracle against the reference's own committed example outputs:
Hi-C_pipline/example_work_dir/output/1/merged/{graph.txt,graph.gfa,tsvs/nodes.tsv} (fixtures in tests/golden)."""
import lzma
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hic_example1")

# the 16 unitigs of the golden GFA
GOLD_LEN = [35, 1800, 32469, 63, 45, 3715, 28251, 158, 15171, 35, 117, 117, 7661, 34, 31, 4350]
GOLD_KC = [3565, 38760, 675269, 2972, 1441, 68020, 574432, 7212, 327103, 3191, 2668, 2927, 154747, 3951, 3441, 96912]


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def _canon(s):
    return min(s, _rc(s))


@pytest.fixture(scope="module")
def graph_txt(tmp_path_factory):
    d = tmp_path_factory.mktemp("gold")
    p = d / "graph.txt"
    with lzma.open(os.path.join(GOLD, "graph.txt.xz"), "rb") as f:
        p.write_bytes(f.read())
    return str(p)


def _s_lines(path):
    out = []
    for line in open(path):
        t = line.rstrip("\n").split("\t")
        if t[0] == "S":
            tags = dict(x.split(":", 2)[::2] for x in t[3:])
            out.append((t[1], t[2], int(tags["LN"]), int(tags["KC"]), tags.get("CL")))
    return out


def _l_lines(path):
    return [tuple(line.rstrip("\n").split("\t")) for line in open(path) if line.startswith("L\t")]


def test_golden_fixture_is_what_survey_says():
    s = _s_lines(os.path.join(GOLD, "graph.gfa"))
    assert [x[2] for x in s] == GOLD_LEN
    assert [x[3] for x in s] == GOLD_KC
    assert len(_l_lines(os.path.join(GOLD, "graph.gfa"))) == 41


def test_java_hashmap_order_reproduces_reference_graph_txt(oracle, graph_txt, tmp_path):
    """graph.txt is written in java.util.HashMap<String,Integer> iteration order
    (OneSequenceCalculator.java:303-305).  Loading it and re-emitting it through the oracle's
    HashMap emulation must give back the same bytes."""
    out = tmp_path / "o"
    oracle.env_from_graph_txt(graph_txt, "", 31, 1, str(out))
    assert (out / "graph.txt").read_bytes() == open(graph_txt, "rb").read()
    assert (out / "env.txt").read_bytes() == open(graph_txt, "rb").read()


def test_java_string_hash(oracle):
    assert oracle.java_hash("") == 0
    assert oracle.java_hash("a") == 97
    assert oracle.java_hash("hello") == 99162322
    assert oracle.java_hash("CAGCTGAGTCGCGGAGGCGGTGCGCGGCAGC") == oracle.java_hash("CAGCTGAGTCGCGGAGGCGGTGCGCGGCAGC")
    # int32 wraparound
    assert oracle.java_hash("The quick brown fox jumps over the lazy dog") == -609428141


def test_stage_d_unitigs_match_golden_gfa(oracle, graph_txt, tmp_path):
    """Replaying initializeStructures+doMerge+GFAWriter on the reference's graph.txt without a gene flag
    must give the golden's 16 unitigs as a multiset of rc-canonical sequences, with the golden KC values
    (KC formula, GFAWriter.java:90-95).  Ids differ: the golden was written by an older revision."""
    out = tmp_path / "o"
    oracle.env_from_graph_txt(graph_txt, "", 31, 1, str(out))
    mine = _s_lines(str(out / "graph.gfa"))
    gold = _s_lines(os.path.join(GOLD, "graph.gfa"))
    assert len(mine) == 16
    assert sorted((_canon(s[1]), s[2], s[3]) for s in mine) == sorted((_canon(s[1]), s[2], s[3]) for s in gold)
    # links: 40 of the golden's 41 L-lines are true (k-1)-overlap links (the 41st is not producible by GFAWriter)
    mine_l = _l_lines(str(out / "graph.gfa"))
    assert len(mine_l) == 40
    # nodes.tsv lists the same sequences
    tsv = [l.rstrip("\n").split("\t") for l in open(out / "tsvs" / "nodes.tsv")][1:]
    assert sorted(_canon(t[2]) for t in tsv) == sorted(_canon(s[1]) for s in gold)
    assert all(int(t[1]) == len(t[2]) for t in tsv)


def test_stage_d_links_are_overlaps_and_symmetric(oracle, graph_txt, tmp_path):
    out = tmp_path / "o"
    oracle.env_from_graph_txt(graph_txt, "", 31, 1, str(out))
    seqs = {s[0]: s[1] for s in _s_lines(str(out / "graph.gfa"))}
    links = _l_lines(str(out / "graph.gfa"))

    def oriented(i, sign):
        return seqs[i] if sign == "+" else _rc(seqs[i])

    pairs = set()
    for _, a, sa, b, sb, ov in links:
        assert ov == "30M"
        assert oriented(a, sa)[-30:] == oriented(b, sb)[:30]
        pairs.add((a, sa, b, sb))
    flip = {"+": "-", "-": "+"}
    for a, sa, b, sb in pairs:   # every edge appears with its reverse-complement twin
        assert (b, flip[sb], a, flip[sa]) in pairs


def test_stage_d_with_current_gene_rule(oracle, graph_txt, tmp_path):
    """With the example target (61 bp) and the CURRENT isGeneNode rule (OneSequenceCalculator.java:421-432)
    the same environment splits into 18 S-lines / 44 L-lines."""
    target = open(os.path.join(GOLD, "seq.fasta")).read().strip()
    out = tmp_path / "o"
    oracle.env_from_graph_txt(graph_txt, target, 31, 10, str(out))
    s = _s_lines(str(out / "graph.gfa"))
    assert len(s) == 18
    assert len(_l_lines(str(out / "graph.gfa"))) == 44
    gene = [x for x in s if x[0].endswith("_start")]
    assert gene and all(x[4] == "GREEN" for x in gene)
    assert all(x[4] is None for x in s if not x[0].endswith("_start"))
    # seqs.fasta honours --chunklength and the header format
    heads = [l for l in open(out / "seqs.fasta") if l.startswith(">")]
    assert all(h.startswith("> Id") and " Length:" in h and " Neighbors:[" in h for h in heads)


def _golden_table_pairs(graph_txt, k=31):
    """(canonical 2-bit keys, counts) of the k-mers the reference's committed graph.txt lists"""
    import numpy as np
    from oracle import pyref
    lines = [l.split(" ") for l in open(graph_txt) if l.strip()]
    keys = np.array([min(pyref.encode(a), pyref.revcomp_code(pyref.encode(a), k)) for a, _ in lines], dtype=np.uint64)
    return keys, np.array([int(b) for _, b in lines], dtype=np.int16), [f"{a} {int(b)}" for a, b in lines]


def test_bfs_from_the_example_target_reaches_exactly_the_golden_environment(oracle, graph_txt, tmp_path):
    """Stage C1 against a reference-produced artefact, at the level of SETS.  The golden graph.txt was written by
    `--k 31 --coverage 5 --seq example/seq.fasta --maxradius 100000 --bothdirs False --chunklength 10 --merge true`
    (Hi-C_pipline/HiCEnvironmentFinder.sh:57) from WGS reads that are not in the repository.  A count table holding exactly the
    golden k-mers with their counts stands in for the reads' table: every k-mer outside the environment was either not solid or not
    reachable, so the two backward / forward passes from the target's 31 seed k-mers must visit every golden k-mer (each listed
    k-mer is reachable by a directed walk through listed k-mers) and write the same 93 572 `KMER COUNT` lines.
    What this does NOT pin: the order of the lines inside a HashMap bucket -- the committed file was written by an older revision
    (its node ids differ too): 56 % of the multi-entry buckets agree with the insertion order the current code implies,
    30 % are reversed; no variant of pass order / keySet-vs-queue insertion reproduces it."""
    keys, counts, want = _golden_table_pairs(graph_txt)
    assert counts.min() == 5                                   # --coverage 5: nothing below the threshold was ever listed
    t = oracle.table(31, threads=2)
    t.add_pairs(keys, counts)
    out = tmp_path / "out"
    t.env_find_ex(os.path.join(GOLD, "seq.fasta"), None, 5, None, 100000, False, 10, True, False, str(out))
    got = [l.rstrip("\n") for l in open(out / "merged" / "graph.txt")]
    assert len(got) == 93572 and sorted(got) == sorted(want)
    # the unitigs of the golden GFA (sequences up to reverse complement, with their KC), whatever the node numbering: 15 of the 16
    # come out unchanged; the current revision's gene flag cuts the 3 715-base one at the borders of the 61-base target into
    # 95 + 61 (GREEN) + 3 619 bases, overlapping by k - 1 = 30
    s_mine = _s_lines(out / "merged" / "graph.gfa")
    mine = set((_canon(s[1]), s[3]) for s in s_mine)
    gold = set((_canon(s[1]), s[3]) for s in _s_lines(os.path.join(GOLD, "graph.gfa")))
    assert len(gold) == 16 and len(s_mine) == 18 and len(mine & gold) == 15
    assert sorted(len(a) for a, _ in gold - mine) == [3715]
    assert sorted(len(a) for a, _ in mine - gold) == [61, 95, 3619] and 61 + 95 + 3619 - 2 * 30 == 3715
    assert [s[2] for s in s_mine if s[4] == "GREEN"] == [61]
