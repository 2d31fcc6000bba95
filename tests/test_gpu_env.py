"""GPU parity, stages C+D and M1: BFS order, environment files, unitig graph and the multi tool against the oracle."""
import lzma
import os
import random

import numpy as np
import pytest

from helpers import rand_seq, small_graph
from oracle import pyref
from oracle_ffi import FMT_FASTA, FMT_FASTQ

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hic_example1")
FILES = ("graph.txt", "env.txt", "seqs.fasta", "graph.gfa", "tsvs/nodes.tsv", "tsvs/edges.tsv")


@pytest.fixture(scope="module")
def ctx():
    from metacherchant_b200 import Context
    c = Context(0)
    yield c
    c.close()


def same_files(a, b, files=FILES):
    for f in files:
        fa, fb = os.path.join(a, f), os.path.join(b, f)
        assert os.path.exists(fa) and os.path.exists(fb), f
        assert open(fa, "rb").read() == open(fb, "rb").read(), f"{f} differs"


@pytest.mark.parametrize("k", [5, 7, 12, 31])
@pytest.mark.parametrize("direction", [-1, 0, 1])
@pytest.mark.parametrize("limits", [(None, 3), (10, None), (40, 6), (100000, None), (1, None), (None, 0), (300, 4)])
def test_bfs_discovery_order_matches_oracle(ctx, oracle, k, direction, limits):
    rng = random.Random(k * 100 + direction)
    genome, data = small_graph(rng, k, genome_len=700, n_reads=600)
    g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
    o = oracle.table(k).count_bytes(data, FMT_FASTA)
    target = genome[100:100 + max(k + 5, 30)]
    target = target + target[:k] + "N" * 0
    max_kmers, max_radius = limits
    for min_cov in (1, 2):
        want = o.bfs(target, min_cov, max_kmers, max_radius, direction)
        kmers, dist, counts = g.bfs(target, min_cov, max_kmers, max_radius, direction)
        if want is None:
            assert kmers.size == 0
            continue
        order, odist = want
        assert [oracle.decode(int(x), k) for x in kmers] == order     # same set AND same discovery order
        assert dist.tolist() == odist.tolist()
        assert counts.tolist() == [o.get(oracle.canonical(int(x), k)) for x in kmers]


@pytest.mark.parametrize("k", [9, 21])
@pytest.mark.parametrize("both", [False, True])
@pytest.mark.parametrize("limits", [(150, None), (None, 12), (100000, None)])
def test_environment_files_identical_to_oracle(ctx, oracle, tmp_path, k, both, limits):
    rng = random.Random(k + 2 * both)
    genome, data = small_graph(rng, k, genome_len=2000, n_reads=2500, read_len=60)
    g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
    o = oracle.table(k).count_bytes(data, FMT_FASTA)
    fa = tmp_path / "seq.fasta"
    t1, t2 = genome[300:380], genome[900:960]
    fa.write_text(f">gene one\n{t1[:30]}\n{t1[30:]}\n>second\n{t2}\n>absent\n{'ACGT' * 12}\n")
    max_kmers, max_radius = limits
    n = o.env_find(str(fa), 2, max_kmers, max_radius, both, 20, str(tmp_path / "orc"))
    targets = ctx.load_targets(fa)
    assert [t[0] for t in targets] == ["gene one", "second", "absent"]
    solid = g.filter(2)
    found = ctx.find_environments(solid, [t[1] for t in targets], [t[0] for t in targets], min_cov=2, max_kmers=max_kmers, max_radius=max_radius,
                                  both_dirs=both, chunk_len=20, out_dir=tmp_path / "gpu")
    assert sum(found) == n
    # "Extending endings by N kmers" (OneSequenceCalculator.java:291)
    assert ctx.extend_counts() == o.extend_counts(str(fa), 2, max_kmers, max_radius, both)
    for name, ok in zip(("gene one", "second", "absent"), found):
        if ok:
            same_files(str(tmp_path / "gpu" / name), str(tmp_path / "orc" / name))
        else:
            assert not (tmp_path / "gpu" / name).exists() and not (tmp_path / "orc" / name).exists()
    # the unfiltered table gives the same answer as the coverage-compacted one
    found2 = ctx.find_environments(g, [targets[0][1]], ["g1"], min_cov=2, max_kmers=max_kmers, max_radius=max_radius, both_dirs=both, chunk_len=20,
                                   out_dir=tmp_path / "gpu2")
    if found2[0]:
        same_files(str(tmp_path / "gpu2" / "g1"), str(tmp_path / "orc" / "gene one"))


def test_small_k_environments_with_hairpins_and_cycles(ctx, oracle, tmp_path):
    """k=5/6: reverse-palindromic overlaps, tiny cycles and (even k) palindromic k-mers are everywhere."""
    for k in (5, 6, 8):
        rng = random.Random(k)
        genome, data = small_graph(rng, k, genome_len=500, n_reads=400, read_len=30)
        g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
        o = oracle.table(k).count_bytes(data, FMT_FASTA)
        fa = tmp_path / f"s{k}.fasta"
        fa.write_text(f">t\n{genome[50:90]}\n")
        o.env_find(str(fa), 1, 400, None, True, 1, str(tmp_path / f"orc{k}"))
        (name, seq), = ctx.load_targets(fa)
        ctx.find_environments(g, [seq], [name], min_cov=1, max_kmers=400, both_dirs=True, out_dir=tmp_path / f"gpu{k}")
        same_files(str(tmp_path / f"gpu{k}" / "t"), str(tmp_path / f"orc{k}" / "t"))


def test_stage_d_on_the_reference_golden_environment(ctx, oracle, tmp_path):
    """The reference's committed graph.txt through the GPU unitig stage: byte-identical to the oracle (which is
    pinned on the golden GFA), 16 unitigs without gene flag, 18 S / 44 L with the example target."""
    p = tmp_path / "graph.txt"
    with lzma.open(os.path.join(GOLD, "graph.txt.xz"), "rb") as f:
        p.write_bytes(f.read())
    target = open(os.path.join(GOLD, "seq.fasta")).read().strip()
    for tgt, tag in (("", "plain"), (target, "gene")):
        ctx.env_from_graph_txt(p, tgt, 31, 10, tmp_path / f"gpu_{tag}")
        oracle.env_from_graph_txt(str(p), tgt, 31, 10, str(tmp_path / f"orc_{tag}"))
        same_files(str(tmp_path / f"gpu_{tag}"), str(tmp_path / f"orc_{tag}"))
    s_lines = [l for l in open(tmp_path / "gpu_plain" / "graph.gfa") if l.startswith("S\t")]
    assert len(s_lines) == 16
    gold = sorted(int(l.split("KC:i:")[1].split()[0]) for l in open(os.path.join(GOLD, "graph.gfa")) if l.startswith("S\t"))
    assert sorted(int(l.split("KC:i:")[1].split()[0]) for l in s_lines) == gold
    assert (tmp_path / "gpu_plain" / "graph.txt").read_bytes() == p.read_bytes()
    gfa = open(tmp_path / "gpu_gene" / "graph.gfa").read().splitlines()
    assert sum(l.startswith("S\t") for l in gfa) == 18 and sum(l.startswith("L\t") for l in gfa) == 44


def test_bfs_from_the_example_target_reaches_exactly_the_golden_environment(ctx, oracle, tmp_path):
    """The GPU's BFS + unitigs on the table of the reference's committed environment (see the oracle's test of the same name in
    tests/test_oracle_golden.py): from the 31 seed k-mers of example/seq.fasta with `--maxradius 100000 --coverage 5 --merge true`
    the two passes visit exactly the 93 572 golden k-mers; every output file is byte-identical to the oracle's."""
    from test_oracle_golden import _golden_table_pairs
    p = tmp_path / "graph.txt"
    with lzma.open(os.path.join(GOLD, "graph.txt.xz"), "rb") as f:
        p.write_bytes(f.read())
    keys, counts, want = _golden_table_pairs(str(p))
    g = ctx.new_table(31)
    g.add_pairs(keys, counts)
    o = oracle.table(31, threads=2)
    o.add_pairs(keys, counts)
    fa = os.path.join(GOLD, "seq.fasta")
    o.env_find_ex(fa, None, 5, None, 100000, False, 10, True, False, str(tmp_path / "orc"))
    targets = ctx.load_targets(fa)
    found = ctx.find_environments_ex(g, [x[1] for x in targets], [x[0] for x in targets], min_cov=5, max_radius=100000, both_dirs=False,
                                     chunk_len=10, merge=True, out_dir=tmp_path / "gpu")
    assert found == [True]
    same_files(str(tmp_path / "gpu" / "merged"), str(tmp_path / "orc" / "merged"))
    got = [l.rstrip("\n") for l in open(tmp_path / "gpu" / "merged" / "graph.txt")]
    assert len(got) == 93572 and sorted(got) == sorted(want)
    g.free()


def test_stage_d_with_treeified_hashmap_bins(ctx, oracle, tmp_path):
    """A graph.txt whose k-mers crowd a few java.util.HashMap bins (9 .. 40 keys with the same residue of the spread hash): the bins
    are treeified and the iteration order -- the line order of graph.txt / env.txt and the node ids of the unitig files -- depends on
    the history of the puts.  The library (JavaHashReplay) and the oracle (a literal JdkHashMap) write the same bytes, and the order
    really differs from the plain bucket order."""
    from test_host_logic import crowded_kmers
    rng = random.Random(77)
    k = 21
    genome = rand_seq(rng, 1500)
    path_kmers = [genome[i:i + k] for i in range(len(genome) - k + 1)]
    # graph.txt holds ASCII-canonical k-mers only (a node's k-mer is looked up in its normal form): the crowd is drawn among those
    crowd = crowded_kmers(rng, k, 600, 3, 30, 4095, accept=lambda x: x == pyref.normalize(x))
    kmers = list(dict.fromkeys([pyref.normalize(x) for x in path_kmers] + crowd))
    rng.shuffle(kmers)
    p = tmp_path / "graph.txt"
    p.write_text("".join(f"{x} {rng.randint(1, 300)}\n" for x in kmers))
    loaded = [l.split(" ")[0] for l in open(p)]
    assert pyref.java_order_needs_replay(loaded)
    ctx.env_from_graph_txt(p, genome[100:400], k, 1, tmp_path / "gpu")
    oracle.env_from_graph_txt(str(p), genome[100:400], k, 1, str(tmp_path / "orc"))
    same_files(str(tmp_path / "gpu"), str(tmp_path / "orc"))
    got = [l.split(" ")[0] for l in open(tmp_path / "gpu" / "graph.txt")]
    assert got == pyref.java_order(loaded)
    cap = 16
    while len(loaded) > 0.75 * cap:
        cap *= 2
    plain = sorted(loaded, key=lambda x: (lambda h: (h ^ (h >> 16)) & (cap - 1))(pyref.java_hash(x)))
    assert got != plain and sorted(got) == sorted(plain)


def test_multi_tool_matches_oracle(ctx, oracle, tmp_path):
    k = 11
    rng = random.Random(3)
    genome, data1 = small_graph(rng, k, genome_len=1500, n_reads=1500, read_len=50)
    # time point 2: part of the genome is gone and coverage differs
    genome2 = genome[:700] + pyref.rc_str(genome[900:])
    reads2 = []
    for i in range(1200):
        s = rng.randrange(0, len(genome2) - 50)
        reads2.append(f">r{i}\n{genome2[s:s + 50]}\n")
    data2 = "".join(reads2).encode()
    target = genome[300:380]
    fa = tmp_path / "seq.fasta"
    fa.write_text(f">first\nACGTACGTACGTAGCTAGCTAGCATCGA\n>gene two\n{target}\n")
    envs = []
    for i, data in enumerate((data1, data2, data1 + data2)):
        g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
        ctx.find_environments(g, [target], [f"tp{i}"], min_cov=2, max_kmers=500, both_dirs=False, out_dir=tmp_path / "envs")
        envs.append(str(tmp_path / "envs" / f"tp{i}" / "env.txt"))
    files = ("seqs.fasta", "graph.gfa", "gene.fasta", "Jacard_sym.txt", "Jacard_alt.txt")
    for n_env in (2, 3, 1):
        ctx.multi_environment(envs[:n_env], fa, 2, tmp_path / f"gpu_m{n_env}")
        oracle.env_multi(envs[:n_env], str(fa), 2, str(tmp_path / f"orc_m{n_env}"))
        same_files(str(tmp_path / f"gpu_m{n_env}"), str(tmp_path / f"orc_m{n_env}"), files)
    gfa = open(tmp_path / "gpu_m2" / "graph.gfa").read()
    assert "CL:Z:#00ff00" in gfa and ("CL:Z:#ff0000" in gfa or "CL:Z:#0000ff" in gfa)


def test_kmers_bin_round_trip(ctx, oracle, tmp_path):
    rng = random.Random(17)
    _, data = small_graph(rng, 13, genome_len=1500, n_reads=1500, read_len=50)
    g = ctx.new_table(13).count_bytes(data, FMT_FASTA)
    o = oracle.table(13).count_bytes(data, FMT_FASTA)
    good = g.print_kmers(0, tmp_path / "g.kmers.bin", tmp_path / "g.stat.txt")
    assert good == o.print_kmers(0, str(tmp_path / "o.kmers.bin"), str(tmp_path / "o.stat.txt")) == g.size()
    assert (tmp_path / "g.kmers.bin").read_bytes() == (tmp_path / "o.kmers.bin").read_bytes()
    assert (tmp_path / "g.stat.txt").read_bytes() == (tmp_path / "o.stat.txt").read_bytes()
    g2 = ctx.new_table(13)
    g2.load_kmers(tmp_path / "g.kmers.bin", 1)       # keep freq > 1
    o2 = oracle.table(13)
    o2.load_kmers(str(tmp_path / "o.kmers.bin"), 1)
    gk, gc = g2.dump()
    ok, oc = o2.dump()
    assert np.array_equal(gk, ok) and np.array_equal(gc, oc)


def test_synthetic_config1_environment(ctx, oracle, tmp_path):
    """A 150k-read slice of BASELINE config 1 (k=31, coverage 3 at this depth): env.txt and graph.gfa identical."""
    s = oracle.synth(genome_unit=20_000)        # 1.1 Mbp community so that 150k reads give ~20x
    data = s.generate(0, 150_000)
    fa = tmp_path / "genes.fasta"
    fa.write_text(s.targets_fasta())
    g = ctx.new_table(31).count_bytes(data, FMT_FASTQ)
    o = oracle.table(31, threads=8).count_bytes(data, FMT_FASTQ)
    for both, mk in ((False, 100000), (True, 3000)):
        tag = f"{int(both)}_{mk}"
        n = o.env_find(str(fa), 3, mk, None, both, 1, str(tmp_path / f"orc{tag}"))
        assert n == 1
        targets = ctx.load_targets(fa)
        found = ctx.find_environments(g.filter(3), [t[1] for t in targets], [t[0] for t in targets], min_cov=3, max_kmers=mk, both_dirs=both,
                                      out_dir=tmp_path / f"gpu{tag}")
        assert found == [True]
        same_files(str(tmp_path / f"gpu{tag}" / "gene0"), str(tmp_path / f"orc{tag}" / "gene0"))
        assert sum(1 for _ in open(tmp_path / f"gpu{tag}" / "gene0" / "env.txt")) > 900


# ---- --merge (+ --hicseq) and --trim ------------------------------------------------------
@pytest.mark.parametrize("k", [7, 11, 31])
@pytest.mark.parametrize("both", [False, True])
@pytest.mark.parametrize("merge,trim,hic", [(True, False, False), (True, False, True), (False, True, False), (True, True, True)])
@pytest.mark.parametrize("limits", [(60, None), (None, 5), (100000, None), (25, 3)])
def test_merge_hicseq_trim_identical_to_oracle(ctx, oracle, tmp_path, k, both, merge, trim, hic, limits):
    rng = random.Random(7 * k + both + 3 * merge + 5 * trim)
    genome, data = small_graph(rng, k, genome_len=1500, n_reads=1500, read_len=max(40, k + 10))
    g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
    o = oracle.table(k).count_bytes(data, FMT_FASTA)
    L = max(k + 8, 40)
    targets = [genome[200:200 + L], genome[200 + L - k - 3:200 + 2 * L], genome[900:900 + L]]      # the first two overlap: duplicated seeds
    hics = [genome[500:500 + L], "ACGT" * 12]
    fa, hf = tmp_path / "seq.fasta", tmp_path / "hic.fasta"
    fa.write_text("".join(f">g{i}\n{s}\n" for i, s in enumerate(targets)))
    hf.write_text("".join(f">h{i}\n{s}\n" for i, s in enumerate(hics)))
    max_kmers, max_radius = limits
    n = o.env_find_ex(str(fa), str(hf) if hic else None, 2, max_kmers, max_radius, both, 1, merge, trim, str(tmp_path / "orc"))
    names = [f"g{i}" for i in range(len(targets))]
    found = ctx.find_environments_ex(g.filter(2), targets, None if merge else names, hics if hic else (), min_cov=2, max_kmers=max_kmers,
                                     max_radius=max_radius, both_dirs=both, merge=merge, trim=trim, out_dir=tmp_path / "gpu")
    assert sum(found) == n
    for d, ok in zip(["merged"] if merge else names, found):
        if ok:
            same_files(str(tmp_path / "gpu" / d), str(tmp_path / "orc" / d))
        else:
            assert not os.path.exists(tmp_path / "orc" / d / "graph.txt")
