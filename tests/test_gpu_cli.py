"""GPU: the `metacherchant` CLI end to end (reads on disk -> output directory) against the oracle."""
import os
import random
import subprocess

import pytest

from helpers import make_fastq, small_graph
from oracle_ffi import FMT_FASTA

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "metacherchant_b200", "csrc", "metacherchant")
FILES = ("graph.txt", "env.txt", "seqs.fasta", "graph.gfa", "tsvs/nodes.tsv", "tsvs/edges.tsv")


def test_environment_finder_cli(oracle, tmp_path):
    k = 15
    rng = random.Random(1)
    genome, data = small_graph(rng, k, genome_len=3000, n_reads=4000, read_len=70)
    (tmp_path / "a.fasta").write_bytes(data[: len(data) // 2 + 30])       # cut inside a record on purpose? no: keep whole records
    # split at a record boundary into two files, one FASTA and one FASTQ built from the same reads
    recs = data.decode().split(">")[1:]
    half = len(recs) // 2
    (tmp_path / "a.fasta").write_text("".join(">" + r for r in recs[:half]))
    fq = "".join(f"@{r.splitlines()[0]}\n{r.splitlines()[1]}\n+\n{'I' * len(r.splitlines()[1])}\n" for r in recs[half:])
    (tmp_path / "b.fq").write_text(fq)
    (tmp_path / "genes.fa").write_text(f">geneA\n{genome[500:600]}\n>geneB\n{genome[1500:1570]}\n")
    r = subprocess.run([CLI, "--tool", "environment-finder", "-k", str(k), "--reads", str(tmp_path / "a.fasta"), str(tmp_path / "b.fq"),
                        "--seq", str(tmp_path / "genes.fa"), "--output", str(tmp_path / "out"), "--work-dir", str(tmp_path / "wd"),
                        "--maxkmers=300", "--coverage", "2", "--bothdirs", "False", "--chunklength", "10", "-p", "4"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Loading file a.fasta..." in r.stderr and "reads added" in r.stderr and "Hashtable size:" in r.stderr
    assert "Finished processing all sequences!" in r.stderr
    o = oracle.table(k, threads=4)
    o.count_file(str(tmp_path / "a.fasta"))
    o.count_file(str(tmp_path / "b.fq"))
    assert f"Hashtable size: {o.size()} kmers" in r.stderr
    assert o.env_find(str(tmp_path / "genes.fa"), 2, 300, None, False, 10, str(tmp_path / "orc")) == 2
    for n_ext in o.extend_counts(str(tmp_path / "genes.fa"), 2, 300, None, False):
        assert f"Extending endings by {n_ext} kmers" in r.stderr
    for gene in ("geneA", "geneB"):
        for f in FILES:
            assert (tmp_path / "out" / gene / f).read_bytes() == (tmp_path / "orc" / gene / f).read_bytes(), (gene, f)
    # multi tool over the two environments
    envs = [str(tmp_path / "out" / g / "env.txt") for g in ("geneA", "geneB")]
    r = subprocess.run([CLI, "-t", "environment-finder-multi", "--env", *envs, "--seq", str(tmp_path / "genes.fa"), "-g", "2", "-o", str(tmp_path / "multi"),
                        "-w", str(tmp_path / "wd")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    oracle.env_multi(envs, str(tmp_path / "genes.fa"), 2, str(tmp_path / "orc_multi"))
    for f in ("graph.gfa", "seqs.fasta", "gene.fasta", "Jacard_sym.txt", "Jacard_alt.txt"):
        assert (tmp_path / "multi" / f).read_bytes() == (tmp_path / "orc_multi" / f).read_bytes(), f
    # kmer-counter
    r = subprocess.run([CLI, "-t", "kmer-counter", "-k", str(k), "-i", str(tmp_path / "a.fasta"), str(tmp_path / "b.fq"), "-w", str(tmp_path / "wd2")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o.print_kmers(0, str(tmp_path / "o.kmers.bin"), str(tmp_path / "o.stat.txt"))
    assert (tmp_path / "wd2" / "kmers" / "a.kmers.bin").read_bytes() == (tmp_path / "o.kmers.bin").read_bytes()


@pytest.mark.parametrize("flags", [["--merge", "true"], ["--merge", "--hicseq", "HIC"], ["--trim"], ["--merge", "true", "--trim", "--hicseq", "HIC", "--bothdirs", "true"],
                                   ["--hicseq", "HIC"]])
def test_cli_merge_hicseq_trim(oracle, tmp_path, flags):
    """SURVEY 8f rows 2 and 3 through the command line, byte for byte against the oracle."""
    k = 13
    rng = random.Random(5)
    genome, data = small_graph(rng, k, genome_len=2500, n_reads=3500, read_len=60)
    (tmp_path / "r.fasta").write_bytes(data)
    (tmp_path / "genes.fa").write_text(f">geneA\n{genome[500:580]}\n>geneB\n{genome[560:640]}\n")
    (tmp_path / "hic.fa").write_text(f">hicA\n{genome[1500:1560]}\n>hicB\n{genome[2000:2050]}\n")
    flags = [str(tmp_path / "hic.fa") if f == "HIC" else f for f in flags]
    merge, trim, hic = "--merge" in flags, "--trim" in flags, "--hicseq" in flags
    both = "--bothdirs" in flags
    r = subprocess.run([CLI, "-k", str(k), "-i", str(tmp_path / "r.fasta"), "--seq", str(tmp_path / "genes.fa"), "-o", str(tmp_path / "out"),
                        "-w", str(tmp_path / "wd"), "--maxkmers", "200", "--coverage", "2", *flags], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    o = oracle.table(k, threads=4)
    o.count_file(str(tmp_path / "r.fasta"))
    n = o.env_find_ex(str(tmp_path / "genes.fa"), str(tmp_path / "hic.fa") if hic else None, 2, 200, None, both, 1, merge, trim, str(tmp_path / "orc"))
    dirs = ["merged"] if merge else (["hicA", "hicB"] if hic else ["geneA", "geneB"])
    assert n == len(dirs)
    if merge:
        assert "Finding single environment for 2 sequences" in r.stderr
    for d in dirs:
        for f in FILES:
            assert (tmp_path / "out" / d / f).read_bytes() == (tmp_path / "orc" / d / f).read_bytes(), (d, f)


def test_cli_reports_reader_errors(tmp_path):
    (tmp_path / "bad.fastq").write_text("@r\nACGT\n+\nIII\n")
    (tmp_path / "g.fa").write_text(">g\nACGTACGTACGTACGTACGT\n")
    r = subprocess.run([CLI, "-k", "5", "-i", str(tmp_path / "bad.fastq"), "--seq", str(tmp_path / "g.fa"), "-o", str(tmp_path / "o"), "--maxkmers", "10",
                        "-w", str(tmp_path / "wd")], capture_output=True, text=True)
    assert r.returncode == 1 and "Bad DnaQ record" in r.stderr
    r = subprocess.run([CLI, "-k", "5", "-i", str(tmp_path / "missing.fastq"), "--seq", str(tmp_path / "g.fa"), "-o", str(tmp_path / "o"), "--maxkmers", "10",
                        "-w", str(tmp_path / "wd")], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to read from file" in r.stderr
    r = subprocess.run([CLI, "-k", "5", "-i", str(tmp_path / "reads.txt"), "--seq", str(tmp_path / "g.fa"), "-o", str(tmp_path / "o"), "--maxkmers", "10",
                        "-w", str(tmp_path / "wd")], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to read from file" in r.stderr
