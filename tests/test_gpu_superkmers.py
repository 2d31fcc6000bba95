"""GPU parity of the super-k-mer front end (k >= 22: k_extract<.., W> -> minimizer bins -> k_sk_expand, mc_sk.cu) against the oracle.

The front end changes HOW the k-mer instances of the reads reach the table (16-byte records of consecutive k-mers, identical records
counted once and unpacked with their multiplicity), never WHICH instances: every table here must be bit-identical to the oracle's, and
to the table the flat-key path builds from the same input."""
import random

import numpy as np
import pytest

from helpers import make_fasta, make_fastq, rand_seq
from oracle import pyref
from oracle_ffi import FMT_FASTA, FMT_FASTQ33

pytestmark = pytest.mark.gpu


def assert_same_table(gpu_table, orc_table):
    gk, gc = gpu_table.dump(0)
    ok, oc = orc_table.dump(-32768)
    assert gk.size == ok.size, f"distinct k-mers: gpu {gk.size} oracle {ok.size}"
    assert np.array_equal(gk, ok)
    assert np.array_equal(gc, oc)
    assert gpu_table.size() == orc_table.size()
    assert gpu_table.stats() == orc_table.stats()


def genome_reads(rng, fmt, genome_len=3000, n_reads=1500, read_len=100, p_err=0.01, p_n=0.002, p_bang=0.002, jitter=0):
    """Reads of both strands over a small genome with a repeat and a homopolymer: deep coverage, so most super-k-mers repeat."""
    g = rand_seq(rng, genome_len)
    g = g[:genome_len // 2] + g[50:250] + "A" * 90 + g[genome_len // 2:] + "ACACACACACACACACACACACACACACACACACACACACACACAC"
    out = []
    for i in range(n_reads):
        L = read_len + (rng.randint(-jitter, jitter) if jitter else 0)
        s = rng.randrange(0, len(g) - L)
        r = list(g[s:s + L])
        for p in range(L):
            x = rng.random()
            if x < p_err:
                r[p] = rng.choice("ACGT")
            elif x < p_err + p_n:
                r[p] = "N"
        r = "".join(r)
        if rng.random() < 0.5:
            r = pyref.rc_str(r) if "N" not in r else r[::-1].translate(str.maketrans("ACGTN", "TGCAN"))
        if fmt == FMT_FASTA:
            out.append(f">r{i}\n{r}\n")
        else:
            q = "".join("!" if rng.random() < p_bang else chr(rng.randint(34, 74)) for _ in range(L))
            out.append(f"@r{i}\n{r}\n+\n{q}\n")
    return "".join(out).encode()


@pytest.fixture(scope="module")
def ctx():
    from metacherchant_b200 import Context
    c = Context(0)
    c.set_tuning("superkmers", 1)   # off by default (slower than the flat path at cfg2, profiles/r2_NOTES.md section 6)
    yield c
    c.close()


@pytest.mark.parametrize("k", [22, 23, 26, 27, 29, 30, 31])
@pytest.mark.parametrize("fmt", [FMT_FASTQ33, FMT_FASTA])
def test_deep_coverage_tables_match_oracle(ctx, oracle, k, fmt):
    rng = random.Random(31 * k + fmt)
    data = genome_reads(rng, fmt, p_n=0.002 if fmt == FMT_FASTQ33 else 0.0005)
    o = oracle.table(k).count_bytes(data, fmt)
    for bits in (-1, 0, 5):
        ctx.set_tuning("sk_bits", bits)
        g = ctx.new_table(k).count_bytes(data, fmt)
        assert_same_table(g, o)
        g.free()
    ctx.set_tuning("sk_bits", -1)


@pytest.mark.parametrize("k", [22, 31])
@pytest.mark.parametrize("fmt", [FMT_FASTQ33, FMT_FASTA])
def test_noisy_random_records_match_oracle(ctx, oracle, k, fmt):
    """The generic builders of the suite: N, '.', zero qualities, lower case, empty and very short records."""
    rng = random.Random(k + fmt)
    data = make_fastq(rng, 800, max_len=160) if fmt == FMT_FASTQ33 else make_fasta(rng, 600, max_len=200)
    g = ctx.new_table(k).count_bytes(data, fmt)
    o = oracle.table(k).count_bytes(data, fmt)
    assert_same_table(g, o)


@pytest.mark.parametrize("parts", [1, 4, 32])
@pytest.mark.parametrize("k", [23, 31])
def test_lanes_per_record_and_long_reads(ctx, oracle, k, parts):
    """Long reads are shared by 4 or 32 lanes: super-k-mers are cut at the lanes' borders; homopolymers fill records to their maximum."""
    rng = random.Random(k * parts)
    data = genome_reads(rng, FMT_FASTQ33, genome_len=6000, n_reads=120, read_len=1500, jitter=700) + \
        genome_reads(rng, FMT_FASTQ33, genome_len=500, n_reads=60, read_len=90)
    ctx.set_tuning("extract_parts", parts)
    try:
        g = ctx.new_table(k).count_bytes(data, FMT_FASTQ33)
        o = oracle.table(k).count_bytes(data, FMT_FASTQ33)
        assert_same_table(g, o)
    finally:
        ctx.set_tuning("extract_parts", 0)


@pytest.mark.parametrize("cap", [1, 8, 64])
def test_full_bins_spill_and_full_spill_lists_fall_back_to_flat_keys(oracle, cap):
    """Tiny bins: most records take the spill list; with cap = 1 that fills up too and the chunk is redone with flat keys."""
    from metacherchant_b200 import Context
    rng = random.Random(cap)
    data = genome_reads(rng, FMT_FASTQ33, n_reads=2500)
    c = Context(0, chunk_bytes=65536)
    try:
        c.set_tuning("superkmers", 1)
        c.set_tuning("sk_bits", 4)
        c.set_tuning("sk_cap", cap)
        g = c.new_table(31).count_bytes(data, FMT_FASTQ33)
        o = oracle.table(31).count_bytes(data, FMT_FASTQ33)
        assert_same_table(g, o)
    finally:
        c.close()


@pytest.mark.parametrize("chunk", [8192, 100000])
def test_chunks_accumulate_in_the_bins_and_the_flat_path_agrees(oracle, chunk):
    from metacherchant_b200 import Context
    rng = random.Random(chunk)
    data = genome_reads(rng, FMT_FASTQ33, n_reads=4000, read_len=150)
    c = Context(0, chunk_bytes=chunk)
    try:
        c.set_tuning("superkmers", 1)
        o = oracle.table(31).count_bytes(data, FMT_FASTQ33)
        g = c.new_table(31).count_bytes(data, FMT_FASTQ33)
        assert_same_table(g, o)
        # a second input into the same table, device resident, small device chunks
        more = genome_reads(rng, FMT_FASTA, n_reads=1000, read_len=120)
        c.set_tuning("device_chunk_bytes", chunk)
        d = c.device_alloc(len(more))
        c.h2d(d, np.frombuffer(more, dtype=np.uint8))
        g.count_device(d, len(more), FMT_FASTA)
        o.count_bytes(more, FMT_FASTA)
        assert_same_table(g, o)
        c.device_free(d)
        # flat keys from the same input
        c.set_tuning("superkmers", 0)
        f = c.new_table(31).count_bytes(data, FMT_FASTQ33)
        f.count_bytes(more, FMT_FASTA)
        fk, fc = f.dump(0)
        gk, gc = g.dump(0)
        assert np.array_equal(fk, gk) and np.array_equal(fc, gc)
    finally:
        c.close()


def test_an_input_error_applies_nothing_of_its_chunk(oracle):
    """The records of a chunk that ends in a reader error leave the bins again (cursor snapshot): what was counted before stays exact."""
    from metacherchant_b200 import Context, McError
    rng = random.Random(5)
    good = genome_reads(rng, FMT_FASTQ33, n_reads=500)
    bad = genome_reads(rng, FMT_FASTQ33, n_reads=300) + b"@x\nACGTRACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
    c = Context(0)
    try:
        c.set_tuning("superkmers", 1)
        g = c.new_table(31).count_bytes(good, FMT_FASTQ33)
        with pytest.raises(McError):
            g.count_bytes(bad, FMT_FASTQ33)
        o = oracle.table(31).count_bytes(good, FMT_FASTQ33)
        gk, gc = g.dump(0)
        ok, oc = o.dump(-32768)
        assert np.array_equal(gk, ok) and np.array_equal(gc, oc)
    finally:
        c.close()


def test_saturation_through_weighted_updates(ctx, oracle):
    """One read repeated 40 000 times: every record of a bin is the same, the weights pass 32767 and the counts must clamp there."""
    rng = random.Random(9)
    r = rand_seq(rng, 60)
    data = (f"@r\n{r}\n+\n{'I' * 60}\n" * 40000).encode()
    g = ctx.new_table(31).count_bytes(data, FMT_FASTQ33)
    o = oracle.table(31).count_bytes(data, FMT_FASTQ33)
    assert_same_table(g, o)
    _, gc = g.dump(0)
    assert gc.max() == 32767
