"""GPU parity, stages A+B: the CUDA parser/extractor/count table (through the C ABI) against the oracle."""
import random

import numpy as np
import pytest

from helpers import make_fasta, make_fastq
from oracle_ffi import FMT_FASTA, FMT_FASTQ, FMT_FASTQ33, FMT_FASTQ64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from metacherchant_b200 import Context
    c = Context(0)
    yield c
    c.close()


def assert_same_table(gpu_table, orc_table, min_count=0):
    gk, gc = gpu_table.dump(min_count)
    ok, oc = orc_table.dump(min_count if min_count else -32768)
    assert gk.size == ok.size, f"distinct k-mers: gpu {gk.size} oracle {ok.size}"
    assert np.array_equal(gk, ok)
    assert np.array_equal(gc, oc)
    if min_count == 0:
        assert gpu_table.size() == orc_table.size()


@pytest.mark.parametrize("k", [1, 3, 5, 11, 21, 31])
@pytest.mark.parametrize("variant", ["plain", "crlf", "blanks", "lower", "nonl"])
def test_fastq_table_matches_oracle(ctx, oracle, k, variant):
    rng = random.Random(1000 * k + len(variant))
    data = make_fastq(rng, 400, max_len=120, crlf=variant == "crlf", blanks=variant == "blanks", lower=variant == "lower",
                      final_newline=False if variant == "nonl" else None)
    g = ctx.new_table(k).count_bytes(data, FMT_FASTQ33)
    o = oracle.table(k).count_bytes(data, FMT_FASTQ33)
    assert_same_table(g, o)
    assert g.stats() == o.stats()


@pytest.mark.parametrize("k", [4, 15, 31])
@pytest.mark.parametrize("multiline", [False, True])
def test_fasta_table_matches_oracle(ctx, oracle, k, multiline):
    rng = random.Random(k + 7 * multiline)
    data = make_fasta(rng, 500, max_len=150, multiline=multiline)
    g = ctx.new_table(k).count_bytes(data, FMT_FASTA)
    o = oracle.table(k).count_bytes(data, FMT_FASTA)
    assert_same_table(g, o)
    assert g.stats() == o.stats()


@pytest.mark.parametrize("chunk", [4096, 5000, 65536])
@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_chunk_boundaries_do_not_matter(oracle, chunk, fmt):
    """Records straddling device chunks are carried over; the table must not depend on the chunk size."""
    from metacherchant_b200 import Context
    rng = random.Random(chunk)
    if fmt == "fastq":
        data, f = make_fastq(rng, 1500, max_len=150, blanks=True, crlf=True), FMT_FASTQ33
    else:
        data, f = make_fasta(rng, 1200, max_len=200), FMT_FASTA
    c = Context(0, chunk_bytes=chunk)
    try:
        g = c.new_table(9).count_bytes(data, f)
        o = oracle.table(9).count_bytes(data, f)
        assert_same_table(g, o)
        assert g.stats() == o.stats()
        # device-resident path with the same small chunks (by default it takes pieces as large as 32-bit offsets allow)
        c.set_tuning("device_chunk_bytes", chunk)
        d = c.device_alloc(len(data))
        c.h2d(d, np.frombuffer(data, dtype=np.uint8))
        g2 = c.new_table(9).count_device(d, len(data), f)
        assert_same_table(g2, o)
        c.device_free(d)
    finally:
        c.close()


def test_quality_offset_sniffing(ctx, oracle):
    rng = random.Random(7)
    d64 = make_fastq(rng, 200, offset=64)
    d33 = make_fastq(rng, 200, offset=33)
    for data in (d64, d33):
        assert_same_table(ctx.new_table(7).count_bytes(data, FMT_FASTQ), oracle.table(7).count_bytes(data, FMT_FASTQ))
    assert_same_table(ctx.new_table(7).count_bytes(d64, FMT_FASTQ64), oracle.table(7).count_bytes(d64, FMT_FASTQ64))


def test_phred_wrap_and_saturation(ctx, oracle):
    data = b"@r\nACGTACGTAC\n+\nIIIIaIIIII\n"
    assert_same_table(ctx.new_table(3).count_bytes(data, FMT_FASTQ33), oracle.table(3).count_bytes(data, FMT_FASTQ33))
    big = (">r\n" + "A" * 70000 + "\n").encode()
    g = ctx.new_table(3).count_bytes(big, FMT_FASTA)
    assert g.get([0, 5]).tolist() == [32767, -1]
    assert g.size() == 1
    assert_same_table(g, oracle.table(3).count_bytes(big, FMT_FASTA))


@pytest.mark.parametrize("data,fmt", [
    (b"@r\nACGT\n+\nIII\n", FMT_FASTQ33),
    (b"@r\nACGT\n+\n", FMT_FASTQ33),
    (b"@r\nACGT\n", FMT_FASTQ33),
    (b"r\nACGT\n+\nIIII\n", FMT_FASTQ33),
    (b"@r\nACXT\n+\nIIII\n", FMT_FASTQ33),
    (b"@r\nACRT\n+\nIIII\n", FMT_FASTQ33),
    (b"@r\nACGT\n+\nII\x1fI\n", FMT_FASTQ33),
    (b"@r\nACGT\n+\nII!I\n", FMT_FASTQ64),
    (b">r\nAC.T\n", FMT_FASTA),
])
def test_malformed_input_fails_like_the_reference(ctx, oracle, data, fmt):
    from metacherchant_b200 import McError
    from oracle_ffi import OracleError
    with pytest.raises(OracleError):
        oracle.table(5).count_bytes(data, fmt)
    with pytest.raises(McError) as e:
        ctx.new_table(5).count_bytes(data, fmt)
    assert e.value.code == -4


def test_tolerated_oddities(ctx, oracle):
    for data, fmt in [(b"@r\nACNT\n+\nII\x1fI\n", FMT_FASTQ33), (b"", FMT_FASTQ33), (b"\n\n\n", FMT_FASTQ33), (b"", FMT_FASTA),
                      (b">only header\n", FMT_FASTA), (b"ACGTACGT", FMT_FASTA), (b">r\nACGNNACGT\n>s\nACGTACGT\n", FMT_FASTA)]:
        assert_same_table(ctx.new_table(3).count_bytes(data, fmt), oracle.table(3).count_bytes(data, fmt))


def test_get_filter_and_add_pairs(ctx, oracle):
    rng = random.Random(11)
    data = make_fastq(rng, 3000, max_len=100, p_n=0.01, p_bang=0.01)
    k = 6
    g = ctx.new_table(k).count_bytes(data, FMT_FASTQ33)
    o = oracle.table(k).count_bytes(data, FMT_FASTQ33)
    keys = np.arange(0, 4 ** k, dtype=np.uint64)
    assert np.array_equal(g.get(keys), o.get_many(keys))
    for c in (1, 2, 5, 40):
        solid = g.filter(c)
        assert_same_table(solid, o, min_count=c)
        want = o.get_many(keys)
        want[want < c] = -1
        assert np.array_equal(solid.get(keys), want)
    # add_pairs == addAndBound(key, freq)
    ok, oc = o.dump()
    g2 = ctx.new_table(k)
    g2.add_pairs(ok, oc)
    g2.add_pairs(ok[:100], np.full(100, 32000, dtype=np.int16))
    o2 = oracle.table(k)
    o2.add_pairs(ok, oc)
    o2.add_pairs(ok[:100], np.full(100, 32000, dtype=np.int16))
    assert_same_table(g2, o2)


def test_table_growth(ctx, oracle):
    """Many distinct k-mers force several rehash doublings of the device table."""
    rng = np.random.default_rng(5)
    n = 6_000_000
    keys = np.unique(rng.integers(0, 1 << 62, size=n, dtype=np.uint64))
    counts = rng.integers(1, 9, size=keys.size).astype(np.int16)
    g = ctx.new_table(31)
    for part in np.array_split(np.arange(keys.size), 3):
        g.add_pairs(keys[part], counts[part])
    assert g.size() == keys.size
    gk, gc = g.dump()
    assert np.array_equal(gk, keys) and np.array_equal(gc, counts)
    probe = np.concatenate([keys[:1000], keys[:1000] ^ np.uint64(1)])
    got = g.get(probe)
    assert np.array_equal(got[:1000], counts[:1000])


def test_synthetic_generator_twins_agree(ctx, oracle):
    for fasta in (False, True):
        s = oracle.synth(n_targets=3, fasta=fasta, genome_unit=20_000)
        n = 5000
        cpu = s.generate(100, n)
        d = ctx.device_alloc(cpu.size)
        ctx.synth_generate_device(s.params_bytes(), 100, n, d)
        gpu = np.empty_like(cpu)
        ctx.d2h(gpu, d)
        ctx.device_free(d)
        assert np.array_equal(cpu, gpu)


def test_synthetic_config1_slice(ctx, oracle):
    """A 60k-read slice of BASELINE config 1 (k=31, 150 bp FASTQ with N and '!' qualities): bit-exact table."""
    s = oracle.synth()
    data = s.generate(0, 60_000)
    g = ctx.new_table(31).count_bytes(data, FMT_FASTQ)
    o = oracle.table(31, threads=8).count_bytes(data, FMT_FASTQ)
    assert_same_table(g, o)
    assert g.stats() == o.stats()
    assert_same_table(g.filter(5), o, min_count=5)


# ---- the binned update path (mc_bins.cu): bins, spill list, several batches, sampled growth ----------------------
BIN_TUNINGS = [
    dict(part_bits=0),
    dict(part_bits=5),
    dict(part_bits=11),
    dict(part_bits=4, batch_keys=3000),                        # tiny batch: most keys take the spill list, many flushes
    dict(part_bits=3, segment_keys=2048),                      # one k_bin_count launch per tile
    dict(part_bits=6, batch_keys=50_000, segment_keys=4096),
    dict(extract_parts=4),
    dict(extract_parts=32, part_bits=2),
    # two-level bins + shared-memory aggregation (k_bin_scatter2 + k_bin_agg), forced on for a small input
    dict(agg=1),
    dict(agg=1, part_bits=5, agg_unit=512),
    dict(agg=1, part_bits=3, agg_unit=64, segment_keys=100_000),    # 2^11 sub-bins per bin, several groups
    dict(agg=1, part_bits=0, agg_unit=1 << 22),                     # one unit far larger than the shared table: direct fallback
    dict(agg=1, part_bits=6, batch_keys=50_000, agg_unit=256),      # many flushes into a table that already holds the keys
]


@pytest.mark.parametrize("tuning", BIN_TUNINGS, ids=lambda t: ",".join(f"{k}={v}" for k, v in t.items()))
def test_binned_update_path_is_exact(oracle, tuning):
    from metacherchant_b200 import Context
    s = oracle.synth(genome_unit=5_000)
    data = s.generate(0, 12_000)                               # 1.4 M k-mer instances, k = 31
    o = oracle.table(31, threads=8).count_bytes(data, FMT_FASTQ)
    c = Context(0, chunk_bytes=1 << 20)                        # 4 chunks
    try:
        for name, v in tuning.items():
            c.set_tuning(name, v)
        g = c.new_table(31).count_bytes(data, FMT_FASTQ)
        assert_same_table(g, o)
        assert g.stats() == o.stats()
        # a second input into the same table (a second file of --reads): counts add up
        g.count_bytes(data, FMT_FASTQ)
        o2 = oracle.table(31, threads=8).count_bytes(data, FMT_FASTQ).count_bytes(data, FMT_FASTQ)
        assert_same_table(g, o2)
    finally:
        c.close()


@pytest.mark.parametrize("agg", [0, 1])
@pytest.mark.parametrize("part_bits", [0, 7])
def test_heavy_hitters_overflow_their_bin(oracle, part_bits, agg):
    """One k-mer repeated far beyond a bin's capacity diverts to the spill list; counts stay exact and saturate at 32767."""
    from metacherchant_b200 import Context
    rng = random.Random(3)
    reads = [">h\n" + "ACGTTGCA" * 5000 + "\n"] + [f">r{i}\n" + "".join(rng.choice("ACGT") for _ in range(80)) + "\n" for i in range(2000)]
    data = "".join(reads).encode()
    c = Context(0)
    try:
        c.set_tuning("part_bits", part_bits)
        c.set_tuning("batch_keys", 20_000 if not agg else 200_000)
        c.set_tuning("agg", agg)
        c.set_tuning("agg_unit", 128)
        g = c.new_table(8).count_bytes(data, FMT_FASTA)
        assert_same_table(g, oracle.table(8).count_bytes(data, FMT_FASTA))
    finally:
        c.close()


@pytest.mark.parametrize("agg", [0, 1])
def test_binned_growth_from_an_empty_hint(oracle, agg):
    """No capacity hint and ~3 M distinct k-mers: the first sampled segment sizes the table, nothing is lost."""
    from metacherchant_b200 import Context
    s = oracle.synth(genome_unit=40_000, err_ppm=20_000)
    data = s.generate(0, 50_000)
    c = Context(0)
    try:
        c.set_tuning("agg", agg)
        g = c.new_table(31).count_bytes(data, FMT_FASTQ)
        o = oracle.table(31, threads=8).count_bytes(data, FMT_FASTQ)
        assert_same_table(g, o)
    finally:
        c.close()


def test_gz_files_through_the_file_api(ctx, oracle, tmp_path):
    """mc_count_files on `.fastq.gz` / `.fa.gz` / `.fq.bz2` / `.fasta.bz2` (host-side inflate) == the oracle on the same files."""
    import gzip
    rng = random.Random(12)
    fq, fa = make_fastq(rng, 3000, max_len=120), make_fasta(rng, 1500)
    import bz2
    paths = [str(tmp_path / "r1.fastq.gz"), str(tmp_path / "r2.fa.gz"), str(tmp_path / "r3.fq"), str(tmp_path / "r4.fq.bz2"), str(tmp_path / "r5.fasta.bz2")]
    open(paths[0], "wb").write(gzip.compress(fq[:50_000]) + gzip.compress(fq[50_000:]))
    open(paths[1], "wb").write(gzip.compress(fa))
    open(paths[2], "wb").write(fq)
    open(paths[3], "wb").write(bz2.compress(fq[:70_000]) + bz2.compress(fq[70_000:]))
    open(paths[4], "wb").write(bz2.compress(fa))
    g = ctx.load_reads(paths, 17)
    o = oracle.table(17, threads=4)
    for p in paths:
        o.count_file(p)
    assert_same_table(g, o)
    assert g.stats() == o.stats()
    from metacherchant_b200._ffi import McError
    open(tmp_path / "x.fastq.bz2", "wb").write(b"BZh9" + b"garbage" * 10)
    with pytest.raises(McError):
        ctx.load_reads([str(tmp_path / "x.fastq.bz2")], 17)


@pytest.mark.parametrize("chunk", [4096, 30000])
def test_packed_files_are_streamed_slice_by_slice(oracle, tmp_path, chunk):
    """.gz / .bz2 inputs are inflated slice by slice (two host buffers of chunk_bytes), never whole: with chunks far smaller than the files
    every slice boundary falls inside a record, the quality sniffer's head spans several slices and concatenated members / streams end
    in the middle of a slice.  Tables and reader statistics equal the oracle's on the same files."""
    import bz2
    import gzip
    from metacherchant_b200 import Context
    rng = random.Random(chunk)
    fq = make_fastq(rng, 4000, max_len=150, offset=64)     # Phred+64: the sniffer must see 1000 records (~300 KB) before deciding
    fa = make_fasta(rng, 1500, max_len=300)
    paths = [str(tmp_path / "a.fastq.gz"), str(tmp_path / "b.fq.bz2"), str(tmp_path / "c.fa.gz"), str(tmp_path / "d.fasta.bz2"), str(tmp_path / "e.fq.gz")]
    open(paths[0], "wb").write(gzip.compress(fq[:123_457]) + gzip.compress(fq[123_457:]))
    open(paths[1], "wb").write(bz2.compress(fq[:200_001]) + bz2.compress(fq[200_001:]))
    open(paths[2], "wb").write(gzip.compress(fa))
    open(paths[3], "wb").write(bz2.compress(fa[:10_000]) + bz2.compress(fa[10_000:]))
    open(paths[4], "wb").write(gzip.compress(b""))
    c = Context(0, chunk_bytes=chunk)
    try:
        g = c.load_reads(paths, 21)
        o = oracle.table(21, threads=4)
        for p in paths:
            o.count_file(p)
        assert_same_table(g, o)
        assert g.stats() == o.stats()
    finally:
        c.close()
