"""CPU: the exactness argument behind the planned single-level partition ("Planned next"): if reads are cut
into super-k-mers at changes of the CANONICAL minimizer and every super-k-mer goes to the bin of its minimizer, then
  (1) unpacking the super-k-mers gives back exactly the k-mer instances of the reads (same multiset of canonical keys), and
  (2) every canonical k-mer lands in ONE bin whatever strand or read it was seen in,
so per-bin counting followed by table updates gives the same table as counting all instances directly.
Plain Python on small random inputs; the 2-bit code and canonical form are the reference's (oracle/pyref.py)."""
import collections
import random

import pytest

from oracle import pyref


def canon(code, k):
    return min(code, pyref.revcomp_code(code, k))


def mix(x):
    x ^= x >> 31
    x = (x * 0x1E3779B97F4A7C15) & ((1 << 62) - 1)
    x ^= x >> 29
    x = (x * 0x3F58476D1CE4E5B9) & ((1 << 62) - 1)
    return x ^ (x >> 32)


def decode(code, k):
    return "".join(pyref.NUC[(code >> (2 * (k - 1 - i))) & 3] for i in range(k))


def kmers_of(read, k):
    return [pyref.encode(read[i:i + k]) for i in range(len(read) - k + 1)]


def minimizer(kmer_str, k, m):
    """Smallest hashed canonical m-mer of the k-mer (ties: the hash is injective on canonical m-mers, so the value is unique)."""
    return min(mix(canon(pyref.encode(kmer_str[i:i + m]), m)) for i in range(k - m + 1))


def super_kmers(read, k, m):
    """[(minimizer, substring)]: maximal runs of consecutive k-mers with the same minimizer; substring covers the run."""
    out, start, cur = [], 0, None
    n = len(read) - k + 1
    for i in range(n):
        mz = minimizer(read[i:i + k], k, m)
        if cur is None:
            cur = mz
        elif mz != cur:
            out.append((cur, read[start:i + k - 1]))
            start, cur = i, mz
    if n > 0:
        out.append((cur, read[start:]))
    return out


@pytest.mark.parametrize("k,m", [(31, 13), (31, 11), (23, 9), (15, 7), (9, 5)])
def test_super_kmers_partition_the_kmer_instances_and_bins_are_strand_symmetric(k, m):
    rng = random.Random(k * 100 + m)
    genome = "".join(rng.choice("ACGT") for _ in range(600))
    genome += genome[100:180] + "A" * 40 + genome[300:340]            # a repeat and a homopolymer
    reads = []
    for _ in range(120):
        s = rng.randrange(0, len(genome) - 100)
        r = genome[s:s + rng.randrange(k, 100)]
        if rng.random() < 0.5:
            r = pyref.rc_str(r)
        if rng.random() < 0.3:
            p = rng.randrange(len(r))
            r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        reads.append(r)
    direct = collections.Counter(canon(c, k) for r in reads for c in kmers_of(r, k))
    n_bins = 64
    per_bin = collections.defaultdict(collections.Counter)
    n_super = 0
    for r in reads:
        pieces = super_kmers(r, k, m)
        n_super += len(pieces)
        assert sum(len(p) - k + 1 for _, p in pieces) == len(r) - k + 1          # (1) every k-mer instance exactly once
        for mz, piece in pieces:
            b = mix(mz ^ 0x2545F4914F6CDD1D) % n_bins
            for c in kmers_of(piece, k):
                assert minimizer(decode(c, k), k, m) == mz                             # all k-mers of a piece share its minimizer
                per_bin[b][canon(c, k)] += 1
    merged = collections.Counter()
    owner = {}
    for b, cnt in per_bin.items():
        for key, v in cnt.items():
            assert owner.setdefault(key, b) == b                                  # (2) one bin per canonical k-mer
            merged[key] += v
    assert merged == direct
    assert n_super < sum(len(r) - k + 1 for r in reads) / 2                       # and the stream really shrinks


def test_minimizer_is_the_same_on_both_strands():
    rng = random.Random(4)
    for _ in range(200):
        s = "".join(rng.choice("ACGT") for _ in range(31))
        assert minimizer(s, 31, 13) == minimizer(pyref.rc_str(s), 31, 13)
