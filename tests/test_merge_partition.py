"""The lockstep walk of k_merge_pairs (metacherchant_b200/csrc/mc_xchg.cu) maps virtual position v of the longest stream to the
positions [f(v), f(v + 1)) of source s with f(v) = (v * ratio_s) >> 32, ratio_s = floor(n_s * 2^32 / n_max), and f(n_max) DEFINED as
n_s.  The kernel is exact only if these ranges partition [0, n_s) -- every received pair is added once -- and if the products stay
inside 64 bits.  Checked here on the CPU with the kernel's integer arithmetic (no device)."""
import random

import pytest

M64 = (1 << 64) - 1


def ranges(n_s, n_max):
    ratio = (n_s << 32) // n_max
    assert ratio <= 1 << 32
    for v in range(n_max):
        lo = ((v * ratio) & M64) >> 32
        hi = n_s if v + 1 == n_max else (((v + 1) * ratio) & M64) >> 32
        assert v * ratio <= M64 and (v + 1) * ratio <= M64          # no wrap-around in the kernel's unsigned 64-bit product
        yield lo, hi


@pytest.mark.parametrize("seed", range(6))
def test_ranges_partition_every_source_exactly(seed):
    rng = random.Random(seed)
    for _ in range(40):
        n_max = rng.choice([1, 2, 3, 7, 64, 1000, 4097, rng.randint(1, 20000)])
        n_s = rng.choice([0, 1, n_max, max(0, n_max - 1), rng.randint(0, n_max)])
        if n_s == 0:
            continue   # (the kernel skips empty sources)
        nxt, multi = 0, 0
        for lo, hi in ranges(n_s, n_max):
            assert lo == nxt and hi >= lo
            multi += hi - lo > 1
            nxt = hi
        assert nxt == n_s
        assert multi <= 1                                            # only the last position can hold more than one pair


def test_products_fit_64_bits_at_the_largest_supported_stream():
    """n_max < 2^32 is what the host checks (MC_ERR_UNSUPPORTED otherwise): the largest products are still below 2^64."""
    n_max = (1 << 32) - 1
    for n_s in (1, n_max // 3, n_max - 1, n_max):
        ratio = (n_s << 32) // n_max
        assert ratio <= 1 << 32 and n_max * ratio <= M64
        # the last positions: exact partition at the end of the stream
        lo = ((n_max - 1) * ratio) >> 32
        assert lo <= n_s and n_s - lo <= 2


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_lockstep_walk_in_batches_of_four_visits_every_pair_once(world):
    """The kernel's loop structure restated: virtual positions x sources taken four at a time, first pair loaded up front, the
    (rare) further pairs of a position in a tail loop.  Every (source, position) pair must be visited exactly once for any mix of
    stream lengths, empty sources included."""
    rng = random.Random(world)
    for _ in range(30):
        n = [rng.choice([0, 1, rng.randint(0, 300), rng.randint(200, 300)]) for _ in range(world)]
        n_max = max(n)
        if n_max == 0:
            continue
        ratio = [(x << 32) // n_max for x in n]
        seen = [[0] * x for x in n]
        for v in range(n_max):
            for s0 in range(0, world, 4):
                lo, hi = [0] * 4, [0] * 4
                first = [None] * 4
                for j in range(4):
                    s = s0 + j
                    if s < world and n[s]:
                        lo[j] = (v * ratio[s]) >> 32
                        hi[j] = n[s] if v + 1 == n_max else ((v + 1) * ratio[s]) >> 32
                        if hi[j] > lo[j]:
                            first[j] = (s, lo[j])
                for j in range(4):
                    if first[j] is not None:
                        seen[first[j][0]][first[j][1]] += 1
                    for i in range(lo[j] + 1, hi[j]):
                        seen[s0 + j][i] += 1
        assert all(c == 1 for row in seen for c in row)
