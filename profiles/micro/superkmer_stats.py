#!/usr/bin/env python3
"""Design evidence for the next structural step (profiles/r1d_NOTES.md, item 2): how much smaller is the stream that has to be
scattered if reads are cut into super-k-mers (maximal runs of consecutive k-mers sharing their canonical minimizer) instead of
one 8-byte hashed key per k-mer, and how evenly do minimizer bins fill?  CPU only (numpy + the oracle's workload generator).

    python profiles/micro/superkmer_stats.py [reads]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ffi  # noqa: E402  (workload generator only)

K, L = 31, 150
M62 = np.uint64((1 << 62) - 1)


def mix(x):
    x = x ^ (x >> np.uint64(31))
    x = (x * np.uint64(0x1E3779B97F4A7C15)) & M62
    x = x ^ (x >> np.uint64(29))
    x = (x * np.uint64(0x3F58476D1CE4E5B9)) & M62
    return x ^ (x >> np.uint64(32))


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
    o = oracle_ffi.load()
    s = o.synth(genome_unit=1_000_000, n_targets=20)           # the cfg2 community (55 Mbp)
    raw = s.generate(0, n_reads).reshape(n_reads, s.record_bytes)
    # record = "@r<idx padded>\n" + 150 bases + "\n+\n" + 150 quals + "\n": find the sequence column by the first newline
    nl = int(np.argmax(raw[0] == 10))
    seq = raw[:, nl + 1: nl + 1 + L]
    code = np.full(256, 0, dtype=np.uint64)
    for ch, c in zip(b"AGCT", range(4)):
        code[ch] = c
    b = code[seq]                                               # N counted as A here: statistics only
    print(f"{n_reads} reads of {L} bp, k = {K}: {L - K + 1} k-mers per read, 8 B per hashed key = {8 * (L - K + 1)} B per read")
    for m in (9, 11, 13, 15):
        w = K - m + 1                                           # m-mers per k-mer
        nm = L - m + 1
        fw = np.zeros((n_reads, nm), dtype=np.uint64)
        rc = np.zeros((n_reads, nm), dtype=np.uint64)
        for i in range(m):
            fw = (fw << np.uint64(2)) | b[:, i: i + nm]
            rc = rc | ((np.uint64(3) - b[:, i: i + nm]) << np.uint64(2 * i))
        canon = np.minimum(fw, rc)
        h = mix(canon)                                          # random order of the m-mers (no poly-A bias)
        win = np.lib.stride_tricks.sliding_window_view(h, w, axis=1)     # (reads, k-mers, w)
        mn = win.min(axis=2)                                    # minimizer hash of every k-mer
        change = np.ones(mn.shape, dtype=bool)
        change[:, 1:] = mn[:, 1:] != mn[:, :-1]
        n_super = int(change.sum())
        n_kmers = mn.size
        per = n_kmers / n_super
        # encoded super-k-mer: 2 bits per base of (k + n - 1) bases + 1 length byte, rounded up to whole bytes
        lens = np.diff(np.append(np.flatnonzero(change.ravel()), n_kmers))
        # (runs do not cross reads: every read starts with change = True)
        bytes_super = int(np.ceil((K + lens - 1) / 4).sum() + lens.size)
        bins = 1 << 14
        # the minimum of w hashes is biased towards small values: re-mix it before taking the bin bits
        fill = np.bincount((mix(mn[change] ^ np.uint64(0x2545F4914F6CDD1D)) >> np.uint64(62 - 14)).astype(np.int64), weights=lens.astype(np.float64),
                           minlength=bins)
        print(f"m = {m:2d}: {per:5.2f} k-mers per super-k-mer, {bytes_super / n_reads:6.1f} B per read instead of {8 * (L - K + 1)} "
              f"({8 * n_kmers / bytes_super:4.1f} x smaller); 2^14 minimizer bins hold k-mers: mean {fill.mean():.0f}, "
              f"max {fill.max():.0f} ({fill.max() / fill.mean():.2f} x mean), empty {int((fill == 0).sum())}")


if __name__ == "__main__":
    main()
