#!/usr/bin/env python3
"""A small stand-alone BFS workload for ncu (k_bfs is latency bound per step: the table size does not matter):
300 k reads over a 1.1 Mbp community, one target, --maxkmers 100000, both passes."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ffi
from metacherchant_b200 import Context
o = oracle_ffi.load()
s = o.synth(genome_unit=20_000, n_targets=2)
data = s.generate(0, 300_000)
ctx = Context(0)
t = ctx.new_table(31).count_bytes(data, oracle_ffi.FMT_FASTQ)
solid = t.filter(3)
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "g.fasta"); open(fa, "w").write(s.targets_fasta())
    targets = ctx.load_targets(fa)
    found = ctx.find_environments(solid, [x[1] for x in targets], [x[0] for x in targets], min_cov=3, max_kmers=100000, both_dirs=False, out_dir=os.path.join(d, "out"))
    print("found", found, sum(1 for _ in open(os.path.join(d, "out", "gene0", "env.txt"))))
ctx.close()
