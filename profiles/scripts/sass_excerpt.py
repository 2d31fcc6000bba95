#!/usr/bin/env python3
"""SASS evidence of the Blackwell-specific paths: which kernels of libmcgpu.so hold bulk-copy / mbarrier / bulk-prefetch instructions.
usage: python profiles/scripts/sass_excerpt.py > profiles/r2_sass_excerpt.txt   (runs cuobjdump -sass on the in-tree library)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "metacherchant_b200", "csrc", "libmcgpu.so")
PAT = ["UBLKCP", "UBLKPF", "UTMACMDFLUSH", "SYNCS", "DEPBAR", "FENCE.VIEW.ASYNC", "UTMALDG", "UTMASTG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
print("# SASS evidence of the bulk-copy (TMA unit) paths -- cuobjdump -sass metacherchant_b200/csrc/libmcgpu.so (sm_100a), made by profiles/scripts/sass_excerpt.py")
print("# k_bin_scatter<true>, k_bin_scatter2<true> (mc_bins.cu): bulk_store_s2g = cp.async.bulk.global.shared::cta.bulk_group -> UBLKCP.G.S;")
print("#   commit_group -> UTMACMDFLUSH; wait_group.read -> DEPBAR; fence.proxy.async -> FENCE.VIEW.ASYNC; cp.async.bulk.prefetch.L2 -> UBLKPF.L2")
print("# k_extract<*> (mc_count.cu): bulk_load_g2s = cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes -> UBLKCP.S.G;")
print("#   mbarrier.arrive.expect_tx -> SYNCS.ARRIVE.TRANS64; mbarrier.try_wait.parity -> SYNCS.PHASECHK.TRANS64.TRYWAIT")
print("# per kernel: count of each mnemonic, then its first occurrences\n")
fn = None
hits = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and "/*" in line:
        for p in PAT:
            if re.search(r"\b" + re.escape(p), line):
                hits.setdefault(fn, collections.OrderedDict()).setdefault(p, []).append(line.rstrip())
total = collections.Counter()
for fn, d in hits.items():
    print("==", fn)
    for p, lines in d.items():
        total[p] += len(lines)
        print(f"  {p}: {len(lines)} instructions")
        for l in lines[:3]:
            print("    ", l.strip()[:120])
print("\n# totals over the library:", dict(total), f"in {len(hits)} kernels", file=sys.stdout)
