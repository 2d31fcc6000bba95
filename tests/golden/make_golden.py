#!/usr/bin/env python3
"""Regenerates tests/golden/hic_example1 from the reference's committed example outputs.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
The fixtures are reference DATA (example outputs of the Hi-C pipeline, produced by an older
revision of the reference with --merge true), not reference source:
  graph.txt.xz  <- Hi-C_pipline/example_work_dir/output/1/merged/graph.txt (93 572 env k-mers, java HashMap order)
  graph.gfa, nodes.tsv, edges.tsv, seqs.fasta.head <- same directory
  seq.fasta     <- Hi-C_pipline/example/seq.fasta (61 bp target, no header line)
"""
import lzma, os, shutil, sys

REF = "/root/reference/Hi-C_pipline"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hic_example1")

def main():
    src = f"{REF}/example_work_dir/output/1/merged"
    os.makedirs(OUT, exist_ok=True)
    with open(f"{src}/graph.txt", "rb") as f, lzma.open(f"{OUT}/graph.txt.xz", "wb", preset=9) as g:
        g.write(f.read())
    shutil.copy(f"{src}/graph.gfa", f"{OUT}/graph.gfa")
    shutil.copy(f"{src}/tsvs/nodes.tsv", f"{OUT}/nodes.tsv")
    shutil.copy(f"{src}/tsvs/edges.tsv", f"{OUT}/edges.tsv")
    shutil.copy(f"{REF}/example/seq.fasta", f"{OUT}/seq.fasta")
    for p in (f"{OUT}/graph.gfa", f"{OUT}/nodes.tsv", f"{OUT}/edges.tsv", f"{OUT}/seq.fasta"):
        os.chmod(p, 0o644)
    print("wrote", sorted(os.listdir(OUT)))

if __name__ == "__main__":
    sys.exit(main())
