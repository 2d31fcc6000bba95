"""Shared input builders for the CPU and GPU tests."""
import random

from oracle import pyref


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def make_fastq(rng, n_reads, max_len=60, min_len=0, p_n=0.03, p_bang=0.03, crlf=False, blanks=False, lower=False, offset=33, final_newline=None):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i in range(n_reads):
        L = rng.randint(min_len, max_len)
        seq = "".join(("N" if rng.random() < p_n else rng.choice("ACGTacgt" if lower else "ACGT")) for _ in range(L))
        if rng.random() < 0.1 and L:
            p = rng.randrange(L)
            seq = seq[:p] + rng.choice(".n") + seq[p + 1:]
        zero = chr(offset)
        qual = "".join((zero if rng.random() < p_bang else chr(rng.randint(offset + 1, min(126, offset + 41)))) for _ in range(L))
        if blanks and rng.random() < 0.2:
            out.append("")
        out.append(f"@r{i} some comment")
        out.append(seq)
        if blanks and rng.random() < 0.2:
            out.append("")
        out.append("+" + (f"r{i}" if rng.random() < 0.3 else ""))
        out.append(qual)
    if final_newline is None:
        final_newline = rng.random() < 0.8
    text = nl.join(out) + (nl if final_newline else "")
    return text.encode()


def make_fasta(rng, n_records, max_len=70, multiline=True, p_n=0.08):
    lines = []
    for i in range(n_records):
        if rng.random() < 0.9:
            lines.append((">" if rng.random() < 0.8 else ";") + f"r{i}")
        if rng.random() < 0.05:
            lines.append(">double header")
        for _ in range(rng.randint(0, 3) if multiline else 1):
            L = rng.randint(0, max_len)
            s = rand_seq(rng, L, "ACGTacgt")
            if rng.random() < p_n and L:
                p = rng.randrange(L)
                s = s[:p] + rng.choice("Nn") + s[p + 1:]
            lines.append(s)
            if multiline and rng.random() < 0.1:
                lines.append("")
    return ("\n".join(lines) + "\n").encode()


def small_graph(rng, k, genome_len=400, n_reads=300, read_len=40):
    """A small genome with a repeat (so the de Bruijn graph branches) and noisy FASTA reads over it."""
    genome = rand_seq(rng, genome_len)
    genome = genome[:150] + genome[20:60] + genome[150:]
    reads = []
    for i in range(n_reads):
        s = rng.randrange(0, len(genome) - read_len)
        r = genome[s:s + read_len]
        if rng.random() < 0.5:
            r = pyref.rc_str(r)
        if rng.random() < 0.2:
            p = rng.randrange(read_len)
            r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        reads.append(f">r{i}\n{r}\n")
    return genome, "".join(reads).encode()
