"""ORACLE (second opinion) — TEST INFRASTRUCTURE ONLY.

A deliberately naive, independent pure-Python transcription of stages A-C of the reference
(read parsing -> k-mer counting -> BFS -> environment map), written from the Java sources
without looking at oracle/mc_oracle.cpp, used only by tests/ to cross-check the C++ oracle on
small inputs (stages A-C are parity-unpinned by any reference fixture).

Citations: src/... = /root/reference/src/..., ITMO!/... = lib/itmo-assembler-src.jar!/ru/ifmo/genetics/...
"""
from __future__ import annotations

NUC = "AGCT"  # ITMO!/dna/DnaTools.java:31
CODE = {"A": 0, "a": 0, "G": 1, "g": 1, "C": 2, "c": 2, "T": 3, "t": 3}  # DnaTools.java:46-64


def java_lines(data: bytes):
    """BufferedReader.readLine(): terminators \\n, \\r, \\r\\n; no phantom line after a final terminator."""
    text = data.decode("latin-1")
    out, cur, i = [], [], 0
    while i < len(text):
        c = text[i]
        if c == "\n" or c == "\r":
            out.append("".join(cur))
            cur = []
            if c == "\r" and i + 1 < len(text) and text[i + 1] == "\n":
                i += 1
        else:
            cur.append(c)
        i += 1
    if cur:
        out.append("".join(cur))
    return out


class QualityError(Exception):
    pass


def fastq_records(data: bytes, offset: int):
    """ITMO!/io/readers/FastqReader.java:53-110 -> list of byte-lists ((phred<<2)|nuc)&0xFF."""
    lines = java_lines(data)
    pos = 0

    def next_data_line():
        nonlocal pos
        while pos < len(lines) and lines[pos] == "":
            pos += 1
        if pos >= len(lines):
            return None
        s = lines[pos]
        pos += 1
        if not (s.startswith("@") or s.startswith("+")):
            raise ValueError("Unknown structure of fastq file!")
        if pos >= len(lines):
            raise ValueError("Unexpected end of file")
        d = lines[pos]
        pos += 1
        return d

    while True:
        seq = next_data_line()
        if seq is None:
            return
        qual = next_data_line()
        if qual is None:
            raise ValueError("Unexpected end of file")
        if len(seq) != len(qual):
            raise ValueError("Bad DnaQ record")
        rec = []
        for s, q in zip(seq, qual):
            if s in "Nn.":
                rec.append(0)
            else:
                if s not in CODE:
                    raise ValueError("Incorrect nucleotide char")
                c = ord(q)
                if c < offset or c > 126:  # Illumina.java:7-12 / Sanger.java:7-12
                    raise QualityError(q)
                rec.append(((c - offset) << 2 | CODE[s]) & 0xFF)
        yield rec


def sniff_offset(data: bytes) -> int:
    """ITMO!/io/ReadersUtils.java:63-77."""
    try:
        for i, _ in enumerate(fastq_records(data, 64)):
            if i + 1 >= 1000:
                break
        return 64
    except QualityError:
        return 33


def fastq_runs(data: bytes, offset: int | None = None):
    """FastaReaderFromXQSourceTrunc.java:61-94 + DnaQ.java:147-150,176-186: cut at stored phred == 0."""
    if offset is None:
        offset = sniff_offset(data)
    runs = []
    for rec in fastq_records(data, offset):
        cur = []
        pieces = []
        i = 0
        n = len(rec)
        start = 0
        while True:
            t = start
            while t < n and ((rec[t] >> 2) & 63) >= 1:
                t += 1
            pieces.append([b & 3 for b in rec[start:t]])
            if n > t + 1:
                start = t + 1
            else:
                break
        runs.extend(pieces)
    return runs


def fasta_runs(data: bytes):
    """ITMO!/io/readers/FastaReader.java:54-108."""
    runs = []
    sb = []
    def flush():
        s = "".join(sb)
        sb.clear()
        if s == "":
            return
        if "N" in s or "n" in s:
            return
        for ch in s:
            if ch not in CODE:
                raise ValueError("Incorrect nucleotide char")
        runs.append([CODE[ch] for ch in s])
    for line in java_lines(data):
        if line.startswith(">") or line.startswith(";"):
            if sb and "".join(sb) != "":
                flush()
        else:
            sb.append(line)
    flush()
    return runs


def revcomp_code(kmer: int, k: int) -> int:
    r = 0
    for _ in range(k):
        r = (r << 2) | (3 - (kmer & 3))
        kmer >>= 2
    return r


def count_runs(runs, k: int):
    """ShortKmer.kmersOf + BigLong2ShortHashMap.addAndBound(.., 1): dict, saturating at 32767."""
    table = {}
    for run in runs:
        for i in range(len(run) - k + 1):
            fw = 0
            for b in run[i:i + k]:
                fw = (fw << 2) | b
            key = min(fw, revcomp_code(fw, k))
            table[key] = min(32767, table.get(key, 0) + 1)
    return table


def encode(s: str) -> int:
    r = 0
    for ch in s:
        r = (r << 2) | CODE[ch]
    return r


def rc_str(s: str) -> str:
    return "".join(NUC[3 - CODE[ch]] for ch in reversed(s))


def normalize(s: str) -> str:  # src/utils/StringUtils.java:34-41
    r = rc_str(s)
    return s if s < r else r


def java_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def java_order(keys, peak=None):
    """Iteration order of a java.util.HashMap<String,_> filled in `keys` order (no treeified bins).
    peak = the largest size the map ever had (removals never shrink the table)."""
    cap = 16
    while max(len(keys), peak or 0) > 0.75 * cap:
        cap *= 2
    def bucket(s):
        h = java_hash(s)
        return (h ^ (h >> 16)) & (cap - 1)
    return sorted(keys, key=bucket)  # python's sort is stable


def neighbors_by_dir(kmer: str, direction: int):
    """getNeighborsByDir (OneSequenceCalculator.java:222-239)."""
    left = [n + kmer[:-1] for n in NUC]
    right = [kmer[1:] + n for n in NUC]
    if direction == -1:
        return left
    if direction == 1:
        return right
    return [x for pair in zip(left, right) for x in pair]


def bfs(table: dict, seq, k: int, min_cov: int, max_kmers, max_radius, direction: int, trim: bool = False, info: dict | None = None):
    """src/algo/OneSequenceCalculator.java:154-220 with TerminationMode.allowsAddition, literal.
    seq: one target, or (--merge) the list of seed sequences; trim: runTrimPaths (:241-262).
    info (optional) receives the peak size of distanceToKmer."""
    def occ(kmer):
        c = encode(kmer)
        return table.get(min(c, revcomp_code(c, k)), -1)
    queue, dist, last = [], {}, set()
    for one in ([seq] if isinstance(seq, str) else seq):
        for i in range(len(one) - k + 1):
            kmer = one[i:i + k]
            if occ(kmer) >= min_cov:
                queue.append(kmer)
                dist[kmer] = 0
    if not queue:
        return None
    head = 0
    while head < len(queue):
        kmer = queue[head]
        head += 1
        d = dist[kmer]
        nbs = neighbors_by_dir(kmer, direction)
        for nb in nbs:
            if occ(nb) >= min_cov:
                ok = nb not in dist
                if ok and max_kmers is not None and len(dist) >= max_kmers:
                    ok = False
                if ok and max_radius is not None and d + 1 > max_radius:
                    ok = False
                if ok:
                    queue.append(nb)
                    dist[nb] = d + 1
                else:
                    last.add(kmer)
    if info is not None:
        info["peak"] = len(dist)
    if trim:
        q2, seen = list(last), set(last)
        head = 0
        while head < len(q2):
            kmer = q2[head]
            head += 1
            for nb in neighbors_by_dir(kmer, -direction):
                if nb in dist and nb not in seen:
                    q2.append(nb)
                    seen.add(nb)
        dist = {kk: v for kk, v in dist.items() if kk in seen}
    return dist  # python dicts keep insertion (= discovery) order


def environment(table: dict, seq, k: int, min_cov: int, max_kmers, max_radius, both_dirs: bool, trim: bool = False):
    """buildEnvironment + addToSubgraph (OneSequenceCalculator.java:137-148,217-219) -> [(kmer, count)] in
    the java iteration order of `subgraph`."""
    sub = {}
    def occ(kmer):
        c = encode(kmer)
        return table.get(min(c, revcomp_code(c, k)), -1)
    for direction in ([0] if both_dirs else [-1, 1]):
        info = {}
        dist = bfs(table, seq, k, min_cov, max_kmers, max_radius, direction, trim, info)
        if dist is None:
            return None
        for kmer in java_order(list(dist.keys()), info["peak"]):
            sub[normalize(kmer)] = occ(kmer)
    return [(kmer, sub[kmer]) for kmer in java_order(list(sub.keys()))]
