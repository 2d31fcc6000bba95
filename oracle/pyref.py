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
    """Iteration order of a java.util.HashMap<String,_> filled in `keys` order.
    peak = the largest size the map ever had (removals never shrink the table).
    Bins that never held 9 entries are plain lists: bucket index at the final capacity, insertion order inside a bucket (resize
    splits preserve the relative order).  When some bin did reach 9 entries (TREEIFY_THRESHOLD) the history matters and the map
    is replayed put by put (JavaHashMapModel below)."""
    if java_order_needs_replay(keys):
        if peak is not None and peak > len(keys):
            raise NotImplementedError("treeified bins followed by removals are not modelled")
        m = JavaHashMapModel()
        for k in keys:
            m.put(k)
        return m.keys()
    cap = 16
    while max(len(keys), peak or 0) > 0.75 * cap:
        cap *= 2
    def bucket(s):
        h = java_hash(s)
        return (h ^ (h >> 16)) & (cap - 1)
    return sorted(keys, key=bucket)  # python's sort is stable


def java_order_needs_replay(keys) -> bool:
    """True when, filling a default HashMap in `keys` order, some bin reaches 9 entries while the table has the capacity it has at
    that moment (java.util.HashMap.putVal: binCount >= TREEIFY_THRESHOLD - 1 after appending)."""
    hs = [(lambda h: h ^ (h >> 16))(java_hash(k)) for k in keys]
    cap, lo = 16, 0
    while lo < len(hs):
        hi = min(len(hs), int(0.75 * cap) + 1)   # the put that makes size > threshold still lands in this table, then it doubles
        cnt = {}
        for i in range(hi):
            b = hs[i] & (cap - 1)
            c = cnt.get(b, 0) + 1
            cnt[b] = c
            if i >= lo and c >= 9:
                return True
        lo = hi
        cap *= 2
    return False


class JavaHashMapModel:
    """java.util.HashMap<String,_> (JDK 8+), insertions only, keys distinct: table of bins, plain bins as linked lists, bins of 9
    nodes turned into red-black trees (treeifyBin / TreeNode.treeify / putTreeVal / split / untreeify / moveRootToFront), iteration
    along the `next` links.  Written from the JDK's algorithm; nodes are dicts of indices."""
    TREEIFY, UNTREEIFY, MIN_TREEIFY_CAP = 8, 6, 64

    def __init__(self):
        self.tab = None
        self.thr = 0
        self.size = 0
        self.key, self.h = [], []
        self.nxt, self.prv, self.left, self.right, self.par, self.red, self.tree = [], [], [], [], [], [], []

    # ---- helpers -------------------------------------------------------------------------
    @staticmethod
    def _sgn(x):   # int32 view of the spread hash (TreeNode ordering compares ints)
        return x - (1 << 32) if x & 0x80000000 else x

    def _dir(self, a, b):   # ordering of node a against node b: by hash, then String.compareTo
        ha, hb = self._sgn(self.h[a]), self._sgn(self.h[b])
        if hb > ha:
            return -1
        if hb < ha:
            return 1
        return -1 if self.key[a] < self.key[b] else 1

    def _new(self, key, h):
        self.key.append(key); self.h.append(h)
        for arr in (self.nxt, self.prv, self.left, self.right, self.par):
            arr.append(-1)
        self.red.append(False); self.tree.append(False)
        return len(self.key) - 1

    # ---- red-black plumbing (TreeNode.rotateLeft/rotateRight/balanceInsertion) ------------
    def _rot_left(self, root, p):
        r = self.right[p] if p != -1 else -1
        if p != -1 and r != -1:
            rl = self.left[r]
            self.right[p] = rl
            if rl != -1:
                self.par[rl] = p
            pp = self.par[p]
            self.par[r] = pp
            if pp == -1:
                root = r
                self.red[r] = False
            elif self.left[pp] == p:
                self.left[pp] = r
            else:
                self.right[pp] = r
            self.left[r] = p
            self.par[p] = r
        return root

    def _rot_right(self, root, p):
        l = self.left[p] if p != -1 else -1
        if p != -1 and l != -1:
            lr = self.right[l]
            self.left[p] = lr
            if lr != -1:
                self.par[lr] = p
            pp = self.par[p]
            self.par[l] = pp
            if pp == -1:
                root = l
                self.red[l] = False
            elif self.right[pp] == p:
                self.right[pp] = l
            else:
                self.left[pp] = l
            self.right[l] = p
            self.par[p] = l
        return root

    def _balance_insertion(self, root, x):
        self.red[x] = True
        while True:
            xp = self.par[x]
            if xp == -1:
                self.red[x] = False
                return x
            xpp = self.par[xp]
            if not self.red[xp] or xpp == -1:
                return root
            xppl = self.left[xpp]
            if xp == xppl:
                xppr = self.right[xpp]
                if xppr != -1 and self.red[xppr]:
                    self.red[xppr] = False; self.red[xp] = False; self.red[xpp] = True
                    x = xpp
                else:
                    if x == self.right[xp]:
                        x = xp
                        root = self._rot_left(root, x)
                        xp = self.par[x]
                        xpp = -1 if xp == -1 else self.par[xp]
                    if xp != -1:
                        self.red[xp] = False
                        if xpp != -1:
                            self.red[xpp] = True
                            root = self._rot_right(root, xpp)
            else:
                if xppl != -1 and self.red[xppl]:
                    self.red[xppl] = False; self.red[xp] = False; self.red[xpp] = True
                    x = xpp
                else:
                    if x == self.left[xp]:
                        x = xp
                        root = self._rot_right(root, x)
                        xp = self.par[x]
                        xpp = -1 if xp == -1 else self.par[xp]
                    if xp != -1:
                        self.red[xp] = False
                        if xpp != -1:
                            self.red[xpp] = True
                            root = self._rot_left(root, xpp)

    def _move_root_to_front(self, tab, root):
        if root == -1 or tab is None:
            return
        idx = self.h[root] & (len(tab) - 1)
        first = tab[idx]
        if root != first:
            tab[idx] = root
            rp, rn = self.prv[root], self.nxt[root]
            if rn != -1:
                self.prv[rn] = rp
            if rp != -1:
                self.nxt[rp] = rn
            if first != -1:
                self.prv[first] = root
            self.nxt[root] = first
            self.prv[root] = -1

    def _treeify(self, tab, head):
        root = -1
        x = head
        while x != -1:
            nx = self.nxt[x]
            self.left[x] = self.right[x] = -1
            if root == -1:
                self.par[x] = -1
                self.red[x] = False
                root = x
            else:
                p = root
                while True:
                    d = self._dir(x, p)
                    xp = p
                    p = self.left[p] if d <= 0 else self.right[p]
                    if p == -1:
                        self.par[x] = xp
                        if d <= 0:
                            self.left[xp] = x
                        else:
                            self.right[xp] = x
                        root = self._balance_insertion(root, x)
                        break
            x = nx
        self._move_root_to_front(tab, root)

    def _treeify_bin(self, h):
        if self.tab is None or len(self.tab) < self.MIN_TREEIFY_CAP:
            self._resize()
            return
        idx = h & (len(self.tab) - 1)
        e = self.tab[idx]
        if e == -1:
            return
        prev = -1
        x = e
        while x != -1:        # replacementTreeNode for every node, same order, prev links set
            self.tree[x] = True
            self.prv[x] = prev
            prev = x
            x = self.nxt[x]
        self._treeify(self.tab, e)

    def _put_tree_val(self, first, n):
        root = first
        while self.par[root] != -1:
            root = self.par[root]
        p = root
        while True:
            d = self._dir(n, p)
            xp = p
            p = self.left[p] if d <= 0 else self.right[p]
            if p == -1:
                xpn = self.nxt[xp]
                self.tree[n] = True
                self.nxt[n] = xpn
                if d <= 0:
                    self.left[xp] = n
                else:
                    self.right[xp] = n
                self.nxt[xp] = n
                self.par[n] = xp
                self.prv[n] = xp
                if xpn != -1:
                    self.prv[xpn] = n
                self._move_root_to_front(self.tab, self._balance_insertion(root, n))
                return

    def _resize(self):
        old = self.tab
        if old is None:
            self.tab = [-1] * 16
            self.thr = 12
            return
        oc = len(old)
        nt = [-1] * (oc * 2)
        self.thr *= 2
        self.tab = nt
        for j in range(oc):
            e = old[j]
            if e == -1:
                continue
            if self.nxt[e] == -1:
                nt[self.h[e] & (2 * oc - 1)] = e
                if self.tree[e]:     # a one-node tree bin goes through split like any other
                    self._split(nt, e, j, oc)
                continue
            if self.tree[e]:
                self._split(nt, e, j, oc)
                continue
            lo_h = lo_t = hi_h = hi_t = -1
            while e != -1:
                nx = self.nxt[e]
                if self.h[e] & oc == 0:
                    if lo_t == -1:
                        lo_h = e
                    else:
                        self.nxt[lo_t] = e
                    lo_t = e
                else:
                    if hi_t == -1:
                        hi_h = e
                    else:
                        self.nxt[hi_t] = e
                    hi_t = e
                e = nx
            if lo_t != -1:
                self.nxt[lo_t] = -1
                nt[j] = lo_h
            if hi_t != -1:
                self.nxt[hi_t] = -1
                nt[j + oc] = hi_h

    def _split(self, tab, b, index, bit):
        lo_h = lo_t = hi_h = hi_t = -1
        lc = hc = 0
        e = b
        while e != -1:
            nx = self.nxt[e]
            self.nxt[e] = -1
            if self.h[e] & bit == 0:
                self.prv[e] = lo_t
                if lo_t == -1:
                    lo_h = e
                else:
                    self.nxt[lo_t] = e
                lo_t = e
                lc += 1
            else:
                self.prv[e] = hi_t
                if hi_t == -1:
                    hi_h = e
                else:
                    self.nxt[hi_t] = e
                hi_t = e
                hc += 1
            e = nx
        tab[index] = -1
        tab[index + bit] = -1
        for head, cnt, other, at in ((lo_h, lc, hi_h, index), (hi_h, hc, lo_h, index + bit)):
            if head == -1:
                continue
            if cnt <= self.UNTREEIFY:
                x = head
                while x != -1:      # untreeify: plain nodes in the same order
                    self.tree[x] = False
                    self.left[x] = self.right[x] = self.par[x] = -1
                    x = self.nxt[x]
                tab[at] = head
            else:
                tab[at] = head
                if other != -1:     # (else the tree is intact: nothing moved out of it)
                    self._treeify(tab, head)

    # ---- public --------------------------------------------------------------------------
    def put(self, key):
        """A NEW key (the callers de-duplicate)."""
        hh = java_hash(key)
        h = hh ^ (hh >> 16)
        if self.tab is None:
            self._resize()
        n = self._new(key, h)
        i = h & (len(self.tab) - 1)
        p = self.tab[i]
        if p == -1:
            self.tab[i] = n
        elif self.tree[p]:
            self._put_tree_val(p, n)
        else:
            cnt = 0
            while self.nxt[p] != -1:
                p = self.nxt[p]
                cnt += 1
            self.nxt[p] = n
            if cnt >= self.TREEIFY - 1:
                self._treeify_bin(h)
        self.size += 1
        if self.size > self.thr:
            self._resize()

    def keys(self):
        out = []
        for e in self.tab or []:
            while e != -1:
                out.append(self.key[e])
                e = self.nxt[e]
        return out


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
