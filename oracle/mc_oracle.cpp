// ORACLE — TEST INFRASTRUCTURE ONLY (see mc_oracle.hpp for the rules and the parity-pinning statement).
// Literal CPU restatement of the reference algorithm; every function cites the
// reference lines it follows.  Not used, linked or executed by the product path.
#include "mc_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <zlib.h>
#include <dlfcn.h>

namespace {

thread_local std::string g_err;

struct OracleError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw OracleError(buf);
}

// =====================================================================================
// DNA arithmetic
// =====================================================================================

// ITMO!/dna/DnaTools.java:31  NUCLEOTIDES = {'A','G','C','T'}  (code 0,1,2,3)
const char NUCLEOTIDES[4] = {'A', 'G', 'C', 'T'};

// ITMO!/dna/DnaTools.java:46-64 fromChar.  IUPAC ambiguity codes are replaced by a *random*
// base in the reference (DnaTools.java:66-117, unseeded Random :35): not reproducible, so the
// oracle rejects them; every other char -> IllegalArgumentException.
int from_char(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'G': case 'g': return 1;
        case 'C': case 'c': return 2;
        case 'T': case 't': return 3;
        default: break;
    }
    const char* iupac = "RrYyMmKkSsWwHhBbVvDd";
    if (c != 0 && strchr(iupac, c))
        fail("IUPAC ambiguity code \"%c\": the reference substitutes a random base (DnaTools.java:66-117); not reproducible", c);
    fail("Incorrect nucleotide char: \"%c\"", c);
}

// ITMO!/dna/DnaTools.java:139-145 reverseComplement(String): complement(char) = toChar(fromChar(c) ^ 3)
std::string rc_string(const std::string& s) {
    std::string r(s.size(), 'A');
    for (size_t i = 0; i < s.size(); i++) r[s.size() - 1 - i] = NUCLEOTIDES[from_char(s[i]) ^ 3];
    return r;
}

// src/utils/StringUtils.java:34-41 normalizeDna: ASCII-lexicographic min of s and rc(s) (ties -> rc, same string)
std::string normalize_dna(const std::string& s) {
    std::string rc = rc_string(s);
    return s.compare(rc) < 0 ? s : rc;
}

// ITMO!/utils/KmerUtils.java:12-22
uint64_t revcomp_bits(uint64_t kmer, int k) {
    kmer = ((kmer & 0x3333333333333333ULL) << 2) | ((kmer & 0xccccccccccccccccULL) >> 2);
    kmer = ((kmer & 0x0f0f0f0f0f0f0f0fULL) << 4) | ((kmer & 0xf0f0f0f0f0f0f0f0ULL) >> 4);
    kmer = ((kmer & 0x00ff00ff00ff00ffULL) << 8) | ((kmer & 0xff00ff00ff00ff00ULL) >> 8);
    kmer = ((kmer & 0x0000ffff0000ffffULL) << 16) | ((kmer & 0xffff0000ffff0000ULL) >> 16);
    kmer = ((kmer & 0x00000000ffffffffULL) << 32) | ((kmer & 0xffffffff00000000ULL) >> 32);
    kmer = ~kmer;
    return kmer >> (64 - 2 * k);
}

// ITMO!/utils/KmerUtils.java:58-60 getKmerKey = Math.min(kmer, rc) (signed; 2k <= 62 bits so same as unsigned)
uint64_t kmer_key(uint64_t kmer, int k) {
    uint64_t rc = revcomp_bits(kmer, k);
    return (int64_t)kmer < (int64_t)rc ? kmer : rc;
}

// ITMO!/dna/DnaTools.java:123-129 toLong(LightDna): r = (r << 2) + nuc
uint64_t to_long(const std::string& s) {
    uint64_t r = 0;
    for (char c : s) r = (r << 2) + (uint64_t)from_char(c);
    return r;
}

// ITMO!/utils/KmerUtils.java:50-57
std::string kmer2string(uint64_t kmer, int k) {
    std::string s;
    for (int i = k - 1; i >= 0; --i) s.push_back(NUCLEOTIDES[(kmer >> (2 * i)) & 3]);
    return s;
}

// java.lang.String.hashCode: s[0]*31^(n-1) + ... (int32 wraparound)
int32_t java_string_hash(const std::string& s) {
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return (int32_t)h;
}

// =====================================================================================
// java.util.HashMap<String, V> iteration-order emulation.
// Buckets are visited in index order; inside a bucket entries keep insertion order (resize
// splits preserve relative order, java.util.HashMap.resize()).  Capacity: 16, doubled whenever
// ++size > 0.75*capacity.  A map in which some bin reached 9 nodes (treeifyBin) is replayed put by put
// through JdkHashMap below, a literal restatement of java.util.HashMap's putVal / resize / treeifyBin /
// TreeNode.{treeify, putTreeVal, split, untreeify, balanceInsertion, rotate*, moveRootToFront}.
// Pinned by tests/golden: the reference's committed graph.txt is in exactly this order.
// =====================================================================================
// java.util.HashMap<String,?> for keys that are all NEW (the caller de-duplicates), statement by statement after the JDK (8u .. 21:
// these methods did not change).  Only what decides the iteration order is kept: no values, no removal.
struct JdkHashMap {
    struct Node {
        int hash;                       // spread hash
        const std::string* key;
        Node* next = nullptr;
        // TreeNode part
        Node *parent = nullptr, *left = nullptr, *right = nullptr, *prev = nullptr;
        bool red = false, is_tree = false;
    };
    static const int TREEIFY_THRESHOLD = 8, UNTREEIFY_THRESHOLD = 6, MIN_TREEIFY_CAPACITY = 64;
    std::deque<Node> pool;
    std::vector<Node*> table;
    size_t threshold = 0, size = 0;

    static int hash_of(const std::string& k) { uint32_t h = (uint32_t)java_string_hash(k); return (int)(h ^ (h >> 16)); }
    static int compare_keys(const Node* x_new, const Node* p) { return x_new->key->compare(*p->key); }   // String.compareTo (sign only)

    void put(const std::string& key) {   // putVal
        if (table.empty()) resize();
        pool.emplace_back();
        Node* x = &pool.back();
        x->hash = hash_of(key);
        x->key = &key;
        const size_t n = table.size(), i = (n - 1) & (size_t)(uint32_t)x->hash;
        Node* p = table[i];
        if (p == nullptr) {
            table[i] = x;
        } else if (p->is_tree) {
            put_tree_val(p, x);
        } else {
            for (int binCount = 0;; ++binCount) {
                Node* e = p->next;
                if (e == nullptr) {
                    p->next = x;
                    if (binCount >= TREEIFY_THRESHOLD - 1) treeify_bin(x->hash);
                    break;
                }
                p = e;
            }
        }
        if (++size > threshold) resize();
    }
    void resize() {
        if (table.empty()) {
            table.assign(16, nullptr);
            threshold = 12;
            return;
        }
        const size_t oldCap = table.size(), newCap = oldCap << 1;
        threshold <<= 1;
        std::vector<Node*> oldTab(newCap, nullptr);
        oldTab.swap(table);   // table = newTab
        for (size_t j = 0; j < oldCap; ++j) {
            Node* e = oldTab[j];
            if (e == nullptr) continue;
            if (e->next == nullptr) {
                table[(size_t)(uint32_t)e->hash & (newCap - 1)] = e;
            } else if (e->is_tree) {
                split(e, j, oldCap);
            } else {   // preserve order
                Node *loHead = nullptr, *loTail = nullptr, *hiHead = nullptr, *hiTail = nullptr, *next;
                do {
                    next = e->next;
                    if (((size_t)(uint32_t)e->hash & oldCap) == 0) {
                        if (loTail == nullptr) loHead = e; else loTail->next = e;
                        loTail = e;
                    } else {
                        if (hiTail == nullptr) hiHead = e; else hiTail->next = e;
                        hiTail = e;
                    }
                } while ((e = next) != nullptr);
                if (loTail != nullptr) { loTail->next = nullptr; table[j] = loHead; }
                if (hiTail != nullptr) { hiTail->next = nullptr; table[j + oldCap] = hiHead; }
            }
        }
    }
    void treeify_bin(int hash) {
        const size_t n = table.size();
        if (n < (size_t)MIN_TREEIFY_CAPACITY) { resize(); return; }
        const size_t index = (n - 1) & (size_t)(uint32_t)hash;
        Node* e = table[index];
        if (e == nullptr) return;
        Node *hd = nullptr, *tl = nullptr;
        do {   // replacementTreeNode(e, null)
            e->is_tree = true;
            e->parent = e->left = e->right = nullptr;
            if (tl == nullptr) hd = e;
            else { e->prev = tl; tl->next = e; }
            tl = e;
        } while ((e = e->next) != nullptr);
        hd->prev = nullptr;
        table[index] = hd;
        treeify(hd);
    }
    static int direction(const Node* x, const Node* p) {   // of x relative to p: hashes as ints, then compareTo
        if (p->hash > x->hash) return -1;
        if (p->hash < x->hash) return 1;
        return compare_keys(x, p) < 0 ? -1 : 1;   // distinct keys: never 0, tieBreakOrder is unreachable
    }
    void treeify(Node* self) {
        Node* root = nullptr;
        for (Node *x = self, *next; x != nullptr; x = next) {
            next = x->next;
            x->left = x->right = nullptr;
            if (root == nullptr) {
                x->parent = nullptr;
                x->red = false;
                root = x;
            } else {
                for (Node* p = root;;) {
                    const int dir = direction(x, p);
                    Node* xp = p;
                    if ((p = (dir <= 0) ? p->left : p->right) == nullptr) {
                        x->parent = xp;
                        if (dir <= 0) xp->left = x; else xp->right = x;
                        root = balance_insertion(root, x);
                        break;
                    }
                }
            }
        }
        move_root_to_front(root);
    }
    void put_tree_val(Node* self, Node* x) {
        Node* root = self;
        while (root->parent != nullptr) root = root->parent;
        for (Node* p = root;;) {
            const int dir = direction(x, p);
            Node* xp = p;
            if ((p = (dir <= 0) ? p->left : p->right) == nullptr) {
                Node* xpn = xp->next;
                x->is_tree = true;
                x->next = xpn;
                if (dir <= 0) xp->left = x; else xp->right = x;
                xp->next = x;
                x->parent = x->prev = xp;
                if (xpn != nullptr) xpn->prev = x;
                move_root_to_front(balance_insertion(root, x));
                return;
            }
        }
    }
    void move_root_to_front(Node* root) {
        if (root == nullptr || table.empty()) return;
        const size_t index = (table.size() - 1) & (size_t)(uint32_t)root->hash;
        Node* first = table[index];
        if (root != first) {
            table[index] = root;
            Node *rp = root->prev, *rn = root->next;
            if (rn != nullptr) rn->prev = rp;
            if (rp != nullptr) rp->next = rn;
            if (first != nullptr) first->prev = root;
            root->next = first;
            root->prev = nullptr;
        }
    }
    static Node* untreeify(Node* self) {   // plain nodes, same order
        for (Node* q = self; q != nullptr; q = q->next) {
            q->is_tree = false;
            q->parent = q->left = q->right = q->prev = nullptr;
        }
        return self;
    }
    void split(Node* b, size_t index, size_t bit) {
        Node *loHead = nullptr, *loTail = nullptr, *hiHead = nullptr, *hiTail = nullptr;
        int lc = 0, hc = 0;
        for (Node *e = b, *next; e != nullptr; e = next) {
            next = e->next;
            e->next = nullptr;
            if (((size_t)(uint32_t)e->hash & bit) == 0) {
                if ((e->prev = loTail) == nullptr) loHead = e; else loTail->next = e;
                loTail = e;
                ++lc;
            } else {
                if ((e->prev = hiTail) == nullptr) hiHead = e; else hiTail->next = e;
                hiTail = e;
                ++hc;
            }
        }
        if (loHead != nullptr) {
            if (lc <= UNTREEIFY_THRESHOLD) table[index] = untreeify(loHead);
            else {
                table[index] = loHead;
                if (hiHead != nullptr) treeify(loHead);   // (else is already treeified)
            }
        }
        if (hiHead != nullptr) {
            if (hc <= UNTREEIFY_THRESHOLD) table[index + bit] = untreeify(hiHead);
            else {
                table[index + bit] = hiHead;
                if (loHead != nullptr) treeify(hiHead);
            }
        }
    }
    static Node* rotate_left(Node* root, Node* p) {
        Node *r, *pp, *rl;
        if (p != nullptr && (r = p->right) != nullptr) {
            if ((rl = p->right = r->left) != nullptr) rl->parent = p;
            if ((pp = r->parent = p->parent) == nullptr) (root = r)->red = false;
            else if (pp->left == p) pp->left = r;
            else pp->right = r;
            r->left = p;
            p->parent = r;
        }
        return root;
    }
    static Node* rotate_right(Node* root, Node* p) {
        Node *l, *pp, *lr;
        if (p != nullptr && (l = p->left) != nullptr) {
            if ((lr = p->left = l->right) != nullptr) lr->parent = p;
            if ((pp = l->parent = p->parent) == nullptr) (root = l)->red = false;
            else if (pp->right == p) pp->right = l;
            else pp->left = l;
            l->right = p;
            p->parent = l;
        }
        return root;
    }
    static Node* balance_insertion(Node* root, Node* x) {
        x->red = true;
        for (Node *xp, *xpp, *xppl, *xppr;;) {
            if ((xp = x->parent) == nullptr) {
                x->red = false;
                return x;
            } else if (!xp->red || (xpp = xp->parent) == nullptr) {
                return root;
            }
            if (xp == (xppl = xpp->left)) {
                if ((xppr = xpp->right) != nullptr && xppr->red) {
                    xppr->red = false;
                    xp->red = false;
                    xpp->red = true;
                    x = xpp;
                } else {
                    if (x == xp->right) {
                        root = rotate_left(root, x = xp);
                        xpp = (xp = x->parent) == nullptr ? nullptr : xp->parent;
                    }
                    if (xp != nullptr) {
                        xp->red = false;
                        if (xpp != nullptr) {
                            xpp->red = true;
                            root = rotate_right(root, xpp);
                        }
                    }
                }
            } else {
                if (xppl != nullptr && xppl->red) {
                    xppl->red = false;
                    xp->red = false;
                    xpp->red = true;
                    x = xpp;
                } else {
                    if (x == xp->left) {
                        root = rotate_right(root, x = xp);
                        xpp = (xp = x->parent) == nullptr ? nullptr : xp->parent;
                    }
                    if (xp != nullptr) {
                        xp->red = false;
                        if (xpp != nullptr) {
                            xpp->red = true;
                            root = rotate_left(root, xpp);
                        }
                    }
                }
            }
        }
    }
};

template <class V>
struct JHashMap {
    std::unordered_map<std::string, size_t> index;   // key -> position in `entries`
    std::vector<std::pair<std::string, V>> entries;  // insertion order
    size_t peak = 0;                                 // largest size ever (java.util.HashMap never shrinks its table on removal)

    bool contains(const std::string& k) const { return index.count(k) != 0; }
    size_t size() const { return entries.size(); }
    void put(const std::string& k, const V& v) {
        auto it = index.find(k);
        if (it == index.end()) {
            index.emplace(k, entries.size());
            entries.emplace_back(k, v);
            peak = std::max(peak, entries.size());
        } else {
            entries[it->second].second = v;
        }
    }
    // keySet().retainAll(keep): survivors keep their relative (bucket-internal) order
    void retain_all(const std::unordered_set<std::string>& keep) {
        std::vector<std::pair<std::string, V>> kept;
        for (auto& e : entries)
            if (keep.count(e.first)) kept.push_back(std::move(e));
        entries.swap(kept);
        index.clear();
        for (size_t i = 0; i < entries.size(); i++) index.emplace(entries[i].first, i);
    }
    const V* get(const std::string& k) const {
        auto it = index.find(k);
        return it == index.end() ? nullptr : &entries[it->second].second;
    }
    V* getm(const std::string& k) {
        auto it = index.find(k);
        return it == index.end() ? nullptr : &entries[it->second].second;
    }
    static uint32_t spread(int32_t h) { return (uint32_t)h ^ ((uint32_t)h >> 16); }
    size_t capacity() const {
        size_t cap = 16;
        while ((double)std::max(peak, entries.size()) > 0.75 * (double)cap) cap <<= 1;
        return cap;
    }
    // true when, filling a default HashMap with the entries in their order, some bin reaches 9 nodes (the order then depends on the
    // history: treeified bins, or an early resize of a table under 64 buckets)
    bool needs_replay() const {
        size_t cap = 16, lo = 0;
        std::vector<uint32_t> cnt;
        while (lo < entries.size()) {
            const size_t hi = std::min(entries.size(), cap / 4 * 3 + 1);
            cnt.assign(cap, 0);
            for (size_t i = 0; i < hi; i++)
                if (++cnt[spread(java_string_hash(entries[i].first)) & (uint32_t)(cap - 1)] >= 9 && i >= lo) return true;
            lo = hi;
            cap <<= 1;
        }
        return false;
    }
    // positions into `entries` in java iteration order
    std::vector<size_t> order() const {
        if (entries.size() >= 9 && peak <= entries.size() && needs_replay()) {
            JdkHashMap m;
            for (auto& e : entries) m.put(e.first);
            std::unordered_map<const std::string*, size_t> where;
            for (size_t i = 0; i < entries.size(); i++) where.emplace(&entries[i].first, i);
            std::vector<size_t> o;
            o.reserve(entries.size());
            for (JdkHashMap::Node* e : m.table)
                for (; e != nullptr; e = e->next) o.push_back(where[e->key]);
            return o;
        }
        size_t cap = capacity();
        std::vector<std::pair<uint32_t, size_t>> v(entries.size());
        for (size_t i = 0; i < entries.size(); i++)
            v[i] = {spread(java_string_hash(entries[i].first)) & (uint32_t)(cap - 1), i};
        std::stable_sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first < b.first; });
        std::vector<size_t> o(v.size());
        for (size_t i = 0; i < v.size(); i++) o[i] = v[i].second;
        return o;
    }
};

// =====================================================================================
// java.io.BufferedReader.readLine over an in-memory byte buffer:
// a line ends at '\n', '\r' or "\r\n"; EOF without terminator still yields the last line.
// =====================================================================================
struct LineReader {
    const uint8_t* p;
    const uint8_t* end;
    LineReader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
    bool ready() const { return p < end; }   // BufferedReader.ready(): more chars available
    bool read_line(std::string& out) {
        if (p >= end) return false;
        const uint8_t* s = p;
        while (p < end && *p != '\n' && *p != '\r') p++;
        out.assign((const char*)s, (size_t)(p - s));
        if (p < end) {
            if (*p == '\r' && p + 1 < end && p[1] == '\n') p += 2;
            else p += 1;
        }
        return true;
    }
};

// =====================================================================================
// A1: FASTQ  (ITMO!/io/readers/FastqReader.java:53-110)
// =====================================================================================
struct IllegalQualityValue : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ITMO!/io/formats/Illumina.java:7-12 / Sanger.java:7-12
int get_phred(int offset, unsigned char c) {
    if (c < offset || c > 126) {
        char b[96];
        snprintf(b, sizeof b, "Invalid quality code char: \"%c\" char code = %d", c, (int)c);
        throw IllegalQualityValue(b);
    }
    return c - offset;
}

struct FastqIterator {
    LineReader lr;
    int offset;  // 33 Sanger, 64 Illumina
    bool closed = false;
    FastqIterator(const uint8_t* b, size_t n, int off) : lr(b, n), offset(off) {}

    // FastqReader.java:84-110.  Returns false at (clean) EOF.
    bool read_next_data_line(std::string& out) {
        if (closed) return false;
        std::string s;
        bool ok = lr.read_line(s);
        while (ok && s.empty()) ok = lr.read_line(s);   // skipping empty lines
        if (!ok) { closed = true; return false; }
        if (!(s[0] == '@' || s[0] == '+'))
            fail("Unknown structure of fastq file! Waiting \"@ID\" or \"+ID\" string, found \"%s\".",
                 (s.size() > 20 ? s.substr(0, 20) + "..." : s).c_str());
        if (!lr.read_line(out)) fail("Unexpected end of file. File is corrupted/Format mismatch.");
        return true;
    }

    // FastqReader.java:53-82.  value[i] = (byte)(nuc | phred << 2)  (DnaQBuilder.java:32-35)
    bool read_next(std::vector<uint8_t>& value) {
        std::string data, qual;
        if (!read_next_data_line(data)) return false;
        if (!read_next_data_line(qual)) fail("Unexpected end of file. File is corrupted/Format mismatch.");
        if (data.size() != qual.size())
            fail("Bad DnaQ record: length of chars and quality is not the same. File is corrupted/Format mismatch.");
        value.resize(data.size());
        for (size_t i = 0; i < data.size(); i++) {
            if (data[i] == 'N' || data[i] == 'n' || data[i] == '.') {
                value[i] = 0;                                  // unsafeAppendUnknown
            } else {
                int nuc = from_char(data[i]);
                int phred = get_phred(offset, (unsigned char)qual[i]);
                value[i] = (uint8_t)((nuc | (phred << 2)) & 0xFF);
            }
        }
        return true;
    }
};

// ITMO!/io/ReadersUtils.java:63-77: try Phred+64 on the first 1000 records; IllegalQualityValue => Phred+33.
int determine_quality_offset(const uint8_t* b, size_t n) {
    try {
        FastqIterator it(b, n, 64);
        std::vector<uint8_t> v;
        for (int i = 0; i < 1000 && it.read_next(v); i++) {}
        return 64;
    } catch (const IllegalQualityValue&) {
        return 33;
    }
}

// ITMO!/io/readers/FastaReaderFromXQSourceTrunc.java:61-94 + ITMO!/dna/DnaQ.java:147-150,176-186:
// stored phred = (value >> 2) & 63; the read is cut at every base whose stored phred < 1; each maximal run
// (also empty ones) is emitted as one Dna (nucleotide = value & 3).  Emission order inside a read is
// irrelevant for counting and not reproduced.
void split_by_quality(const std::vector<uint8_t>& value, const std::function<void(const uint8_t*, size_t)>& emit,
                      std::vector<uint8_t>& scratch) {
    size_t n = value.size(), start = 0;
    while (true) {
        size_t t = start;
        while (t < n && ((value[t] >> 2) & 63) >= 1) t++;   // inplaceTruncateByQuality(1)
        scratch.resize(t - start);
        for (size_t i = start; i < t; i++) scratch[i - start] = value[i] & 3;
        emit(scratch.data(), scratch.size());
        if (n > t + 1) start = t + 1;   // "dnaQ.length() > goodDnaQ.length() + 1" -> nextPart
        else break;
    }
}

// =====================================================================================
// A1': FASTA (ITMO!/io/readers/FastaReader.java:54-108)
// =====================================================================================
struct FastaIterator {
    LineReader lr;
    bool closed = false;
    FastaIterator(const uint8_t* b, size_t n) : lr(b, n) {}
    static bool is_comment(const std::string& s) { return !s.empty() && (s[0] == '>' || s[0] == ';'); }
    // FastaReader.java:78-104
    bool read_next_data_line(std::string& sb) {
        if (closed) return false;
        sb.clear();
        std::string s;
        while (true) {
            if (!lr.read_line(s)) { closed = true; break; }
            if (is_comment(s)) {
                if (!sb.empty()) break;
            } else {
                sb += s;
            }
        }
        return !sb.empty();
    }
    // FastaReader.java:54-76: records containing 'N' or 'n' are skipped whole; then new Dna(s) (Dna.java:59-64)
    bool read_next(std::vector<uint8_t>& nucs) {
        std::string s;
        while (read_next_data_line(s)) {
            if (s.find('N') != std::string::npos || s.find('n') != std::string::npos) continue;
            nucs.resize(s.size());
            for (size_t i = 0; i < s.size(); i++) nucs[i] = (uint8_t)from_char(s[i]);
            return true;
        }
        return false;
    }
};

// =====================================================================================
// B1: the count map.  Semantics of BigLong2ShortHashMap (ITMO!/structures/map/BigLong2ShortHashMap.java:40-77)
// over Long2ShortHashMap.addAndBound/get (Long2ShortHashMap.java:118-175) and NumUtils.addAndBound(short,short)
// (ITMO!/utils/NumUtils.java:21-26): exact count saturating at 32767, get() = -1 for absent keys.
// Structure kept: 2^(floor(log2 P)+4) small maps picked by murmurHash3((int)key) (src/io/IOUtils.java:220-221),
// each open-addressing with linear probing from murmurHash3(key) (LongHashSet.java:153-165), 2^12 initial slots,
// doubling at load 0.75, one lock per small map.  FREE key = 0 has a side slot (Long2ShortHashMap.java:120-134).
// Placement only affects iteration order, never results.
// =====================================================================================
inline uint32_t murmur3_int(uint32_t x) {   // fastutil HashCommon.murmurHash3(int)
    x ^= x >> 16; x *= 0x85ebca6bU; x ^= x >> 13; x *= 0xc2b2ae35U; x ^= x >> 16;
    return x;
}
inline uint64_t murmur3_long(uint64_t x) {  // fastutil HashCommon.murmurHash3(long)
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
inline int16_t add_and_bound(int16_t value, int16_t inc) {   // NumUtils.java:21-26
    if (value > 32767 - inc) return 32767;
    return (int16_t)(value + inc);
}

struct SmallMap {
    std::mutex lock;
    std::vector<uint64_t> keys;
    std::vector<int16_t> values;
    size_t capacity, size = 0, max_fill;
    bool contains_free = false;
    int16_t value_for_free = 0;
    explicit SmallMap(int log_cap = 12) : capacity((size_t)1 << log_cap) {
        keys.assign(capacity, 0);
        values.assign(capacity, 0);
        max_fill = (size_t)(capacity * 0.75);
    }
    size_t position(uint64_t key) const {   // LongHashSet.getPositionInt
        size_t mask = capacity - 1, pos = (size_t)murmur3_long(key) & mask;
        while (keys[pos] != 0 && keys[pos] != key) pos = (pos + 1) & mask;
        return pos;
    }
    void enlarge() {   // Long2ShortHashMap.java:191-214
        std::vector<uint64_t> ok;
        std::vector<int16_t> ov;
        ok.swap(keys);
        ov.swap(values);
        capacity *= 2;
        keys.assign(capacity, 0);
        values.assign(capacity, 0);
        max_fill = (size_t)(capacity * 0.75);
        for (size_t i = 0; i < ok.size(); i++)
            if (ok[i] != 0) {
                size_t pos = position(ok[i]);
                keys[pos] = ok[i];
                values[pos] = ov[i];
            }
    }
    void add(uint64_t key, int16_t inc) {   // addAndBound under the write lock
        std::lock_guard<std::mutex> g(lock);
        if (key == 0) {
            value_for_free = add_and_bound(value_for_free, inc);
            if (!contains_free) { contains_free = true; size++; }
            return;
        }
        size_t pos = position(key);
        values[pos] = add_and_bound(values[pos], inc);
        if (keys[pos] == 0) {
            keys[pos] = key;
            if (++size >= max_fill) enlarge();
        }
    }
    int16_t get(uint64_t key) const {   // Long2ShortHashMap.java:159-175
        if (key == 0) return contains_free ? value_for_free : (int16_t)-1;
        size_t pos = position(key);
        return keys[pos] == key ? values[pos] : (int16_t)-1;
    }
};

}  // namespace

struct orc_table {
    int k;
    int threads;
    uint32_t mask;
    std::vector<std::unique_ptr<SmallMap>> maps;
    std::atomic<uint64_t> n_reads{0}, n_runs{0}, n_bases{0}, n_kmers{0};

    orc_table(int k_, int threads_) : k(k_), threads(std::max(1, threads_)) {
        int log_n = (int)(std::log((double)threads) / std::log(2.0)) + 4;   // src/io/IOUtils.java:220-221
        size_t n = (size_t)1 << log_n;
        mask = (uint32_t)(n - 1);
        for (size_t i = 0; i < n; i++) maps.emplace_back(new SmallMap(12));
    }
    void add(uint64_t key, int16_t inc) { maps[murmur3_int((uint32_t)key) & mask]->add(key, inc); }
    int16_t get(uint64_t key) const { return maps[murmur3_int((uint32_t)key) & mask]->get(key); }
    uint64_t size() const {
        uint64_t s = 0;
        for (auto& m : maps) s += m->size;
        return s;
    }
    template <class F>
    void for_each(F f) const {   // entryIterator(): map by map, free key first is irrelevant for sorted use
        for (auto& m : maps) {
            if (m->contains_free) f((uint64_t)0, m->value_for_free);
            for (size_t i = 0; i < m->capacity; i++)
                if (m->keys[i] != 0) f(m->keys[i], m->values[i]);
        }
    }
};

namespace {

// A2: ShortKmer.kmersOf (ITMO!/dna/kmers/ShortKmer.java:104-149): first k-mer from bases 0..k-1, then
// shiftRight (:68-71) fw=((fw<<2)|nuc)&mask, rc=(rc>>2)|((3-nuc)<<(2k-2)); key = min(fw,rc) (:54-56).
// ReadsLoadWorker.process (src/io/IOUtils.java:201-214) with minDnaLen = 0.
void count_run(orc_table* t, const uint8_t* nucs, size_t n) {
    int k = t->k;
    t->n_runs++;
    t->n_bases += n;
    if (n < (size_t)k) return;
    uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t fw = 0;
    for (int i = 0; i < k; i++) fw = (fw << 2) | nucs[i];
    uint64_t rc = revcomp_bits(fw, k);
    t->add(std::min<int64_t>((int64_t)fw, (int64_t)rc), 1);
    uint64_t cnt = 1;
    for (size_t i = k; i < n; i++) {
        uint64_t nuc = nucs[i];
        fw = ((fw << 2) | nuc) & mask;
        rc = (rc >> 2) | ((3ULL - nuc) << (2 * k - 2));
        t->add((uint64_t)std::min<int64_t>((int64_t)fw, (int64_t)rc), 1);
        cnt++;
    }
    t->n_kmers += cnt;
}

// One source -> sequence of Dna runs.  fmt: 0 fasta, 1 fastq(sniff), 33/64 forced.
struct RunSource {
    int fmt;
    std::unique_ptr<FastqIterator> fq;
    std::unique_ptr<FastaIterator> fa;
    std::deque<std::vector<uint8_t>> pending;
    std::vector<uint8_t> value, scratch;
    uint64_t reads = 0;
    RunSource(const uint8_t* b, size_t n, int fmt_) : fmt(fmt_) {
        if (fmt == 0) fa.reset(new FastaIterator(b, n));
        else {
            int off = fmt == 1 ? determine_quality_offset(b, n) : fmt;
            fq.reset(new FastqIterator(b, n, off));
        }
    }
    bool next(std::vector<uint8_t>& run) {
        while (pending.empty()) {
            if (fa) {
                if (!fa->read_next(run)) return false;
                reads++;
                return true;
            }
            if (!fq->read_next(value)) return false;
            reads++;
            split_by_quality(value, [&](const uint8_t* p, size_t n) { pending.emplace_back(p, p + n); }, scratch);
        }
        run.swap(pending.front());
        pending.pop_front();
        return true;
    }
};

// src/io/IOUtils.java:283-316 run() + src/io/ReadsDispatcher.java:34-53 + src/io/ReadsWorker.java:29-41:
// P workers pull ranges of 32768 Dna from a synchronized dispatcher (parsing happens inside the monitor).
void count_source(orc_table* t, const uint8_t* buf, size_t len, int fmt) {
    RunSource src(buf, len, fmt);
    const size_t RANGE = 1 << 15;   // READS_WORK_RANGE_SIZE, src/io/IOUtils.java:25
    int P = t->threads;
    if (P <= 1) {
        std::vector<uint8_t> run;
        while (src.next(run)) count_run(t, run.data(), run.size());
        t->n_reads += src.reads;
        return;
    }
    std::mutex monitor;
    std::string first_error;
    auto get_work_range = [&](std::vector<std::vector<uint8_t>>& list) {
        std::lock_guard<std::mutex> g(monitor);
        list.clear();
        if (!first_error.empty()) return false;
        try {
            std::vector<uint8_t> run;
            while (list.size() < RANGE && src.next(run)) list.emplace_back(std::move(run));
        } catch (const std::exception& e) {
            first_error = e.what();
            return false;
        }
        return !list.empty();
    };
    std::vector<std::thread> workers;
    for (int i = 0; i < P; i++)
        workers.emplace_back([&] {
            std::vector<std::vector<uint8_t>> list;
            while (get_work_range(list))
                for (auto& r : list) count_run(t, r.data(), r.size());
        });
    for (auto& w : workers) w.join();
    t->n_reads += src.reads;
    if (!first_error.empty()) throw OracleError(first_error);
}

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail("Failed to read from file %s", path.c_str());   // src/io/IOUtils.java:292
    std::vector<uint8_t> b;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    b.resize((size_t)n);
    if (n > 0 && fread(b.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); fail("Failed to read from file %s", path.c_str()); }
    fclose(f);
    return b;
}

// FastqGZReader / FastaGZReader (ITMO!/io/readers/FastqGZReader.java:26-34, FastaGZReader.java:24-32): the plain readers over a
// java.util.zip.GZIPInputStream, which also reads concatenated members -- like zlib's gzread
std::vector<uint8_t> read_gz_file(const std::string& path) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) fail("Failed to read from file %s", path.c_str());
    gzbuffer(g, 1 << 20);
    std::vector<uint8_t> b;
    std::vector<uint8_t> piece(8 << 20);
    for (;;) {
        int got = gzread(g, piece.data(), (unsigned)piece.size());
        if (got < 0) { gzclose(g); fail("Failed to read from file %s", path.c_str()); }
        if (got == 0) break;
        b.insert(b.end(), piece.begin(), piece.begin() + got);
    }
    gzclose(g);
    return b;
}


// libbz2 through dlopen: the image ships libbz2.so.1.0 but no bzlib.h; the structure and the three entry points below are
// the library's stable public ABI (bzlib.h 1.0.x).  Concatenated streams are read one after the other, like
// org.apache.hadoop.io.compress.BZip2Codec (ITMO!/io/readers/FastqBZ2Reader.java:26-34, FastaBZ2Reader.java).
struct OrcBzStream {
    char* next_in;
    unsigned int avail_in, total_in_lo32, total_in_hi32;
    char* next_out;
    unsigned int avail_out, total_out_lo32, total_out_hi32;
    void* state;
    void* (*bzalloc)(void*, int, int);
    void (*bzfree)(void*, void*);
    void* opaque;
};
std::vector<uint8_t> read_bz2_file(const std::string& path) {
    typedef int (*init_t)(OrcBzStream*, int, int);
    typedef int (*step_t)(OrcBzStream*);
    void* lib = dlopen("libbz2.so.1.0", RTLD_NOW);
    if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW);
    if (!lib) fail("Failed to read from file %s (libbz2 is not installed)", path.c_str());
    init_t init = (init_t)dlsym(lib, "BZ2_bzDecompressInit");
    step_t step = (step_t)dlsym(lib, "BZ2_bzDecompress"), end = (step_t)dlsym(lib, "BZ2_bzDecompressEnd");
    if (!init || !step || !end) fail("Failed to read from file %s (libbz2 lacks the decompression entry points)", path.c_str());
    std::vector<uint8_t> in = read_file(path), out;
    std::vector<uint8_t> piece(8 << 20);
    size_t pos = 0;
    while (pos < in.size()) {
        OrcBzStream z{};
        if (init(&z, 0, 0) != 0) fail("Failed to read from file %s", path.c_str());
        z.next_in = (char*)in.data() + pos;
        z.avail_in = (unsigned)std::min<size_t>(in.size() - pos, 1u << 30);
        int rc = 0;
        while (rc != 4) {   // BZ_STREAM_END
            z.next_out = (char*)piece.data();
            z.avail_out = (unsigned)piece.size();
            const unsigned in_before = z.avail_in;
            rc = step(&z);
            if (rc != 0 && rc != 4) { end(&z); fail("Failed to read from file %s (corrupt bzip2 stream)", path.c_str()); }
            out.insert(out.end(), piece.begin(), piece.begin() + (piece.size() - z.avail_out));
            pos += in_before - z.avail_in;
            if (rc == 0 && z.avail_in == 0) {
                if (pos >= in.size()) { end(&z); fail("Failed to read from file %s (truncated bzip2 stream)", path.c_str()); }
                z.next_in = (char*)in.data() + pos;
                z.avail_in = (unsigned)std::min<size_t>(in.size() - pos, 1u << 30);
            }
        }
        end(&z);
    }
    return out;
}

// ITMO!/io/ReadersUtils.java:27-54; *gz = 1 / 2: the ".gz" / ".bz2" suffix was stripped first (:36-43)
int detect_format(const std::string& path, int* gz = nullptr) {
    size_t slash = path.find_last_of('/');
    std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    for (auto& c : name) c = (char)tolower((unsigned char)c);
    auto ends = [&](const char* s) { size_t n = strlen(s); return name.size() >= n && name.compare(name.size() - n, n, s) == 0; };
    if (gz) *gz = 0;
    if (gz && ends(".gz")) {
        *gz = 1;
        name.resize(name.size() - 3);
    }
    if (gz && ends(".bz2")) {   // ReadersUtils strips ".gz" and then ".bz2" (a ".bz2.gz" name reads as bz2: kept)
        *gz = 2;
        name.resize(name.size() - 4);
    }
    if (ends(".gz") || ends(".bz2")) fail("compressed input is out of scope for the hot path: '%s'", name.c_str());
    if (ends(".binq")) fail("binq input is out of scope for the hot path: '%s'", name.c_str());
    if (ends(".fastq") || ends(".fq")) return 1;
    if (ends(".fasta") || ends(".fa") || ends(".fn") || ends(".fna")) return 0;
    fail("Can't detect file format for file '%s'", name.c_str());
}

void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i < path.size(); i++) {
        cur.push_back(path[i]);
        if (path[i] == '/' || i + 1 == path.size()) mkdir(cur.c_str(), 0777);
    }
}

// =====================================================================================
// C0: src/io/RichFastaReader.java:38-76 (readAll) + new DnaQ(str, 0) (ITMO!/dna/DnaQ.java:21-30) + toString
// =====================================================================================
struct RichFasta {
    std::vector<std::string> dnas, comments;
};
std::string dnaq_string(const std::string& raw) {   // N/n/. -> nuc 0 ('A'); others through fromChar; printed upper-case
    std::string s(raw.size(), 'A');
    for (size_t i = 0; i < raw.size(); i++) {
        char c = raw[i];
        if (c == 'N' || c == 'n' || c == '.') s[i] = 'A';
        else s[i] = NUCLEOTIDES[from_char(c)];
    }
    return s;
}
RichFasta rich_fasta_read(const std::string& path) {
    std::vector<uint8_t> b = read_file(path);
    LineReader lr(b.data(), b.size());
    RichFasta r;
    bool last_comment = true;
    std::string cur_comment, cur_dna, line;
    while (lr.ready()) {
        lr.read_line(line);
        if (!line.empty() && (line[0] == '>' || line[0] == ';')) {
            if (!last_comment) {
                r.dnas.push_back(dnaq_string(cur_dna));
                cur_dna.clear();
                cur_comment.clear();
            }
            cur_comment += line.substr(1);
            last_comment = true;
        } else {
            if (last_comment) {
                r.comments.push_back(cur_comment);
                cur_dna.clear();
                cur_comment.clear();
            }
            cur_dna += line;
            last_comment = false;
        }
    }
    if (!cur_comment.empty()) r.comments.push_back(cur_comment);
    if (!cur_dna.empty()) r.dnas.push_back(dnaq_string(cur_dna));
    return r;
}

// =====================================================================================
// C1: BFS (src/algo/OneSequenceCalculator.java:154-239, src/algo/TerminationMode.java:31-47,
//          src/utils/StringUtils.java:8-32)
// =====================================================================================
std::vector<std::string> neighbors_by_dir(int dir, const std::string& kmer) {
    std::vector<std::string> left(4), right(4);
    for (int i = 0; i < 4; i++) {
        left[i] = std::string(1, NUCLEOTIDES[i]) + kmer.substr(0, kmer.size() - 1);
        right[i] = kmer.substr(1) + NUCLEOTIDES[i];
    }
    if (dir == -1) return left;
    if (dir == 1) return right;
    std::vector<std::string> all(8);
    for (int i = 0; i < 4; i++) {
        all[2 * i] = left[i];
        all[2 * i + 1] = right[i];
    }
    return all;
}

struct Termination {
    int64_t max_kmers = -1, max_radius = -1;   // EnvironmentFinderMain.java:171-183: MAX_KMERS added first, then MAX_RADIUS
    bool allows(const JHashMap<int>& dist, const std::string& kmer, int new_distance) const {
        if (dist.contains(kmer)) return false;
        if (max_kmers >= 0 && (int64_t)dist.size() >= max_kmers) return false;
        if (max_radius >= 0 && new_distance > max_radius) return false;
        return true;
    }
};

struct EnvCalc {
    const orc_table* reads;
    std::vector<std::string> sequences;   // one entry unless --merge
    int k, min_occ;
    bool both_dirs;
    int chunk_len;
    Termination term;
    JHashMap<int> subgraph;
    bool failed = false;
    int extend_additions = 0;
    std::vector<std::string> hic_sequences;   // --hicseq (only used with --merge, :180-191)
    bool trim = false;                        // --trim

    int occs(const std::string& kmer) const { return reads->get(kmer_key(to_long(kmer), k)); }   // :89-96

    // runTrimPaths (:241-262)
    void run_trim_paths(const std::unordered_set<std::string>& last_kmers, JHashMap<int>& distance, int dir) {
        std::vector<std::string> queue(last_kmers.begin(), last_kmers.end());
        std::unordered_set<std::string> visited(last_kmers.begin(), last_kmers.end());
        size_t head = 0;
        while (head < queue.size()) {
            std::string kmer = queue[head++];
            for (const std::string& nb : neighbors_by_dir(-dir, kmer))
                if (distance.contains(nb) && !visited.count(nb)) {
                    queue.push_back(nb);
                    visited.insert(nb);
                }
        }
        distance.retain_all(visited);
    }

    // runBfs (:154-220).  distanceToKmer is a java HashMap (iteration order matters for subgraph order).
    void run_bfs(int dir, std::vector<std::string>* order_out = nullptr, std::vector<int>* dist_out = nullptr) {
        std::vector<std::string> queue;
        JHashMap<int> distance;
        std::unordered_set<std::string> last_kmers;
        auto seed = [&](const std::vector<std::string>& seqs) {
            for (const std::string& sequence : seqs)
                for (size_t i = 0; i + k <= sequence.size(); i++) {
                    std::string kmer = sequence.substr(i, k);
                    if (occs(kmer) >= min_occ) {
                        queue.push_back(kmer);   // duplicates are queued again (:164,175,187)
                        distance.put(kmer, 0);
                    }
                }
        };
        seed(sequences);
        seed(hic_sequences);
        if (queue.empty()) { failed = true; return; }
        size_t head = 0;
        while (head < queue.size()) {
            std::string kmer = queue[head++];
            int d = *distance.get(kmer);
            for (const std::string& nb : neighbors_by_dir(dir, kmer)) {
                if (occs(nb) >= min_occ) {
                    if (term.allows(distance, nb, d + 1)) {
                        queue.push_back(nb);
                        distance.put(nb, d + 1);
                    } else {
                        last_kmers.insert(kmer);
                    }
                }
            }
        }
        if (trim) run_trim_paths(last_kmers, distance, dir);
        if (order_out) {   // discovery order = insertion order of the map
            for (auto& e : distance.entries) {
                order_out->push_back(e.first);
                dist_out->push_back(e.second);
            }
        }
        for (size_t pos : distance.order())   // for (String kmer : distanceToKmer.keySet()) addToSubgraph(kmer)
            add_to_subgraph(distance.entries[pos].first);
    }
    void add_to_subgraph(const std::string& kmer) { subgraph.put(normalize_dna(kmer), (int)occs(kmer)); }   // :146-148

    void build_environment() {   // :137-144
        if (both_dirs) run_bfs(0);
        else { run_bfs(-1); if (!failed) run_bfs(1); }
    }

    // extendEnvironment (:265-295): `cont = kmer` (:276) makes it a no-op on `subgraph`; only the log count remains.
    void extend_environment() {
        std::set<std::string> additions;
        for (size_t pos : subgraph.order()) {
            std::string kmer = subgraph.entries[pos].first;
            while (true) {
                std::string cont;
                bool has = false, multi = false;
                for (const std::string& nb : neighbors_by_dir(0, kmer))
                    if (!subgraph.contains(normalize_dna(nb)) && occs(nb) >= min_occ) {
                        if (!has) { has = true; cont = kmer; }
                        else multi = true;
                    }
                if (has && !multi && !additions.count(cont)) { additions.insert(cont); kmer = cont; }
                else break;
            }
        }
        extend_additions = (int)additions.size();
        for (const std::string& kmer : additions) add_to_subgraph(kmer);   // re-puts existing keys with the same value
    }
};

// =====================================================================================
// C4-C6: unitig graph + writers
// =====================================================================================
struct Node {
    std::string sequence;
    int id;
    bool gene, deleted = false;
    int rc;
    std::vector<int> neighbors;
};

struct Picture {
    int k;
    std::vector<Node> nodes;
    const JHashMap<int>* subgraph;
    std::vector<std::string> targets;

    bool is_gene(const std::string& seq, const std::string& rc) const {   // :421-432
        for (const std::string& t : targets)
            if (t.find(seq) != std::string::npos || t.find(rc) != std::string::npos) return true;
        return false;
    }
    // initializeStructures (:387-419)
    void initialize() {
        std::vector<size_t> ord = subgraph->order();
        nodes.resize(ord.size() * 2);
        int id = 0;
        for (size_t pos : ord) {
            const std::string& seq = subgraph->entries[pos].first;
            std::string rc = rc_string(seq);
            bool g = is_gene(seq, rc);
            nodes[id] = Node{seq, id, g, false, id + 1, {}};
            nodes[id + 1] = Node{rc, id + 1, g, false, id, {}};
            id += 2;
        }
        std::unordered_map<std::string, std::vector<int>> by_prefix;
        for (size_t i = 0; i < nodes.size(); i++) by_prefix[nodes[i].sequence.substr(0, k - 1)].push_back((int)i);
        for (size_t i = 0; i < nodes.size(); i++) {
            auto it = by_prefix.find(nodes[i].sequence.substr(1));
            if (it != by_prefix.end()) {
                auto& nb = nodes[nodes[i].rc].neighbors;
                nb.insert(nb.end(), it->second.begin(), it->second.end());
            }
        }
    }
    static std::string merge_labels(const std::string& a, const std::string& b, int k) {   // :453-462
        if (a.substr(a.size() - (k - 1)) != b.substr(0, k - 1)) fail("Labels should be merged, but can not: %s and %s", a.c_str(), b.c_str());
        return a + b.substr(k - 1);
    }
    void merge_nodes(int first_plus, int second_minus) {   // :312-324
        int first_minus = nodes[first_plus].rc, second_plus = nodes[second_minus].rc;
        std::string new_seq = merge_labels(nodes[second_plus].sequence, nodes[first_plus].sequence, k);
        std::string new_rc = merge_labels(nodes[first_minus].sequence, nodes[second_minus].sequence, k);
        nodes[second_plus].sequence = new_seq;
        nodes[first_minus].sequence = new_rc;
        nodes[second_plus].rc = first_minus;
        nodes[first_minus].rc = second_plus;
        nodes[first_plus].deleted = nodes[second_minus].deleted = true;
    }
    void do_merge() {   // :434-451
        while (true) {
            bool acted = false;
            for (size_t i = 0; i < nodes.size(); i++)
                if (!nodes[i].deleted && nodes[i].neighbors.size() == 1) {
                    int other = nodes[i].neighbors[0];
                    if (nodes[other].neighbors.size() != 1 || nodes[i].gene != nodes[other].gene) continue;
                    merge_nodes((int)i, other);
                    acted = true;
                }
            if (!acted) break;
        }
    }
    std::string node_id(const Node& a) const {   // getNodeId (:464-466; GFAWriter.java:84-86)
        return std::to_string(std::min(nodes[a.rc].id, a.id) + 1) + (a.gene ? "_start" : "");
    }
    std::string neighbor_ids(const Node& node) const {   // :375-385; TreeSet<Integer>.toString()
        std::set<int> r;
        for (int n : node.neighbors) r.insert(std::min(nodes[n].id, nodes[nodes[n].rc].id) + 1);
        for (int n : nodes[node.rc].neighbors) r.insert(std::min(nodes[n].id, nodes[nodes[n].rc].id) + 1);
        r.erase(std::min(node.id, nodes[node.rc].id) + 1);
        std::string s = "[";
        bool first = true;
        for (int v : r) {
            if (!first) s += ", ";
            s += std::to_string(v);
            first = false;
        }
        return s + "]";
    }
    void output_node_sequences(const std::string& prefix, int chunk_len) const {   // :354-373
        mkdirs(prefix);
        std::ofstream out(prefix + "/seqs.fasta");
        for (const Node& n : nodes)
            if (!n.deleted && n.id < nodes[n.rc].id && (int)n.sequence.size() >= chunk_len)
                out << "> Id" << node_id(n) << " Length:" << n.sequence.size() << " Neighbors:" << neighbor_ids(n) << "\n"
                    << n.sequence << "\n";
    }
    long subgraph_get(const std::string& kmer) const {
        const int* v = subgraph->get(normalize_dna(kmer));
        if (!v) fail("NullPointerException: k-mer %s of a node is not in the environment", kmer.c_str());
        return *v;
    }
    void write_gfa(const std::string& prefix) const {   // src/io/writers/GFAWriter.java:47-99
        mkdirs(prefix);
        std::ofstream out(prefix + "/graph.gfa");
        for (const Node& n : nodes)
            if (!n.deleted && n.sequence.compare(nodes[n.rc].sequence) <= 0) {
                long coverage = 0;
                for (size_t i = 0; i + k <= n.sequence.size(); i++) coverage += subgraph_get(n.sequence.substr(i, k));
                coverage += subgraph_get(n.sequence.substr(n.sequence.size() - k)) * (k - 1);
                out << "S\t" << node_id(n) << "\t" << n.sequence << "\tLN:i:" << n.sequence.size() << "\tKC:i:" << coverage
                    << (n.gene ? "\tCL:Z:GREEN" : "") << "\n";
            }
        for (const Node& i : nodes)
            if (!i.deleted)
                for (int jn : i.neighbors) {
                    const Node& j = nodes[jn];
                    if (j.deleted) continue;
                    out << "L\t" << node_id(i) << "\t" << (i.sequence.compare(nodes[i.rc].sequence) >= 0 ? "+" : "-") << "\t"
                        << node_id(j) << "\t" << (j.sequence.compare(nodes[j.rc].sequence) <= 0 ? "+" : "-") << "\t" << (k - 1)
                        << "M\n";
                }
    }
    std::string tsv_id(const Node& n) const {   // src/io/writers/TSVWriter.java:77-79
        bool fwd = n.sequence.compare(nodes[n.rc].sequence) <= 0;
        return (fwd ? std::to_string(n.id + 1) : "-" + std::to_string(nodes[n.rc].id + 1)) + (n.gene ? "_start" : "");
    }
    void write_tsv(const std::string& prefix) const {   // src/io/writers/TSVWriter.java:27-78
        mkdirs(prefix + "/tsvs");
        {
            std::ofstream out(prefix + "/tsvs/edges.tsv");
            out << "source\ttarget\n";
            for (const Node& i : nodes)
                if (!i.deleted)
                    for (int jn : i.neighbors)
                        if (!nodes[jn].deleted) out << tsv_id(nodes[i.rc]) << "\t" << tsv_id(nodes[jn]) << "\tpp\n";
        }
        {
            std::ofstream out(prefix + "/tsvs/nodes.tsv");
            out << "id\tlength\tseq\n";
            for (size_t i = 0; i < nodes.size(); i++)
                if (!nodes[i].deleted && nodes[i].sequence.compare(nodes[nodes[i].rc].sequence) <= 0)
                    out << (i + 1) << "\t" << nodes[i].sequence.size() << "\t" << nodes[i].sequence << "\n";
        }
    }
};

// printEnvironment (:297-310).  The reference writes graph.txt; README/BASELINE call it env.txt: both are written.
void print_environment(const JHashMap<int>& subgraph, const std::string& prefix) {
    mkdirs(prefix);
    std::string body;
    for (size_t pos : subgraph.order()) body += subgraph.entries[pos].first + " " + std::to_string(subgraph.entries[pos].second) + "\n";
    for (const char* name : {"/graph.txt", "/env.txt"}) {
        std::ofstream out(prefix + name);
        out << body;
    }
}

void create_picture(const JHashMap<int>& subgraph, const std::vector<std::string>& targets, int k, int chunk_len,
                    const std::string& prefix) {   // createPicture (:326-339)
    Picture p;
    p.k = k;
    p.subgraph = &subgraph;
    p.targets = targets;
    p.initialize();
    p.do_merge();
    p.output_node_sequences(prefix, chunk_len);
    p.write_gfa(prefix);
    p.write_tsv(prefix);
}

// =====================================================================================
// M1: environment-finder-multi
// =====================================================================================
struct MultiNode {
    std::string sequence;
    int id;
    bool gene, deleted = false;
    int rc;
    std::vector<int> neighbors;
    std::set<int> graphs;
};

// src/io/graph/DeBruijnGraphUtils.java:13-27: lines "KMER DEPTH" -> HashMap
JHashMap<int> load_graph(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail("Couldn't load graph from file %s", path.c_str());
    fclose(f);
    std::vector<uint8_t> b = read_file(path);
    LineReader lr(b.data(), b.size());
    JHashMap<int> g;
    std::string line;
    while (lr.ready()) {
        lr.read_line(line);
        size_t sp = line.find(' ');
        if (sp == std::string::npos) fail("ArrayIndexOutOfBoundsException: malformed environment line \"%s\"", line.c_str());
        std::string rest = line.substr(sp + 1);
        size_t sp2 = rest.find(' ');
        if (sp2 != std::string::npos) rest = rest.substr(0, sp2);
        char* endp = nullptr;
        long depth = strtol(rest.c_str(), &endp, 10);
        if (rest.empty() || *endp) fail("NumberFormatException: For input string: \"%s\"", rest.c_str());
        g.put(line.substr(0, sp), (int)depth);
    }
    return g;
}

// String.format("%6.2f ", floatValue): java.util.Formatter widens the float to double, takes the SHORTEST decimal
// digits that round-trip that double (FloatingDecimal) and rounds them HALF_UP to 2 decimals (not the C rule).
std::string java_format_6_2f(float v) {
    if (std::isnan(v)) return "   NaN ";
    if (std::isinf(v)) return v > 0 ? "Infinity " : "-Infinity ";
    double dv = (double)v;
    bool neg = std::signbit(dv);
    dv = std::fabs(dv);
    char sci[64];
    for (int prec = 0; prec <= 17; prec++) {
        snprintf(sci, sizeof sci, "%.*e", prec, dv);
        if (strtod(sci, nullptr) == dv) break;
    }
    // sci = d.ddddde[+-]XX -> digit string + decimal exponent
    std::string m(sci);
    size_t epos = m.find('e');
    int exp10 = atoi(m.c_str() + epos + 1);
    std::string digits;
    for (size_t i = 0; i < epos; i++)
        if (m[i] != '.') digits.push_back(m[i]);
    // value = 0.digits * 10^(exp10+1); build integer part and fraction part strings
    int point = exp10 + 1;   // number of digits before the decimal point
    std::string ip, fp;
    if (point <= 0) { ip = "0"; fp = std::string((size_t)(-point), '0') + digits; }
    else if ((size_t)point >= digits.size()) { ip = digits + std::string((size_t)point - digits.size(), '0'); fp = ""; }
    else { ip = digits.substr(0, (size_t)point); fp = digits.substr((size_t)point); }
    while (fp.size() < 3) fp.push_back('0');
    bool up = fp[2] >= '5';
    std::string all = ip + fp.substr(0, 2);
    if (up) {
        int i = (int)all.size() - 1;
        while (i >= 0) {
            if (all[i] == '9') { all[i] = '0'; i--; }
            else { all[i]++; break; }
        }
        if (i < 0) all = "1" + all;
    }
    std::string r = all.substr(0, all.size() - 2) + "." + all.substr(all.size() - 2);
    if (neg) r = "-" + r;
    while (r.size() < 6) r = " " + r;
    return r + " ";
}

void env_multi(const std::vector<std::string>& env_paths, const std::string& seq_fasta, int gene_id, const std::string& out_dir) {
    // loadInput (src/tools/EnvironmentFinderMultiMain.java:49-78)
    std::vector<JHashMap<int>> graphs;
    for (auto& p : env_paths) graphs.push_back(load_graph(p));
    if (graphs.empty()) fail("Zero environments given");
    if (graphs[0].size() == 0) fail("NoSuchElementException: first environment is empty");
    int k = (int)graphs[0].entries[graphs[0].order()[0]].first.size();
    for (auto& g : graphs)
        for (size_t pos : g.order())
            if ((int)g.entries[pos].first.size() != k)
                fail("K-mers of different lengths encountered: %d and %d", k, (int)g.entries[pos].first.size());
    RichFasta rf = rich_fasta_read(seq_fasta);
    if (gene_id < 1 || gene_id > (int)rf.dnas.size() || gene_id > (int)rf.comments.size())
        fail("IndexOutOfBoundsException: --geneid %d", gene_id);
    std::string sequence = rf.dnas[gene_id - 1], comment = rf.comments[gene_id - 1];

    // MultiSequenceCalculator.initializeStructures (src/algo/MultiSequenceCalculator.java:51-100)
    JHashMap<int> node_by_kmer;   // value = node index, -1 = null
    for (auto& g : graphs)
        for (size_t pos : g.order()) {
            const std::string& kmer = g.entries[pos].first;
            node_by_kmer.put(kmer, -1);
            node_by_kmer.put(rc_string(kmer), -1);
        }
    std::vector<MultiNode> nodes(node_by_kmer.size());
    {
        int ptr = 0;
        for (size_t pos : node_by_kmer.order()) {   // order fixed before the value updates below (puts of existing keys)
            std::string kmer = node_by_kmer.entries[pos].first;
            std::string rc = rc_string(kmer);
            if (kmer.compare(rc) > 0) continue;
            bool g = sequence.find(kmer) != std::string::npos || sequence.find(rc) != std::string::npos;
            nodes[ptr] = MultiNode{kmer, ptr, g, false, ptr + 1, {}, {}};
            nodes[ptr + 1] = MultiNode{rc, ptr + 1, g, false, ptr, {}, {}};
            node_by_kmer.put(kmer, ptr);
            node_by_kmer.put(rc, ptr + 1);
            ptr += 2;
        }
        if (ptr != (int)nodes.size()) fail("NullPointerException: palindromic k-mers (even k) leave null nodes in the reference");
    }
    for (size_t i = 0; i < graphs.size(); i++)
        for (size_t pos : graphs[i].order()) {
            int n = *node_by_kmer.get(graphs[i].entries[pos].first);
            nodes[n].graphs.insert((int)i);
            nodes[nodes[n].rc].graphs.insert((int)i);
        }
    for (size_t i = 0; i < nodes.size(); i++)
        for (int j = 0; j < 4; j++) {
            const int* nb = node_by_kmer.get(nodes[i].sequence.substr(1) + NUCLEOTIDES[j]);
            if (nb && *nb >= 0) nodes[nodes[i].rc].neighbors.push_back(*nb);
        }
    // doMerge / canBeMerged / mergeNodes (:102-137)
    while (true) {
        bool acted = false;
        for (size_t i = 0; i < nodes.size(); i++)
            if (!nodes[i].deleted && nodes[i].neighbors.size() == 1) {
                int o = nodes[i].neighbors[0];
                if (nodes[i].neighbors.size() == 1 && nodes[o].neighbors.size() == 1 && nodes[i].gene == nodes[o].gene &&
                    nodes[i].graphs == nodes[o].graphs) {
                    int first_plus = (int)i, second_minus = o;
                    int first_minus = nodes[first_plus].rc, second_plus = nodes[second_minus].rc;
                    std::string a = Picture::merge_labels(nodes[second_plus].sequence, nodes[first_plus].sequence, k);
                    std::string b = Picture::merge_labels(nodes[first_minus].sequence, nodes[second_minus].sequence, k);
                    nodes[second_plus].sequence = a;
                    nodes[first_minus].sequence = b;
                    nodes[second_plus].rc = first_minus;
                    nodes[first_minus].rc = second_plus;
                    nodes[first_plus].deleted = nodes[second_minus].deleted = true;
                    acted = true;
                }
            }
        if (!acted) break;
    }
    mkdirs(out_dir);
    auto min_id = [&](const MultiNode& n) { return std::min(n.id, nodes[n.rc].id) + 1; };
    {   // outputNodeSequences (:139-175)
        std::ofstream out(out_dir + "/seqs.fasta");
        for (const MultiNode& n : nodes)
            if (!n.deleted && n.id < nodes[n.rc].id) {
                std::set<int> r;
                for (int x : n.neighbors) r.insert(min_id(nodes[x]));
                for (int x : nodes[n.rc].neighbors) r.insert(min_id(nodes[x]));
                r.erase(min_id(n));
                std::string s = "[";
                bool first = true;
                for (int v : r) { if (!first) s += ", "; s += std::to_string(v); first = false; }
                s += "]";
                out << "> Id" << min_id(n) << (n.gene ? "_start" : "") << " Length:" << n.sequence.size() << " Neighbors:" << s << "\n"
                    << n.sequence << "\n";
            }
    }
    {   // GFAWriterMulti.printGraph (src/io/writers/GFAWriterMulti.java:39-133)
        std::ofstream out(out_dir + "/graph.gfa");
        size_t G = graphs.size();
        for (const MultiNode& n : nodes)
            if (!n.deleted && n.id < nodes[n.rc].id) {
                long coverage = 0;
                for (auto& g : graphs)
                    for (size_t i = 0; i + k <= n.sequence.size(); i++) {
                        const int* cov = g.get(normalize_dna(n.sequence.substr(i, k)));
                        coverage += cov ? *cov : 0;
                    }
                std::string color;
                size_t gs = n.graphs.size();
                if (G == 2) {
                    color = n.gene ? "#00ff00" : gs == 1 ? "#ff0000" : gs == 2 ? "#0000ff" : "#000000";
                } else if (G == 3) {
                    color = n.gene ? "#00ff00" : gs == 1 ? "#ff0000" : gs == 2 ? "#0000ff" : gs == 3 ? "#ff00ff" : "#000000";
                } else if (n.gene) {
                    color = "#00ff00";
                } else {
                    int value = (int)(256 * gs / G);
                    char b[32];
                    snprintf(b, sizeof b, "#%02X%02X%02X", value, value, value);
                    color = b;
                }
                out << "S\t" << (n.id < nodes[n.rc].id ? "" : "-") << min_id(n) << (n.gene ? "_start" : "") << "\t" << n.sequence
                    << "\tLN:i:" << n.sequence.size() << "\tKC:i:" << coverage << "\tCL:Z:" << color << "\tC2:Z:" << color << "\n";
            }
        for (const MultiNode& i : nodes)
            if (!i.deleted)
                for (int jn : i.neighbors) {   // no deleted-target check (GFAWriterMulti.java:51-57)
                    const MultiNode& j = nodes[jn];
                    out << "L\t" << min_id(i) << (i.gene ? "_start" : "") << "\t" << (i.id < nodes[i.rc].id ? "+" : "-") << "\t" << min_id(j)
                        << (j.gene ? "_start" : "") << "\t" << (j.id > nodes[j.rc].id ? "+" : "-") << "\t" << (k - 1) << "M\n";
                }
    }
    {   // printGene (EnvironmentFinderMultiMain.java:93-104)
        std::ofstream out(out_dir + "/gene.fasta");
        out << ">" << comment << "\n" << sequence << "\n";
    }
    {   // printProbability (:106-168); int32 accumulators, float division, "%6.2f "
        size_t G = graphs.size();
        std::vector<std::vector<int32_t>> diff(G, std::vector<int32_t>(G, 0)), diff_alt = diff, uni = diff;
        for (size_t i = 0; i < G; i++)
            for (size_t j = 0; j < G; j++) {
                for (auto& e : graphs[i].entries) {
                    const int* o = graphs[j].get(e.first);
                    if (!o) {
                        diff[i][j] += e.second; diff_alt[i][j] += e.second; uni[i][j] += e.second;
                    } else {
                        diff[i][j] += std::abs(e.second - *o); diff_alt[i][j] += std::abs(e.second - *o);
                        uni[i][j] += std::max(e.second, *o);
                    }
                }
                for (auto& e : graphs[j].entries)
                    if (!graphs[i].contains(e.first)) { diff[i][j] += e.second; uni[i][j] += e.second; }
            }
        std::ofstream sym(out_dir + "/Jacard_sym.txt"), alt(out_dir + "/Jacard_alt.txt");
        sym << "The[31mWarning! symmetric <<Jaccard distance>> (1 - AB/AUB):\n\n";
        alt << "The[31mWarning! alternative <<Jaccard distance>> (1 - AB/A):\n\n";
        for (size_t i = 0; i < G; i++) {
            sym << env_paths[i];
            alt << env_paths[i];
            for (size_t j = 0; j < G; j++) {
                int32_t inter = uni[i][j] - diff[i][j];
                sym << java_format_6_2f(1 - inter / (float)uni[i][j]);
                alt << java_format_6_2f(1 - inter / (float)(uni[i][j] - diff_alt[i][j]));
            }
            sym << "\n";
            alt << "\n";
        }
    }
}

// =====================================================================================
// f-4b: reads-classifier (src/tools/ReadsClassifier.java, src/algo/ReadsFinderInGraph.java, src/algo/PairFinder.java)
// =====================================================================================
// A read as LazyDnaQReaderTool yields it (ReadersUtils.readDnaQLazy -> FastqReader, NO truncation): value[i] = nuc | phred << 2
// (N, n, . = nuc 0 with phred 0); nucAt = value & 3, phredAt = (value >> 2) & 63 (ITMO!/dna/DnaQ.java:147-150).
struct ClassifyParams {
    int k;
    double z;                 // 1, or 1.96 with --interval95 (ReadsClassifier.java:171-172)
    double found_threshold;   // --found-threshold / 100 (:173)
    bool correction;          // --correction
};
// getCoverage (ReadsFinderInGraph.java:51-74), k <= 31: graph.getWithZero(min(fw, rc)) for every k-mer of the read
static void get_coverage(const orc_table* g, const std::vector<uint8_t>& value, int k, std::vector<int>& cov) {
    cov.clear();
    const size_t n = value.size();
    uint64_t mask = (1ULL << (2 * k)) - 1, fw = 0, rc = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t nuc = value[i] & 3;
        fw = ((fw << 2) | nuc) & mask;
        rc = (rc >> 2) | ((3ULL - nuc) << (2 * k - 2));
        if (i + 1 >= (size_t)k) {
            int16_t c = g->get(std::min(fw, rc));
            cov.push_back(c < 0 ? 0 : c);   // getWithZero: 0 for an absent key
        }
    }
}
// the test shared by findRead (:37-49, threshold = found_threshold) and findReadWithCorrection (:113-119, threshold 0.9);
// *margin (optional) = distance of the decisive comparison from its boundary (tests skip reads that sit on it)
static bool width_test(const std::vector<int>& cov, int k, size_t length, double z, double threshold, double* margin) {
    int sum = 0, nonzero = 0;
    for (int c : cov) { sum += c; nonzero += c > 0 ? 1 : 0; }
    const int last = cov.back();
    const double cov_mean = (double)(sum + last * (k - 1)) / (double)length;
    const double width = (double)(nonzero + (last > 0 ? 1 : 0) * (k - 1)) / (double)length;
    const double theory_width = 1.0 - std::exp(-cov_mean);                                    // getTheoryWidth (:76-78)
    const double sd = z * std::sqrt(std::exp(-cov_mean) * (1 - std::exp(-cov_mean)) / (double)length);   // delta (:81-87)
    const double d = width - theory_width;
    // only exp / sqrt differ between math libraries (division and comparisons are exact IEEE): the comparison with sd is the one
    // that can flip, and only when it decides the outcome
    if (margin) *margin = (width == 1 || width == 0 || width < threshold) ? 1.0 : std::fabs(std::fabs(d) - sd);
    const bool delta = width == 1 || (width != 0 && -sd <= d && d <= sd);
    return !(width < threshold) && delta;
}
static bool find_read(const orc_table* g, const std::vector<uint8_t>& value, const ClassifyParams& p, double* margin) {
    if (margin) *margin = 1.0;
    if (value.size() < (size_t)p.k) return false;
    std::vector<int> cov;
    get_coverage(g, value, p.k, cov);
    return width_test(cov, p.k, value.size(), p.z, p.found_threshold, margin);
}
// findReadWithCorrection (:89-123): exactly one base with phred < 10 -> try the four nucleotides there (threshold 0.9, as coded)
static bool find_read_with_correction(const orc_table* g, const std::vector<uint8_t>& value, const ClassifyParams& p, double* margin) {
    if (margin) *margin = 1.0;
    if (value.size() < (size_t)p.k) return false;
    std::vector<size_t> bad;
    for (size_t i = 0; i < value.size(); i++)
        if (((value[i] >> 2) & 63) < 10) bad.push_back(i);
    if (bad.size() != 1) return find_read(g, value, p, margin);
    std::vector<int> cov;
    double worst = 1.0;
    for (int nuc = 0; nuc < 4; nuc++) {
        std::vector<uint8_t> corrected = value;
        corrected[bad[0]] = (uint8_t)((corrected[bad[0]] & ~3u) | (unsigned)nuc);   // DnaQ.setNuc
        get_coverage(g, corrected, p.k, cov);
        double m = 1.0;
        const bool ok = width_test(cov, p.k, corrected.size(), p.z, 0.9, &m);
        worst = std::min(worst, m);
        if (ok) { if (margin) *margin = worst; return true; }
    }
    if (margin) *margin = worst;
    return false;
}
static bool classify_one(const orc_table* g, const std::vector<uint8_t>& value, const ClassifyParams& p, double* margin) {
    return p.correction ? find_read_with_correction(g, value, p, margin) : find_read(g, value, p, margin);
}
static std::vector<std::vector<uint8_t>> read_dnaqs(const uint8_t* buf, size_t len, int fmt) {
    if (fmt == 0) fail("reads-classifier restatement: FASTQ input only");
    const int off = fmt == 33 || fmt == 64 ? fmt : determine_quality_offset(buf, len);
    FastqIterator it(buf, len, off);
    std::vector<std::vector<uint8_t>> out;
    std::vector<uint8_t> v;
    while (it.read_next(v)) out.push_back(v);
    return out;
}
// WritersUtils.writeDnaQsToFastqFile(data, file) (ITMO!/io/writers/WritersUtils.java + FastqDedicatedWriter.writeData): comments
// 1, 2, 3, ... (DataCounter), bases by nucleotide (N was stored as A), qualities re-encoded Phred+64 (Illumina.instance)
static void write_dnaqs(const std::vector<const std::vector<uint8_t>*>& data, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) fail("cannot write %s", path.c_str());
    const char NUCS[4] = {'A', 'G', 'C', 'T'};
    long n = 0;
    std::string line;
    for (const std::vector<uint8_t>* d : data) {
        if (d->empty()) { fclose(f); fail("Empty DnaQ!"); }
        fprintf(f, "@%ld\n", ++n);
        line.assign(d->size(), 'A');
        for (size_t i = 0; i < d->size(); i++) line[i] = NUCS[(*d)[i] & 3];
        fprintf(f, "%s\n+\n", line.c_str());
        for (size_t i = 0; i < d->size(); i++) line[i] = (char)(64 + (((*d)[i] >> 2) & 63));
        fprintf(f, "%s\n", line.c_str());
    }
    fclose(f);
}

template <class F>
int guarded(F f) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_detect_format(const char* path) {
    return guarded([&] { return detect_format(path); });
}

orc_table* orc_table_new(int k, int threads) {
    if (k < 1 || k > 31) { g_err = "k must be in 1..31"; return nullptr; }
    return new orc_table(k, threads);
}
void orc_table_free(orc_table* t) { delete t; }

int orc_count_bytes(orc_table* t, const uint8_t* buf, size_t len, int fmt) {
    return guarded([&] { count_source(t, buf, len, fmt); return 0; });
}
int orc_count_file(orc_table* t, const char* path) {
    return guarded([&] {
        int gz = 0;
        int fmt = detect_format(path, &gz);
        std::vector<uint8_t> b = gz == 1 ? read_gz_file(path) : gz == 2 ? read_bz2_file(path) : read_file(path);
        count_source(t, b.data(), b.size(), fmt);
        return 0;
    });
}
int orc_table_add_pairs(orc_table* t, const uint64_t* keys, const int16_t* counts, size_t n) {
    for (size_t i = 0; i < n; i++) t->add(keys[i], counts[i]);
    return 0;
}
uint64_t orc_table_size(const orc_table* t) { return t->size(); }
int16_t orc_table_get(const orc_table* t, uint64_t key) { return t->get(key); }
void orc_table_get_many(const orc_table* t, const uint64_t* keys, size_t n, int16_t* out) {
    for (size_t i = 0; i < n; i++) out[i] = t->get(keys[i]);
}
uint64_t orc_table_dump(const orc_table* t, int min_count, uint64_t* keys, int16_t* counts, uint64_t cap, uint64_t* total) {
    std::vector<std::pair<uint64_t, int16_t>> v;
    t->for_each([&](uint64_t k, int16_t c) { if (c >= min_count) v.emplace_back(k, c); });
    std::sort(v.begin(), v.end());
    if (total) *total = v.size();
    uint64_t n = std::min<uint64_t>(cap, v.size());
    for (uint64_t i = 0; i < n; i++) { keys[i] = v[i].first; counts[i] = v[i].second; }
    return n;
}
// Order-independent digest of the entries with count >= min_count -- the formula of the library's mc_table_checksum
// (include/mc_b200.h): out = {entries, sum of counts, sum of digest(key) * count mod 2^64, largest count},
// digest(x) = (x ^ (x >> 29)) * 0xBF58476D1CE4E5B9 + 1.  Lets a test compare tables far too large to dump and sort.
void orc_table_checksum(const orc_table* t, int min_count, uint64_t out[4]) {
    uint64_t n = 0, sum = 0, dig = 0, mx = 0;
    t->for_each([&](uint64_t k, int16_t c) {
        if (c <= 0 || c < min_count) return;
        n++;
        sum += (uint64_t)c;
        dig += ((k ^ (k >> 29)) * 0xBF58476D1CE4E5B9ULL + 1ULL) * (uint64_t)c;
        mx = std::max<uint64_t>(mx, (uint64_t)c);
    });
    out[0] = n; out[1] = sum; out[2] = dig; out[3] = mx;
}
void orc_table_stats(const orc_table* t, uint64_t* reads, uint64_t* runs, uint64_t* bases, uint64_t* kmers) {
    if (reads) *reads = t->n_reads;
    if (runs) *runs = t->n_runs;
    if (bases) *bases = t->n_bases;
    if (kmers) *kmers = t->n_kmers;
}

uint64_t orc_revcomp(uint64_t kmer, int k) { return revcomp_bits(kmer, k); }
uint64_t orc_canonical(uint64_t kmer, int k) { return kmer_key(kmer, k); }
uint64_t orc_encode(const char* s, int k) {
    uint64_t r = 0;
    guarded([&] { r = to_long(std::string(s, (size_t)k)); return 0; });
    return r;
}
void orc_decode(uint64_t kmer, int k, char* out) {
    std::string s = kmer2string(kmer, k);
    memcpy(out, s.c_str(), s.size() + 1);
}
int32_t orc_java_string_hash(const char* s) { return java_string_hash(s); }
// iteration order of a java.util.HashMap<String,_> filled with the n distinct keys in their order: order[i] = index of the i-th key visited
void orc_java_hashmap_order(const char* const* keys, size_t n, uint64_t* order) {
    JHashMap<int> m;
    for (size_t i = 0; i < n; i++) m.put(keys[i], (int)i);
    const std::vector<size_t> o = m.order();
    for (size_t i = 0; i < o.size(); i++) order[i] = (uint64_t)m.entries[o[i]].second;
}

int64_t orc_bfs(const orc_table* t, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int dir,
                char* out_kmers, int32_t* out_dist, uint64_t cap) {
    int64_t result = -1;
    int rc = guarded([&] {
        EnvCalc c{t, {std::string(seq)}, k, min_cov, dir == 0, 1, {max_kmers, max_radius}, {}, false, 0, {}, false};
        std::vector<std::string> order;
        std::vector<int> dist;
        c.run_bfs(dir, &order, &dist);
        if (c.failed) { result = -1; return 0; }
        if (order.size() > cap) { result = -2; return 0; }
        for (size_t i = 0; i < order.size(); i++) {
            memcpy(out_kmers + i * (size_t)k, order[i].data(), (size_t)k);
            out_dist[i] = dist[i];
        }
        result = (int64_t)order.size();
        return 0;
    });
    return rc < 0 ? -3 : result;
}

int orc_env_find(const orc_table* t, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                 int chunk_len, const char* out_dir) {
    return guarded([&] {
        if (max_kmers < 0 && max_radius < 0) fail("At least one of --maxkmers and --maxradius parameters should be set");
        FILE* f = fopen(seq_fasta, "rb");
        if (!f) fail("Could not load sequences from %s", seq_fasta);
        fclose(f);
        RichFasta rf = rich_fasta_read(seq_fasta);
        int done = 0;
        for (size_t i = 0; i < rf.dnas.size(); i++) {
            if (i >= rf.comments.size()) fail("IndexOutOfBoundsException: comments.get(%zu)", i);
            std::string prefix = std::string(out_dir) + "/" + rf.comments[i] + "/";   // getOutputPrefix (:245-249)
            EnvCalc c{t, {rf.dnas[i]}, k, min_cov, both_dirs != 0, chunk_len, {max_kmers, max_radius}, {}, false, 0, {}, false};
            c.build_environment();
            if (c.failed) continue;   // "Could not find any k-mers of the target gene in the input, halting."
            c.extend_environment();
            print_environment(c.subgraph, prefix);
            create_picture(c.subgraph, c.sequences, k, chunk_len, prefix);
            done++;
        }
        return done;
    });
}

// environment-finder with --merge / --hicseq / --trim (EnvironmentFinderMain.java:126-156,218-233)
int orc_env_find_ex(const orc_table* t, const char* seq_fasta, const char* hic_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                    int both_dirs, int chunk_len, int merge, int trim, const char* out_dir) {
    return guarded([&] {
        if (max_kmers < 0 && max_radius < 0) fail("At least one of --maxkmers and --maxradius parameters should be set");
        FILE* f = fopen(seq_fasta, "rb");
        if (!f) fail("Could not load sequences from %s", seq_fasta);
        fclose(f);
        RichFasta rf = rich_fasta_read(seq_fasta);
        std::vector<std::string> comments = rf.comments, hic;
        if (hic_fasta && *hic_fasta) {
            FILE* g = fopen(hic_fasta, "rb");
            if (!g) fail("Could not load Hi-C sequences from %s", hic_fasta);
            fclose(g);
            RichFasta hf = rich_fasta_read(hic_fasta);
            hic = hf.dnas;
            comments = hf.comments;   // loadInput overwrites this.comments with the Hi-C file's (:146-150)
        }
        int done = 0;
        auto run_one = [&](EnvCalc& c, const std::string& prefix) {
            c.trim = trim != 0;
            c.build_environment();
            if (c.failed) return;
            c.extend_environment();
            print_environment(c.subgraph, prefix);
            create_picture(c.subgraph, c.sequences, k, chunk_len, prefix);
            done++;
        };
        if (!merge) {
            for (size_t i = 0; i < rf.dnas.size(); i++) {
                if (i >= comments.size()) fail("IndexOutOfBoundsException: comments.get(%zu)", i);
                EnvCalc c{t, {rf.dnas[i]}, k, min_cov, both_dirs != 0, chunk_len, {max_kmers, max_radius}, {}, false, 0, {}, false};
                run_one(c, std::string(out_dir) + "/" + comments[i] + "/");
            }
        } else {
            EnvCalc c{t, rf.dnas, k, min_cov, both_dirs != 0, chunk_len, {max_kmers, max_radius}, {}, false, 0, {}, false};
            c.hic_sequences = hic;
            run_one(c, std::string(out_dir) + "/merged/");
        }
        return done;
    });
}

int orc_env_from_graph_txt(const char* graph_txt, const char* target_seq, int k, int chunk_len, const char* out_dir) {
    return guarded([&] {
        JHashMap<int> loaded = load_graph(graph_txt);
        // keep FILE order as the iteration order: the committed graph.txt already is in java order, and the
        // emulated order() of a map built in that order reproduces it (stable by bucket) — asserted by the tests.
        create_picture(loaded, {std::string(target_seq)}, k, chunk_len, out_dir);
        print_environment(loaded, out_dir);
        return 0;
    });
}

int orc_env_multi(const char* const* env_paths, int n, const char* seq_fasta, int gene_id, const char* out_dir) {
    return guarded([&] {
        std::vector<std::string> p(env_paths, env_paths + n);
        env_multi(p, seq_fasta, gene_id, out_dir);
        return 0;
    });
}

// src/io/IOUtils.java:39-65 printKmers: DataOutputStream big-endian long + short for value > threshold;
// stat file: QuickQuantitativeStatistics.printToFile (frequency histogram, ascending).
int64_t orc_print_kmers(const orc_table* t, int threshold, const char* out_file, const char* stat_file) {
    int64_t good = 0;
    int rc = guarded([&] {
        FILE* f = fopen(out_file, "wb");
        if (!f) fail("cannot open %s", out_file);
        std::map<int, uint64_t> hist;
        std::vector<std::pair<uint64_t, int16_t>> v;
        t->for_each([&](uint64_t k, int16_t c) { v.emplace_back(k, c); });
        std::sort(v.begin(), v.end());   // the reference writes in map order; consumers treat the file as a set
        for (auto& e : v) {
            hist[e.second]++;
            if (e.second > threshold) {
                uint8_t rec[10];
                for (int i = 0; i < 8; i++) rec[i] = (uint8_t)(e.first >> (56 - 8 * i));
                rec[8] = (uint8_t)((uint16_t)e.second >> 8);
                rec[9] = (uint8_t)e.second;
                fwrite(rec, 1, 10, f);
                good++;
            }
        }
        fclose(f);
        if (stat_file) {
            std::ofstream st(stat_file);
            st << "# k-mer frequency\tnumber of such k-mers\n";
            for (auto& h : hist) st << h.first << "\t" << h.second << "\n";
        }
        return 0;
    });
    return rc < 0 ? -1 : good;
}

int orc_load_kmers(orc_table* t, const char* path, int freq_threshold) {
    return guarded([&] {
        std::vector<uint8_t> b = read_file(path);
        if (b.size() % 10) fail("Can't load k-mers file");
        for (size_t i = 0; i < b.size(); i += 10) {
            uint64_t key = 0;
            for (int j = 0; j < 8; j++) key = (key << 8) | b[i + j];
            int16_t freq = (int16_t)((b[i + 8] << 8) | b[i + 9]);
            if (freq > freq_threshold) t->add(key, freq);
        }
        return 0;
    });
}

double orc_bench_count(const uint8_t* buf, size_t len, int fmt, int k, int threads, uint64_t* bases, uint64_t* distinct) {
    double secs = -1;
    guarded([&] {
        orc_table t(k, threads);
        auto t0 = std::chrono::steady_clock::now();
        count_source(&t, buf, len, fmt);
        auto t1 = std::chrono::steady_clock::now();
        secs = std::chrono::duration<double>(t1 - t0).count();
        if (bases) *bases = t.n_bases;
        if (distinct) *distinct = t.size();
        return 0;
    });
    return secs;
}

// f-4b: found[i] = findRead / findReadWithCorrection of read i of a FASTQ buffer; margins[i] (optional) = how far the decisive
// floating-point comparison was from its boundary.  Returns the number of reads, < 0 on error.
long orc_classify_bytes(const orc_table* g, const uint8_t* buf, size_t len, int fmt, double z, double found_threshold, int correction,
                        uint8_t* found, double* margins, size_t cap) {
    long n = -1;
    guarded([&] {
        const ClassifyParams p{g->k, z, found_threshold, correction != 0};
        std::vector<std::vector<uint8_t>> reads = read_dnaqs(buf, len, fmt);
        if (reads.size() > cap) fail("room for %zu reads, the input holds %zu", cap, reads.size());
        for (size_t i = 0; i < reads.size(); i++) {
            double m = 1.0;
            found[i] = classify_one(g, reads[i], p, &m) ? 1 : 0;
            if (margins) margins[i] = m;
        }
        n = (long)reads.size();
        return 0;
    });
    return n;
}
// The tool's run (ReadsClassifier.runImpl :141-224 + PairFinder.run :30-55): one or two read files, the four classes, six
// output files.  The reference appends to its class queues from pool threads, i.e. in no fixed order; here every file lists
// its reads in INPUT order.  counts[4] = {both found, first only, second only, none}.
int orc_reads_classifier(const orc_table* g, const char* reads1, const char* reads2, double z, double found_threshold, int correction,
                         const char* out_dir, long counts[4]) {
    return guarded([&] {
        const ClassifyParams p{g->k, z, found_threshold, correction != 0};
        std::vector<uint8_t> b1 = read_file(reads1), b2;
        std::vector<std::vector<uint8_t>> r1 = read_dnaqs(b1.data(), b1.size(), 1), r2;
        const bool paired = reads2 && *reads2;
        if (paired) {
            b2 = read_file(reads2);
            r2 = read_dnaqs(b2.data(), b2.size(), 1);
        }
        const std::vector<uint8_t> empty;
        const size_t n = paired ? std::min(r1.size(), r2.size()) : r1.size();   // PairSource stops at the shorter file
        std::vector<const std::vector<uint8_t>*> found_1, found_2, not_found_1, not_found_2, found_s, not_found_s, tail_found, tail_not;
        for (int i = 0; i < 4; i++) counts[i] = 0;
        for (size_t i = 0; i < n; i++) {
            const std::vector<uint8_t>& a = r1[i];
            const std::vector<uint8_t>& b = paired ? r2[i] : empty;
            bool f1 = classify_one(g, a, p, nullptr), f2 = classify_one(g, b, p, nullptr);
            if (b.empty()) f2 = !f1;                                              // PairFinder.java:42-44
            if (f1 && f2) { counts[0]++; found_1.push_back(&a); found_2.push_back(&b); }
            else if (f1) { counts[1]++; if (!a.empty()) found_s.push_back(&a); if (!b.empty()) not_found_s.push_back(&b); }
            else if (f2) { counts[2]++; if (!b.empty()) tail_found.push_back(&b); if (!a.empty()) tail_not.push_back(&a); }
            else { counts[3]++; not_found_1.push_back(&a); not_found_2.push_back(&b); }
        }
        // found_s = first_found.first ++ second_found.second; not_found_s = first_found.second ++ second_found.first (:212-217)
        found_s.insert(found_s.end(), tail_found.begin(), tail_found.end());
        not_found_s.insert(not_found_s.end(), tail_not.begin(), tail_not.end());
        mkdirs(out_dir);
        const std::string o(out_dir);
        write_dnaqs(found_1, o + "/found_1.fastq");
        write_dnaqs(found_2, o + "/found_2.fastq");
        write_dnaqs(not_found_1, o + "/not_found_1.fastq");
        write_dnaqs(not_found_2, o + "/not_found_2.fastq");
        write_dnaqs(found_s, o + "/found_s.fastq");
        write_dnaqs(not_found_s, o + "/not_found_s.fastq");
        return 0;
    });
}

// "Extending endings by N kmers" (OneSequenceCalculator.java:291) for every target of a FASTA, in order (-1: no seed k-mer)
int orc_env_extend_counts(const orc_table* t, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                          long long* out, int cap) {
    int n = -1;
    guarded([&] {
        RichFasta rf = rich_fasta_read(seq_fasta);
        n = (int)rf.dnas.size();
        for (size_t i = 0; i < rf.dnas.size() && (int)i < cap; i++) {
            EnvCalc c{t, {rf.dnas[i]}, k, min_cov, both_dirs != 0, 1, {max_kmers, max_radius}, {}, false, 0, {}, false};
            c.build_environment();
            if (c.failed) { out[i] = -1; continue; }
            c.extend_environment();
            out[i] = c.extend_additions;
        }
        return 0;
    });
    return n;
}

// test surface: the Jaccard cell text (tests/test_host_logic.py compares three independent implementations)
void orc_java_format_6_2f(float v, char* out) {
    const std::string s = java_format_6_2f(v);
    memcpy(out, s.c_str(), s.size() + 1);
}

// CPU baseline of the environment stage (bench.py only): every target as its own OneSequenceCalculator on a pool of
// `threads` workers, the way EnvironmentFinderMain.runImpl submits them (src/tools/EnvironmentFinderMain.java:188,218-225);
// the table is shared read-only.  Returns wall seconds (BFS + unitigs + every output file), < 0 on error; *done = targets
// that had a seed.
double orc_bench_env(const orc_table* t, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                     int chunk_len, int threads, const char* out_dir, int* done_out) {
    double secs = -1;
    guarded([&] {
        if (max_kmers < 0 && max_radius < 0) fail("At least one of --maxkmers and --maxradius parameters should be set");
        RichFasta rf = rich_fasta_read(seq_fasta);
        if (rf.comments.size() < rf.dnas.size()) fail("IndexOutOfBoundsException: comments.get(%zu)", rf.comments.size());
        std::atomic<size_t> next{0};
        std::atomic<int> done{0};
        auto t0 = std::chrono::steady_clock::now();
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < rf.dnas.size();) {
                std::string prefix = std::string(out_dir) + "/" + rf.comments[i] + "/";
                EnvCalc c{t, {rf.dnas[i]}, k, min_cov, both_dirs != 0, chunk_len, {max_kmers, max_radius}, {}, false, 0, {}, false};
                c.build_environment();
                if (c.failed) continue;
                c.extend_environment();
                print_environment(c.subgraph, prefix);
                create_picture(c.subgraph, c.sequences, k, chunk_len, prefix);
                done++;
            }
        };
        std::vector<std::thread> pool;
        for (int i = 1; i < std::max(1, threads); i++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (done_out) *done_out = done.load();
        return 0;
    });
    return secs;
}

}  // extern "C"
