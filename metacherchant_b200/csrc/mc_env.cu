// Seam 3 / Seam 4 host side: target loading, environment assembly in the reference's iteration order, the merge
// sweep over the device-computed neighbour lists, and the writers (graph.txt/env.txt, seqs.fasta, graph.gfa, tsvs/).
//
// What runs where:
//   device  BFS passes for all targets at once (mc_bfs.cu); node hash, neighbour lists, degrees, gene flags and
//           chain ranks of every environment (mc_unitig.cu)
//   host    string formatting, java.util.HashMap iteration order (it fixes node ids and line order of every output),
//           the O(n) merge sweep over the device's neighbour lists, file writing
#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <atomic>
#include <exception>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <cstdlib>
#include <fstream>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "mc_env.hpp"

namespace {

// MC_TRACE=1: accumulated host/device time of every environment stage on stderr (profiles/)
struct EnvTrace {
    static bool on() { static const bool v = getenv("MC_TRACE") != nullptr; return v; }
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6; }
    static std::vector<std::pair<const char*, double>>& acc() { static std::vector<std::pair<const char*, double>> a; return a; }
    const char* name;
    double t0;
    explicit EnvTrace(const char* n) : name(n), t0(on() ? now() : 0) {}
    static std::mutex& mu() { static std::mutex m; return m; }
    ~EnvTrace() {
        if (!on()) return;
        const double dt = now() - t0;   // summed over the worker threads for the per-environment stages
        std::lock_guard<std::mutex> lk(mu());
        for (auto& e : acc()) if (e.first == name) { e.second += dt; return; }
        acc().emplace_back(name, dt);
    }
    static void report() {
        if (!on()) return;
        for (auto& e : acc()) fprintf(stderr, "[mc trace] %-28s %10.1f ms\n", e.first, e.second);
        acc().clear();
    }
};

// The reference runs one OneSequenceCalculator per target on a pool of --available-processors threads
// (EnvironmentFinderMain.java:188,218-225); the host half of an environment (HashMap order, merge sweep, writers) does the
// same here.  MC_THREADS overrides the default (all host cores).
unsigned env_threads(size_t n_items) {
    unsigned n = std::thread::hardware_concurrency();
    if (const char* e = getenv("MC_THREADS")) n = (unsigned)std::max(1, atoi(e));
    if (n == 0) n = 1;
    return (unsigned)std::min<size_t>(n, std::max<size_t>(n_items, 1));
}
template <class F>
void parallel_for(size_t n, F&& body) {
    const unsigned nt = env_threads(n);
    if (nt <= 1 || n <= 1) {
        for (size_t i = 0; i < n; i++) body(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex err_mu;
    auto worker = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) return;
            try {
                body(i);
            } catch (...) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (!err) err = std::current_exception();
                next.store(n);
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < nt; w++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    if (err) std::rethrow_exception(err);
}

const char NUC[4] = {'A', 'G', 'C', 'T'};   // ITMO!/dna/DnaTools.java:31

std::string kmer_string(uint64_t code, int k) {   // KmerUtils.kmer2String
    std::string s((size_t)k, 'A');
    for (int i = 0; i < k; i++) s[(size_t)i] = NUC[(code >> (2 * (k - 1 - i))) & 3];
    return s;
}
uint64_t kmer_code(const std::string& s) {
    uint64_t r = 0;
    for (char c : s) r = (r << 2) | (uint64_t)(mc_nuc_code((uint8_t)c) & 3);
    return r;
}
std::string rc_string(const std::string& s) {
    std::string r(s.size(), 'A');
    for (size_t i = 0; i < s.size(); i++) {
        int c = mc_nuc_code((uint8_t)s[i]);
        if (c >= 4) mc_throw(MC_ERR_FORMAT, "Incorrect nucleotide char: \"%c\"", s[i]);
        r[s.size() - 1 - i] = NUC[c ^ 3];
    }
    return r;
}
int32_t java_string_hash(const std::string& s) {
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return (int32_t)h;
}

// ------------------------------------------------------------------------------------------------
// java.util.HashMap iteration order (JDK 8+, default constructor), insertions of DISTINCT keys in index order.
//
// Bins that never hold 9 nodes are plain lists: the order is "bucket index at the final capacity, insertion order inside a
// bucket" (resize splits keep the relative order) -- a stable counting sort.  A bin that reaches 9 nodes while the table has >= 64
// buckets becomes a red-black tree of TreeNodes (HashMap.treeifyBin): iteration still follows the `next` links, but treeify moves
// the tree's root to the front of the bin, putTreeVal links a new node right after its tree PARENT, and a resize splits the bin
// into two lists that are re-treeified (root to the front again) or, at <= 6 nodes, turned back into plain nodes in their current
// order.  With fewer than 64 buckets the table is resized instead.  Either way the final order depends on the history, so such a
// map is replayed put by put (JavaHashReplay).  `needs_replay` finds out in one counting pass per capacity stage (2n steps).
// h[i] = the key's SPREAD hash (hashCode ^ hashCode >>> 16); less(a, b) = key a < key b by String.compareTo.
// ------------------------------------------------------------------------------------------------
static bool java_hashmap_needs_replay(const std::vector<uint32_t>& h) {
    std::vector<uint16_t> cnt;
    size_t cap = 16, lo = 0;
    while (lo < h.size()) {
        const size_t hi = std::min(h.size(), cap / 4 * 3 + 1);   // the put that exceeds the threshold still lands in this table
        cnt.assign(cap, 0);
        for (size_t i = 0; i < hi; i++) {
            uint16_t& c = cnt[h[i] & (cap - 1)];
            if (c < 0xFFFF) c++;
            if (i >= lo && c >= 9) return true;
        }
        lo = hi;
        cap <<= 1;
    }
    return false;
}

template <class Less>
struct JavaHashReplay {
    struct Node {
        int32_t next = -1, prev = -1, left = -1, right = -1, parent = -1;
        bool red = false, tree = false;
    };
    const std::vector<uint32_t>& h;
    Less less;
    std::vector<Node> nd;
    std::vector<int32_t> tab;
    size_t threshold = 0, size = 0;
    JavaHashReplay(const std::vector<uint32_t>& hashes, Less l) : h(hashes), less(l), nd(hashes.size()) {}

    // TreeNode ordering: int compare of the hashes, then compareTo (keys are distinct: never 0)
    int dir(int32_t x, int32_t p) const {
        const int32_t hx = (int32_t)h[(size_t)x], hp = (int32_t)h[(size_t)p];
        if (hp > hx) return -1;
        if (hp < hx) return 1;
        return less((uint32_t)x, (uint32_t)p) ? -1 : 1;
    }
    int32_t rotate_left(int32_t root, int32_t p) {
        if (p < 0) return root;
        const int32_t r = nd[p].right;
        if (r < 0) return root;
        const int32_t rl = nd[p].right = nd[r].left;
        if (rl >= 0) nd[rl].parent = p;
        const int32_t pp = nd[r].parent = nd[p].parent;
        if (pp < 0) { root = r; nd[r].red = false; }
        else if (nd[pp].left == p) nd[pp].left = r;
        else nd[pp].right = r;
        nd[r].left = p;
        nd[p].parent = r;
        return root;
    }
    int32_t rotate_right(int32_t root, int32_t p) {
        if (p < 0) return root;
        const int32_t l = nd[p].left;
        if (l < 0) return root;
        const int32_t lr = nd[p].left = nd[l].right;
        if (lr >= 0) nd[lr].parent = p;
        const int32_t pp = nd[l].parent = nd[p].parent;
        if (pp < 0) { root = l; nd[l].red = false; }
        else if (nd[pp].right == p) nd[pp].right = l;
        else nd[pp].left = l;
        nd[l].right = p;
        nd[p].parent = l;
        return root;
    }
    int32_t balance_insertion(int32_t root, int32_t x) {
        nd[x].red = true;
        for (;;) {
            int32_t xp = nd[x].parent;
            if (xp < 0) { nd[x].red = false; return x; }
            int32_t xpp = nd[xp].parent;
            if (!nd[xp].red || xpp < 0) return root;
            const int32_t xppl = nd[xpp].left;
            if (xp == xppl) {
                const int32_t xppr = nd[xpp].right;
                if (xppr >= 0 && nd[xppr].red) {
                    nd[xppr].red = false; nd[xp].red = false; nd[xpp].red = true;
                    x = xpp;
                } else {
                    if (x == nd[xp].right) {
                        root = rotate_left(root, x = xp);
                        xp = nd[x].parent;
                        xpp = xp < 0 ? -1 : nd[xp].parent;
                    }
                    if (xp >= 0) {
                        nd[xp].red = false;
                        if (xpp >= 0) { nd[xpp].red = true; root = rotate_right(root, xpp); }
                    }
                }
            } else {
                if (xppl >= 0 && nd[xppl].red) {
                    nd[xppl].red = false; nd[xp].red = false; nd[xpp].red = true;
                    x = xpp;
                } else {
                    if (x == nd[xp].left) {
                        root = rotate_right(root, x = xp);
                        xp = nd[x].parent;
                        xpp = xp < 0 ? -1 : nd[xp].parent;
                    }
                    if (xp >= 0) {
                        nd[xp].red = false;
                        if (xpp >= 0) { nd[xpp].red = true; root = rotate_left(root, xpp); }
                    }
                }
            }
        }
    }
    void move_root_to_front(int32_t root) {
        if (root < 0 || tab.empty()) return;
        const size_t idx = h[(size_t)root] & (tab.size() - 1);
        const int32_t first = tab[idx];
        if (root == first) return;
        tab[idx] = root;
        const int32_t rp = nd[root].prev, rn = nd[root].next;
        if (rn >= 0) nd[rn].prev = rp;
        if (rp >= 0) nd[rp].next = rn;
        if (first >= 0) nd[first].prev = root;
        nd[root].next = first;
        nd[root].prev = -1;
    }
    void treeify(int32_t head) {   // TreeNode.treeify: the bin's nodes, in list order, into a fresh tree; root to the front
        int32_t root = -1;
        for (int32_t x = head, nx; x >= 0; x = nx) {
            nx = nd[x].next;
            nd[x].left = nd[x].right = -1;
            if (root < 0) {
                nd[x].parent = -1;
                nd[x].red = false;
                root = x;
                continue;
            }
            for (int32_t p = root;;) {
                const int d = dir(x, p);
                const int32_t xp = p;
                p = d <= 0 ? nd[p].left : nd[p].right;
                if (p < 0) {
                    nd[x].parent = xp;
                    (d <= 0 ? nd[xp].left : nd[xp].right) = x;
                    root = balance_insertion(root, x);
                    break;
                }
            }
        }
        move_root_to_front(root);
    }
    void split(int32_t b, size_t index, size_t bit) {   // TreeNode.split during a resize
        int32_t head[2] = {-1, -1}, tail[2] = {-1, -1};
        size_t count[2] = {0, 0};
        for (int32_t e = b, nx; e >= 0; e = nx) {
            nx = nd[e].next;
            nd[e].next = -1;
            const int s = (h[(size_t)e] & bit) ? 1 : 0;
            nd[e].prev = tail[s];
            if (tail[s] < 0) head[s] = e;
            else nd[tail[s]].next = e;
            tail[s] = e;
            count[s]++;
        }
        for (int s = 0; s < 2; s++) {
            const size_t at = index + (s ? bit : 0);
            tab[at] = head[s];
            if (head[s] < 0) continue;
            if (count[s] <= 6) {   // UNTREEIFY_THRESHOLD: plain nodes again, in the list's current order
                for (int32_t x = head[s]; x >= 0; x = nd[x].next) {
                    nd[x].tree = false;
                    nd[x].left = nd[x].right = nd[x].parent = -1;
                }
            } else if (head[1 - s] >= 0) {
                treeify(head[s]);   // (nothing left the bin otherwise: its tree is intact)
            }
        }
    }
    void resize() {
        if (tab.empty()) {
            tab.assign(16, -1);
            threshold = 12;
            return;
        }
        const size_t oc = tab.size();
        std::vector<int32_t> old((size_t)oc * 2, -1);
        old.swap(tab);   // tab = the new table, old = the previous one
        threshold *= 2;
        for (size_t j = 0; j < oc; j++) {
            int32_t e = old[j];
            if (e < 0) continue;
            if (nd[e].tree) { split(e, j, oc); continue; }
            int32_t head[2] = {-1, -1}, tail[2] = {-1, -1};
            for (int32_t nx; e >= 0; e = nx) {
                nx = nd[e].next;
                nd[e].next = -1;
                const int s = (h[(size_t)e] & oc) ? 1 : 0;
                if (tail[s] < 0) head[s] = e;
                else nd[tail[s]].next = e;
                tail[s] = e;
            }
            tab[j] = head[0];
            tab[j + oc] = head[1];
        }
    }
    void put(int32_t n) {
        if (tab.empty()) resize();
        const size_t i = h[(size_t)n] & (tab.size() - 1);
        int32_t p = tab[i];
        if (p < 0) {
            tab[i] = n;
        } else if (nd[p].tree) {   // putTreeVal: a leaf of the tree, linked right after its parent
            int32_t root = p;
            while (nd[root].parent >= 0) root = nd[root].parent;
            for (int32_t q = root;;) {
                const int d = dir(n, q);
                const int32_t xp = q;
                q = d <= 0 ? nd[q].left : nd[q].right;
                if (q < 0) {
                    const int32_t xpn = nd[xp].next;
                    nd[n].tree = true;
                    nd[n].next = xpn;
                    (d <= 0 ? nd[xp].left : nd[xp].right) = n;
                    nd[xp].next = n;
                    nd[n].parent = nd[n].prev = xp;
                    if (xpn >= 0) nd[xpn].prev = n;
                    move_root_to_front(balance_insertion(root, n));
                    break;
                }
            }
        } else {
            int walked = 0;
            while (nd[p].next >= 0) { p = nd[p].next; walked++; }
            nd[p].next = n;
            if (walked >= 7) {   // TREEIFY_THRESHOLD - 1: the bin now holds 9 nodes
                if (tab.size() < 64) {
                    resize();    // MIN_TREEIFY_CAPACITY
                } else {
                    int32_t prev = -1;
                    for (int32_t x = tab[i]; x >= 0; x = nd[x].next) { nd[x].tree = true; nd[x].prev = prev; prev = x; }
                    treeify(tab[i]);
                }
            }
        }
        if (++size > threshold) resize();
    }
    std::vector<uint32_t> run() {
        for (size_t i = 0; i < h.size(); i++) put((int32_t)i);
        std::vector<uint32_t> o;
        o.reserve(h.size());
        for (int32_t e : tab)
            for (; e >= 0; e = nd[e].next) o.push_back((uint32_t)e);
        return o;
    }
};

// positions 0 .. n-1 in iteration order.  peak = the largest size the map ever had (removals never shrink the table).
template <class Less>
static std::vector<uint32_t> java_hashmap_order(const std::vector<uint32_t>& h, size_t peak, Less less) {
    const size_t n = h.size();
    if (n >= 9 && java_hashmap_needs_replay(h)) {
        if (peak <= n) return JavaHashReplay<Less>(h, less).run();
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "WARN: a java.util.HashMap bin of this environment was treeified and keys were removed afterwards (--trim): the order inside "
                            "that bin is not modelled and may differ from the JVM's\n");
        }
    }
    size_t cap = 16;
    while ((double)std::max(peak, n) > 0.75 * (double)cap) cap <<= 1;
    std::vector<uint32_t> start(cap + 1, 0), o(n);
    for (size_t i = 0; i < n; i++) start[(h[i] & (uint32_t)(cap - 1)) + 1]++;
    for (size_t b = 0; b < cap; b++) start[b + 1] += start[b];
    for (size_t i = 0; i < n; i++) o[start[h[i] & (uint32_t)(cap - 1)]++] = (uint32_t)i;
    return o;
}

// A java.util.HashMap<String,V> by its iteration order (java_hashmap_order above).  Pinned on the reference's committed graph.txt by the tests.
template <class V>
struct JavaMap {
    std::unordered_map<std::string, uint32_t> index;
    std::vector<std::pair<std::string, V>> entries;
    size_t peak = 0;   // largest size ever: a java HashMap never shrinks its table when keys are removed
    void put(const std::string& k, const V& v) {
        auto it = index.find(k);
        if (it == index.end()) {
            index.emplace(k, (uint32_t)entries.size());
            entries.emplace_back(k, v);
            peak = std::max(peak, entries.size());
        } else {
            entries[it->second].second = v;
        }
    }
    const V* get(const std::string& k) const {
        auto it = index.find(k);
        return it == index.end() ? nullptr : &entries[it->second].second;
    }
    size_t size() const { return entries.size(); }
    // keySet().retainAll(...): the survivors keep their relative order
    void retain(const std::vector<uint8_t>& keep) {
        std::vector<std::pair<std::string, V>> kept;
        for (size_t i = 0; i < entries.size(); i++)
            if (keep[i]) kept.push_back(std::move(entries[i]));
        entries.swap(kept);
        index.clear();
        for (size_t i = 0; i < entries.size(); i++) index.emplace(entries[i].first, (uint32_t)i);
    }
    std::vector<uint32_t> order() const {
        std::vector<uint32_t> h(entries.size());
        for (size_t i = 0; i < entries.size(); i++) {
            const uint32_t x = (uint32_t)java_string_hash(entries[i].first);
            h[i] = x ^ (x >> 16);
        }
        return java_hashmap_order(h, peak, [this](uint32_t a, uint32_t b) { return entries[a].first < entries[b].first; });
    }
};

// The same java.util.HashMap<String,Integer>, for maps whose keys are all k-mers of one length: the key is held as its 2-bit
// code (A0 G1 C2 T3, first base most significant -- one string <-> one code), String.hashCode is computed from the code, and
// entries live in flat arrays, so an environment costs no per-k-mer allocation.  Same iteration order as JavaMap (tests pin
// both on the reference's committed graph.txt).
struct KmerMap {
    int k = 0;
    std::vector<uint64_t> keys;          // insertion order
    std::vector<int32_t> vals;
    std::vector<uint32_t> slots;         // open addressing over `keys`: index + 1, 0 = empty
    size_t peak = 0;
    uint32_t hash_tab[32][4];            // ascii(NUC[c]) * 31^(k-1-i) mod 2^32

    void init(int k_, size_t expect) {
        k = k_;
        keys.clear();
        vals.clear();
        peak = 0;
        keys.reserve(expect);
        vals.reserve(expect);
        size_t cap = 64;
        while (cap < expect * 2 + 2) cap <<= 1;
        slots.assign(cap, 0);
        uint32_t pw = 1;
        for (int i = k - 1; i >= 0; i--) {
            for (int c = 0; c < 4; c++) hash_tab[i][c] = (uint32_t)(unsigned char)"AGCT"[c] * pw;
            pw *= 31u;
        }
    }
    static uint64_t mix(uint64_t x) {
        x ^= x >> 33;
        x *= 0xff51afd7ed558ccdULL;
        x ^= x >> 29;
        return x;
    }
    void grow() {
        std::vector<uint32_t> ns(slots.size() * 2, 0);
        const size_t m = ns.size() - 1;
        for (size_t i = 0; i < keys.size(); i++) {
            size_t p = mix(keys[i]) & m;
            while (ns[p]) p = (p + 1) & m;
            ns[p] = (uint32_t)i + 1;
        }
        slots.swap(ns);
    }
    // index of `code`, or -1
    long find(uint64_t code) const {
        const size_t m = slots.size() - 1;
        for (size_t p = mix(code) & m;; p = (p + 1) & m) {
            const uint32_t v = slots[p];
            if (!v) return -1;
            if (keys[v - 1] == code) return (long)v - 1;
        }
    }
    void put(uint64_t code, int32_t v) {
        const size_t m = slots.size() - 1;
        size_t p = mix(code) & m;
        for (;; p = (p + 1) & m) {
            const uint32_t s = slots[p];
            if (!s) break;
            if (keys[s - 1] == code) { vals[s - 1] = v; return; }
        }
        keys.push_back(code);
        vals.push_back(v);
        slots[p] = (uint32_t)keys.size();
        peak = std::max(peak, keys.size());
        if (keys.size() * 2 > slots.size()) grow();
    }
    const int32_t* get(uint64_t code) const {
        long i = find(code);
        return i < 0 ? nullptr : &vals[(size_t)i];
    }
    void retain(const std::vector<uint8_t>& keep) {   // keySet().retainAll(...): the survivors keep their relative order
        size_t n = 0;
        for (size_t i = 0; i < keys.size(); i++)
            if (keep[i]) { keys[n] = keys[i]; vals[n] = vals[i]; n++; }
        keys.resize(n);
        vals.resize(n);
        std::fill(slots.begin(), slots.end(), 0u);
        const size_t m = slots.size() - 1;
        for (size_t i = 0; i < n; i++) {
            size_t p = mix(keys[i]) & m;
            while (slots[p]) p = (p + 1) & m;
            slots[p] = (uint32_t)i + 1;
        }
    }
    uint32_t java_hash(uint64_t code) const {         // String.hashCode() of the k-mer string
        uint32_t h = 0;
        for (int i = k - 1; i >= 0; i--, code >>= 2) h += hash_tab[i][code & 3];
        return h;
    }
    std::vector<uint32_t> order() const {             // java_hashmap_order: a stable counting sort by bucket unless a bin was treeified
        std::vector<uint32_t> h(keys.size());
        for (size_t i = 0; i < keys.size(); i++) {
            const uint32_t x = java_hash(keys[i]);
            h[i] = x ^ (x >> 16);
        }
        // String.compareTo on k-mers of one length = numeric order of the codes with G (1) and C (2) exchanged (ASCII A < C < G < T)
        auto rank = [](uint64_t x) { return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL); };
        return java_hashmap_order(h, peak, [this, rank](uint32_t a, uint32_t b) { return rank(keys[a]) < rank(keys[b]); });
    }
};

// StringUtils.normalizeDna (src/utils/StringUtils.java:34-41) on codes: the ASCII order A < C < G < T is the code order with G (1) and C (2) exchanged
inline uint64_t ascii_rank(uint64_t x) { return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL); }
inline uint64_t normalize_code(uint64_t code, int k) {
    const uint64_t rc = mc_revcomp(code, k);
    return ascii_rank(code) <= ascii_rank(rc) ? code : rc;   // s.compareTo(rc) < 0 ? s : rc  (equal strings are the same string)
}
inline void kmer_chars(uint64_t code, int k, char* out) {
    for (int i = k - 1; i >= 0; i--, code >>= 2) out[i] = NUC[code & 3];
}

void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i < path.size(); i++) {
        cur.push_back(path[i]);
        if (path[i] == '/' || i + 1 == path.size()) mkdir(cur.c_str(), 0777);
    }
}

// BufferedReader.readLine over a whole file
std::vector<std::string> read_lines(const std::string& path, const char* what) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) mc_throw(MC_ERR_IO, "%s %s", what, path.c_str());
    std::string data;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, got);
    fclose(f);
    std::vector<std::string> lines;
    size_t p = 0;
    while (p < data.size()) {
        size_t s = p;
        while (p < data.size() && data[p] != '\n' && data[p] != '\r') p++;
        lines.emplace_back(data, s, p - s);
        if (p < data.size()) p += (data[p] == '\r' && p + 1 < data.size() && data[p + 1] == '\n') ? 2 : 1;
    }
    return lines;
}

// ------------------------------------------------------------------------------------------------
// The merge sweep (doMerge, OneSequenceCalculator.java:434-451 / MultiSequenceCalculator.java:102-137) over the
// device's neighbour lists.  One ascending sweep does all the work (a node that is alive when visited and has a
// mergeable start is merged on the spot; the merge conditions are static), the reference's second sweep finds nothing.
// Sequences are kept as linked lists of original nodes, so a merge is O(1).
// ------------------------------------------------------------------------------------------------
struct Graph {
    int k = 0;
    uint32_t n2 = 0;
    struct NbrList {                          // at most 4 neighbours, a palindromic one counts twice
        int32_t v[8];
        uint8_t n = 0;
        size_t size() const { return n; }
        int32_t operator[](size_t i) const { return v[i]; }
        const int32_t* begin() const { return v; }
        const int32_t* end() const { return v + n; }
        void push_back(int32_t x) { v[n++] = x; }
    };
    std::vector<NbrList> nbrs;                // expanded neighbour lists (palindromes as two ids)
    std::vector<uint8_t> gene, deleted;
    std::vector<int32_t> rc;                  // current reverse-complement partner (stale for deleted nodes, like the reference)
    // A node's sequence is an immutable concatenation tree over the original nodes (leaf v = original node v, id >= n2 =
    // cat[id - n2]); mergeLabels(a, b) = a + b[k-1:] becomes a new tree node, so assignments behave exactly like the
    // reference's String assignments (also in the self-referential corner cases) and a merge stays O(1).
    std::vector<std::pair<int32_t, int32_t>> cat;
    std::vector<int32_t> seq_id;
    std::vector<std::string> seq;             // materialised for alive nodes and for the rc partners of alive nodes
    std::vector<uint64_t> node_code;          // oriented k-mer of every original node
    std::vector<long> kc_pre;                 // from_device: KC of every alive node (GFAWriter.java:90-95), else empty

    // The unitigs as the device assembled them (mc_unitig.cu: chains of mergeable links ranked by pointer jumping, sequences and
    // KC sums gathered along the chains): the host only sets up what the writers read -- which nodes are alive, their partners,
    // their sequences, the neighbour lists of the alive nodes.  Used when the environment is `regular` (McUnitigNodes).
    void from_device(const McUnitigInput& in, const McUnitigNodes& nodes) {
        k = in.k;
        n2 = (uint32_t)in.kmers.size() * 2;
        nbrs.assign(n2, NbrList());
        gene = nodes.gene;
        deleted.assign(n2, 1);
        rc.assign(n2, -1);
        seq.assign(n2, std::string());
        kc_pre.assign(n2, 0);
        for (size_t i = 0; i < nodes.u_node.size(); i++) {
            const uint32_t v = (uint32_t)nodes.u_node[i];
            deleted[v] = 0;
            rc[v] = nodes.u_rc[i];
            seq[v].assign(nodes.u_chars, nodes.u_off[i], nodes.u_len[i]);
            kc_pre[v] = (long)nodes.u_kc[i];
            for (int s4 = 0; s4 < 4; s4++) {
                const int32_t x = nodes.nbr[4 * (size_t)v + s4];
                if (x >= 0) nbrs[v].push_back(x);   // (a regular environment has no palindromic neighbour)
            }
        }
    }

    void build(const McUnitigInput& in, const McUnitigNodes& nodes) {
        k = in.k;
        n2 = (uint32_t)in.kmers.size() * 2;
        nbrs.assign(n2, NbrList());
        gene = nodes.gene;
        deleted.assign(n2, 0);
        rc.resize(n2);
        seq_id.resize(n2);
        node_code.resize(n2);
        for (uint32_t v = 0; v < n2; v++) {
            rc[v] = (int32_t)(v ^ 1);
            seq_id[v] = (int32_t)v;
            node_code[v] = (v & 1) ? mc_revcomp(in.kmers[v >> 1], k) : in.kmers[v >> 1];
            for (int s = 0; s < 4; s++) {
                int32_t x = nodes.nbr[4 * (size_t)v + s];
                if (x < 0) continue;
                nbrs[v].push_back(x);
                if (nodes.nbr_pal[v] & (1u << s)) nbrs[v].push_back(x + 1);
            }
        }
    }
    int32_t concat(int32_t a, int32_t b) {
        cat.emplace_back(a, b);
        return (int32_t)(n2 + cat.size() - 1);
    }
    void merge(int32_t first_plus, int32_t second_minus) {   // mergeNodes (:312-324), statement by statement
        int32_t first_minus = rc[first_plus], second_plus = rc[second_minus];
        int32_t new_seq = concat(seq_id[second_plus], seq_id[first_plus]);
        int32_t new_seq_rc = concat(seq_id[first_minus], seq_id[second_minus]);
        seq_id[second_plus] = new_seq;
        seq_id[first_minus] = new_seq_rc;
        rc[second_plus] = first_minus;
        rc[first_minus] = second_plus;
        deleted[first_plus] = deleted[second_minus] = 1;
    }
    void sweep(const std::vector<uint64_t>* colors) {
        for (uint32_t i = 0; i < n2; i++) {
            if (deleted[i] || nbrs[i].size() != 1) continue;
            int32_t other = nbrs[i][0];
            if (nbrs[other].size() != 1 || gene[i] != gene[other]) continue;
            if (colors && (*colors)[i >> 1] != (*colors)[(uint32_t)other >> 1]) continue;
            merge((int32_t)i, other);
        }
    }
    // leaves of a concatenation tree, left to right
    void leaves(int32_t id, std::vector<int32_t>& out) const {
        std::vector<int32_t> stack{id};
        while (!stack.empty()) {
            int32_t x = stack.back();
            stack.pop_back();
            if ((uint32_t)x < n2) { out.push_back(x); continue; }
            const std::pair<int32_t, int32_t>& c = cat[(uint32_t)x - n2];
            stack.push_back(c.second);
            stack.push_back(c.first);
        }
    }
    std::string materialise(int32_t v) const {
        std::vector<int32_t> lv;
        leaves(seq_id[v], lv);
        std::string s = kmer_string(node_code[lv[0]], k);
        for (size_t i = 1; i < lv.size(); i++) s.push_back(NUC[node_code[lv[i]] & 3]);
        return s;
    }
    void assemble() {
        seq.assign(n2, std::string());
        std::vector<uint8_t> need(n2, 0);
        for (uint32_t v = 0; v < n2; v++)
            if (!deleted[v]) need[v] = need[(uint32_t)rc[v]] = 1;
        for (uint32_t v = 0; v < n2; v++)
            if (need[v]) seq[v] = materialise((int32_t)v);
    }
    // original nodes whose k-mers make up node v's sequence, in order
    std::vector<int32_t> path(int32_t v) const {
        std::vector<int32_t> lv;
        leaves(seq_id[v], lv);
        return lv;
    }
    int32_t min_id(int32_t v) const { return std::min(v, rc[v]) + 1; }
    std::string neighbor_ids(int32_t v) const {   // getNeighborIds (:375-385): TreeSet<Integer>.toString()
        std::set<int32_t> r;
        for (int32_t x : nbrs[v]) r.insert(min_id(x));
        for (int32_t x : nbrs[rc[v]]) r.insert(min_id(x));
        r.erase(min_id(v));
        std::string s = "[";
        bool first = true;
        for (int32_t x : r) {
            if (!first) s += ", ";
            s += std::to_string(x);
            first = false;
        }
        return s + "]";
    }
};

std::vector<uint64_t> target_windows(const std::vector<std::string>& targets, int k) {
    std::vector<uint64_t> w;
    for (const std::string& t : targets)
        for (size_t i = 0; i + k <= t.size(); i++) w.push_back(kmer_code(t.substr(i, k)));
    std::sort(w.begin(), w.end());
    w.erase(std::unique(w.begin(), w.end()), w.end());
    return w;
}

// ------------------------------------------------------------------------------------------------
// single-environment writers
// ------------------------------------------------------------------------------------------------
void write_environment(const KmerMap& subgraph, const std::vector<uint32_t>& ord, const std::string& prefix) {   // printEnvironment (:297-310)
    mkdirs(prefix);
    const int k = subgraph.k;
    std::string body;
    body.resize(ord.size() * (size_t)(k + 14));   // k-mer, space, up to 10 digits (+ sign), newline
    char* out = &body[0];
    for (uint32_t p : ord) {
        kmer_chars(subgraph.keys[p], k, out);
        out += k;
        *out++ = ' ';
        {   // Integer.toString of a count (<= 32767; negative values cannot occur: digits by hand, snprintf was a third of this stage)
            int32_t sv = subgraph.vals[p];
            if (sv < 0) *out++ = '-';   // (only through a hand-made graph.txt)
            uint32_t v = sv < 0 ? 0u - (uint32_t)sv : (uint32_t)sv;
            char tmp[12];
            int nd = 0;
            do { tmp[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (nd) *out++ = tmp[--nd];
        }
        *out++ = '\n';
    }
    body.resize((size_t)(out - &body[0]));
    for (const char* name : {"/graph.txt", "/env.txt"}) {   // the reference writes graph.txt; README/BASELINE call it env.txt
        FILE* f = fopen((prefix + name).c_str(), "wb");
        if (!f) mc_throw(MC_ERR_IO, "cannot write %s%s", prefix.c_str(), name);
        fwrite(body.data(), 1, body.size(), f);
        fclose(f);
    }
}

void write_single_outputs(const Graph& g, const KmerMap& subgraph, int chunk_len, const std::string& prefix) {
    const int k = g.k;
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    // KC (GFAWriter.java:90-95): sum of the coverage of every k-window + coverage of the last window * (k - 1); the windows are
    // looked up in `subgraph` by their normalised form, rolling the 2-bit code along the sequence
    auto kc = [&](const std::string& seq) -> long {
        long total = 0, last = 0;
        uint64_t code = 0;
        for (size_t i = 0; i < seq.size(); i++) {
            code = ((code << 2) | (uint64_t)(mc_nuc_code((uint8_t)seq[i]) & 3)) & kmask;
            if (i + 1 < (size_t)k) continue;
            const int32_t* c = subgraph.get(normalize_code(code, k));
            if (!c) mc_throw(MC_ERR_FORMAT, "k-mer %s of a unitig is not in the environment", seq.substr(i + 1 - k, k).c_str());
            last = *c;
            total += last;
        }
        return total + last * (k - 1);
    };
    auto node_id = [&](int32_t v) { return std::to_string(g.min_id(v)) + (g.gene[v] ? "_start" : ""); };
    auto fwd = [&](int32_t v) { return g.seq[v].compare(g.seq[g.rc[v]]) <= 0; };
    mkdirs(prefix);
    {   // outputNodeSequences (:354-373)
        std::ofstream out(prefix + "/seqs.fasta", std::ios::binary);
        for (uint32_t v = 0; v < g.n2; v++)
            if (!g.deleted[v] && (int32_t)v < g.rc[v] && (int)g.seq[v].size() >= chunk_len)
                out << "> Id" << node_id((int32_t)v) << " Length:" << g.seq[v].size() << " Neighbors:" << g.neighbor_ids((int32_t)v) << "\n" << g.seq[v] << "\n";
    }
    {   // GFAWriter.printGraph (src/io/writers/GFAWriter.java:47-99)
        std::ofstream out(prefix + "/graph.gfa", std::ios::binary);
        for (uint32_t v = 0; v < g.n2; v++)
            if (!g.deleted[v] && fwd((int32_t)v)) {
                const std::string& s = g.seq[v];
                const long coverage = g.kc_pre.empty() ? kc(s) : g.kc_pre[v];
                out << "S\t" << node_id((int32_t)v) << "\t" << s << "\tLN:i:" << s.size() << "\tKC:i:" << coverage << (g.gene[v] ? "\tCL:Z:GREEN" : "") << "\n";
            }
        for (uint32_t i = 0; i < g.n2; i++)
            if (!g.deleted[i])
                for (int32_t j : g.nbrs[i])
                    if (!g.deleted[j])
                        out << "L\t" << node_id((int32_t)i) << "\t" << (g.seq[i].compare(g.seq[g.rc[i]]) >= 0 ? "+" : "-") << "\t" << node_id(j) << "\t"
                            << (fwd(j) ? "+" : "-") << "\t" << (k - 1) << "M\n";
    }
    {   // TSVWriter (src/io/writers/TSVWriter.java:27-78)
        mkdirs(prefix + "/tsvs");
        auto tsv_id = [&](int32_t v) { return (fwd(v) ? std::to_string(v + 1) : "-" + std::to_string(g.rc[v] + 1)) + (g.gene[v] ? "_start" : ""); };
        std::ofstream edges(prefix + "/tsvs/edges.tsv", std::ios::binary);
        edges << "source\ttarget\n";
        for (uint32_t i = 0; i < g.n2; i++)
            if (!g.deleted[i])
                for (int32_t j : g.nbrs[i])
                    if (!g.deleted[j]) edges << tsv_id(g.rc[i]) << "\t" << tsv_id(j) << "\tpp\n";
        std::ofstream nodes(prefix + "/tsvs/nodes.tsv", std::ios::binary);
        nodes << "id\tlength\tseq\n";
        for (uint32_t v = 0; v < g.n2; v++)
            if (!g.deleted[v] && fwd((int32_t)v)) nodes << (v + 1) << "\t" << g.seq[v].size() << "\t" << g.seq[v] << "\n";
    }
}

// subgraph of one target from its BFS passes: for (kmer : distanceToKmer.keySet()) subgraph.put(normalize(kmer), count) (:217-219,146-148)
// runTrimPaths (OneSequenceCalculator.java:241-262): from the lastKmers walk AGAINST the pass direction inside the visited
// set; what is not reached is dropped (keySet().retainAll).  The visited set is at most seeds + --maxkmers k-mers: host work.
static std::vector<uint8_t> trim_paths(const McBfsResult& r, int dir, int k) {
    const size_t n = r.kmers.size();
    std::unordered_map<uint64_t, uint32_t> at;
    at.reserve(n * 2);
    for (size_t i = 0; i < n; i++) at.emplace(r.kmers[i], (uint32_t)i);
    std::vector<uint8_t> keep(n, 0);
    std::vector<uint32_t> queue;
    for (size_t i = 0; i < n; i++)
        if (r.last[i] & 1) {
            keep[i] = 1;
            queue.push_back((uint32_t)i);
        }
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    for (size_t head = 0; head < queue.size(); head++) {
        const uint64_t kmer = r.kmers[queue[head]];
        for (int s = 0; s < 8; s++) {
            const bool left = (s & 1) == 0;
            // getNeighborsByDir(-dir): dir -1 -> right neighbours, dir +1 -> left neighbours, dir 0 -> both
            if ((dir == -1 && left) || (dir == 1 && !left)) continue;
            const uint64_t nuc = (uint64_t)(s >> 1);
            const uint64_t nb = left ? ((nuc << (2 * k - 2)) | (kmer >> 2)) : (((kmer << 2) | nuc) & kmask);
            auto it = at.find(nb);
            if (it != at.end() && !keep[it->second]) {
                keep[it->second] = 1;
                queue.push_back(it->second);
            }
        }
    }
    return keep;
}

bool build_subgraph(const std::vector<const McBfsResult*>& passes, const std::vector<int>& dirs, bool trim, int k, KmerMap& subgraph) {
    size_t total = 0;
    for (const McBfsResult* r : passes) total += r->kmers.size();
    subgraph.init(k, total);
    for (size_t pi = 0; pi < passes.size(); pi++) {
        const McBfsResult* r = passes[pi];
        if (r->kmers.empty()) return false;   // "fail": no seed k-mer in the reads
        KmerMap dist;                         // distanceToKmer: insertion order = discovery order, value = queue index
        dist.init(k, r->kmers.size());
        for (size_t i = 0; i < r->kmers.size(); i++) dist.put(r->kmers[i], (int32_t)i);
        if (trim) dist.retain(trim_paths(*r, dirs[pi], k));
        for (uint32_t p : dist.order()) subgraph.put(normalize_code(dist.keys[p], k), (int32_t)r->counts[(size_t)dist.vals[p]]);
    }
    return true;
}

void unitig_input_from_subgraph(const KmerMap& subgraph, const std::vector<uint32_t>& ord, const std::vector<std::string>& targets, int k,
                                McUnitigInput& in) {
    in.k = k;
    in.kmers.resize(ord.size());
    in.counts.resize(ord.size());
    for (size_t i = 0; i < ord.size(); i++) {
        in.kmers[i] = subgraph.keys[ord[i]];
        in.counts[i] = subgraph.vals[ord[i]];
    }
    in.gene_windows = target_windows(targets, k);
}

struct Target {
    std::string seq, name;
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// targets (RichFastaReader.readAll, src/io/RichFastaReader.java:38-76; DnaQ(str,0).toString())
// ------------------------------------------------------------------------------------------------
struct mc_targets {
    std::vector<std::string> dnas, comments;
};

static std::string dnaq_to_string(const std::string& raw) {
    std::string s(raw.size(), 'A');
    for (size_t i = 0; i < raw.size(); i++) {
        char c = raw[i];
        if (c == 'N' || c == 'n' || c == '.') { s[i] = 'A'; continue; }
        int code = mc_nuc_code((uint8_t)c);
        if (code >= 4) mc_throw(MC_ERR_FORMAT, "Incorrect nucleotide char: \"%c\" in a target sequence", c);
        s[i] = NUC[code];
    }
    return s;
}

static void load_targets(const std::string& path, mc_targets& t) {
    std::vector<std::string> lines = read_lines(path, "Could not load sequences from");
    bool last_comment = true;
    std::string cur_comment, cur_dna;
    for (const std::string& line : lines) {
        if (!line.empty() && (line[0] == '>' || line[0] == ';')) {
            if (!last_comment) {
                t.dnas.push_back(dnaq_to_string(cur_dna));
                cur_dna.clear();
                cur_comment.clear();
            }
            cur_comment += line.substr(1);
            last_comment = true;
        } else {
            if (last_comment) {
                t.comments.push_back(cur_comment);
                cur_dna.clear();
                cur_comment.clear();
            }
            cur_dna += line;
            last_comment = false;
        }
    }
    if (!cur_comment.empty()) t.comments.push_back(cur_comment);
    if (!cur_dna.empty()) t.dnas.push_back(dnaq_to_string(cur_dna));
}

extern "C" int mc_targets_load(const char* fasta_path, mc_targets** out) {
    try {
        if (!fasta_path || !out) mc_throw(MC_ERR_ARG, "null argument");
        mc_targets* t = new mc_targets();
        try {
            load_targets(fasta_path, *t);
        } catch (...) {
            delete t;
            throw;
        }
        *out = t;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
extern "C" int mc_targets_count(const mc_targets* t) { return t ? (int)t->dnas.size() : 0; }
extern "C" const char* mc_targets_seq(const mc_targets* t, int i) { return (t && i >= 0 && i < (int)t->dnas.size()) ? t->dnas[(size_t)i].c_str() : nullptr; }
extern "C" const char* mc_targets_name(const mc_targets* t, int i) {
    return (t && i >= 0 && i < (int)t->comments.size()) ? t->comments[(size_t)i].c_str() : nullptr;
}
extern "C" void mc_targets_free(mc_targets* t) { delete t; }

// ------------------------------------------------------------------------------------------------
// Seam 3
// ------------------------------------------------------------------------------------------------
// One job = one environment: its seed sequences (one target, or with --merge every target followed by the Hi-C sequences).
struct EnvJob {
    std::vector<std::string> seeds;
};
static void run_jobs(mc_ctx* ctx, const mc_table* t, const std::vector<EnvJob>& jobs, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                     bool both_dirs, bool trim, std::vector<KmerMap>& subgraphs, std::vector<int>& found) {
    std::vector<McBfsRequest> reqs;
    for (const EnvJob& job : jobs) {
        std::vector<uint8_t> codes;
        for (size_t i = 0; i < job.seeds.size(); i++) {
            if (i) codes.push_back(4);   // no seed window spans two sequences (:168-192 loop over the sequences one by one)
            std::vector<uint8_t> c = mc_codes_from_target(job.seeds[i]);
            codes.insert(codes.end(), c.begin(), c.end());
        }
        if (both_dirs) {
            reqs.push_back(McBfsRequest{codes, 0, trim});
        } else {   // buildEnvironment (:137-144): runBfs(-1) then runBfs(+1), each with its own visited set and budget
            reqs.push_back(McBfsRequest{codes, -1, trim});
            reqs.push_back(McBfsRequest{codes, 1, trim});
        }
    }
    std::vector<McBfsResult> res;
    {
        EnvTrace tr("bfs passes (device)");
        mc_run_bfs_passes(ctx, t, k, min_cov, max_kmers, max_radius, reqs, res);
    }
    EnvTrace tr("build_subgraph");
    const size_t per = both_dirs ? 1 : 2;
    subgraphs.assign(jobs.size(), KmerMap());
    found.assign(jobs.size(), 0);
    parallel_for(jobs.size(), [&](size_t i) {
        std::vector<const McBfsResult*> passes;
        std::vector<int> dirs;
        for (size_t p = 0; p < per; p++) {
            passes.push_back(&res[i * per + p]);
            dirs.push_back(reqs[i * per + p].dir);
        }
        found[i] = build_subgraph(passes, dirs, trim, k, subgraphs[i]) ? 1 : 0;
    });
}
static void run_targets(mc_ctx* ctx, const mc_table* t, const std::vector<Target>& targets, int k, int min_cov, int64_t max_kmers,
                        int64_t max_radius, bool both_dirs, std::vector<KmerMap>& subgraphs, std::vector<int>& found) {
    std::vector<EnvJob> jobs;
    for (const Target& tg : targets) jobs.push_back(EnvJob{{tg.seq}});
    run_jobs(ctx, t, jobs, k, min_cov, max_kmers, max_radius, both_dirs, false, subgraphs, found);
}

// graph.txt/env.txt + unitig stage + writers of the found environments; gene nodes are decided against `gene_seqs` (isGeneNode
// :421-432).  One task per environment on the thread pool: host stages run concurrently, the device stage in between (node
// hash, neighbour lists, chain ranks: mc_unitig.cu) is taken in turns (one stream per context), so it overlaps the other
// workers' host stages.
static void write_environments(mc_ctx* ctx, const std::vector<KmerMap>& subs, const std::vector<int>& found,
                               const std::vector<std::vector<std::string>>& gene_seqs, int k, int chunk_len, const std::vector<std::string>& prefixes,
                               const mc_table* table = nullptr, int min_cov = 0) {
    ctx->last_extend_counts.assign(subs.size(), -1);
    std::vector<size_t> todo;
    for (size_t i = 0; i < subs.size(); i++)
        if (found[i]) todo.push_back(i);   // else: "Could not find any k-mers of the target gene in the input, halting."
    std::mutex device_turn;
    parallel_for(todo.size(), [&](size_t w) {
        const size_t job = todo[w];
        const KmerMap& sub = subs[job];
        std::vector<uint32_t> ord;
        McUnitigInput in;
        McUnitigNodes nodes;
        { EnvTrace tr("subgraph order"); ord = sub.order(); }
        { EnvTrace tr("write env.txt/graph.txt"); write_environment(sub, ord, prefixes[job]); }
        { EnvTrace tr("unitig input"); unitig_input_from_subgraph(sub, ord, gene_seqs[job], k, in); }
        in.table = table;
        in.min_cov = min_cov;
        {
            std::lock_guard<std::mutex> turn(device_turn);
            EnvTrace tr("unitig nodes (device)");
            mc_unitig_nodes(ctx, in, nodes);
            ctx->last_extend_counts[job] = nodes.extend_count;
        }
        Graph g;
        if (nodes.regular && !getenv("MC_HOST_UNITIGS")) {
            EnvTrace tr("unitigs from the device");
            g.from_device(in, nodes);
        } else {   // palindromes, hairpins, cyclic chains: the literal sweep (its outcome depends on the reference's sweep order)
            { EnvTrace tr("graph build"); g.build(in, nodes); }
            { EnvTrace tr("merge sweep"); g.sweep(nullptr); }
            { EnvTrace tr("assemble"); g.assemble(); }
        }
        { EnvTrace tr("write gfa/fasta/tsv"); write_single_outputs(g, sub, chunk_len, prefixes[job]); }
    });
}

// --merge / --hicseq / --trim (EnvironmentFinderMain.java:218-233, OneSequenceCalculator.java:168-192,208-216,241-262)
extern "C" int mc_env_find_ex(mc_ctx* ctx, const mc_table* t, const char* const* seqs, const char* const* names, int n_targets,
                              const char* const* hic_seqs, int n_hic, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                              int chunk_len, int merge, int trim, const char* out_dir, int* found_out) {
    try {
        if (!ctx || !t || !out_dir || (n_targets && !seqs) || (n_hic && !hic_seqs) || (!merge && n_targets && !names)) mc_throw(MC_ERR_ARG, "null argument");
        std::vector<EnvJob> jobs;
        std::vector<std::string> prefixes;
        std::vector<std::vector<std::string>> gene_seqs;
        if (merge) {   // one environment seeded by every target and every Hi-C sequence, written to <out>/merged/
            EnvJob job;
            for (int i = 0; i < n_targets; i++) job.seeds.push_back(seqs[i]);
            gene_seqs.push_back(job.seeds);            // isGeneNode looks at `sequences` only, not at the Hi-C reads
            for (int i = 0; i < n_hic; i++) job.seeds.push_back(hic_seqs[i]);
            jobs.push_back(job);
            prefixes.push_back(std::string(out_dir) + "/merged/");
        } else {
            for (int i = 0; i < n_targets; i++) {
                jobs.push_back(EnvJob{{seqs[i]}});
                gene_seqs.push_back({seqs[i]});
                prefixes.push_back(std::string(out_dir) + "/" + names[i] + "/");   // getOutputPrefix (EnvironmentFinderMain.java:245-249)
            }
        }
        std::vector<KmerMap> subgraphs;
        std::vector<int> found;
        run_jobs(ctx, t, jobs, k, min_cov, max_kmers, max_radius, both_dirs != 0, trim != 0, subgraphs, found);
        for (size_t i = 0; i < jobs.size(); i++)
            if (found_out) found_out[i] = found[i];
        {
            EnvTrace tr("environments -> files (wall)");
            write_environments(ctx, subgraphs, found, gene_seqs, k, chunk_len, prefixes, t, min_cov);
        }
        EnvTrace::report();
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// "Extending endings by N kmers" (OneSequenceCalculator.java:291) of every environment of the last mc_env_find(_ex) call on this
// context, in the order of its targets (-1: no environment was built for that target)
extern "C" int mc_env_extend_counts(mc_ctx* ctx, long long* out, int cap, int* n) {
    if (!ctx || !n || (cap && !out)) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    *n = (int)ctx->last_extend_counts.size();
    for (int i = 0; i < *n && i < cap; i++) out[i] = ctx->last_extend_counts[(size_t)i];
    return MC_OK;
}

extern "C" int mc_env_find(mc_ctx* ctx, const mc_table* t, const char* const* seqs, const char* const* names, int n_targets, int k, int min_cov,
                           int64_t max_kmers, int64_t max_radius, int both_dirs, int chunk_len, const char* out_dir, int* found_out) {
    return mc_env_find_ex(ctx, t, seqs, names, n_targets, nullptr, 0, k, min_cov, max_kmers, max_radius, both_dirs, chunk_len, 0, 0, out_dir, found_out);
}

extern "C" int mc_env_kmers(mc_ctx* ctx, const mc_table* t, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                            int both_dirs, uint64_t* kmers, int32_t* counts, size_t cap, size_t* n) {
    try {
        if (!ctx || !t || !seq || !n) mc_throw(MC_ERR_ARG, "null argument");
        std::vector<KmerMap> subgraphs;
        std::vector<int> found;
        run_targets(ctx, t, {Target{seq, ""}}, k, min_cov, max_kmers, max_radius, both_dirs != 0, subgraphs, found);
        if (!found[0]) {
            *n = 0;
            return MC_OK;
        }
        std::vector<uint32_t> ord = subgraphs[0].order();
        *n = ord.size();
        if (ord.size() > cap) mc_throw(MC_ERR_CAPACITY, "environment has %zu k-mers, room for %zu", ord.size(), cap);
        for (size_t i = 0; i < ord.size(); i++) {
            if (kmers) kmers[i] = subgraphs[0].keys[ord[i]];
            if (counts) counts[i] = subgraphs[0].vals[ord[i]];
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// stage D alone, from an environment file (test surface for the reference's committed graph.txt): exported for tests/CLI
extern "C" int mc_env_from_graph_txt(mc_ctx* ctx, const char* graph_txt, const char* target_seq, int k, int chunk_len, const char* out_dir) {
    try {
        if (!ctx || !graph_txt || !target_seq || !out_dir) mc_throw(MC_ERR_ARG, "null argument");
        KmerMap sub;
        const std::vector<std::string> lines = read_lines(graph_txt, "Couldn't load graph from file");
        sub.init(k, lines.size());
        for (const std::string& line : lines) {
            size_t sp = line.find(' ');
            if (sp == std::string::npos) mc_throw(MC_ERR_FORMAT, "malformed environment line \"%s\"", line.c_str());
            if (sp != (size_t)k) mc_throw(MC_ERR_FORMAT, "environment k-mer \"%s\" is not %d bases long", line.substr(0, sp).c_str(), k);
            for (size_t i = 0; i < sp; i++)
                if (line[i] != 'A' && line[i] != 'C' && line[i] != 'G' && line[i] != 'T') mc_throw(MC_ERR_FORMAT, "Incorrect nucleotide char: \"%c\"", line[i]);
            sub.put(kmer_code(line.substr(0, sp)), atoi(line.c_str() + sp + 1));
        }
        std::vector<uint32_t> ord = sub.order();
        std::string prefix = std::string(out_dir) + "/";
        write_environment(sub, ord, prefix);
        McUnitigInput in;
        unitig_input_from_subgraph(sub, ord, {std::string(target_seq)}, k, in);
        McUnitigNodes nodes;
        mc_unitig_nodes(ctx, in, nodes);
        Graph g;
        if (nodes.regular && !getenv("MC_HOST_UNITIGS")) {
            g.from_device(in, nodes);
        } else {
            g.build(in, nodes);
            g.sweep(nullptr);
            g.assemble();
        }
        write_single_outputs(g, sub, chunk_len, prefix);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// host-only test surface: the HashMap order emulation of the environment stage (KmerMap)
extern "C" int mc_kmer_hashmap_order(const uint64_t* codes, size_t n, int k, size_t peak, uint32_t* order, size_t* n_out) {
    try {
        if ((n && (!codes || !order)) || !n_out || k < 1 || k > 31) mc_throw(MC_ERR_ARG, "bad argument");
        KmerMap m;
        m.init(k, n);
        std::vector<uint32_t> first_index;
        first_index.reserve(n);
        for (size_t i = 0; i < n; i++) {
            const size_t before = m.keys.size();
            m.put(codes[i], (int32_t)i);
            if (m.keys.size() != before) first_index.push_back((uint32_t)i);
        }
        m.peak = std::max(m.peak, peak);
        const std::vector<uint32_t> ord = m.order();
        for (size_t i = 0; i < ord.size(); i++) order[i] = first_index[ord[i]];
        *n_out = ord.size();
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// ------------------------------------------------------------------------------------------------
// Seam 4: environment-finder-multi
// ------------------------------------------------------------------------------------------------
// String.format("%6.2f ", floatValue) (src/tools/EnvironmentFinderMultiMain.java:141-142,163-164).  java.util.Formatter widens
// the float to double, takes the digits of Double.toString -- the shortest decimal that reads back as the same double -- and
// rounds THAT digit string HALF_UP to two decimals (sun.misc.FormattedFloatingDecimal.applyPrecision); C's printf rounds the
// exact binary value half-to-even instead, so 0.125f prints " 0.13" in Java and " 0.12" in C.
// Written on integers (shortest digits from std::to_chars, then one division) and pinned on hand-derived Java outputs in
// tests/test_host_logic.py; the oracle's own formatter works on digit strings, so the two do not share an algorithm.
static std::string java_format_6_2f(float v) {
    if (std::isnan(v)) return "   NaN ";
    if (std::isinf(v)) return v > 0 ? "Infinity " : "-Infinity ";
    const double dv = std::fabs((double)v);
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, dv, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest round-trip digits
    *r.ptr = 0;
    unsigned long long mant = 0;   // the digits as one integer (<= 17 digits)
    int n_digits = 0, exp10 = 0;
    for (const char* p = buf; *p; p++) {
        if (*p == 'e') { exp10 = atoi(p + 1); break; }
        if (*p >= '0' && *p <= '9') { mant = mant * 10 + (unsigned)(*p - '0'); n_digits++; }
    }
    // dv = mant * 10^(exp10 - n_digits + 1); hundredths = dv * 100, HALF_UP: drop `drop` digits, look at the first dropped one
    const int scale = exp10 - n_digits + 1 + 2;
    unsigned __int128 hundredths;
    if (scale >= 0) {
        hundredths = mant;
        for (int i = 0; i < scale && i < 30; i++) hundredths *= 10;   // a float is < 3.5e38: fits 128 bits
    } else if (-scale > 19) {
        hundredths = 0;                                                // smaller than 0.5e-2 by far
    } else {
        unsigned long long p10 = 1;
        for (int i = 0; i < -scale - 1; i++) p10 *= 10;
        hundredths = (mant / p10 + 5) / 10;                            // first dropped digit >= 5 rounds up
    }
    std::string digits;
    do { digits.insert(digits.begin(), (char)('0' + (int)(hundredths % 10))); hundredths /= 10; } while (hundredths != 0);
    while (digits.size() < 3) digits.insert(digits.begin(), '0');
    std::string out = digits.substr(0, digits.size() - 2) + "." + digits.substr(digits.size() - 2);
    if (std::signbit(v)) out.insert(out.begin(), '-');                 // Java prints -0.00 for a negative value that rounds to zero
    if (out.size() < 6) out.insert(0, 6 - out.size(), ' ');
    return out + " ";
}
// host-only test surface (tests/test_host_logic.py): out receives the NUL-terminated text, at most 48 characters
extern "C" int mc_java_format_6_2f(float v, char* out) {
    if (!out) return MC_ERR_ARG;
    const std::string s = java_format_6_2f(v);
    memcpy(out, s.c_str(), s.size() + 1);
    return MC_OK;
}

extern "C" int mc_env_multi(mc_ctx* ctx, const char* const* env_paths, int n_envs, const char* seq_fasta, int gene_id, const char* out_dir) {
    try {
        if (!ctx || !seq_fasta || !out_dir || (n_envs && !env_paths)) mc_throw(MC_ERR_ARG, "null argument");
        // loadInput (src/tools/EnvironmentFinderMultiMain.java:49-78)
        std::vector<JavaMap<int>> graphs((size_t)n_envs);
        for (int i = 0; i < n_envs; i++)
            for (const std::string& line : read_lines(env_paths[i], "Couldn't load graph from file")) {   // DeBruijnGraphUtils.loadGraph
                size_t sp = line.find(' ');
                if (sp == std::string::npos) mc_throw(MC_ERR_FORMAT, "malformed environment line \"%s\" in %s", line.c_str(), env_paths[i]);
                std::string rest = line.substr(sp + 1);
                size_t sp2 = rest.find(' ');
                if (sp2 != std::string::npos) rest = rest.substr(0, sp2);
                char* endp = nullptr;
                long depth = strtol(rest.c_str(), &endp, 10);
                if (rest.empty() || *endp) mc_throw(MC_ERR_FORMAT, "For input string: \"%s\" in %s", rest.c_str(), env_paths[i]);
                graphs[(size_t)i].put(line.substr(0, sp), (int)depth);
            }
        if (graphs.empty()) mc_throw(MC_ERR_ARG, "Zero environments given");
        if (graphs[0].size() == 0) mc_throw(MC_ERR_FORMAT, "first environment is empty");
        const int k = (int)graphs[0].entries[graphs[0].order()[0]].first.size();
        if (k < 1 || k > 31) mc_throw(MC_ERR_UNSUPPORTED, "k = %d is outside the GPU path's 1..31", k);
        for (auto& g : graphs)
            for (auto& e : g.entries)
                if ((int)e.first.size() != k) mc_throw(MC_ERR_FORMAT, "K-mers of different lengths encountered: %d and %d", k, (int)e.first.size());
        mc_targets tg;
        load_targets(seq_fasta, tg);
        if (gene_id < 1 || gene_id > (int)tg.dnas.size() || gene_id > (int)tg.comments.size()) mc_throw(MC_ERR_ARG, "--geneid %d is out of range", gene_id);
        const std::string sequence = tg.dnas[(size_t)gene_id - 1], comment = tg.comments[(size_t)gene_id - 1];

        // MultiSequenceCalculator.initializeStructures (src/algo/MultiSequenceCalculator.java:51-100): node order
        JavaMap<int> node_by_kmer;
        for (auto& g : graphs)
            for (uint32_t p : g.order()) {
                const std::string& kmer = g.entries[p].first;
                node_by_kmer.put(kmer, -1);
                node_by_kmer.put(rc_string(kmer), -1);
            }
        McUnitigInput in;
        in.k = k;
        in.order_by_nucleotide = true;
        std::unordered_map<std::string, uint32_t> pair_of;   // either strand -> env index
        for (uint32_t p : node_by_kmer.order()) {
            const std::string& kmer = node_by_kmer.entries[p].first;
            std::string rc = rc_string(kmer);
            if (kmer.compare(rc) > 0) continue;
            if (kmer == rc) mc_throw(MC_ERR_UNSUPPORTED, "palindromic k-mer %s: the reference dereferences a null node here (even k)", kmer.c_str());
            pair_of[kmer] = pair_of[rc] = (uint32_t)in.kmers.size();
            in.kmers.push_back(kmer_code(kmer));
        }
        const size_t n = in.kmers.size();
        in.counts.assign(n, 0);
        in.colors.assign(n, 0);
        std::vector<std::vector<uint8_t>> member(n_envs > 64 ? n : 0);
        std::vector<uint32_t> n_graphs(n, 0);
        for (int gi = 0; gi < n_envs; gi++) {
            std::vector<uint8_t> seen(n, 0);
            for (auto& e : graphs[(size_t)gi].entries) {
                uint32_t idx = pair_of.at(e.first);
                if (!seen[idx]) { seen[idx] = 1; n_graphs[idx]++; }
                // colour = the SET of environments; 64-bit FNV-style fold keeps sets of up to 64 exact, beyond that hashed
                if (n_envs <= 64) in.colors[idx] |= 1ULL << gi;
                else in.colors[idx] = (in.colors[idx] ^ (uint64_t)(gi + 1)) * 0x100000001B3ULL + (uint64_t)gi;
            }
        }
        // KC of a node = sum over environments of graph.get(normalizeDna(kmer)) (GFAWriterMulti.java:80-88)
        for (size_t i = 0; i < n; i++) {
            std::string canon = kmer_string(in.kmers[i], k);
            for (auto& g : graphs) {
                const int* c = g.get(canon);
                if (c) in.counts[i] += *c;
            }
        }
        in.gene_windows = target_windows({sequence}, k);
        McUnitigNodes nodes;
        mc_unitig_nodes(ctx, in, nodes);
        Graph g;
        g.build(in, nodes);
        g.sweep(&in.colors);   // canBeMerged: equal gene flag AND equal environment set (:120-122)
        g.assemble();

        const std::string prefix = out_dir;
        mkdirs(prefix);
        auto gene_sfx = [&](int32_t v) { return g.gene[v] ? "_start" : ""; };
        {   // outputNodeSequences (:139-175)
            std::ofstream out(prefix + "/seqs.fasta", std::ios::binary);
            for (uint32_t v = 0; v < g.n2; v++)
                if (!g.deleted[v] && (int32_t)v < g.rc[v])
                    out << "> Id" << g.min_id((int32_t)v) << gene_sfx((int32_t)v) << " Length:" << g.seq[v].size() << " Neighbors:" << g.neighbor_ids((int32_t)v) << "\n" << g.seq[v] << "\n";
        }
        {   // GFAWriterMulti.printGraph (src/io/writers/GFAWriterMulti.java:39-133)
            std::ofstream out(prefix + "/graph.gfa", std::ios::binary);
            const size_t G = graphs.size();
            for (uint32_t v = 0; v < g.n2; v++)
                if (!g.deleted[v] && (int32_t)v < g.rc[v]) {
                    long coverage = 0;
                    for (int32_t w : g.path((int32_t)v)) coverage += in.counts[(uint32_t)w >> 1];
                    size_t gs = n_graphs[v >> 1];
                    std::string color;
                    if (G == 2) color = g.gene[v] ? "#00ff00" : gs == 1 ? "#ff0000" : gs == 2 ? "#0000ff" : "#000000";
                    else if (G == 3) color = g.gene[v] ? "#00ff00" : gs == 1 ? "#ff0000" : gs == 2 ? "#0000ff" : gs == 3 ? "#ff00ff" : "#000000";
                    else if (g.gene[v]) color = "#00ff00";
                    else {
                        int value = (int)(256 * gs / G);
                        char b[32];
                        snprintf(b, sizeof b, "#%02X%02X%02X", value, value, value);
                        color = b;
                    }
                    out << "S\t" << g.min_id((int32_t)v) << gene_sfx((int32_t)v) << "\t" << g.seq[v] << "\tLN:i:" << g.seq[v].size() << "\tKC:i:" << coverage
                        << "\tCL:Z:" << color << "\tC2:Z:" << color << "\n";
                }
            for (uint32_t i = 0; i < g.n2; i++)
                if (!g.deleted[i])
                    for (int32_t j : g.nbrs[i])   // no deleted-target check (GFAWriterMulti.java:51-57)
                        out << "L\t" << g.min_id((int32_t)i) << gene_sfx((int32_t)i) << "\t" << ((int32_t)i < g.rc[i] ? "+" : "-") << "\t" << g.min_id(j) << gene_sfx(j)
                            << "\t" << (j > g.rc[j] ? "+" : "-") << "\t" << (k - 1) << "M\n";
        }
        {   // printGene (EnvironmentFinderMultiMain.java:93-104)
            std::ofstream out(prefix + "/gene.fasta", std::ios::binary);
            out << ">" << comment << "\n" << sequence << "\n";
        }
        {   // printProbability (:106-168): int32 sums, float division, "%6.2f "
            const size_t G = graphs.size();
            std::ofstream sym(prefix + "/Jacard_sym.txt", std::ios::binary), alt(prefix + "/Jacard_alt.txt", std::ios::binary);
            sym << "The[31mWarning! symmetric <<Jaccard distance>> (1 - AB/AUB):\n\n";
            alt << "The[31mWarning! alternative <<Jaccard distance>> (1 - AB/A):\n\n";
            for (size_t i = 0; i < G; i++) {
                sym << env_paths[i];
                alt << env_paths[i];
                for (size_t j = 0; j < G; j++) {
                    int32_t diff = 0, diff_alt = 0, uni = 0;
                    for (auto& e : graphs[i].entries) {
                        const int* o = graphs[j].get(e.first);
                        if (!o) { diff += e.second; diff_alt += e.second; uni += e.second; }
                        else { diff += std::abs(e.second - *o); diff_alt += std::abs(e.second - *o); uni += std::max(e.second, *o); }
                    }
                    for (auto& e : graphs[j].entries)
                        if (!graphs[i].get(e.first)) { diff += e.second; uni += e.second; }
                    int32_t inter = uni - diff;
                    sym << java_format_6_2f(1 - inter / (float)uni);
                    alt << java_format_6_2f(1 - inter / (float)(uni - diff_alt));
                }
                sym << "\n";
                alt << "\n";
            }
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// ------------------------------------------------------------------------------------------------
// "next" row f-1: kmer-counter's .kmers.bin (src/io/IOUtils.java:39-65) and loadKmers (:94-126)
// ------------------------------------------------------------------------------------------------
extern "C" int mc_table_print_kmers(const mc_table* t, int threshold, const char* out_file, const char* stat_file, uint64_t* good_out) {
    try {
        if (!t || !out_file) mc_throw(MC_ERR_ARG, "null argument");
        size_t n = 0;
        int rc = mc_table_dump(t, 0, nullptr, nullptr, 0, &n);
        if (rc != MC_OK) return rc;
        std::vector<uint64_t> keys(n);
        std::vector<int16_t> counts(n);
        if (n) {
            rc = mc_table_dump(t, 0, keys.data(), counts.data(), n, &n);
            if (rc != MC_OK) return rc;
        }
        FILE* f = fopen(out_file, "wb");
        if (!f) mc_throw(MC_ERR_IO, "cannot open %s", out_file);
        std::vector<uint64_t> hist(32768, 0);
        uint64_t good = 0;
        std::vector<uint8_t> buf;
        buf.reserve(10 << 16);
        for (size_t i = 0; i < n; i++) {
            hist[(size_t)counts[i]]++;
            if (counts[i] > threshold) {
                for (int b = 0; b < 8; b++) buf.push_back((uint8_t)(keys[i] >> (56 - 8 * b)));   // DataOutputStream: big-endian
                buf.push_back((uint8_t)((uint16_t)counts[i] >> 8));
                buf.push_back((uint8_t)counts[i]);
                good++;
                if (buf.size() >= (10u << 16)) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
            }
        }
        if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
        fclose(f);
        if (stat_file) {
            std::ofstream st(stat_file, std::ios::binary);
            st << "# k-mer frequency\tnumber of such k-mers\n";
            for (size_t c = 0; c < hist.size(); c++)
                if (hist[c]) st << c << "\t" << hist[c] << "\n";
        }
        if (good_out) *good_out = good;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_table_load_kmers(mc_table* t, const char* path, int freq_threshold) {
    try {
        if (!t || !path) mc_throw(MC_ERR_ARG, "null argument");
        FILE* f = fopen(path, "rb");
        if (!f) mc_throw(MC_ERR_IO, "Can't load k-mers file %s", path);
        std::vector<uint64_t> keys;
        std::vector<int16_t> counts;
        uint8_t rec[10];
        while (fread(rec, 1, 10, f) == 10) {
            uint64_t key = 0;
            for (int b = 0; b < 8; b++) key = (key << 8) | rec[b];
            int16_t freq = (int16_t)((rec[8] << 8) | rec[9]);
            if (freq > freq_threshold) { keys.push_back(key); counts.push_back(freq); }   // Kmers2HMWorker.processKmer (:84-91)
        }
        fclose(f);
        return mc_table_add_pairs(t, keys.data(), counts.data(), keys.size());
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
