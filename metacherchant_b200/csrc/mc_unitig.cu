// K6: unitig compaction of an environment on the device.
//
// Restates OneSequenceCalculator.initializeStructures / isGeneNode / doMerge / mergeNodes
// (src/algo/OneSequenceCalculator.java:387-451,312-324) and, with `colors`, MultiSequenceCalculator's variant
// (src/algo/MultiSequenceCalculator.java:51-122):
//   * nodes 2i / 2i+1 = env k-mer i as printed / its reverse complement;
//   * X.neighbors = nodes whose (k-1)-prefix equals the (k-1)-suffix of rc(X), in node-id order (:406-418): four
//     cooperating lanes probe the four successors of rc(X) in a hash of the oriented k-mers, a quad ballot gives the
//     degree and each lane ranks its hit among the quad's hits;
//   * X and O = X.neighbors[0] merge iff both have exactly one neighbour and equal gene flags (and equal environment
//     sets in multi mode) (:438-444, MultiSequenceCalculator.java:120-122).  Every merge keeps the two OUTER ends of the
//     joined path alive, so the result of the reference's repeated ascending sweeps does not depend on the sweep order:
//     the alive nodes are the heads of the maximal chains of mergeable links, and a chain that closes into a cycle
//     deletes all of its nodes.  Chains are ranked with pointer jumping (ceil(log2 2n)+1 rounds).
//   * Deviation kept on purpose: a link from a node's start to ITS OWN start (possible only when a (k-1)-mer is its own
//     reverse complement) is not merged; the reference's outcome for that case depends on its sweep order.
#include <vector>

#include "mc_env.hpp"

struct NodeHash {
    uint64_t* keys;   // oriented code + 1
    int32_t* vals;    // node id (the even id for a palindromic k-mer)
    uint32_t mask;
};

__device__ __forceinline__ int32_t node_lookup(const NodeHash& h, uint64_t code) {
    uint32_t pos = (uint32_t)(mc_mix62(code) >> 18) & h.mask;
    while (true) {
        uint64_t cur = h.keys[pos];
        if (cur == 0) return -1;
        if (cur == code + 1) return h.vals[pos];
        pos = (pos + 1) & h.mask;
    }
}

__global__ void __launch_bounds__(256) k_node_insert(NodeHash h, const uint64_t* __restrict__ kmers, uint32_t n, int k) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;   // oriented node id
    if (v >= 2 * n) return;
    uint64_t code = kmers[v >> 1];
    uint64_t rc = mc_revcomp(code, k);
    if (v & 1) {
        if (rc == code) return;   // palindrome: one entry, under the even id
        code = rc;
    }
    uint32_t pos = (uint32_t)(mc_mix62(code) >> 18) & h.mask;
    while (true) {
        uint64_t cur = atomicCAS((unsigned long long*)&h.keys[pos], 0ULL, (unsigned long long)(code + 1));
        if (cur == 0) { h.vals[pos] = (int32_t)v; return; }
        if (cur == code + 1) return;   // duplicate k-mer in the input list: keep the first inserted
        pos = (pos + 1) & h.mask;
    }
}

__device__ __forceinline__ bool in_sorted(const uint64_t* __restrict__ a, uint32_t n, uint64_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && a[lo] == x;
}

// 4 lanes per oriented node
__global__ void __launch_bounds__(256) k_neighbors(NodeHash h, const uint64_t* __restrict__ kmers, uint32_t n, int k,
                                                  const uint64_t* __restrict__ gene_windows, uint32_t n_windows,
                                                  uint8_t* __restrict__ deg, int32_t* __restrict__ nbr, uint8_t* __restrict__ nbr_pal,
                                                  uint8_t* __restrict__ gene, int order_by_nuc) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t v = gt >> 2, c = gt & 3;
    const int lane = threadIdx.x & 31;
    const bool live = v < 2 * n;
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    uint64_t code = 0, rc = 0;
    int32_t hit = -1;
    bool pal = false;
    if (live) {
        code = kmers[v >> 1];
        rc = mc_revcomp(code, k);
        if (v & 1) { uint64_t t = code; code = rc; rc = t; }
        uint64_t succ = ((rc << 2) | c) & kmask;   // c-th successor of rc(X): shares rc(X)'s (k-1)-suffix as its prefix
        hit = node_lookup(h, succ);
        pal = hit >= 0 && mc_revcomp(succ, k) == succ;
    }
    // quad-level degree and rank (ids ascending; successor codes are distinct so ids are distinct)
    const unsigned quad_shift = lane & ~3;
    int32_t h0 = __shfl_sync(0xffffffffu, hit, quad_shift + 0), h1 = __shfl_sync(0xffffffffu, hit, quad_shift + 1);
    int32_t h2 = __shfl_sync(0xffffffffu, hit, quad_shift + 2), h3 = __shfl_sync(0xffffffffu, hit, quad_shift + 3);
    unsigned palm = (__ballot_sync(0xffffffffu, pal) >> quad_shift) & 0xF;
    unsigned hitm = (__ballot_sync(0xffffffffu, hit >= 0) >> quad_shift) & 0xF;
    if (!live) return;
    if (hit >= 0) {
        int r = order_by_nuc ? __popc(hitm & ((1u << c) - 1))
                             : (h0 >= 0 && h0 < hit) + (h1 >= 0 && h1 < hit) + (h2 >= 0 && h2 < hit) + (h3 >= 0 && h3 < hit);
        nbr[4 * v + r] = hit;
        if (pal) atomicOr((unsigned int*)(nbr_pal + (v & ~3u)), (1u << r) << (8 * (v & 3)));   // byte v of a 4-byte word
    }
    if (c == 0) {
        deg[v] = (uint8_t)(__popc(hitm) + __popc(palm));
        gene[v] = (n_windows && (in_sorted(gene_windows, n_windows, code) || in_sorted(gene_windows, n_windows, rc))) ? 1 : 0;
    }
}

// pred[v] = the node that precedes v inside its merged path, or -1
__global__ void __launch_bounds__(256) k_links(uint32_t n2, const uint8_t* __restrict__ deg, const int32_t* __restrict__ nbr,
                                              const uint8_t* __restrict__ gene, const uint64_t* __restrict__ colors,
                                              int32_t* __restrict__ pred, int32_t* __restrict__ jump, uint32_t* __restrict__ rank) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n2) return;
    int32_t p = -1;
    if (deg[v] == 1) {
        int32_t o = nbr[4 * v];
        if (o >= 0 && o != (int32_t)v && deg[o] == 1 && gene[o] == gene[v] && (!colors || colors[o >> 1] == colors[v >> 1])) p = o ^ 1;
    }
    pred[v] = p;
    jump[v] = p >= 0 ? p : (int32_t)v;
    rank[v] = p >= 0 ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_jump(uint32_t n2, const int32_t* __restrict__ jump_in, const uint32_t* __restrict__ rank_in,
                                             int32_t* __restrict__ jump_out, uint32_t* __restrict__ rank_out) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n2) return;
    int32_t j = jump_in[v];
    uint32_t r = rank_in[v] + rank_in[j];
    jump_out[v] = jump_in[j];
    rank_out[v] = r < rank_in[v] ? 0xFFFFFFFFu : r;   // saturate (only cycles can get here)
}

__global__ void __launch_bounds__(256) k_finish(uint32_t n2, const int32_t* __restrict__ pred, const int32_t* __restrict__ jump,
                                               const uint32_t* __restrict__ rank, int32_t* __restrict__ head, int32_t* __restrict__ out_rank) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n2) return;
    int32_t j = jump[v];
    bool cycle = pred[j] >= 0;   // after enough rounds a path node points at its head; a cycle node never does
    head[v] = cycle ? -1 : j;
    out_rank[v] = cycle ? 0 : (int32_t)rank[v];
}

// ---- unitig assembly on the device --------------------------------------------------------------------------------------
// Every chain (head h, members ranked 0 .. len-1 along the merge links) is one unitig: its sequence is the head's k-mer
// followed by the last base of every later member (mergeLabels: a + b[k-1:]), its KC the sum of the members' coverages plus
// the last member's coverage * (k - 1), its reverse-complement partner the tail's partner.
// flags: bit 0 palindromic k-mer among the neighbours, bit 1 start-to-own-start link, bit 2 cyclic chain.
__global__ void __launch_bounds__(256) k_chain_stats(uint32_t n2, const int32_t* __restrict__ head, const int32_t* __restrict__ rank,
                                                    const uint8_t* __restrict__ deg, const int32_t* __restrict__ nbr, const uint8_t* __restrict__ nbr_pal,
                                                    const int32_t* __restrict__ counts, uint32_t* __restrict__ len, unsigned long long* __restrict__ kcsum,
                                                    unsigned int* __restrict__ flags) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n2) return;
    if (nbr_pal[v]) atomicOr(flags, 1u);
#pragma unroll
    for (int s = 0; s < 4; s++) {   // a link to myself or to my own reverse complement (hairpin): the reference's outcome depends on its sweep order
        const int32_t o = nbr[4 * v + s];
        if (o == (int32_t)v || o == (int32_t)(v ^ 1u)) atomicOr(flags, 2u);
    }
    const int32_t h = head[v];
    if (h < 0) { atomicOr(flags, 4u); return; }
    atomicMax(&len[h], (uint32_t)rank[v] + 1u);
    atomicAdd(&kcsum[h], (unsigned long long)(uint32_t)counts[v >> 1]);
}
// single CTA: exclusive prefix of the unitig lengths (k + len - 1 for heads, 0 otherwise) -> off[], total -> *total, heads -> *n_heads
__global__ void __launch_bounds__(1024) k_unitig_offsets(uint32_t n2, int k, const int32_t* __restrict__ head, const uint32_t* __restrict__ len,
                                                        uint32_t* __restrict__ off, uint32_t* __restrict__ hidx, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_a[1024], s_b[1024];
    const uint32_t per = (n2 + 1023) / 1024;
    const uint32_t lo = min(threadIdx.x * per, n2), hi = min(lo + per, n2);
    uint32_t sum = 0, cnt = 0;
    for (uint32_t v = lo; v < hi; v++)
        if (head[v] == (int32_t)v) { sum += (uint32_t)k + len[v] - 1; cnt++; }
    s_a[threadIdx.x] = sum;
    s_b[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t va = threadIdx.x >= (uint32_t)o ? s_a[threadIdx.x - o] : 0, vb = threadIdx.x >= (uint32_t)o ? s_b[threadIdx.x - o] : 0;
        __syncthreads();
        s_a[threadIdx.x] += va;
        s_b[threadIdx.x] += vb;
        __syncthreads();
    }
    uint32_t run = s_a[threadIdx.x] - sum, idx = s_b[threadIdx.x] - cnt;
    for (uint32_t v = lo; v < hi; v++) {
        if (head[v] == (int32_t)v) {
            off[v] = run;
            hidx[v] = idx++;
            run += (uint32_t)k + len[v] - 1;
        }
    }
    if (threadIdx.x == 1023) { totals[0] = s_a[1023]; totals[1] = s_b[1023]; }
}
struct McUnitigRec { int32_t node, rc; uint32_t off, len; long long kc; };
__global__ void __launch_bounds__(256) k_unitig_fill(uint32_t n2, int k, const uint64_t* __restrict__ kmers, const int32_t* __restrict__ counts,
                                                    const int32_t* __restrict__ head, const int32_t* __restrict__ rank, const uint32_t* __restrict__ len,
                                                    const uint32_t* __restrict__ off, const uint32_t* __restrict__ hidx,
                                                    const unsigned long long* __restrict__ kcsum, char* __restrict__ chars, McUnitigRec* __restrict__ recs) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n2) return;
    const int32_t h = head[v];
    if (h < 0) return;
    uint64_t code = kmers[v >> 1];
    if (v & 1) code = mc_revcomp(code, k);
    const uint32_t r = (uint32_t)rank[v], o = off[h], L = len[h];
    const char NUCS[4] = {'A', 'G', 'C', 'T'};
    if (r == 0) {
        for (int i = 0; i < k; i++) chars[o + i] = NUCS[(code >> (2 * (k - 1 - i))) & 3];
    } else {
        chars[o + (uint32_t)k - 1 + r] = NUCS[code & 3];
    }
    if (r == L - 1) {   // the tail: closes the record of its chain
        McUnitigRec rec;
        rec.node = h;
        rec.rc = (int32_t)(v ^ 1u);
        rec.off = o;
        rec.len = (uint32_t)k + L - 1;
        rec.kc = (long long)kcsum[h] + (long long)counts[v >> 1] * (k - 1);
        recs[hidx[h]] = rec;
    }
}

// extendEnvironment (src/algo/OneSequenceCalculator.java:265-295) changes nothing in the environment (`cont = kmer`, :276) but logs how
// many of its k-mers have exactly one solid neighbour OUTSIDE the environment: 8 lanes per environment k-mer (allNeighbors: left
// and right neighbours of the k-mer as printed), membership through the node hash (it holds both strands), coverage through the
// count table.
__global__ void __launch_bounds__(256) k_extend_count(NodeHash h, const uint64_t* __restrict__ kmers, uint32_t n, int k, McTableDev table, int min_cov,
                                                     unsigned long long* __restrict__ out) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = gt >> 3, s = gt & 7;
    const int lane = threadIdx.x & 31;
    bool outside_solid = false;
    if (i < n) {
        const uint64_t kmask = (1ULL << (2 * k)) - 1, kmer = kmers[i], nuc = s >> 1;
        const uint64_t x = (s & 1) ? (((kmer << 2) | nuc) & kmask) : ((nuc << (2 * k - 2)) | (kmer >> 2));
        if (node_lookup(h, x) < 0 && node_lookup(h, mc_revcomp(x, k)) < 0) outside_solid = mc_table_lookup(table, mc_mix62(mc_canonical(x, k))) >= min_cov;
    }
    const unsigned m = (__ballot_sync(0xffffffffu, outside_solid) >> (lane & ~7)) & 0xFFu;
    if (i < n && s == 0 && __popc(m) == 1) atomicAdd(out, 1ULL);
}

void mc_unitig_nodes(mc_ctx* ctx, const McUnitigInput& in, McUnitigNodes& out) {
    MC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)in.kmers.size(), n2 = 2 * n;
    out.deg.assign(n2, 0);
    out.nbr.assign((size_t)4 * n2, -1);
    out.nbr_pal.assign(n2 + 4, 0);
    out.gene.assign(n2, 0);
    out.head.assign(n2, -1);
    out.rank.assign(n2, 0);
    if (n == 0) return;
    uint32_t hcap = 1024;
    while (hcap < (uint64_t)n2 * 2) hcap <<= 1;
    const uint32_t nw = (uint32_t)in.gene_windows.size();
    const bool multi = !in.colors.empty();
    size_t bytes = 0;
    auto take = [&](size_t sz) { size_t o = bytes; bytes += (sz + 255) / 256 * 256; return o; };
    size_t o_kmers = take((size_t)n * 8), o_hk = take((size_t)hcap * 8), o_hv = take((size_t)hcap * 4), o_win = take((size_t)nw * 8 + 8),
           o_col = take(multi ? (size_t)n * 8 : 8), o_deg = take(n2), o_nbr = take((size_t)n2 * 16), o_pal = take(n2 + 4), o_gene = take(n2),
           o_pred = take((size_t)n2 * 4), o_j0 = take((size_t)n2 * 4), o_j1 = take((size_t)n2 * 4), o_r0 = take((size_t)n2 * 4),
           o_r1 = take((size_t)n2 * 4), o_head = take((size_t)n2 * 4), o_rank = take((size_t)n2 * 4),
           o_cnt = take((size_t)n * 4), o_len = take((size_t)n2 * 4), o_kc = take((size_t)n2 * 8), o_off = take((size_t)n2 * 4), o_hidx = take((size_t)n2 * 4),
           o_flags = take(16), o_chars = take((size_t)n2 * (size_t)(in.k + 1) + 64), o_recs = take((size_t)n2 * sizeof(McUnitigRec));
    uint8_t* d = (uint8_t*)mc_dev_alloc(ctx, bytes, "unitig work memory");   // the context's arena: no driver call per environment
    struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } free_d{ctx, d};
    {
        MC_CUDA(cudaMemcpyAsync(d + o_kmers, in.kmers.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        if (nw) MC_CUDA(cudaMemcpyAsync(d + o_win, in.gene_windows.data(), (size_t)nw * 8, cudaMemcpyHostToDevice, st));
        if (multi) MC_CUDA(cudaMemcpyAsync(d + o_col, in.colors.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        MC_CUDA(cudaMemsetAsync(d + o_hk, 0, (size_t)hcap * 8, st));
        MC_CUDA(cudaMemsetAsync(d + o_nbr, 0xFF, (size_t)n2 * 16, st));
        MC_CUDA(cudaMemsetAsync(d + o_pal, 0, n2 + 4, st));
        NodeHash h{(uint64_t*)(d + o_hk), (int32_t*)(d + o_hv), hcap - 1};
        const unsigned g1 = (n2 + 255) / 256, g4 = (unsigned)(((uint64_t)n2 * 4 + 255) / 256);
        k_node_insert<<<g1, 256, 0, st>>>(h, (const uint64_t*)(d + o_kmers), n, in.k);
        k_neighbors<<<g4, 256, 0, st>>>(h, (const uint64_t*)(d + o_kmers), n, in.k, (const uint64_t*)(d + o_win), nw, d + o_deg,
                                        (int32_t*)(d + o_nbr), d + o_pal, d + o_gene, in.order_by_nucleotide ? 1 : 0);
        k_links<<<g1, 256, 0, st>>>(n2, d + o_deg, (const int32_t*)(d + o_nbr), d + o_gene, multi ? (const uint64_t*)(d + o_col) : nullptr,
                                    (int32_t*)(d + o_pred), (int32_t*)(d + o_j0), (uint32_t*)(d + o_r0));
        ctx->launches += 3;
        out.extend_count = -1;
        if (in.table && !multi) {
            unsigned long long* d_ext = (unsigned long long*)(d + o_flags) + 1;   // (the flag block is zeroed again before the assembly uses it)
            MC_CUDA(cudaMemsetAsync(d_ext, 0, sizeof(unsigned long long), st));
            k_extend_count<<<(unsigned)(((uint64_t)n * 8 + 255) / 256), 256, 0, st>>>(h, (const uint64_t*)(d + o_kmers), n, in.k, in.table->dev, in.min_cov, d_ext);
            MC_LAUNCH(ctx);
            unsigned long long ext = 0;
            MC_CUDA(cudaMemcpyAsync(&ext, d_ext, sizeof ext, cudaMemcpyDeviceToHost, st));
            MC_CUDA(cudaStreamSynchronize(st));
            out.extend_count = (long long)ext;
        }
        int rounds = 1;
        while ((1ull << rounds) < (uint64_t)n2 + 1) rounds++;
        rounds += 1;
        int32_t* j[2] = {(int32_t*)(d + o_j0), (int32_t*)(d + o_j1)};
        uint32_t* r[2] = {(uint32_t*)(d + o_r0), (uint32_t*)(d + o_r1)};
        int cur = 0;
        for (int it = 0; it < rounds; it++, cur ^= 1) {
            k_jump<<<g1, 256, 0, st>>>(n2, j[cur], r[cur], j[cur ^ 1], r[cur ^ 1]);
            MC_LAUNCH(ctx);
        }
        k_finish<<<g1, 256, 0, st>>>(n2, (const int32_t*)(d + o_pred), j[cur], r[cur], (int32_t*)(d + o_head), (int32_t*)(d + o_rank));
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        // ---- unitig assembly (single environments; the multi tool keeps per-colour paths on the host) ----
        out.regular = false;
        const bool assemble = !multi && in.counts.size() == (size_t)n;
        uint32_t h_tot[4] = {0, 0, 0, 0};
        if (assemble) {
            MC_CUDA(cudaMemcpyAsync(d + o_cnt, in.counts.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
            MC_CUDA(cudaMemsetAsync(d + o_len, 0, (size_t)n2 * 4, st));
            MC_CUDA(cudaMemsetAsync(d + o_kc, 0, (size_t)n2 * 8, st));
            MC_CUDA(cudaMemsetAsync(d + o_flags, 0, 16, st));
            k_chain_stats<<<g1, 256, 0, st>>>(n2, (const int32_t*)(d + o_head), (const int32_t*)(d + o_rank), d + o_deg, (const int32_t*)(d + o_nbr), d + o_pal,
                                              (const int32_t*)(d + o_cnt), (uint32_t*)(d + o_len), (unsigned long long*)(d + o_kc), (unsigned int*)(d + o_flags));
            k_unitig_offsets<<<1, 1024, 0, st>>>(n2, in.k, (const int32_t*)(d + o_head), (const uint32_t*)(d + o_len), (uint32_t*)(d + o_off),
                                                 (uint32_t*)(d + o_hidx), (uint32_t*)(d + o_flags) + 1);
            k_unitig_fill<<<g1, 256, 0, st>>>(n2, in.k, (const uint64_t*)(d + o_kmers), (const int32_t*)(d + o_cnt), (const int32_t*)(d + o_head),
                                              (const int32_t*)(d + o_rank), (const uint32_t*)(d + o_len), (const uint32_t*)(d + o_off),
                                              (const uint32_t*)(d + o_hidx), (const unsigned long long*)(d + o_kc), (char*)(d + o_chars), (McUnitigRec*)(d + o_recs));
            ctx->launches += 3;
            MC_CUDA(cudaGetLastError());
            MC_CUDA(cudaMemcpyAsync(h_tot, d + o_flags, 12, cudaMemcpyDeviceToHost, st));   // flags, total chars, unitigs
            MC_CUDA(cudaStreamSynchronize(st));
            out.regular = h_tot[0] == 0;
        }
        if (out.regular) {
            const uint32_t n_chars = h_tot[1], n_u = h_tot[2];
            std::vector<McUnitigRec> recs(n_u);
            out.u_chars.resize(n_chars);
            if (n_u) MC_CUDA(cudaMemcpyAsync(recs.data(), d + o_recs, (size_t)n_u * sizeof(McUnitigRec), cudaMemcpyDeviceToHost, st));
            if (n_chars) MC_CUDA(cudaMemcpyAsync(&out.u_chars[0], d + o_chars, n_chars, cudaMemcpyDeviceToHost, st));
            MC_CUDA(cudaMemcpyAsync(out.nbr.data(), d + o_nbr, (size_t)n2 * 16, cudaMemcpyDeviceToHost, st));
            MC_CUDA(cudaMemcpyAsync(out.gene.data(), d + o_gene, n2, cudaMemcpyDeviceToHost, st));
            MC_CUDA(cudaStreamSynchronize(st));
            out.u_node.resize(n_u); out.u_rc.resize(n_u); out.u_off.resize(n_u); out.u_len.resize(n_u); out.u_kc.resize(n_u);
            for (uint32_t i = 0; i < n_u; i++) {
                out.u_node[i] = recs[i].node; out.u_rc[i] = recs[i].rc; out.u_off[i] = recs[i].off; out.u_len[i] = recs[i].len; out.u_kc[i] = recs[i].kc;
            }
            return;   // (the guard object returns the work memory to the arena)
        }
        MC_CUDA(cudaMemcpyAsync(out.deg.data(), d + o_deg, n2, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(out.nbr.data(), d + o_nbr, (size_t)n2 * 16, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(out.nbr_pal.data(), d + o_pal, n2, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(out.gene.data(), d + o_gene, n2, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(out.head.data(), d + o_head, (size_t)n2 * 4, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(out.rank.data(), d + o_rank, (size_t)n2 * 4, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
    }
}
