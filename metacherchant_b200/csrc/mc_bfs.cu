// K5: the BFS that grows an environment from the target's k-mers.
//
// Restates OneSequenceCalculator.runBfs (src/algo/OneSequenceCalculator.java:154-220), getNeighborsByDir (:222-239),
// TerminationMode.allowsAddition (src/algo/TerminationMode.java:31-47) and StringUtils.{left,right,all}Neighbors
// (src/utils/StringUtils.java:8-32) as a batch-synchronous frontier expansion, one persistent CTA per (target, pass):
//
//   * the queue holds ORIENTED k-mers (2-bit); the visited set is keyed by the oriented k-mer
//     (distanceToKmer is a HashMap<String,..> over oriented strings, :156,165,207);
//   * a CTA step takes the next P queue entries, generates their 4 (or 8) neighbours in the reference's order
//     A,G,C,T (dir 0: L[A],R[A],L[G],R[G],...), probes the count table with the CANONICAL key (getKmerKey :89-96),
//     drops those below --coverage or beyond --maxradius, and orders the survivors by (parent queue index, slot);
//   * first occurrence wins (atomicMin of the candidate rank on the visited slot), a block scan gives every winner its
//     queue position, and winners are accepted while |visited| < --maxkmers.
// This is the sequential FIFO rule, batch by batch: same visited set AND same discovery order, also when --maxkmers
// truncates (the equivalence is simulated against the literal loop in tests/, test_level_synchronous_rule_equals_fifo_rule).
// Once a candidate is refused for the budget nothing can ever be added again (|visited| never shrinks), so the pass ends.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mc_common.cuh"
#include "mc_env.hpp"

#ifndef BFS_THREADS
#define BFS_THREADS 256
#endif
#define BFS_OVERFLOW 0xFFFFFFFFu

struct McBfsJobDev {
    const uint8_t* seq;    // target as nucleotide codes (A0 G1 C2 T3), one byte per base
    uint32_t seq_len;
    int dir;               // -1 left, +1 right, 0 both
    uint64_t* queue;       // oriented k-mers in discovery order
    int32_t* dist;
    int16_t* count;
    uint64_t* vis_keys;    // open addressing: oriented k-mer + 1 (0 = empty)
    uint32_t* vis_val;     // < 2^31: queue index; >= 2^31: claim of candidate (val & 0x7FFFFFFF) in the running step; ~0: untouched
    uint32_t cap;          // queue capacity
    uint32_t vis_mask;
    uint32_t* n_out;       // visited count, or BFS_OVERFLOW when cap was too small
    uint8_t* last;         // --trim: 1 = "lastKmers" member (some solid neighbour was refused by allowsAddition), else null
};
#define BFS_SEPARATOR 4    // code between the sequences of a merged seed list: no seed window may span it

// home slot of an oriented k-mer in the pass's visited set (Fibonacci hashing: one multiply on the step's critical path)
__device__ __forceinline__ uint32_t vis_home(const McBfsJobDev& j, uint64_t kmer) {
    return (uint32_t)((kmer * 0x9E3779B97F4A7C15ULL) >> 32) & j.vis_mask;
}
__device__ __forceinline__ uint32_t vis_find_or_claim(const McBfsJobDev& j, uint64_t kmer, uint32_t code) {
    const uint64_t want = kmer + 1;
    uint32_t pos = vis_home(j, kmer);
    while (true) {
        uint64_t cur = atomicCAS((unsigned long long*)&j.vis_keys[pos], 0ULL, (unsigned long long)want);
        if (cur == 0 || cur == want) {
            atomicMin(&j.vis_val[pos], code);
            return pos;
        }
        pos = (pos + 1) & j.vis_mask;   // table is at most half full: terminates
    }
}

// block-wide exclusive scan of a 0/1 flag in thread order; returns my offset, total in *total
__device__ __forceinline__ uint32_t block_scan_flag(bool flag, uint32_t* total) {
    __shared__ uint32_t s_w[BFS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < BFS_THREADS / 32; i++) {
        uint32_t c = s_w[i];
        if (i < warp) off += c;
        tot += c;
    }
    __syncthreads();
    *total = tot;
    return off + __popc(m & ((1u << lane) - 1));
}

__global__ void __launch_bounds__(BFS_THREADS) k_bfs(McTableDev table, uint32_t shard_id, int k, int min_cov, long long max_kmers, long long max_radius,
                                                    const McBfsJobDev* __restrict__ jobs) {
    const McBfsJobDev j = jobs[blockIdx.x];
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    // MC_TRACE: SM cycles and wall nanoseconds of this pass (their ratio is the clock the SM really ran at)
    long long t_clk = 0;
    unsigned long long t_ns = 0;
    if (threadIdx.x == 0) {
        t_clk = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ns));
    }
    __shared__ uint32_t s_head, s_tail, s_stop;
    if (threadIdx.x == 0) { s_head = 0; s_tail = 0; s_stop = 0; }
    __syncthreads();
    (void)shard_id;

    // ---- seeds: every window of the target whose count >= min_cov, in order, first occurrence only (:160-167) ----
    const uint32_t n_win = j.seq_len >= (uint32_t)k ? j.seq_len - k + 1 : 0;
    for (uint32_t base = 0; base < n_win; base += BFS_THREADS) {
        uint32_t i = base + threadIdx.x;
        bool cand = false;
        uint64_t kmer = 0;
        int c = -1;
        uint32_t slot = 0;
        if (i < n_win) {
            bool whole = true;
            for (int t = 0; t < k; t++) {
                whole = whole && j.seq[i + t] < BFS_SEPARATOR;
                kmer = (kmer << 2) | (uint64_t)(j.seq[i + t] & 3);
            }
            if (whole) {
                c = mc_table_lookup(table, mc_mix62(mc_canonical(kmer, k)));
                cand = c >= min_cov;
            }
        }
        if (cand) slot = vis_find_or_claim(j, kmer, 0x80000000u | threadIdx.x);
        __syncthreads();
        bool win = cand && __ldcg(&j.vis_val[slot]) == (0x80000000u | threadIdx.x);
        uint32_t total;
        uint32_t pos = block_scan_flag(win, &total);
        uint32_t tail = s_tail;
        if (tail + total > j.cap) {
            if (threadIdx.x == 0) *j.n_out = BFS_OVERFLOW;
            return;
        }
        if (win) {
            j.queue[tail + pos] = kmer;
            j.dist[tail + pos] = 0;
            j.count[tail + pos] = (int16_t)c;
            j.vis_val[slot] = tail + pos;
        }
        __syncthreads();
        // A k-mer that occurs twice among the seed windows is queued twice by the reference (:160-167); when it is popped the
        // second time all its solid neighbours are refused, which makes it a lastKmers member: remember the duplicate (bit 1).
        if (j.last && cand && !win) j.last[__ldcg(&j.vis_val[slot])] = 2;
        if (threadIdx.x == 0) s_tail = tail + total;
        __syncthreads();
    }

    // ---- queue loop (:197-213) ----------------------------------------------------------------------------------
    const int NB = j.dir == 0 ? 8 : 4;
    const uint32_t PARENTS = BFS_THREADS / NB;
    uint32_t steps = 0;
    while (true) {
        steps++;
        const uint32_t head = s_head, tail = s_tail;
        // --trim (runBfs :197-213 keeps popping after the budget is spent, only to collect lastKmers): once nothing can be
        // added any more the loop goes on "frozen" over the rest of the queue, looking only at which parents have solid neighbours
        const bool frozen = s_stop != 0;
        if (head >= tail || (frozen && !j.last)) break;
        // ---- linear stretch: the queue holds ONE unprocessed k-mer.  The general step below costs five block barriers and ~130
        // instructions per warp for that one parent; here the first warp walks the chain alone -- same candidates, same order (slot
        // = lane), same visited set, same budget and radius rules -- until the frontier widens, empties or the pass stops.
        if (tail - head == 1 && !frozen && !j.last) {
            if (threadIdx.x < 32) {
                const uint32_t lane = threadIdx.x;
                uint32_t h = head, tl = tail;
                uint64_t kmer = j.queue[h];
                int32_t dist = j.dist[h];
                bool stop = false, overflow = false;
                while (true) {
                    const int32_t nd = dist + 1;
                    const bool within = max_radius < 0 || nd <= max_radius;
                    bool cand = false;
                    uint64_t nb = 0xFFFFFFFFFFFFFF00ULL | lane;   // distinct per lane: never matches a k-mer or another lane
                    int c = -1;
                    uint32_t vpos = 0;
                    uint64_t vk = 0;
                    if (lane < (uint32_t)NB && within) {
                        const uint64_t nuc = (NB == 8) ? (lane >> 1) : lane;
                        const bool left = (j.dir == -1) || (j.dir == 0 && (lane & 1) == 0);
                        const uint64_t x = left ? ((nuc << (2 * k - 2)) | (kmer >> 2)) : (((kmer << 2) | nuc) & kmask);
                        // The visited-set slot of the neighbour is read BEFORE the table answers, so the two global round trips of a
                        // step overlap instead of following each other (3 700 -> ~2 000 cycles per step of a linear stretch).  A plain
                        // load is enough: while the walk lasts this warp is the only writer of the pass's visited set, and its own
                        // earlier stores are ordered before this load by the __syncwarp() that ends every step.
                        vpos = vis_home(j, x);
                        vk = __ldcg(&j.vis_keys[vpos]);
                        c = mc_table_lookup(table, mc_mix62(mc_canonical(x, k)));
                        if (c >= min_cov) { cand = true; nb = x; }
                    }
                    // (Letting the idle lanes prefetch the home buckets of the neighbours' neighbours into L2 was measured twice, in
                    // round 1 and again on this loop: 2 971 against 2 777 cycles per step -- the extra divergent path costs more issue
                    // time on the critical path than the L2 hit saves.)
                    // first occurrence among the neighbours of this parent (left and right neighbours can only coincide when both
                    // directions are expanded), then "not visited before"; vis_val gets the queue index so that the general step sees
                    // the k-mer as visited.
                    const unsigned same = NB == 8 ? __match_any_sync(0xffffffffu, nb) : (1u << lane);
                    bool win = cand && (uint32_t)(__ffs(same) - 1) == lane;
                    const unsigned win_mask = __ballot_sync(0xffffffffu, win);
                    if (win) {
                        const uint64_t want = nb + 1;
                        // two candidates of this step whose home slots coincide cannot both take it blindly: the first one may, the
                        // others walk on from the NEXT slot (the first one's plain store and their CAS must not meet on one slot)
                        const unsigned clash = __match_any_sync(win_mask, vpos);
                        const bool first_here = (uint32_t)(__ffs(clash) - 1) == lane;
                        if (vk == want) {
                            win = false;                                   // visited before
                        } else if (vk == 0 && first_here) {
                            j.vis_keys[vpos] = want;                       // new: claimed with a plain store (no other writer, see above)
                        } else {                                           // occupied by another k-mer, or contested: walk on with the CAS protocol
                            if (vk == 0) vpos = (vpos + 1) & j.vis_mask;   // (contested and free: the first candidate is storing there right now)
                            while (true) {
                                const uint64_t cur = atomicCAS((unsigned long long*)&j.vis_keys[vpos], 0ULL, (unsigned long long)want);
                                if (cur == 0) break;
                                if (cur == want) { win = false; break; }
                                vpos = (vpos + 1) & j.vis_mask;
                            }
                        }
                    }
                    const unsigned winners = __ballot_sync(0xffffffffu, win);
                    const uint32_t total = __popc(winners), pos = __popc(winners & ((1u << lane) - 1u));
                    uint32_t accept = total;
                    if (max_kmers >= 0) {
                        long long room = max_kmers - (long long)tl;
                        if (room < 0) room = 0;
                        if ((long long)total > room) { accept = (uint32_t)room; stop = true; }
                    }
                    if (tl + accept > j.cap) { overflow = true; break; }
                    const bool added = win && pos < accept;
                    if (added) {
                        j.queue[tl + pos] = nb;
                        j.dist[tl + pos] = nd;
                        j.count[tl + pos] = (int16_t)c;
                        j.vis_val[vpos] = tl + pos;
                    } else if (win) {
                        j.vis_val[vpos] = 0x80000000u | lane;   // claimed but refused by the budget: what the general step leaves behind
                    }
                    if (max_radius >= 0 && (long long)dist + 1 > max_radius) stop = true;
                    h += 1;
                    tl += accept;
                    steps++;
                    if (stop || tl - h != 1) break;
                    __syncwarp();   // orders this step's visited-set stores before the next step's loads (any lane)
                    // exactly one k-mer was appended and it is the whole frontier: it is the next parent
                    const int src = __ffs(__ballot_sync(0xffffffffu, added)) - 1;
                    kmer = __shfl_sync(0xffffffffu, nb, src);
                    dist = nd;
                }
                if (lane == 0) {
                    s_head = h;
                    s_tail = tl;
                    if (stop) s_stop = 1;
                    if (overflow) s_stop = 2;
                }
            }
            __syncthreads();
            if (s_stop == 2) {
                if (threadIdx.x == 0) *j.n_out = BFS_OVERFLOW;
                return;
            }
            continue;
        }
        const uint32_t P = min(tail - head, PARENTS);
        const uint32_t t = threadIdx.x;
        bool cand = false, solid = false;
        uint64_t nb = 0;
        int c = -1;
        int32_t nd = 0;
        uint32_t slot = 0, pi = 0;
        if (t < P * NB) {
            pi = head + t / NB;
            const uint32_t s = t % NB;
            const uint64_t kmer = j.queue[pi];
            nd = j.dist[pi] + 1;
            const uint64_t nuc = (NB == 8) ? (s >> 1) : s;
            const bool left = (j.dir == -1) || (j.dir == 0 && (s & 1) == 0);
            nb = left ? ((nuc << (2 * k - 2)) | (kmer >> 2)) : (((kmer << 2) | nuc) & kmask);
            const bool within = max_radius < 0 || nd <= max_radius;
            if (within || j.last) {
                c = mc_table_lookup(table, mc_mix62(mc_canonical(nb, k)));
                solid = c >= min_cov;
                cand = solid && within && !frozen;
            }
        }
        if (cand) slot = vis_find_or_claim(j, nb, 0x80000000u | t);
        __syncthreads();
        bool win = cand && __ldcg(&j.vis_val[slot]) == (0x80000000u | t);
        uint32_t total;
        uint32_t pos = block_scan_flag(win, &total);
        uint32_t accept = total;
        bool stop = false;
        if (max_kmers >= 0) {
            long long room = max_kmers - (long long)tail;   // allowsAddition: refuse when distanceToKmer.size() >= threshold
            if (room < 0) room = 0;
            if ((long long)total > room) { accept = (uint32_t)room; stop = true; }
        }
        if (tail + accept > j.cap) {
            if (threadIdx.x == 0) *j.n_out = BFS_OVERFLOW;
            return;
        }
        const bool added = win && pos < accept;
        if (added) {
            j.queue[tail + pos] = nb;
            j.dist[tail + pos] = nd;
            j.count[tail + pos] = (int16_t)c;
            j.vis_val[slot] = tail + pos;
        }
        // allowsAddition said no (already visited, budget or radius) -> lastKmers.add(kmer); a duplicated seed is popped again later
        // and then every solid neighbour is refused
        if (j.last && solid && (!added || (j.last[pi] & 2))) j.last[pi] |= 1;
        // all remaining parents are at least as far: once a whole step is beyond the radius the pass is over
        bool radius_done = max_radius >= 0 && (long long)j.dist[head] + 1 > max_radius;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_head = head + P;
            s_tail = tail + accept;
            if (stop || radius_done) s_stop = 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        j.n_out[0] = s_tail;
        j.n_out[1] = steps;
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        *(unsigned long long*)(j.n_out + 2) = (unsigned long long)(clock64() - t_clk);
        *(unsigned long long*)(j.n_out + 4) = ns1 - t_ns;
    }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
std::vector<uint8_t> mc_codes_from_target(const std::string& seq) {
    // DnaQ(str, 0) (ITMO!/dna/DnaQ.java:21-30): N, n, '.' become nucleotide 0; other chars go through DnaTools.fromChar
    std::vector<uint8_t> codes(seq.size());
    for (size_t i = 0; i < seq.size(); i++) {
        char ch = seq[i];
        if (ch == 'N' || ch == 'n' || ch == '.') { codes[i] = 0; continue; }
        int c = mc_nuc_code((uint8_t)ch);
        if (c >= 4) mc_throw(MC_ERR_FORMAT, "Incorrect nucleotide char: \"%c\" in a target sequence", ch);
        codes[i] = (uint8_t)c;
    }
    return codes;
}

static uint32_t pow2_at_least(uint64_t x) {
    uint32_t p = 1024;
    while (p < x) p <<= 1;
    return p;
}

void mc_run_bfs_passes(mc_ctx* ctx, const mc_table* t, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                       const std::vector<McBfsRequest>& reqs, std::vector<McBfsResult>& results) {
    if (max_kmers < 0 && max_radius < 0) mc_throw(MC_ERR_ARG, "At least one of --maxkmers and --maxradius parameters should be set");
    if (k != t->k) mc_throw(MC_ERR_ARG, "k (%d) differs from the table's k (%d)", k, t->k);
    // a shard only holds the k-mers it owns: every other neighbour would read as absent.  The environment stage runs on a
    // complete (gathered) solid table (all-gather of the solid set, then targets split over the ranks).
    if (t->dev.shard_bits != 0) mc_throw(MC_ERR_UNSUPPORTED, "the environment stage needs a complete table, not one shard of %d (gather the solid set first)", 1 << t->dev.shard_bits);
    MC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    mc_table_flush(const_cast<mc_table*>(t));   // pending binned updates first
    const size_t n = reqs.size();
    results.assign(n, McBfsResult{});
    std::vector<uint32_t> caps(n);
    std::vector<size_t> todo(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t seeds = reqs[i].codes.size() >= (size_t)k ? reqs[i].codes.size() - k + 1 : 0;
        uint64_t cap = max_kmers >= 0 ? seeds + (uint64_t)max_kmers + 8 : std::max<uint64_t>(seeds * 2 + 1024, 1u << 16);
        if (cap > 0x3FFFFFFFull) mc_throw(MC_ERR_ARG, "--maxkmers too large for the GPU path");
        caps[i] = (uint32_t)cap;
        todo[i] = i;
    }
    while (!todo.empty()) {
        // one allocation for the whole wave
        std::vector<McBfsJobDev> jobs(todo.size());
        size_t bytes = 0;
        auto take = [&](size_t sz) { size_t o = bytes; bytes += (sz + 255) / 256 * 256; return o; };
        struct Off { size_t seq, queue, dist, count, vk, vv, nout, last; uint32_t vis_cap; };
        std::vector<Off> offs(todo.size());
        for (size_t w = 0; w < todo.size(); w++) {
            size_t i = todo[w];
            uint32_t cap = caps[i];
            uint32_t vis_cap = pow2_at_least((uint64_t)cap * 2);
            offs[w] = Off{take(reqs[i].codes.size() + 1), take((size_t)cap * 8), take((size_t)cap * 4), take((size_t)cap * 2),
                          take((size_t)vis_cap * 8), take((size_t)vis_cap * 4), take(32), take(reqs[i].trim ? (size_t)cap : 1), vis_cap};
        }
        size_t jobs_off = take(jobs.size() * sizeof(McBfsJobDev));
        uint8_t* d = (uint8_t*)mc_dev_alloc(ctx, bytes, "BFS work memory");   // from the context's arena: reused call after call
        struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } free_d{ctx, d};
        for (size_t w = 0; w < todo.size(); w++) {
            size_t i = todo[w];
            const Off& o = offs[w];
            if (!reqs[i].codes.empty())
                MC_CUDA(cudaMemcpyAsync(d + o.seq, reqs[i].codes.data(), reqs[i].codes.size(), cudaMemcpyHostToDevice, st));
            MC_CUDA(cudaMemsetAsync(d + o.vk, 0, (size_t)o.vis_cap * 8, st));
            MC_CUDA(cudaMemsetAsync(d + o.vv, 0xFF, (size_t)o.vis_cap * 4, st));
            MC_CUDA(cudaMemsetAsync(d + o.nout, 0, 32, st));   // [0] visited count, [1] steps of the queue loop, [2..3] cycles, [4..5] ns (MC_TRACE)
            if (reqs[i].trim) MC_CUDA(cudaMemsetAsync(d + o.last, 0, caps[i], st));
            jobs[w] = McBfsJobDev{d + o.seq, (uint32_t)reqs[i].codes.size(), reqs[i].dir, (uint64_t*)(d + o.queue), (int32_t*)(d + o.dist),
                                  (int16_t*)(d + o.count), (uint64_t*)(d + o.vk), (uint32_t*)(d + o.vv), caps[i], o.vis_cap - 1,
                                  (uint32_t*)(d + o.nout), reqs[i].trim ? d + o.last : nullptr};
        }
        MC_CUDA(cudaMemcpyAsync(d + jobs_off, jobs.data(), jobs.size() * sizeof(McBfsJobDev), cudaMemcpyHostToDevice, st));
        cudaEvent_t ev_a = nullptr, ev_b = nullptr;
        const bool trace = getenv("MC_TRACE") != nullptr;   // device time of the kernel alone on stderr
        if (trace) {
            cudaEventCreate(&ev_a);
            cudaEventCreate(&ev_b);
            cudaEventRecord(ev_a, st);
        }
        k_bfs<<<(unsigned)jobs.size(), BFS_THREADS, 0, st>>>(t->dev, t->shard_id, k, min_cov, (long long)max_kmers, (long long)max_radius,
                                                            (const McBfsJobDev*)(d + jobs_off));
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        if (trace) {
            cudaEventRecord(ev_b, st);
            cudaEventSynchronize(ev_b);
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_a, ev_b);
            uint64_t steps = 0, visited = 0, cyc_most = 0, ns_most = 0;
            uint32_t most = 0;
            for (size_t w = 0; w < todo.size(); w++) {
                uint32_t v[8] = {0};
                MC_CUDA(cudaMemcpyAsync(v, d + offs[w].nout, 32, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaStreamSynchronize(st));
                if (v[0] != BFS_OVERFLOW) visited += v[0];
                steps += v[1];
                if (v[1] > most) {
                    most = v[1];
                    cyc_most = (uint64_t)v[2] | ((uint64_t)v[3] << 32);
                    ns_most = (uint64_t)v[4] | ((uint64_t)v[5] << 32);
                }
            }
            fprintf(stderr, "[mc trace] k_bfs: %zu passes, %.1f ms on the device, %.1f MiB of work memory; %llu k-mers visited in %llu steps (longest pass %u steps: %.2f us "
                            "= %.0f SM cycles per step, SM clock during the pass %.0f MHz)\n",
                    jobs.size(), ms, (double)bytes / (1 << 20), (unsigned long long)visited, (unsigned long long)steps, most, most ? 1e3 * ms / most : 0.0,
                    most ? (double)cyc_most / most : 0.0, ns_most ? 1e3 * (double)cyc_most / (double)ns_most : 0.0);
            cudaEventDestroy(ev_a);
            cudaEventDestroy(ev_b);
        }
        std::vector<uint32_t> n_out(todo.size());
        for (size_t w = 0; w < todo.size(); w++)
            MC_CUDA(cudaMemcpyAsync(&n_out[w], d + offs[w].nout, 4, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        std::vector<size_t> again;
        for (size_t w = 0; w < todo.size(); w++) {
            size_t i = todo[w];
            if (n_out[w] == BFS_OVERFLOW) {
                if (caps[i] > 0x1FFFFFFFu) mc_throw(MC_ERR_NOMEM, "environment larger than 2^29 k-mers");
                caps[i] *= 4;
                again.push_back(i);
                continue;
            }
            McBfsResult& r = results[i];
            r.kmers.resize(n_out[w]);
            r.dist.resize(n_out[w]);
            r.counts.resize(n_out[w]);
            if (n_out[w]) {
                MC_CUDA(cudaMemcpyAsync(r.kmers.data(), d + offs[w].queue, (size_t)n_out[w] * 8, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaMemcpyAsync(r.dist.data(), d + offs[w].dist, (size_t)n_out[w] * 4, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaMemcpyAsync(r.counts.data(), d + offs[w].count, (size_t)n_out[w] * 2, cudaMemcpyDeviceToHost, st));
                if (reqs[i].trim) {
                    r.last.resize(n_out[w]);
                    MC_CUDA(cudaMemcpyAsync(r.last.data(), d + offs[w].last, (size_t)n_out[w], cudaMemcpyDeviceToHost, st));
                }
            }
        }
        MC_CUDA(cudaStreamSynchronize(st));
        todo.swap(again);
    }
}

extern "C" int mc_bfs(mc_ctx* ctx, const mc_table* t, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int dir,
                      uint64_t* kmers, int32_t* dist, int16_t* counts, size_t cap, size_t* n) {
    try {
        if (!ctx || !t || !seq || !n) mc_throw(MC_ERR_ARG, "null argument");
        if (dir < -1 || dir > 1) mc_throw(MC_ERR_ARG, "dir must be -1, 0 or +1");
        std::vector<McBfsRequest> reqs(1);
        reqs[0].codes = mc_codes_from_target(seq);
        reqs[0].dir = dir;
        std::vector<McBfsResult> res;
        mc_run_bfs_passes(ctx, t, k, min_cov, max_kmers, max_radius, reqs, res);
        *n = res[0].kmers.size();
        if (*n > cap) mc_throw(MC_ERR_CAPACITY, "BFS visited %zu k-mers, room for %zu", *n, cap);
        for (size_t i = 0; i < *n; i++) {
            if (kmers) kmers[i] = res[0].kmers[i];
            if (dist) dist[i] = res[0].dist[i];
            if (counts) counts[i] = res[0].counts[i];
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
