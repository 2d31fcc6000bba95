// Super-k-mer front end of the count table: minimizer bins of 16-byte records, de-duplicated and unpacked at flush time.
//
// Why: the flat path moves every k-mer INSTANCE (8 bytes) through two scatter levels and the shared-memory aggregation: ~57 bytes of
// DRAM traffic per instance and, more to the point, three instruction-bound kernels per instance.  A read's consecutive k-mers mostly
// share their minimizer, so the extractor (mc_count.cu, k_extract<.., W>) appends one 16-byte record per super-k-mer (~9 k-mers at
// k = 31) to the bin of that minimizer.  Every copy of a canonical k-mer lands in the same bin whatever read or strand it came from,
// and identical stretches of the sequenced genomes give IDENTICAL records (the cut points depend on the sequence only, except at
// read ends): k_sk_expand first counts identical records of a bin in a shared-memory table and then unpacks every DISTINCT record
// once, with its multiplicity c as the weight of its k-mers.  Weight-1 k-mers (sequencing errors, read-end fragments) continue as
// flat hashed keys through the key bins of mc_bins.cu (a stream several times smaller than the instance stream); k-mers with
// c >= 2 are added to the table directly as (key, c) pairs.  add(key, c) saturating at 32767 commutes
// (BigLong2ShortHashMap.addAndBound, ITMO!/structures/map/BigLong2ShortHashMap.java:67-77), so the table is the same as with one
// update per instance; the parity tests compare it with the oracle's.
//
// Nothing here can fail for lack of room: a full bin diverts to the spill list, a full spill list makes the host withdraw the
// chunk (cursor snapshot) and redo it with flat keys; the unpacked lists are sized from the exact k-mer totals the bins carry.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mc_common.cuh"

#define SKX_THREADS 256
#define SKX_SLOTS 4096u       // shared-memory record table: 16 B record + 4 B multiplicity per slot
#define SKX_ROUND 2048u       // records de-duplicated at a time (load <= 0.5: the probe loop always ends)
#define SKX_PER (SKX_ROUND / SKX_THREADS)
static size_t skx_smem_bytes() { return (size_t)SKX_SLOTS * 20 + (size_t)SKX_ROUND * 6 + 64; }

struct McSkExpandArgs {
    McSkDev sk;
    uint32_t b_lo, n_units;       // bins [b_lo, b_lo + n_units)  --  or, from_spill: units of SKX_ROUND records of the spill list
    uint64_t spill_n;
    int from_spill;
    int k;
    uint64_t* singles;            // hashed keys of weight 1
    uint64_t* pair_keys;          // hashed keys of weight >= 2 ...
    uint16_t* pair_counts;        // ... and their weights (<= SKX_ROUND)
    unsigned long long* out;      // [0] singles written, [1] pairs written, [2] unit ticket
};

__global__ void __launch_bounds__(SKX_THREADS, 2) k_sk_expand(McSkExpandArgs a) {
    extern __shared__ unsigned long long x_mem[];
    unsigned long long* s_hi = x_mem;                       // record word with the length (never 0 for a record): the claim word
    unsigned long long* s_lo = x_mem + SKX_SLOTS;
    uint32_t* s_cnt = (uint32_t*)(s_lo + SKX_SLOTS);        // multiplicity; > 0 also says that s_lo is in place
    uint32_t* s_off = s_cnt + SKX_SLOTS;                    // [SKX_ROUND] output offset of every distinct record
    uint16_t* s_list = (uint16_t*)(s_off + SKX_ROUND);      // [SKX_ROUND] claimed slots in claim order
    __shared__ unsigned long long s_unit, s_base[2];
    __shared__ uint32_t s_used;
    __shared__ unsigned long long s_warp[SKX_THREADS / 32];
    const int k = a.k;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int s0 = 122 - 2 * k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < SKX_SLOTS; i += SKX_THREADS) {
        s_hi[i] = 0;
        s_lo[i] = 0;
        s_cnt[i] = 0;
    }
    if (threadIdx.x == 0) s_used = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_unit = atomicAdd(&a.out[2], 1ULL);
        __syncthreads();
        if (s_unit >= a.n_units) break;
        const uint32_t unit = (uint32_t)s_unit;
        const uint4* src;
        uint32_t n_rec;
        if (a.from_spill) {
            const uint64_t lo = (uint64_t)unit * SKX_ROUND;
            src = a.sk.spill + lo;
            n_rec = (uint32_t)(a.spill_n - lo < SKX_ROUND ? a.spill_n - lo : SKX_ROUND);
        } else {
            const uint32_t bin = a.b_lo + unit;
            const uint32_t attempts = (uint32_t)a.sk.cursors[bin];
            n_rec = attempts < a.sk.cap ? attempts : (uint32_t)a.sk.cap;
            src = a.sk.recs + (uint64_t)bin * a.sk.cap;
        }
        for (uint32_t r0 = 0; r0 < n_rec; r0 += SKX_ROUND) {
            const uint32_t rn = n_rec - r0 < SKX_ROUND ? n_rec - r0 : SKX_ROUND;
            // ---- A: identical records of the round are counted in the shared-memory table -------------------------------------
            for (uint32_t i = threadIdx.x; i < rn; i += SKX_THREADS) {
                const uint4 r = __ldcs(src + r0 + i);
                const unsigned long long lo = (unsigned long long)r.x | ((unsigned long long)r.y << 32);
                const unsigned long long hi = (unsigned long long)r.z | ((unsigned long long)r.w << 32);
                unsigned long long x = hi ^ (lo * 0x9E3779B97F4A7C15ULL);
                x ^= x >> 29;
                x *= 0xBF58476D1CE4E5B9ULL;
                uint32_t slot = (uint32_t)(x >> 40) & (SKX_SLOTS - 1);
                for (;;) {
                    unsigned long long cur = *(volatile unsigned long long*)&s_hi[slot];
                    if (cur == 0) {
                        cur = atomicCAS(&s_hi[slot], 0ULL, hi);
                        if (cur == 0) {   // mine: the low word, then the count that publishes it
                            *(volatile unsigned long long*)&s_lo[slot] = lo;
                            __threadfence_block();
                            atomicAdd(&s_cnt[slot], 1u);
                            s_list[atomicAdd(&s_used, 1u)] = (uint16_t)slot;
                            break;
                        }
                    }
                    // a slot whose low word is not published yet is passed over: the record then claims a slot of its own further on,
                    // which only splits the multiplicity of that record over two entries
                    if (cur == hi && *(volatile uint32_t*)&s_cnt[slot] != 0 && *(volatile unsigned long long*)&s_lo[slot] == lo) {
                        atomicAdd(&s_cnt[slot], 1u);
                        break;
                    }
                    slot = (slot + 1) & (SKX_SLOTS - 1);
                }
            }
            __syncthreads();
            // ---- B: output offsets (block scan over the distinct records), one reservation per list and round ---------------------
            const uint32_t used = s_used;
            unsigned long long mine = 0;   // k-mers of my entries: singles in the low word, pairs in the high word
            uint32_t nn[SKX_PER];
#pragma unroll
            for (int i = 0; i < (int)SKX_PER; i++) {
                const uint32_t e = threadIdx.x * SKX_PER + i;
                nn[i] = 0;
                if (e < used) {
                    const uint32_t slot = s_list[e];
                    const uint32_t n = (uint32_t)(s_hi[slot] >> 58);
                    nn[i] = s_cnt[slot] > 1 ? n | 0x80000000u : n;
                    mine += s_cnt[slot] > 1 ? (unsigned long long)n << 32 : (unsigned long long)n;
                }
            }
            unsigned long long incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            unsigned long long before = incl - mine, total = 0;
            for (int w = 0; w < SKX_THREADS / 32; w++) {
                if (w < warp) before += s_warp[w];
                total += s_warp[w];
            }
            if (threadIdx.x == 0) {
                const unsigned long long ts = total & 0xFFFFFFFFULL, tp = total >> 32;
                s_base[0] = ts ? atomicAdd(&a.out[0], ts) : 0ULL;
                s_base[1] = tp ? atomicAdd(&a.out[1], tp) : 0ULL;
            }
            {
                uint32_t os = (uint32_t)before, op = (uint32_t)(before >> 32);
#pragma unroll
                for (int i = 0; i < (int)SKX_PER; i++) {
                    const uint32_t e = threadIdx.x * SKX_PER + i;
                    if (e < used) {
                        const uint32_t n = nn[i] & 0x7FFFFFFFu;
                        if (nn[i] & 0x80000000u) { s_off[e] = op; op += n; }
                        else { s_off[e] = os; os += n; }
                    }
                }
            }
            __syncthreads();
            // ---- C: every distinct record is unpacked once; 16 lanes per record -------------------------------------------------
            {
                const uint32_t hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
                uint64_t* const singles = a.singles + s_base[0];
                uint64_t* const pkeys = a.pair_keys + s_base[1];
                uint16_t* const pcnt = a.pair_counts + s_base[1];
                for (uint32_t e = hw; e < used; e += SKX_THREADS / 16) {
                    const uint32_t slot = s_list[e];
                    const unsigned long long hi = s_hi[slot], lo = s_lo[slot];
                    const uint32_t c = s_cnt[slot], off = s_off[e];
                    const uint32_t n = (uint32_t)(hi >> 58);
                    const unsigned long long bh = hi & ((1ULL << 58) - 1);
                    for (uint32_t j = l16; j < n; j += 16) {
                        const int sh = s0 - 2 * (int)j;   // >= 0
                        const unsigned long long v = sh >= 64 ? bh >> (sh - 64) : (sh == 0 ? lo : (lo >> sh) | (bh << (64 - sh)));
                        const uint64_t h = mc_mix62(mc_canonical(v & kmask, k));
                        if (c == 1) {
                            singles[off + j] = h;
                        } else {
                            pkeys[off + j] = h;
                            pcnt[off + j] = (uint16_t)c;
                        }
                    }
                }
            }
            __syncthreads();
            for (uint32_t e = threadIdx.x; e < used; e += SKX_THREADS) {
                const uint32_t slot = s_list[e];
                s_hi[slot] = 0;
                s_lo[slot] = 0;
                s_cnt[slot] = 0;
            }
            if (threadIdx.x == 0) s_used = 0;
            __syncthreads();
        }
    }
}

// (hashed key, weight) pairs straight into the table: the k-mers of records seen more than once
__global__ void __launch_bounds__(256) k_add_hashed_pairs(McTableDev t, const uint64_t* keys, const uint16_t* counts, const unsigned long long* n_ptr, uint32_t shard_id) {
    const unsigned long long n = *n_ptr;
    const uint64_t lmask = (1ULL << (62 - t.shard_bits)) - 1;
    unsigned long long added = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = keys[i];
        if (t.shard_bits && (uint32_t)(h >> (62 - t.shard_bits)) != shard_id) continue;
        if (mc_table_add(t, h & lmask, counts[i])) added++;
    }
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&t.counters[MC_CNT_KEYS], added);
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static size_t sk_cursor_words(uint32_t nb) { return 2 * ((size_t)nb + 4) + 4; }

void mc_sk_release(mc_table* t) {
    McBatch* B = t->batch;
    if (!B || !B->sk_nb) return;
    mc_ctx* ctx = t->ctx;
    mc_dev_free(ctx, B->sk.recs);
    mc_dev_free(ctx, B->sk.spill);
    mc_dev_free(ctx, B->sk.cursors);
    B->sk = McSkDev{};
    B->sk_nb = 0;
    B->sk_room = 0;
}

extern uint64_t mc_available_device_bytes(mc_ctx* ctx);

static void sk_create(mc_table* t, uint64_t ub_kmers, uint64_t total_hint) {
    mc_ctx* ctx = t->ctx;
    McBatch* B = t->batch ? t->batch : (t->batch = new McBatch());
    const int W = mc_sk_window(t->k);
    // records expected: 2 / (W + 1) of the k-mers start a new minimizer in random sequence, + 25 %; every read adds one
    const uint64_t kmers = std::max<uint64_t>(total_hint ? std::max(total_hint, ub_kmers) : ub_kmers * 4, 1024);
    uint64_t records = (uint64_t)((double)kmers * 2.0 / (W + 1) * 1.25) + 4096;
    // at most a fifth of what the device can still give (16 bytes per record, x 1.6 of slack in the bins, + the spill list)
    const uint64_t budget = mc_available_device_bytes(ctx) / 5;
    records = std::min<uint64_t>(records, std::max<uint64_t>(budget / 32, 4096));
    int bits = 0;
    while (bits < MC_SK_MAX_BITS && (records >> bits) > 2048) bits++;
    if (ctx->tune_sk_bits >= 0) bits = std::min(ctx->tune_sk_bits, MC_SK_MAX_BITS);
    const uint32_t nb = 1u << bits;
    uint64_t cap = (records >> bits) + (records >> bits) * 3 / 5 + 64;
    if (ctx->tune_sk_cap) cap = ctx->tune_sk_cap;
    cap = std::min<uint64_t>(cap, 0x7FFFFFFFu);
    const uint64_t spill_cap = ctx->tune_sk_cap ? std::max<uint64_t>(ctx->tune_sk_cap * 4, 64) : records / 8 + 4096;
    B->sk.recs = (uint4*)mc_dev_alloc(ctx, (size_t)nb * cap * 16, "super-k-mer bins");
    B->sk.spill = (uint4*)mc_dev_alloc(ctx, (size_t)spill_cap * 16, "super-k-mer spill list");
    B->sk.cursors = (unsigned long long*)mc_dev_alloc(ctx, sk_cursor_words(nb) * 8, "super-k-mer cursors");
    B->sk.cap = cap;
    B->sk.spill_cap = spill_cap;
    B->sk.nb_mask = nb - 1;
    B->sk.n_max = (uint32_t)std::min(62 - t->k, 33);
    B->sk.flags = (unsigned int*)(B->sk.cursors + nb + 3);
    MC_CUDA(cudaMemsetAsync(B->sk.cursors, 0, sk_cursor_words(nb) * 8, ctx->stream));
    B->sk_nb = nb;
    B->sk_kmers = 0;
    B->sk_records_est = 0;
    B->sk_room = ctx->tune_sk_cap ? (uint64_t)nb * cap : records;
}

bool mc_sk_reserve(mc_table* t, uint64_t ub_kmers, uint64_t total_hint) {
    McBatch* B = t->batch;
    const int W = mc_sk_window(t->k);
    const uint64_t expect = (uint64_t)((double)ub_kmers * 2.0 / (W + 1)) + 1024;   // records this chunk should add
    if (B && B->sk_nb) {
        if (B->sk_records_est + expect <= B->sk_room) return true;
        mc_sk_flush(t);                               // the bins are (expected to be) full: unpack them, start again empty
        if (expect <= B->sk_room) return true;
        mc_sk_release(t);                             // a larger chunk than the bins were made for
    }
    sk_create(t, ub_kmers, total_hint);
    return true;
}

// the cursors (and the table's reader statistics) as they were before the chunk, kept on the device behind the cursors
void mc_sk_snapshot(mc_table* t) {
    McBatch* B = t->batch;
    cudaStream_t st = t->ctx->stream;
    const uint32_t nb = B->sk_nb;
    MC_CUDA(cudaMemsetAsync(B->sk.cursors + nb + 2, 0, 2 * 8, st));   // k-mers of the launch, flag word
    MC_CUDA(cudaMemcpyAsync(B->sk.cursors + nb + 4, B->sk.cursors, ((size_t)nb + 4) * 8, cudaMemcpyDeviceToDevice, st));
    MC_CUDA(cudaMemcpyAsync(B->sk.cursors + 2 * ((size_t)nb + 4), &t->dev.counters[MC_CNT_READS], 4 * 8, cudaMemcpyDeviceToDevice, st));
}
void mc_sk_rollback(mc_table* t) {
    McBatch* B = t->batch;
    cudaStream_t st = t->ctx->stream;
    const uint32_t nb = B->sk_nb;
    MC_CUDA(cudaMemcpyAsync(B->sk.cursors, B->sk.cursors + nb + 4, ((size_t)nb + 4) * 8, cudaMemcpyDeviceToDevice, st));
    MC_CUDA(cudaMemcpyAsync(&t->dev.counters[MC_CNT_READS], B->sk.cursors + 2 * ((size_t)nb + 4), 4 * 8, cudaMemcpyDeviceToDevice, st));
}

// one launch of k_sk_expand over `n_units` units holding at most `kmers` k-mer instances, then its two lists go on: pairs into the
// table, singles into the key bins
static void sk_expand_segment(mc_table* t, McSkExpandArgs a, uint64_t kmers, uint64_t total_kmers, uint64_t* done_kmers, uint64_t* singles_out) {
    mc_ctx* ctx = t->ctx;
    cudaStream_t st = ctx->stream;
    if (a.n_units == 0 || kmers == 0) return;
    const uint64_t pair_cap = kmers / 2 + 16;
    a.singles = (uint64_t*)mc_scratch(ctx, 2, (kmers + 16) * sizeof(uint64_t));
    a.pair_keys = (uint64_t*)mc_dev_alloc(ctx, pair_cap * sizeof(uint64_t), "weighted k-mers");
    a.pair_counts = (uint16_t*)mc_dev_alloc(ctx, pair_cap * sizeof(uint16_t), "weights");
    a.out = (unsigned long long*)mc_dev_alloc(ctx, 4 * sizeof(unsigned long long), "scratch");
    struct Free { mc_ctx* c; void* p[3]; ~Free() { for (void* q : p) mc_dev_free(c, q); } } fr{ctx, {a.pair_keys, a.pair_counts, a.out}};
    MC_CUDA(cudaMemsetAsync(a.out, 0, 4 * sizeof(unsigned long long), st));
    if (mc_attr_once(ctx, MC_ATTR_SK_EXPAND))
        MC_CUDA(cudaFuncSetAttribute(k_sk_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skx_smem_bytes()));
    {
        McStageTimer tm(ctx, MC_ST_EXPAND);
        const int grid = (int)std::min<uint64_t>(a.n_units, (uint64_t)ctx->sm_count * 2);
        k_sk_expand<<<grid, SKX_THREADS, skx_smem_bytes(), st>>>(a);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
    }
    unsigned long long* h = (unsigned long long*)((uint8_t*)ctx->h_pinned_small + 768);
    MC_CUDA(cudaMemcpyAsync(h, a.out, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_s = h[0], n_p = h[1];
    if (n_s > kmers || n_p > pair_cap) mc_throw(MC_ERR_CUDA, "internal: super-k-mer bins unpacked into more k-mers than they hold");
    if (getenv("MC_TRACE"))
        fprintf(stderr, "[mc trace] sk expand: %u units, %llu k-mer instances -> %llu singles + %llu weighted\n", a.n_units, (unsigned long long)kmers,
                (unsigned long long)n_s, (unsigned long long)n_p);
    if (n_p) {
        mc_table_reserve(t, n_p);
        McStageTimer tm(ctx, MC_ST_COUNT);
        const int grid = (int)std::min<uint64_t>((n_p + 255) / 256, (uint64_t)ctx->sm_count * 8);
        k_add_hashed_pairs<<<grid, 256, 0, st>>>(t->dev, a.pair_keys, a.pair_counts, a.out + 1, t->shard_id);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
    }
    if (n_s) {
        // what this flush will put into the key bins in total: the share of singles seen so far (the minimizer bins are a uniform sample
        // of the input) applied to the k-mers still packed, + 10 %
        const uint64_t seen = *done_kmers + kmers;
        const uint64_t rest = total_kmers > seen ? total_kmers - seen : 0;
        const uint64_t hint = n_s + (uint64_t)((double)rest * ((double)(*singles_out + n_s) / (double)seen) * 1.1) + 4096;
        {
            McHostTimer ht(ctx, MC_ST_GROW);
            mc_batch_reserve(t, n_s, hint);
        }
        McHostTimer ht(ctx, MC_ST_SCATTER);
        mc_batch_scatter(t, a.singles, n_s);
        t->batch->inst_equiv += kmers;
    }
    *singles_out += n_s;
    *done_kmers += kmers;
}

void mc_sk_flush(mc_table* t) {
    McBatch* B = t->batch;
    if (!B || !B->sk_nb || B->sk_flushing || B->sk_kmers == 0) return;
    mc_ctx* ctx = t->ctx;
    MC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    struct Guard { McBatch* b; ~Guard() { b->sk_flushing = false; } } guard{B};
    B->sk_flushing = true;
    const uint32_t nb = B->sk_nb;
    std::vector<unsigned long long> cur((size_t)nb + 2);
    MC_CUDA(cudaMemcpyAsync(cur.data(), B->sk.cursors, cur.size() * 8, cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    const uint64_t spill_n = std::min<uint64_t>(cur[nb], B->sk.spill_cap);
    uint64_t total_kmers = cur[nb + 1];
    for (uint32_t b = 0; b < nb; b++) total_kmers += cur[b] >> 32;
    // segments of bins: the unpacked lists of one segment are bounded by its k-mer total
    uint64_t seg_kmers = ctx->tune_segment_keys ? ctx->tune_segment_keys : ((uint64_t)1 << 29);
    seg_kmers = std::max<uint64_t>(std::min<uint64_t>(seg_kmers, mc_available_device_bytes(ctx) / 40), 4096);
    McSkExpandArgs a{};
    a.sk = B->sk;
    a.k = t->k;
    uint64_t singles = 0, done_kmers = 0;
    for (uint32_t b0 = 0; b0 < nb;) {
        uint64_t km = 0;
        uint32_t b1 = b0;
        while (b1 < nb && (b1 == b0 || km + (cur[b1] >> 32) <= seg_kmers)) km += cur[b1++] >> 32;
        a.b_lo = b0;
        a.n_units = b1 - b0;
        a.from_spill = 0;
        sk_expand_segment(t, a, km, total_kmers, &done_kmers, &singles);
        b0 = b1;
    }
    if (spill_n) {
        a.from_spill = 1;
        a.spill_n = spill_n;
        a.n_units = (uint32_t)((spill_n + SKX_ROUND - 1) / SKX_ROUND);
        sk_expand_segment(t, a, cur[nb + 1], total_kmers, &done_kmers, &singles);
    }
    MC_CUDA(cudaMemsetAsync(B->sk.cursors, 0, ((size_t)nb + 4) * 8, st));
    B->sk_kmers = 0;
    B->sk_records_est = 0;
}
