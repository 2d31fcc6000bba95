// Shared device/host definitions of libmcgpu (sm_100a).  See include/mc_b200.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mc_b200.h"

// --------------------------------------------------------------------------------------------
// errors
// --------------------------------------------------------------------------------------------
struct McError : std::runtime_error {
    int code;
    McError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void mc_throw(int code, const char* fmt, ...);
void mc_set_error(const std::string& msg);

#define MC_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            mc_throw(_e == cudaErrorMemoryAllocation ? MC_ERR_NOMEM : MC_ERR_CUDA, "CUDA error %s at %s:%d: %s", \
                     cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e));                 \
    } while (0)

// --------------------------------------------------------------------------------------------
// k-mer arithmetic (2-bit, A0 G1 C2 T3, first base most significant)
// --------------------------------------------------------------------------------------------
// ITMO!/utils/KmerUtils.java:12-22 restated: reverse the 2-bit groups, complement, align.
__host__ __device__ __forceinline__ uint64_t mc_revcomp(uint64_t x, int k) {
#ifdef __CUDA_ARCH__
    x = __brevll(x);                                                                       // reverse all 64 bits
    x = ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);           // un-reverse inside each pair
#else
    x = ((x & 0x3333333333333333ULL) << 2) | ((x & 0xccccccccccccccccULL) >> 2);
    x = ((x & 0x0f0f0f0f0f0f0f0fULL) << 4) | ((x & 0xf0f0f0f0f0f0f0f0ULL) >> 4);
    x = ((x & 0x00ff00ff00ff00ffULL) << 8) | ((x & 0xff00ff00ff00ff00ULL) >> 8);
    x = ((x & 0x0000ffff0000ffffULL) << 16) | ((x & 0xffff0000ffff0000ULL) >> 16);
    x = (x << 32) | (x >> 32);
#endif
    return (~x) >> (64 - 2 * k);
}
__host__ __device__ __forceinline__ uint64_t mc_canonical(uint64_t fw, int k) {
    uint64_t rc = mc_revcomp(fw, k);
    return fw < rc ? fw : rc;   // 2k <= 62 bits: unsigned == Java's signed Math.min
}

// nucleotide code of an input byte: 0..3, 4 = N/n (and '.' in FASTQ mode, decided by the caller), 5 = invalid
__host__ __device__ __forceinline__ int mc_nuc_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'G': case 'g': return 1;
        case 'C': case 'c': return 2;
        case 'T': case 't': return 3;
        case 'N': case 'n': return 4;
        default: return 5;
    }
}

// --------------------------------------------------------------------------------------------
// The count table (replaces BigLong2ShortHashMap).
//
// key -> h = mix62(key), a BIJECTION on 62-bit values, so (bucket, tag) identifies the key and the key
// itself is never stored (quotienting).  Layout in HBM: nb = 2^q buckets of 4 x u64 (one 32-byte sector).
//   h (62 bits) = [ shard : shard_bits | bucket : q | tag : tag_bits ]
//   slot word   = [ tag : <=41 | disp : 3 | count : 20 ]      0 = empty (count >= 1 for live slots)
// disp = how many buckets past its home bucket the entry sits (linear probing over <= 8 buckets); the
// rare entry that finds 8 full buckets goes to a small full-key stash.  count saturates at 32767
// (NumUtils.addAndBound, ITMO!/utils/NumUtils.java:21-26): an update reads the word, and only adds
// while count < 32767; concurrent adders can overshoot by at most the number of resident threads
// (< 2^19), which 20 bits absorb; readers clamp to 32767.
// --------------------------------------------------------------------------------------------
#define MC_COUNT_BITS 20
#define MC_DISP_BITS 3
#define MC_MAX_DISP 8
#define MC_COUNT_MASK ((1ULL << MC_COUNT_BITS) - 1)
#define MC_TAG_SHIFT (MC_COUNT_BITS + MC_DISP_BITS)
#define MC_MAX_TAG_BITS (64 - MC_TAG_SHIFT)   // 41
#define MC_SAT 32767u
#define MC_M62 ((1ULL << 62) - 1)
#define MC_HASH_SENTINEL 0xFFFFFFFFFFFFFFFFULL   // never a valid 62-bit hash

#define MC_MIX_C1 0x1E3779B97F4A7C15ULL   // odd, < 2^62
#define MC_MIX_C2 0x3F58476D1CE4E5B9ULL   // odd, < 2^62

__host__ __device__ __forceinline__ uint64_t mc_mix62(uint64_t x) {
    x ^= x >> 31;
    x = (x * MC_MIX_C1) & MC_M62;
    x ^= x >> 29;
    x = (x * MC_MIX_C2) & MC_M62;
    x ^= x >> 32;
    return x;
}
uint64_t mc_unmix62(uint64_t h);   // host + device versions in mc_table.cu
__device__ uint64_t mc_unmix62_dev(uint64_t h);

struct McTableDev {
    uint64_t* slots;          // nb * 4 words
    uint64_t nb_mask;         // nb - 1
    int q;                    // log2(nb)
    int tag_bits;             // 62 - shard_bits - q  (<= 41)
    int shard_bits;
    uint64_t tag_mask;
    uint64_t* stash_keys;     // hashed key + 1 (0 = empty), stash_mask + 1 entries
    uint32_t* stash_counts;
    uint32_t stash_mask;
    unsigned long long* counters;   // [0] distinct keys, [1] stash entries, [2] reads, [3] runs, [4] bases, [5] kmers, [6] overflow-list entries, [7] overflow-list pushes
    unsigned int* flags;            // bit 0: an update could not be stored anywhere (fatal)
    // Overflow list (only attached while reads are being counted): hashed keys whose 8-bucket window AND stash
    // neighbourhood were full.  The host grows the table and re-inserts them, so no update is ever lost and the table
    // never has to be grown pessimistically ahead of a chunk.
    uint64_t* ovf_keys;
    unsigned long long ovf_cap;
};
enum { MC_CNT_KEYS = 0, MC_CNT_STASH = 1, MC_CNT_READS = 2, MC_CNT_RUNS = 3, MC_CNT_BASES = 4, MC_CNT_KMERS = 5, MC_CNT_OVF = 6, MC_CNT_OVF_KEYS = 7, MC_CNT_N = 8 };
// the counters are the head of one 128-byte status block: [MC_CNT_N] = flag word, [MC_CNT_N + 1] = cursor of the sub-bin spill
// list of the flush group in flight (mc_bins.cu), so that one device-to-host copy tells the host everything about a group
#define MC_STATUS_FLAGS MC_CNT_N
#define MC_STATUS_SPILL2 (MC_CNT_N + 1)
#define MC_STATUS_WORDS 16
#define MC_FLAG_OVERFLOW 1u
#define MC_STASH_PROBES 64u

__device__ __forceinline__ void mc_load_bucket(const uint64_t* p, uint64_t w[4]) {
    // one 32-byte sector, straight from L2 (atomics live there): LDG.E.256
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
}

__device__ __forceinline__ bool mc_stash_add(const McTableDev& t, uint64_t h, uint32_t inc) {
    uint64_t want = h + 1;
    uint32_t pos = (uint32_t)(h * 0x9E3779B97F4A7C15ULL >> 40) & t.stash_mask;
    for (uint32_t probe = 0; probe < MC_STASH_PROBES; probe++, pos = (pos + 1) & t.stash_mask) {
        uint64_t cur = atomicCAS((unsigned long long*)&t.stash_keys[pos], 0ULL, (unsigned long long)want);
        if (cur == 0 || cur == want) {
            if (cur == 0) atomicAdd(&t.counters[MC_CNT_STASH], 1ULL);
            uint32_t old = atomicAdd(&t.stash_counts[pos], inc);
            if (old > 0x7FFFFFFFu - inc) atomicSub(&t.stash_counts[pos], inc);   // keep far away from wrap
            return cur == 0;
        }
    }
    if (t.ovf_keys) {   // `inc` copies: the list is sized for every key INSTANCE of the segment being applied
        unsigned long long pos2 = atomicAdd(&t.counters[MC_CNT_OVF], (unsigned long long)inc);
        if (pos2 + inc <= t.ovf_cap) {
            atomicAdd(&t.counters[MC_CNT_OVF_KEYS], 1ULL);   // list entries are instances; this one counts the calls (sizing estimate)
            for (uint32_t i = 0; i < inc; i++) t.ovf_keys[pos2 + i] = h;
            return false;
        }
    }
    atomicOr(t.flags, MC_FLAG_OVERFLOW);
    return false;
}
__device__ __forceinline__ int mc_stash_get(const McTableDev& t, uint64_t h) {
    uint64_t want = h + 1;
    uint32_t pos = (uint32_t)(h * 0x9E3779B97F4A7C15ULL >> 40) & t.stash_mask;
    for (uint32_t probe = 0; probe < MC_STASH_PROBES; probe++, pos = (pos + 1) & t.stash_mask) {
        uint64_t cur = __ldcg(&t.stash_keys[pos]);
        if (cur == 0) return -1;
        if (cur == want) {
            uint32_t c = __ldcg(&t.stash_counts[pos]);
            return c > MC_SAT ? (int)MC_SAT : (int)c;
        }
    }
    return -1;
}

// Adds `inc` (1 on the hot path) to the count of hashed key h (shard bits already stripped by the caller's
// mask inside).  Returns true when the key was new.  Lock-free: slots only ever go empty -> occupied and the
// tag|disp part is immutable, so probing in a fixed order and resolving stale "empty" reads through
// atomicCAS never creates a duplicate.
// `w0` (optional) = the words of the home bucket, loaded earlier by the caller so that several probes of one
// thread are in flight at once (k_bin_count); a stale copy is fine for the same reason a stale load is.
// FEW_ADDERS: the caller guarantees that at most 16 updates of ONE key, each <= 32767, can ever be in flight together (the merge of
// the multi-GPU exchange: one pair per source rank).  The count then goes up with a plain RED even for inc > 1: the 20-bit field
// cannot overflow (32766 + 16 x 32767 < 2^20), every reader clamps at 32767, and min(sum, 32767) is what sequential saturating adds
// give.  The general path keeps the saturating CAS loop, which retries when several sources bring the same k-mer at the same time.
template <bool PREFETCHED, bool FEW_ADDERS = false>
__device__ __forceinline__ bool mc_table_add_impl(const McTableDev& t, uint64_t hl, uint32_t inc, const uint64_t* w0) {
    const uint64_t tag = hl & t.tag_mask;
    const uint64_t b0 = hl >> t.tag_bits;
#pragma unroll 1
    for (uint32_t d = 0; d < MC_MAX_DISP; d++) {
        const uint64_t b = (b0 + d) & t.nb_mask;
        uint64_t* bucket = t.slots + b * 4;
        uint64_t w[4];
        if (PREFETCHED && d == 0) { w[0] = w0[0]; w[1] = w0[1]; w[2] = w0[2]; w[3] = w0[3]; }
        else mc_load_bucket(bucket, w);
        const uint64_t want = (tag << MC_TAG_SHIFT) | ((uint64_t)d << MC_COUNT_BITS);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            uint64_t cur = w[s];
            if (cur == 0) {
                cur = atomicCAS((unsigned long long*)&bucket[s], 0ULL, (unsigned long long)(want | inc));
                if (cur == 0) return true;   // claimed: new key
            }
            if ((cur >> MC_COUNT_BITS) == (want >> MC_COUNT_BITS)) {
                if (inc == 1 || FEW_ADDERS) {
                    if ((cur & MC_COUNT_MASK) < MC_SAT) atomicAdd((unsigned long long*)&bucket[s], (unsigned long long)inc);
                } else {
                    // rare path (add_pairs, rehash): saturating CAS loop
                    while ((cur & MC_COUNT_MASK) < MC_SAT) {
                        uint64_t c = (cur & MC_COUNT_MASK) + inc;
                        if (c > MC_SAT) c = MC_SAT;
                        uint64_t nv = (cur & ~MC_COUNT_MASK) | c;
                        uint64_t seen = atomicCAS((unsigned long long*)&bucket[s], (unsigned long long)cur, (unsigned long long)nv);
                        if (seen == cur) break;
                        cur = seen;
                    }
                }
                return false;
            }
        }
    }
    return mc_stash_add(t, hl, inc);
}
__device__ __forceinline__ bool mc_table_add(const McTableDev& t, uint64_t h, uint32_t inc) {
    const uint64_t hl = h & ((1ULL << (62 - t.shard_bits)) - 1);
    return mc_table_add_impl<false>(t, hl, inc, nullptr);
}

// count of hashed key h clamped to 32767, or -1 when absent (Long2ShortHashMap.get)
__device__ __forceinline__ int mc_table_lookup(const McTableDev& t, uint64_t h) {
    const uint64_t hl = h & ((1ULL << (62 - t.shard_bits)) - 1);
    const uint64_t tag = hl & t.tag_mask;
    const uint64_t b0 = hl >> t.tag_bits;
#pragma unroll 1
    for (uint32_t d = 0; d < MC_MAX_DISP; d++) {
        const uint64_t b = (b0 + d) & t.nb_mask;
        uint64_t w[4];
        mc_load_bucket(t.slots + b * 4, w);
        const uint64_t want = ((tag << MC_TAG_SHIFT) | ((uint64_t)d << MC_COUNT_BITS)) >> MC_COUNT_BITS;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (w[s] == 0) return -1;   // an empty slot ends the probe sequence (no deletions)
            if ((w[s] >> MC_COUNT_BITS) == want) {
                uint32_t c = (uint32_t)(w[s] & MC_COUNT_MASK);
                return c > MC_SAT ? (int)MC_SAT : (int)c;
            }
        }
    }
    return mc_stash_get(t, hl);
}

// --------------------------------------------------------------------------------------------
// Binned updates (the write-combining front end of the table, mc_bins.cu).
//
// A random probe of an HBM-sized table costs a 128-byte line fill + a write-back per update (profiles/r1a), which
// caps direct counting near 20 G updates/s.  Instead, hashed keys are first scattered into P = 2^pbits bins by the
// TOP bits of their local hash -- bin b covers the contiguous table slice of buckets [b*nb/P, (b+1)*nb/P) -- and the
// bins are then applied in order, so that the resident CTAs of k_bin_count work on one or two slices at a time and
// every probe and atomic hits L2.  Each slice is read from and written back to HBM once per batch.
// A bin that is full (heavy hitters) diverts to the shared spill list, which is applied with plain direct updates.
// --------------------------------------------------------------------------------------------
#define MC_MAX_PBITS 11
struct McBinsDev {
    uint64_t* keys;                 // P * cap hashed LOCAL keys (shard bits stripped)
    unsigned long long* cursors;    // [0..P): append attempts per bin (may exceed cap), [P]: spill cursor
    uint64_t* spill;
    uint64_t cap;                   // entries per bin
    uint64_t spill_cap;
    int pbits;
    int hbits;                      // 62 - shard_bits: keys are stored as hl = h & (2^hbits - 1)
    int shift;                      // bin = (hl >> shift) & (P - 1); hbits - pbits for the first level
    unsigned long long* spill_cursor;   // &cursors[P] for the first level; shared by all sub-bin groups of the second level
};

// --------------------------------------------------------------------------------------------
// Super-k-mer front end (mc_sk.cu).  For k >= 22 the extractor does not emit one hashed key per k-mer: it cuts every run of valid
// bases into SUPER-K-MERS -- maximal stretches of consecutive k-mers whose window minimum (the smallest hashed canonical m-mer among
// the W = k - m + 1 m-mers of the k-mer) falls into the same minimizer bin -- and appends each as one 16-byte record to that bin.
// A k-mer and its reverse complement hold the same canonical m-mers, so every instance of a canonical k-mer lands in the same bin,
// and identical stretches of the sequenced genomes give identical records: at flush time a bin is de-duplicated in shared memory
// (record -> multiplicity) before its records are unpacked into (hashed k-mer, weight) updates.
//
//   record (128 bits):  [122,128) n = k-mers in the record (1 .. MC_SK_NMAX)
//                       [0,122)   bases, 2 bits each, first base highest: base i at bit 2 * (60 - i); unused low bits are zero
//   k-mer j of a record = (bases >> (122 - 2k - 2j)) & (4^k - 1), the reference's packing (ShortKmer.java:68-71)
// --------------------------------------------------------------------------------------------
#define MC_SK_MIN_K 22
#define MC_SK_MAX_BITS 18
struct McSkDev {
    uint4* recs;                    // NB * cap records
    unsigned long long* cursors;    // [NB]: (k-mers << 32 | records) appended to the bin (attempts: may exceed cap);
                                    // [NB] spill records, [NB + 1] spill k-mers, [NB + 2] k-mers of the current extractor launch
    uint4* spill;                   // records of full bins
    uint64_t cap;                   // records per bin
    uint64_t spill_cap;
    uint32_t nb_mask;               // NB - 1
    uint32_t n_max;                 // k-mers per record: min(62 - k, 33)
    unsigned int* flags;            // MC_FLAG_OVERFLOW: the spill list is full too (the host redoes the chunk with flat keys)
};
__host__ __device__ __forceinline__ int mc_sk_window(int k) { return k >= 30 ? 20 : 12; }   // W; m = k - W + 1
// the m-mers are rolled in 32-bit registers (m <= 16) and must be long enough to spread over the bins (m >= 11): k in [22,27] or [30,31]
__host__ __device__ __forceinline__ bool mc_sk_supported(int k) { const int m = k - mc_sk_window(k) + 1; return k >= MC_SK_MIN_K && k <= 31 && m >= 11 && m <= 16; }

#define MC_MAX_XCHG_RANKS 16
struct McBinDst {   // per-bin destination override of k_bin_scatter (owner groups written straight into peer memory)
    uint64_t* base[MC_MAX_XCHG_RANKS];
    int use;
};
// pads (sentinels) a sector-aligned run may carry: k_bin_scatter rounds every bin's run of a tile up to 4 keys
#define MC_SCATTER_TILE 8192u

// --------------------------------------------------------------------------------------------
// host-side objects behind the opaque handles
// --------------------------------------------------------------------------------------------
enum { MC_ST_PARSE = 0, MC_ST_EXTRACT = 1, MC_ST_GROW = 2, MC_ST_SCATTER = 3, MC_ST_COUNT = 4, MC_ST_FILTER = 5, MC_ST_EXPAND = 7, MC_N_STAGES = 8 };
struct McPendingTimer { int stage; cudaEvent_t a, b; };
struct mc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // second compute stream: the sub-bin scatter of flush group g + 1 runs beside the aggregation of group g
    cudaEvent_t ev[16] = {};
    cudaEvent_t stage_ev[8] = {};
    size_t chunk_bytes = (size_t)256 << 20;
    uint64_t launches = 0;
    float stage_ms[MC_N_STAGES] = {};
    double host_ms[MC_N_STAGES] = {};     // host wall time spent inside the same phases (launch + synchronisation included): where a step's
                                          // time goes when it is NOT in the kernels (McHostTimer)
    std::vector<McPendingTimer> timers;   // recorded, not yet resolved (mc_resolve_timers)
    // tuning knobs (mc_ctx_set_tuning); 0 = automatic
    long tune_part_bits = -1;        // P = 2^bits bins
    uint64_t tune_batch_keys = 0;    // capacity of one batch of binned keys
    uint64_t tune_segment_keys = 0;  // keys applied per k_bin_count launch
    uint32_t tune_count_tile = 0;    // keys per k_bin_count work unit
    int tune_count_variant = 0;      // k_bin_count instantiation (probes in flight / occupancy), for measurements
    uint32_t tune_host_batch_chunks = 0;   // host input: chunks binned between two table sweeps (0 = default 1)
    int tune_extract_parts = 0;      // lanes per record in k_extract: 0 = by read length, else 1 / 4 / 32
    int tune_agg = -1;               // two-level bins + shared-memory aggregation in the flush: -1 = by batch size, 0 = off, 1 = on
    uint32_t tune_agg_unit = 0;      // target keys per aggregation unit (sub-bin); 0 = default
    size_t tune_device_chunk_bytes = 0;   // device-resident input: bytes per parse/extract pass (0 = as large as 32-bit offsets allow)
    int tune_defer_flush = 0;        // mc_count_* leave their last batch of bins pending (the next reader of the table, or
                                     // mc_xchg_flush_and_dump, applies it)
    int tune_extract_tma = 1;        // k_extract stages the warps' record spans with cp.async.bulk + mbarrier
    int tune_flush_overlap = 0;      // aggregated flush: scatter2(g + 1) on aux_stream concurrent with agg(g), one CTA per SM each
                                     // (measured: each kernel slows down 1.5x at half occupancy, nothing is gained: profiles/r2_NOTES.md)
    int tune_stage_timing = 1;       // 0: McStageTimer records nothing (four CUDA API calls less per kernel launch)
    int tune_superkmers = 0;         // 1 (and k supported): the extractor emits super-k-mer records into minimizer bins.  OFF: measured 3.3x slower
                                     // than the flat path at cfg2 (509 vs 153 ms per step, profiles/r2_NOTES.md section 6); kept for the parity tests
    int tune_sk_bits = -1;           // log2 of the number of minimizer bins (-1: by input size)
    uint64_t tune_sk_cap = 0;        // records per minimizer bin (0: by input size) -- tests force spills with tiny bins
    int tune_agg_lane_loop = 0;      // k_bin_agg: one probe loop per lane over its keys of a round (0: one loop per key, round 2's first form)
    int tune_scatter_tma = 1;        // bin runs leave the sorted tile through bulk copies (cp.async.bulk) instead of per-sector stores
    int keep_stage_ms = 0;           // 1: mc_count_* do not reset the stage timers (a step made of several count calls)
    int sm_count = 148;
    // reusable device scratch for the parser (grown on demand)
    void* d_chunk[2] = {nullptr, nullptr};
    size_t d_chunk_cap[2] = {0, 0};
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_cap[2] = {0, 0};
    void* d_scratch[4] = {nullptr, nullptr, nullptr, nullptr};   // slot 0: parser arrays, 1: tile counts + header, 2: extract keys, 3: misc
    size_t d_scratch_cap[4] = {0, 0, 0, 0};
    void* h_pinned_small = nullptr;   // 4 KiB pinned for small read-backs
    // Device-block arena (mc_dev_alloc / mc_dev_free): freed blocks stay with the context and are handed out again to
    // requests of (nearly) the same size, so a steady stream of count / filter / free calls issues no driver
    // allocation at all after its first pass.  Everything allocated here is used on ctx->stream only, which is what
    // makes the immediate reuse of a freed block safe (stream order).  mc_ctx_trim returns the cache to the driver.
    struct Block { void* p; size_t bytes; };
    std::mutex arena_mu;
    std::vector<Block> arena_free;
    std::unordered_map<void*, size_t> arena_live;
    size_t arena_cached_bytes = 0;
    uint64_t arena_driver_allocs = 0;   // cudaMalloc / cudaFree calls made by the arena (bench/test visibility)
    uint64_t scratch_allocs = 0;        // (re)allocations of the scratch slots and chunk buffers
    size_t free_bytes_cached = 0;       // cudaMemGetInfo's answer, valid while free_bytes_epoch == arena_driver_allocs + scratch_allocs
    uint64_t free_bytes_epoch = 0;
    std::vector<long long> last_extend_counts;   // mc_env_extend_counts: per environment of the last mc_env_find call
    // pooled timing events (McStageTimer)
    std::vector<cudaEvent_t> event_pool;
    // one kernel-attribute opt-in per context (= per device): bit i = the i-th MC_ATTR_* kernel was configured
    uint32_t attr_done = 0;
};
enum { MC_ATTR_EXTRACT = 0, MC_ATTR_SCATTER = 8, MC_ATTR_SCATTER2 = 9, MC_ATTR_AGG = 10, MC_ATTR_MISC = 11, MC_ATTR_EXTRACT_SK = 12 /* .. 23 */, MC_ATTR_SK_EXPAND = 24 };
static inline bool mc_attr_once(mc_ctx* ctx, int bit) {   // true the first time `bit` is asked for on this context
    if (ctx->attr_done & (1u << bit)) return false;
    ctx->attr_done |= 1u << bit;
    return true;
}
cudaEvent_t mc_event_get(mc_ctx* ctx);
void mc_event_put(mc_ctx* ctx, cudaEvent_t e);
size_t mc_arena_cached_bytes(mc_ctx* ctx);
void mc_arena_trim(mc_ctx* ctx);

// Observer of the aggregated flush (mc_bins.cu): called on the host right after the kernels of a group of first-level bins have been
// queued on ctx->stream, i.e. once everything in the table whose HOME bucket lies in that hash range is final for this flush.  The
// multi-GPU owner dump uses it to ship a finished range to its owners while the next groups are still being aggregated (mc_xchg.cu).
struct McFlushHook {
    void (*group_done)(void* user, struct mc_table* t, uint32_t b0, uint32_t G, uint32_t P1);
    void (*before_grow)(void* user, struct mc_table* t);   // the table's memory is about to be replaced (rehash)
    void* user;
    bool failed;    // the flush did something a per-range consumer cannot follow (first-level spill list, overflow re-insertion, direct sweep)
    uint32_t groups;
};
struct mc_table {
    mc_ctx* ctx = nullptr;
    McFlushHook* hook = nullptr;
    int k = 0;
    uint32_t shard_id = 0;   // this table owns hashes whose top dev.shard_bits bits equal shard_id
    McTableDev dev{};
    uint64_t nb = 0;
    uint64_t capacity_slots() const { return nb * 4; }
    struct McBatch* batch = nullptr;   // pending binned updates (mc_bins.cu); applied by mc_table_flush
    uint64_t expect_instances = 0;     // mc_table_expect_keys: k-mer instances the caller still expects to add (0 = unknown)
    uint64_t seen_instances = 0;       // instances applied since the expectation was set
    uint32_t stream_flushes = 0;       // flushes since the expectation was set
};

// kernels / host helpers shared across translation units
void mc_table_reserve(mc_table* t, uint64_t additional_keys_upper_bound);   // grow (rehash) until it fits at load <= 0.55
void mc_table_prepare_chunk(mc_table* t, uint64_t chunk_kmers_upper_bound);  // amortised growth + attach the overflow list
void mc_table_finish_chunk(mc_table* t, bool grown_already = false);         // re-insert overflowed keys (growing), detach the list
void* mc_dev_alloc(mc_ctx* ctx, size_t bytes, const char* what);            // stream-ordered pool allocation on ctx->stream
void mc_dev_free(mc_ctx* ctx, void* p);
uint64_t mc_table_keys(const mc_table* t);                                  // synchronising read of the key counter
void* mc_scratch(mc_ctx* ctx, int slot, size_t bytes);                      // grows ctx->d_scratch[slot]; contents are lost on growth
#define MC_LAUNCH(ctx) ((ctx)->launches++)

// pending binned updates of one table (host side)
struct McBatch {
    McBinsDev dev{};
    uint32_t P = 1;
    uint64_t capacity = 0;      // total keys the batch is meant to hold (P * cap has slack on top)
    uint64_t used = 0;          // keys scattered so far (upper bound: includes sentinels and foreign-shard keys)
    uint64_t spill_used = 0;
    // super-k-mer bins in front of the key bins (mc_sk.cu)
    McSkDev sk{};
    uint32_t sk_nb = 0;                 // 0: not allocated
    uint64_t sk_kmers = 0;              // k-mer instances held by the records (exact)
    uint64_t sk_records_est = 0;        // records appended so far (estimated from the bytes seen; the exact figure is on the device)
    uint64_t sk_room = 0;               // records the bins were sized for
    bool sk_flushing = false;           // mc_sk_flush is feeding the key bins: a nested mc_table_flush leaves the records alone
    uint64_t inst_equiv = 0;            // k-mer instances the keys in the key bins stand for (de-duplicated records weigh more than one)
    uint64_t ub_sum = 0, last_ub = 0;   // what the callers RESERVED for those scatters (worst case per chunk): used / ub_sum is the input's
                                        // real k-mer density, which plans the next chunk better than its worst case
};
// make sure the table's batch has room for `ub` more keys (flushing if needed); total_hint = all keys the caller expects to add
void mc_batch_reserve(mc_table* t, uint64_t ub, uint64_t total_hint);
// scatter n flat hashed keys (global hashes; sentinels and keys of other shards are dropped) into the batch
void mc_batch_scatter(mc_table* t, const uint64_t* d_flat, uint64_t n);
// apply every pending update to the table (growing it as needed) and keep the buffers for reuse
void mc_table_flush(mc_table* t);
// super-k-mer bins (mc_sk.cu): room for a chunk of at most ub_kmers k-mer instances (false: use flat keys for this chunk);
// roll a failed chunk back; unpack every pending record into the key bins / the table
bool mc_sk_reserve(mc_table* t, uint64_t ub_kmers, uint64_t total_hint);
void mc_sk_snapshot(mc_table* t);
void mc_sk_rollback(mc_table* t);
void mc_sk_flush(mc_table* t);
void mc_sk_release(mc_table* t);
// flush + release the batch buffers
void mc_batch_release(mc_table* t);
void mc_table_grow_to(mc_table* t, int new_q);   // timed rehash into 2^new_q buckets
uint64_t* mc_extract_flat_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, int k, mc_table* stats_sink, uint64_t* n_flat);
void mc_launch_owner_scatter(mc_ctx* ctx, const uint64_t* d_flat, uint64_t n, const McBinsDev& owners, const McBinDst& dst, unsigned int* flags);
int mc_q_for_keys(uint64_t keys, int shard_bits, double load);

// accumulate the device time between two points of ctx->stream into ctx->stage_ms[stage] (resolved lazily)
struct McStageTimer {
    mc_ctx* ctx;
    cudaStream_t on;
    McPendingTimer t;
    McStageTimer(mc_ctx* c, int stage, cudaStream_t s = nullptr) : ctx(c), on(s ? s : c->stream) {
        t.stage = stage;
        t.a = t.b = nullptr;
        if (!c->tune_stage_timing) return;
        t.a = mc_event_get(c);
        t.b = mc_event_get(c);
        cudaEventRecord(t.a, on);
    }
    ~McStageTimer() {
        if (!t.a) return;
        cudaEventRecord(t.b, on);
        ctx->timers.push_back(t);
    }
};
void mc_resolve_timers(mc_ctx* ctx);
#include <chrono>
struct McHostTimer {   // wall time of a host-side phase, accumulated into ctx->host_ms[stage]
    mc_ctx* ctx;
    int stage;
    std::chrono::steady_clock::time_point t0;
    McHostTimer(mc_ctx* c, int s) : ctx(c), stage(s), t0(std::chrono::steady_clock::now()) {}
    ~McHostTimer() { ctx->host_ms[stage] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};   // synchronises with the recorded events and folds them into ctx->stage_ms
