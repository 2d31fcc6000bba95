// Multi-GPU key exchange over peer memory: compute and transfer in ONE kernel.
//
// Every rank owns a receive buffer of `world` regions per slot (two slots: double buffering).  Region (slot, src) of rank
// d's buffer is written ONLY by rank src: k_bin_scatter, run with the owner ranks as its bins, sorts a tile of hashed keys
// by owner in shared memory and stores each owner's run, whole 32-byte sectors, straight into that owner's region through a
// CUDA-IPC mapping -- the stores travel over NVLink while the next tile is being sorted.  No NCCL call moves keys; the
// ranks only need a barrier per batch (torch.distributed, a few bytes) so that a receiver knows its regions are complete.
// Cursors are local to the writer (nobody else appends to its regions), so there are no cross-GPU atomics; the final
// cursor values are posted into the receiver's mailbox after the kernel.
#include <algorithm>
#include <cstring>
#include <vector>

#include "mc_common.cuh"

struct mc_xchg {
    mc_ctx* ctx = nullptr;
    int world = 1, rank = 0, shard_bits = 0;
    uint64_t region_keys = 0;       // capacity of one (slot, src) region (pads included)
    uint64_t max_batch_keys = 0;    // largest batch (in k-mers) a region is guaranteed to hold
    uint8_t* local = nullptr;       // my receive buffer: [2 slots][world regions][region_keys] keys, then the mailbox
    size_t local_bytes = 0;
    uint8_t* peer[MC_MAX_XCHG_RANKS] = {};   // mapped base of every rank's buffer (peer[rank] == local)
    unsigned long long* d_cursors = nullptr; // [world + 1] local append cursors of the batch being scattered
    unsigned int* d_flags = nullptr;
    uint64_t mailbox_off = 0;
    // "merge" exchange: a region holds pair_cap() hashed keys (8 bytes each) followed by as many 16-bit counts
    uint64_t pair_cap() const { return (region_keys * 8 / 10) & ~15ull; }
    uint64_t* region(uint8_t* base, int slot, int src) const { return (uint64_t*)base + ((uint64_t)slot * world + src) * region_keys; }
    unsigned long long* mailbox(uint8_t* base, int slot) const { return (unsigned long long*)(base + mailbox_off) + (uint64_t)slot * MC_MAX_XCHG_RANKS; }
};

extern "C" int mc_xchg_create(mc_ctx* ctx, int world, int rank, uint64_t region_keys, mc_xchg** out) {
    try {
        if (!ctx || !out) mc_throw(MC_ERR_ARG, "null argument");
        if (world < 1 || world > MC_MAX_XCHG_RANKS || (world & (world - 1)) || rank < 0 || rank >= world) mc_throw(MC_ERR_ARG, "world must be a power of two <= %d", MC_MAX_XCHG_RANKS);
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_xchg* x = new mc_xchg();
        x->ctx = ctx;
        x->world = world;
        x->rank = rank;
        while ((1 << x->shard_bits) < world) x->shard_bits++;
        x->max_batch_keys = region_keys;
        // pads: every tile of the scatter may round a region's run up by 3 keys
        region_keys += 3 * (region_keys / MC_SCATTER_TILE + 2) + 64;
        x->region_keys = (region_keys + 15) & ~15ull;
        x->mailbox_off = 2ull * world * x->region_keys * 8;
        x->local_bytes = x->mailbox_off + 2 * MC_MAX_XCHG_RANKS * sizeof(unsigned long long);
        cudaError_t e = cudaMalloc(&x->local, x->local_bytes);   // plain cudaMalloc: IPC cannot export pool memory
        if (e != cudaSuccess) {
            cudaGetLastError();
            const double gib = (double)x->local_bytes / (1 << 30);
            delete x;
            mc_throw(MC_ERR_NOMEM, "exchange buffer of %.2f GiB does not fit in device memory", gib);
        }
        MC_CUDA(cudaMemset(x->local + x->mailbox_off, 0, 2 * MC_MAX_XCHG_RANKS * sizeof(unsigned long long)));
        MC_CUDA(cudaMalloc(&x->d_cursors, (MC_MAX_XCHG_RANKS + 1) * sizeof(unsigned long long)));
        MC_CUDA(cudaMalloc(&x->d_flags, sizeof(unsigned int)));
        MC_CUDA(cudaMemset(x->d_flags, 0, sizeof(unsigned int)));
        x->peer[rank] = x->local;
        *out = x;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_xchg_handle(mc_xchg* x, void* out64) {
    try {
        if (!x || !out64) mc_throw(MC_ERR_ARG, "null argument");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        MC_CUDA(cudaIpcGetMemHandle(&h, x->local));
        memcpy(out64, &h, 64);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_xchg_open(mc_xchg* x, int peer_rank, const void* handle64) {
    try {
        if (!x || !handle64 || peer_rank < 0 || peer_rank >= x->world) mc_throw(MC_ERR_ARG, "bad argument");
        if (peer_rank == x->rank) return MC_OK;
        MC_CUDA(cudaSetDevice(x->ctx->device));
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void* p = nullptr;
        MC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[peer_rank] = (uint8_t*)p;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" void mc_xchg_destroy(mc_xchg* x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->stream);
    for (int r = 0; r < x->world; r++)
        if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->local);
    cudaFree(x->d_cursors);
    cudaFree(x->d_flags);
    cudaGetLastError();
    delete x;
}

// posts my final cursors into every owner's mailbox (one 8-byte store per peer, system scope)
__global__ void k_post_counts(const unsigned long long* __restrict__ cursors, McBinDst mail, int world, int rank) {
    const int o = threadIdx.x;
    if (o < world) {
        unsigned long long* slot = (unsigned long long*)mail.base[o] + rank;
        asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(slot), "l"(cursors[o]) : "memory");
    }
}

// parse + extract the batch, then sort its keys by owner and store every owner's run into that owner's region (slot, my rank)
extern "C" int mc_xchg_scatter_device(mc_xchg* x, const void* dbuf, size_t len, int fmt, int k, int slot, mc_table* stats_sink) {
    try {
        if (!x || !stats_sink || (len && !dbuf) || slot < 0 || slot > 1) mc_throw(MC_ERR_ARG, "bad argument");
        mc_ctx* ctx = x->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        for (int r = 0; r < x->world; r++)
            if (!x->peer[r]) mc_throw(MC_ERR_ARG, "exchange: rank %d's buffer has not been opened", r);
        if (mc_max_kmers(len, fmt, k) > x->max_batch_keys) mc_throw(MC_ERR_ARG, "exchange: batch of %zu bytes exceeds the region capacity", len);
        MC_CUDA(cudaMemsetAsync(x->d_cursors, 0, (MC_MAX_XCHG_RANKS + 1) * sizeof(unsigned long long), st));
        uint64_t n_flat = 0;
        uint64_t* d_flat = len ? mc_extract_flat_device(ctx, dbuf, len, fmt, k, stats_sink, &n_flat) : nullptr;
        McBinDst dst{}, mail{};
        dst.use = 1;
        for (int o = 0; o < x->world; o++) {
            dst.base[o] = x->region(x->peer[o], slot, x->rank);
            mail.base[o] = (uint64_t*)x->mailbox(x->peer[o], slot);
        }
        if (d_flat && n_flat) {
            McStageTimer tm(ctx, MC_ST_SCATTER);
            McBinsDev owners{};
            owners.cursors = x->d_cursors;
            owners.cap = x->region_keys;
            owners.spill_cap = 0;
            owners.pbits = x->shard_bits;
            owners.hbits = 62;
            owners.shift = 62 - x->shard_bits;
            owners.spill_cursor = x->d_cursors + ((size_t)1 << x->shard_bits);
            mc_launch_owner_scatter(ctx, d_flat, n_flat, owners, dst, x->d_flags);
        }
        k_post_counts<<<1, 32, 0, st>>>(x->d_cursors, mail, x->world, x->rank);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned int f = 0;
        MC_CUDA(cudaMemcpyAsync(&f, x->d_flags, sizeof f, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        if (f) mc_throw(MC_ERR_NOMEM, "exchange: a receive region overflowed");
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// after the ranks' barrier: bin what every source wrote into my regions of `slot`
extern "C" int mc_xchg_insert(mc_xchg* x, int slot, mc_table* t, uint64_t* n_received) {
    try {
        if (!x || !t || slot < 0 || slot > 1) mc_throw(MC_ERR_ARG, "bad argument");
        mc_ctx* ctx = x->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        unsigned long long counts[MC_MAX_XCHG_RANKS];
        MC_CUDA(cudaMemcpyAsync(counts, x->mailbox(x->local, slot), sizeof counts, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        uint64_t total = 0;
        for (int src = 0; src < x->world; src++) {
            if (counts[src] > x->region_keys) mc_throw(MC_ERR_CUDA, "exchange: mailbox of rank %d is corrupt", src);
            if (counts[src] == 0) continue;
            int rc = mc_table_insert_hashed_device(t, x->region(x->local, slot, src), counts[src]);
            if (rc != MC_OK) return rc;
            total += counts[src];
        }
        if (n_received) *n_received = total;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// --------------------------------------------------------------------------------------------
// "count locally, then merge" (include/mc_b200.h): distinct (hashed key, count) pairs instead of key instances cross NVLink
// --------------------------------------------------------------------------------------------
// One 32-byte bucket per thread and step (plus the stash), like k_dump; every live entry is appended to its owner's region: the
// hashed key (8 bytes) to the region's key array, the count (16 bits) to its count array at the same position -- 10 bytes per
// entry on NVLink.  Appends are aggregated per CTA: one atomicAdd per owner, CTA and step on a LOCAL cursor (the region
// (slot, my rank) of every peer has this rank as its only writer).  Per-warp appends were measured first: tens of millions of
// atomics on ONE address, which the L2 serialises (20 ms for a 16 GB table, 17 % of DRAM peak).  The threads of a CTA store
// to consecutive positions: whole 32-byte sectors on NVLink.  The grid walks the table in bucket order = hash order, so a
// receiver sees its pairs in nearly ascending order of ITS bucket index and the merge streams through its shard table.
#define DUMP_THREADS 512
// Range form: the scan covers buckets [scan_lo, scan_lo + scan_n) (mod nb) and emits the entries whose HOME bucket lies in
// [home_lo, home_hi) -- an entry sits up to 7 buckets past its home, so a range is scanned 8 buckets beyond its end and entries that
// were displaced INTO it from the range before are left to that range.  Every entry is thereby emitted exactly once whatever the
// order of the ranges and whatever the table's geometry at the time (the fused flush + dump below ships a hash range as soon as
// the flush has finished it).  The whole table is the range [0, nb) scanned once.
__global__ void __launch_bounds__(DUMP_THREADS, 2) k_dump_owners(McTableDev t, int sb, McBinDst dst, unsigned long long* __restrict__ cursors, uint64_t cap_pairs,
                                                             unsigned int* flags, uint64_t scan_lo, uint64_t scan_n, uint64_t home_lo, uint64_t home_hi) {
    __shared__ uint32_t s_tot[MC_MAX_XCHG_RANKS][DUMP_THREADS / 32];
    __shared__ unsigned long long s_base[MC_MAX_XCHG_RANKS];
    __shared__ uint32_t s_otot[MC_MAX_XCHG_RANKS], s_oreal[MC_MAX_XCHG_RANKS];   // entries per owner in this step: padded, real
    // every live entry of the CTA's 512 buckets, grouped by owner; every group padded to a multiple of 4 entries (count 0 = padding)
    __shared__ __align__(16) uint64_t s_keys[DUMP_THREADS * 4 + 4 * MC_MAX_XCHG_RANKS];
    __shared__ __align__(16) uint16_t s_cnts[DUMP_THREADS * 4 + 4 * MC_MAX_XCHG_RANKS];
    const uint64_t nb = scan_n;   // buckets scanned by this launch
    const uint64_t stash_units = ((uint64_t)t.stash_mask + 1 + 3) / 4;
    const uint64_t total = nb + stash_units;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t iters = (total + stride - 1) / stride;   // the same trip count for every thread: the barriers below are convergent
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int world = 1 << sb;
    uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t w4[4] = {0, 0, 0, 0};
    if (u < nb) mc_load_bucket(t.slots + ((scan_lo + u) & t.nb_mask) * 4, w4);
    for (uint64_t it = 0; it < iters; it++, u += stride) {
        uint64_t h[4];
        uint32_t c[4] = {0, 0, 0, 0};
        if (u < nb) {
            const uint64_t b = (scan_lo + u) & t.nb_mask;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint64_t w = w4[s];
                const uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
                const uint64_t home = (b - d) & t.nb_mask;
                h[s] = (home << t.tag_bits) | (w >> MC_TAG_SHIFT);
                c[s] = (w && home >= home_lo && home < home_hi) ? (uint32_t)(w & MC_COUNT_MASK) : 0;
            }
        } else if (u < total) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint64_t i = (u - nb) * 4 + s;
                const uint64_t kx = i <= t.stash_mask ? t.stash_keys[i] : 0;
                h[s] = kx - 1;
                const uint64_t home = ((kx - 1) >> t.tag_bits) & t.nb_mask;
                c[s] = (kx && home >= home_lo && home < home_hi) ? t.stash_counts[i] : 0;
            }
        }
        if (u + stride < nb) mc_load_bucket(t.slots + ((scan_lo + u + stride) & t.nb_mask) * 4, w4);   // the next step's bucket is in flight during this one
        uint32_t own = 0;   // owner of entry s in bits [8s, 8s+8), 0xFF = no entry
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (c[s] > MC_SAT) c[s] = MC_SAT;
            own |= (c[s] ? (uint32_t)(sb ? h[s] >> (62 - sb) : 0) : 0xFFu) << (8 * s);
        }
        // pass 1: entries per owner and warp (the offsets inside the warp are recomputed in pass 2: cheaper than keeping 16 of them).
        // The table is walked in bucket order = hash order and the owner is the top of the hash: the 128 entries of a warp's 32
        // buckets belong to ONE owner (two at a shard border, any in the stash), so the warp only visits the owners it holds --
        // a loop over all `world` owners with a warp scan each made the dump twice as slow at 8 ranks as at 2.
        uint32_t present = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const uint32_t o = (own >> (8 * s)) & 0xFFu;
            if (o != 0xFFu) present |= 1u << o;
        }
        present = __reduce_or_sync(0xffffffffu, present);
        if (lane < world) s_tot[lane][warp] = 0;
        __syncwarp();
        for (uint32_t m = present; m; m &= m - 1) {
            const int o = __ffs(m) - 1;
            uint32_t mine = 0;
#pragma unroll
            for (int s = 0; s < 4; s++) mine += ((own >> (8 * s)) & 0xFFu) == (uint32_t)o;
            uint32_t incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) s_tot[o][warp] = incl;
        }
        __syncthreads();
        if (threadIdx.x < world) {
            uint32_t tot = 0;
            for (int w = 0; w < DUMP_THREADS / 32; w++) tot += s_tot[threadIdx.x][w];
            const uint32_t tot4 = (tot + 3u) & ~3u;   // reservations in units of 4 entries: 32-byte key granules, 8-byte count granules
            unsigned long long base = tot4 ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)tot4) : 0ULL;
            if (base + tot4 > cap_pairs) {
                atomicOr(flags, 1u);
                base = ~0ULL;
            }
            s_base[threadIdx.x] = base;
            s_otot[threadIdx.x] = tot4;
            s_oreal[threadIdx.x] = tot;
        }
        __syncthreads();
        if (threadIdx.x < world) {   // the pads of my owner's group
            uint32_t ooff = 0;
            for (int p = 0; p < (int)threadIdx.x; p++) ooff += s_otot[p];
            for (uint32_t i = s_oreal[threadIdx.x]; i < s_otot[threadIdx.x]; i++) {
                s_keys[ooff + i] = 0;
                s_cnts[ooff + i] = 0;
            }
        }
        // pass 2: the CTA's pairs, grouped by owner, into shared memory ...
        for (uint32_t m = present; m; m &= m - 1) {
            const int o = __ffs(m) - 1;
            uint32_t mine = 0;
#pragma unroll
            for (int s = 0; s < 4; s++) mine += ((own >> (8 * s)) & 0xFFu) == (uint32_t)o;
            uint32_t incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (mine) {
                uint32_t ooff = 0;   // start of owner o's group inside s_keys / s_cnts
                for (int p = 0; p < o; p++) ooff += s_otot[p];
                uint32_t before = 0;
                for (int w = 0; w < warp; w++) before += s_tot[o][w];
                uint32_t pos = ooff + before + incl - mine;
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if (((own >> (8 * s)) & 0xFFu) == (uint32_t)o) {
                        s_keys[pos] = h[s];
                        s_cnts[pos++] = (uint16_t)c[s];   // <= 32767
                    }
            }
        }
        __syncthreads();
        // ... and out in whole runs, keys and counts as two arrays of the owner's region (10 bytes per entry on NVLink instead of the
        // 16-byte pairs of the first form): consecutive lanes store consecutive 16-byte key granules, then consecutive 8-byte count
        // granules -- full sectors (round 1 stored every lane's pairs from registers: 16-byte writes strided by the lanes' own counts,
        // ~250 GB/s to the peers)
        uint32_t ooff = 0;
        for (int o = 0; o < world; o++) {
            const uint32_t tot = s_otot[o];
            const unsigned long long base = s_base[o];
            if (tot && base != ~0ULL) {
                ulonglong2* outk = (ulonglong2*)(dst.base[o] + base);
                const ulonglong2* ink = (const ulonglong2*)(s_keys + ooff);
                for (uint32_t i = threadIdx.x; i < tot / 2; i += DUMP_THREADS) outk[i] = ink[i];
                uint64_t* outc = (uint64_t*)((uint16_t*)(dst.base[o] + cap_pairs) + base);
                const uint64_t* inc = (const uint64_t*)(s_cnts + ooff);
                for (uint32_t i = threadIdx.x; i < tot / 4; i += DUMP_THREADS) outc[i] = inc[i];
            }
            ooff += tot;
        }
        __syncthreads();   // s_tot / s_base / s_keys / s_cnts are rewritten by the next step
    }
}

// Adds the (hashed key, count) pairs every source left in this rank's regions to the shard; count-0 pairs are padding.
// Each source's pairs arrive in hash order, i.e. in the order of the buckets they update, and every source covers the shard's
// whole hash range.  Round 1 ran one launch per source, each sweeping the complete shard table (N x 8-16 GB of table traffic).
// Now ONE launch walks all sources in lockstep: virtual position v of the longest stream maps to position v * n_s / n_max of
// source s, so at any moment the grid updates the same narrow window of the table for all sources and every table line crosses
// HBM once.  Neighbouring lanes still update neighbouring buckets (the bucket is read right before its CAS: a version with
// several prefetched buckets per thread lost more to stale copies and failed CAS retries than it gained, and a version with a
// run of 8 consecutive pairs per thread was 3.6x slower at N = 4: profiles/r2_NOTES.md).
struct McMergeSrc {
    const uint64_t* keys[MC_MAX_XCHG_RANKS];    // hashed keys ...
    const uint16_t* cnts[MC_MAX_XCHG_RANKS];    // ... and their counts (0 = padding), same positions
    unsigned long long n[MC_MAX_XCHG_RANKS];
    unsigned long long ratio[MC_MAX_XCHG_RANKS];   // floor(n[s] * 2^32 / n_max)
    unsigned long long n_max;
    int world;
};
__global__ void __launch_bounds__(256) k_merge_pairs(McTableDev t, McMergeSrc src) {
    unsigned long long added = 0;
    const uint64_t lmask = (1ULL << (62 - t.shard_bits)) - 1;
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < src.n_max; v += (uint64_t)gridDim.x * blockDim.x) {
        // positions [lo, hi) of source s that fall to virtual position v (0 or 1 pair: n_s <= n_max).  f(v) = (v * ratio_s) >> 32
        // with ratio_s = floor(n_s 2^32 / n_max) <= 2^32 is monotone, f(0) = 0 and f(v) <= n_s; f(n_max) is DEFINED as n_s, so the
        // ranges [f(v), f(v + 1)) partition [0, n_s) exactly.  (v < 2^31: the product fits 64 bits; the host checks n_max.)
        // Round 2's first form divided two 128-bit products per source and position: 1.4 G long divisions per step.
        // The sources are taken four at a time: their (key, count) loads are all issued before the first add (one exposed DRAM
        // round trip per four sources instead of two dependent ones per source).
        for (int s0 = 0; s0 < src.world; s0 += 4) {
            unsigned long long lo[4], hi[4];
            uint64_t key[4];
            uint32_t c[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = s0 + j;
                lo[j] = hi[j] = 0;
                c[j] = 0;
                key[j] = 0;
                if (s < src.world && src.n[s]) {
                    lo[j] = (v * src.ratio[s]) >> 32;
                    hi[j] = v + 1 == src.n_max ? src.n[s] : ((v + 1) * src.ratio[s]) >> 32;
                    if (hi[j] > lo[j]) {
                        c[j] = __ldcs(src.cnts[s] + lo[j]);
                        key[j] = __ldcs(src.keys[s] + lo[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // (one pair per source and key, counts <= 32767: the RED form of the add, mc_common.cuh; count 0 = padding)
                if (c[j] && mc_table_add_impl<false, true>(t, key[j] & lmask, c[j], nullptr)) added++;
                for (unsigned long long i = lo[j] + 1; i < hi[j]; i++) {   // (only the last position of a shorter stream holds more than one)
                    const uint32_t ci = __ldcs(src.cnts[s0 + j] + i);
                    const uint64_t ki = __ldcs(src.keys[s0 + j] + i);
                    if (ci && mc_table_add_impl<false, true>(t, ki & lmask, ci, nullptr)) added++;
                }
            }
        }
    }
    __shared__ unsigned long long s_added;
    if (threadIdx.x == 0) s_added = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&s_added, added);
    __syncthreads();
    if (threadIdx.x == 0 && s_added) atomicAdd(&t.counters[MC_CNT_KEYS], s_added);
}

extern "C" int mc_xchg_dump_owners(mc_xchg* x, const mc_table* local, int slot, uint64_t* n_pairs) {
    try {
        if (!x || !local || slot < 0 || slot > 1) mc_throw(MC_ERR_ARG, "bad argument");
        if (local->dev.shard_bits != 0) mc_throw(MC_ERR_ARG, "exchange: the local table must be unsharded");
        mc_ctx* ctx = x->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        mc_batch_release(const_cast<mc_table*>(local));   // applies pending updates and returns the bins (tens of GB) to the pool: the shard table needs the room
        for (int r = 0; r < x->world; r++)
            if (!x->peer[r]) mc_throw(MC_ERR_ARG, "exchange: rank %d's buffer has not been opened", r);
        MC_CUDA(cudaMemsetAsync(x->d_cursors, 0, (MC_MAX_XCHG_RANKS + 1) * sizeof(unsigned long long), st));
        McBinDst dst{}, mail{};
        dst.use = 1;
        for (int o = 0; o < x->world; o++) {
            dst.base[o] = x->region(x->peer[o], slot, x->rank);
            mail.base[o] = (uint64_t*)x->mailbox(x->peer[o], slot);
        }
        {
            McStageTimer tm(ctx, MC_ST_SCATTER);
            const uint64_t units = local->nb;
            const int grid = (int)std::min<uint64_t>((units + DUMP_THREADS - 1) / DUMP_THREADS, (uint64_t)ctx->sm_count * 4);
            // The scan walks the table in hash order = owner order, and at any moment the whole grid is inside ONE owner's range: started
            // at bucket 0 on every rank, all ranks would send to owner 0 first, then to owner 1, ... -- one receiver's NVLink ingress
            // shared by all senders while the other links idle.  Rank r starts at the range of owner r + 1 instead (the ranges are
            // scanned in the same ascending order inside, so every receiver still sees each source's pairs in its bucket order).
            const uint64_t scan_lo = (uint64_t)((x->rank + 1) % x->world) * (local->nb / (uint64_t)x->world);
            k_dump_owners<<<grid, DUMP_THREADS, 0, st>>>(local->dev, x->shard_bits, dst, x->d_cursors, x->pair_cap(), x->d_flags, scan_lo, local->nb, 0, local->nb);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
        }
        k_post_counts<<<1, 32, 0, st>>>(x->d_cursors, mail, x->world, x->rank);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned int f = 0;
        unsigned long long cur[MC_MAX_XCHG_RANKS + 1];
        MC_CUDA(cudaMemcpyAsync(&f, x->d_flags, sizeof f, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(cur, x->d_cursors, sizeof cur, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        if (f) mc_throw(MC_ERR_NOMEM, "exchange: a receive region overflowed");
        if (n_pairs) {
            *n_pairs = 0;
            for (int o = 0; o < x->world; o++) *n_pairs += cur[o];
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// ---- the dump fused into the local table's last flush -------------------------------------------------------------------------
// The flush applies the bins group by group in hash order; once a group's kernels are queued, every table entry whose home bucket
// lies in the group's hash range is final.  The hook below ships that range to its owners on the context's second stream while the
// main stream goes on with the next group: an NVLink-bound scan beside issue- and shared-memory-bound kernels.  Anything the hook
// cannot follow (first-level spill list, re-insertion of overflowed keys, a batch small enough for the direct sweep) marks the
// attempt failed and the whole table is dumped afterwards instead, as in round 1.
struct McFusedDump {
    mc_xchg* x;
    McBinDst dst;
    cudaEvent_t ready;
};
static void fused_group_done(void* user, mc_table* t, uint32_t b0, uint32_t G, uint32_t P1) {
    McFusedDump* f = (McFusedDump*)user;
    mc_ctx* ctx = t->ctx;
    const uint64_t per = t->nb / P1;                       // nb >= 2^10 > P1: whole buckets per first-level bin
    const uint64_t lo = (uint64_t)b0 * per, hi = (uint64_t)(b0 + G) * per;
    const uint64_t scan_n = std::min<uint64_t>(t->nb, hi - lo + MC_MAX_DISP);
    MC_CUDA(cudaEventRecord(f->ready, ctx->stream));
    MC_CUDA(cudaStreamWaitEvent(ctx->aux_stream, f->ready, 0));
    const int grid = (int)std::min<uint64_t>((scan_n + DUMP_THREADS - 1) / DUMP_THREADS, (uint64_t)ctx->sm_count);   // one CTA per SM: it shares the SM with the flush
    k_dump_owners<<<grid, DUMP_THREADS, 0, ctx->aux_stream>>>(t->dev, f->x->shard_bits, f->dst, f->x->d_cursors, f->x->pair_cap(), f->x->d_flags, lo, scan_n, lo, hi);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
}
static void fused_before_grow(void* user, mc_table* t) {   // the range dumps in flight read the memory the rehash is about to give back
    (void)user;
    MC_CUDA(cudaStreamSynchronize(t->ctx->aux_stream));
}

extern "C" int mc_xchg_flush_and_dump(mc_xchg* x, mc_table* local, int slot, uint64_t* n_pairs, int* fused_out) {
    try {
        if (!x || !local || slot < 0 || slot > 1) mc_throw(MC_ERR_ARG, "bad argument");
        if (local->dev.shard_bits != 0) mc_throw(MC_ERR_ARG, "exchange: the local table must be unsharded");
        mc_ctx* ctx = x->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        for (int r = 0; r < x->world; r++)
            if (!x->peer[r]) mc_throw(MC_ERR_ARG, "exchange: rank %d's buffer has not been opened", r);
        MC_CUDA(cudaMemsetAsync(x->d_cursors, 0, (MC_MAX_XCHG_RANKS + 1) * sizeof(unsigned long long), st));
        MC_CUDA(cudaMemsetAsync(x->d_flags, 0, sizeof(unsigned int), st));
        McFusedDump f{};
        f.x = x;
        f.dst.use = 1;
        McBinDst mail{};
        for (int o = 0; o < x->world; o++) {
            f.dst.base[o] = x->region(x->peer[o], slot, x->rank);
            mail.base[o] = (uint64_t*)x->mailbox(x->peer[o], slot);
        }
        f.ready = mc_event_get(ctx);
        McFlushHook hook{fused_group_done, fused_before_grow, &f, false, 0};
        local->hook = &hook;
        try {
            mc_table_flush(local);
        } catch (...) {
            local->hook = nullptr;
            cudaStreamSynchronize(ctx->aux_stream);
            mc_event_put(ctx, f.ready);
            throw;
        }
        local->hook = nullptr;
        MC_CUDA(cudaStreamSynchronize(ctx->aux_stream));
        mc_event_put(ctx, f.ready);
        mc_batch_release(local);   // returns the bins (tens of GB) to the arena: the shard table needs the room
        unsigned int fl = 0;
        MC_CUDA(cudaMemcpyAsync(&fl, x->d_flags, sizeof fl, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        const bool fused = !hook.failed && hook.groups > 0 && fl == 0;
        if (fused_out) *fused_out = fused ? 1 : 0;
        if (fl) {   // a receive region was too small for this table: the caller agrees on larger regions with the other ranks and dumps again
            if (fused_out) *fused_out = -1;
            if (n_pairs) *n_pairs = 0;
            return MC_OK;
        }
        if (!fused) return mc_xchg_dump_owners(x, local, slot, n_pairs);
        k_post_counts<<<1, 32, 0, st>>>(x->d_cursors, mail, x->world, x->rank);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned long long cur[MC_MAX_XCHG_RANKS + 1];
        MC_CUDA(cudaMemcpyAsync(cur, x->d_cursors, sizeof cur, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        if (n_pairs) {
            *n_pairs = 0;
            for (int o = 0; o < x->world; o++) *n_pairs += cur[o];
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_xchg_merge(mc_xchg* x, int slot, mc_table* shard, uint64_t* n_pairs) {
    try {
        if (!x || !shard || slot < 0 || slot > 1) mc_throw(MC_ERR_ARG, "bad argument");
        mc_ctx* ctx = x->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        mc_table_flush(shard);
        unsigned long long counts[MC_MAX_XCHG_RANKS];
        MC_CUDA(cudaMemcpyAsync(counts, x->mailbox(x->local, slot), sizeof counts, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        uint64_t total = 0;
        for (int src = 0; src < x->world; src++) {
            if (counts[src] > x->pair_cap()) mc_throw(MC_ERR_CUDA, "exchange: mailbox of rank %d is corrupt", src);
            total += counts[src];
        }
        // every pair may be a new key and no overflow list is attached here: room for all of them at load <= 0.55 (at 0.7 -- one rank
        // merging 0.72 G pairs into 2^28 buckets -- the 8-bucket windows fill up and the stash overflows)
        {
            const uint64_t have = mc_table_keys(shard);
            const int q = mc_q_for_keys(have + total, shard->dev.shard_bits, 0.55);
            if (q > shard->dev.q) mc_table_grow_to(shard, q);
        }
        if (total) {
            McMergeSrc ms{};
            ms.world = x->world;
            for (int src = 0; src < x->world; src++) {
                ms.keys[src] = x->region(x->local, slot, src);
                ms.cnts[src] = (const uint16_t*)(ms.keys[src] + x->pair_cap());
                ms.n[src] = counts[src];
                ms.n_max = std::max<unsigned long long>(ms.n_max, counts[src]);
            }
            if (ms.n_max >> 32) mc_throw(MC_ERR_UNSUPPORTED, "exchange: more than 2^32 pairs from one rank in one merge");
            for (int src = 0; src < x->world; src++) ms.ratio[src] = (unsigned long long)(((unsigned __int128)ms.n[src] << 32) / ms.n_max);
            McStageTimer tm(ctx, MC_ST_COUNT);
            const int grid = (int)std::min<uint64_t>((ms.n_max + 255) / 256, (uint64_t)ctx->sm_count * 8);
            k_merge_pairs<<<grid, 256, 0, st>>>(shard->dev, ms);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
        }
        unsigned int f = 0;
        MC_CUDA(cudaMemcpyAsync(&f, shard->dev.flags, sizeof f, cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        if (f & MC_FLAG_OVERFLOW) mc_throw(MC_ERR_NOMEM, "k-mer table overflow while merging the received pairs");
        if (n_pairs) *n_pairs = total;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_table_copy_stats(mc_table* dst, const mc_table* src) {
    try {
        if (!dst || !src) mc_throw(MC_ERR_ARG, "null argument");
        MC_CUDA(cudaSetDevice(dst->ctx->device));
        // counters [2..5]: reads, runs, bases, k-mer instances (mc_table_stats)
        MC_CUDA(cudaMemcpyAsync(&dst->dev.counters[MC_CNT_READS], &src->dev.counters[MC_CNT_READS], 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice,
                                dst->ctx->stream));
        MC_CUDA(cudaStreamSynchronize(dst->ctx->stream));
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
