// Binned table updates: the write-combining front end of the count table (K3 in two phases).
//
//   phase 1  k_bin_scatter : hashed keys -> P bins by the top bits of the hash (one tile of 4096 keys per CTA step,
//                            shared-memory histogram, one global reservation per non-empty bin, scattered 8-byte stores
//                            that the L2 merges into full sectors at each bin's append frontier)
//   phase 2  k_bin_count   : bins applied in order; the table slice of a bin (table_bytes / P) is L2 resident while
//                            the resident CTAs sweep it, so the bucket probe (LDG.E.256) and the count update (RED / CAS)
//                            are L2 hits and each slice crosses the HBM bus once per batch in each direction.
//
// Semantics are those of mc_table_add (BigLong2ShortHashMap.addAndBound(key, 1), ITMO!/structures/map/
// BigLong2ShortHashMap.java:67-77): exact counts saturating at 32767, any order of application gives the same table.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mc_common.cuh"

#define BIN_TILE_DEFAULT 2048u  // keys per k_bin_count work unit (tunable: count_tile)
#ifndef SC_THREADS
#define SC_THREADS 512
#define SC_TILE 8192u           // keys per k_bin_scatter tile (sorted by bin in shared memory)
#define SC_MINB 2               // resident CTAs per SM the register budget is cut for
#endif
#define SC_KPT (SC_TILE / SC_THREADS)   // keys per thread, held in registers between the histogram and the sort

void mc_resolve_timers(mc_ctx* ctx) {
    for (auto& t : ctx->timers) {
        float ms = 0;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) ctx->stage_ms[t.stage] += ms;
        mc_event_put(ctx, t.a);
        mc_event_put(ctx, t.b);
    }
    cudaGetLastError();
    ctx->timers.clear();
}

// --------------------------------------------------------------------------------------------
// phase 1
//
// Measured on B200 (profiles/micro/l2ops.cu): a divergent 8-byte store costs as much as a divergent 32-byte one
// (~6 cycles per request and SM when the target lines are not resident), so the bins are appended to in whole,
// aligned 32-byte sectors only: a tile of SC_TILE keys is counting-sorted by bin in shared memory, every bin's run
// is padded to a multiple of four keys with sentinels, and each lane then moves one full sector.
// --------------------------------------------------------------------------------------------
struct McScatterSmem {   // carved out of dynamic shared memory; P = number of bins
    uint32_t* hist;      // [P]   keys per bin in this tile, then reused as the insertion cursor
    uint32_t* off;       // [P+1] start of each bin's padded run inside `sorted`
    unsigned long long* gbase;   // [P] position reserved in the bin's global run
    uint16_t* secbin;    // [(SC_TILE + 3P) / 4] bin of every sector of `sorted`
    uint64_t* sorted;    // [SC_TILE + 3P]
};
static size_t scatter_smem_bytes(uint32_t P) {
    size_t keys = SC_TILE + 3 * (size_t)P;
    return keys * 8 + (size_t)P * 8 + (size_t)P * 4 + ((size_t)P + 1) * 4 + (keys / 4) * 2 + 64;
}

// ---- bulk asynchronous copies (the TMA unit, sm_90+/sm_100a): shared memory -> global memory ----------------------------------
// A bin's run of the sorted tile is one contiguous, 32-byte aligned piece of shared memory that goes to one contiguous, 32-byte
// aligned piece of the bin: ONE cp.async.bulk per (tile, non-empty bin) hands the whole run to the copy engine, instead of every
// lane moving one sector through registers (2 x LDS.128 + three table look-ups + STG.256 per sector in round 1, on the
// shared-memory pipe that bounds this kernel).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the copies this thread committed have READ their shared-memory source (it may be overwritten); they need not have landed yet
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory (the sorted tile) become visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the copy engine pulls a tile this CTA will read one iteration from now into L2 (no destination: the later LDGs hit L2 instead of HBM)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
#ifndef SC_PREFETCH
#define SC_PREFETCH 1
#endif
#define SC_MAX_PER ((1u << MC_MAX_PBITS) / SC_THREADS)   // bins owned by one thread in the scan / bulk-store phases

// `D.use` = the bins are the OWNER ranks of the multi-GPU exchange: bin o's run goes to D.base[o], the receive region this rank
// owns inside rank o's buffer (peer memory mapped through CUDA IPC: the sector stores travel over NVLink), mc_xchg.cu.
// tiles first_tile, first_tile + tile_stride, ... of `flat` (one CTA).  TMA: runs leave through bulk copies (never with D.use:
// peer-memory destinations keep the per-sector stores).
template <bool TMA>
__device__ __forceinline__ void scatter_tiles(const uint64_t* __restrict__ flat, uint64_t n, const McBinsDev& B, int shard_bits, uint32_t shard_id,
                                              unsigned int* flags, const McBinDst& D, uint32_t first_tile, uint32_t tile_stride) {
    extern __shared__ unsigned long long s_mem[];
    const uint32_t P = 1u << B.pbits;
    McScatterSmem S;
    S.sorted = (uint64_t*)s_mem;
    S.gbase = s_mem + SC_TILE + 3 * P;
    S.hist = (uint32_t*)(S.gbase + P);
    S.off = S.hist + P;
    S.secbin = (uint16_t*)(S.off + P + 1 + ((P + 1) & 1));
    __shared__ uint32_t s_warp_tot[SC_THREADS / 32];
    // A run is padded with sentinels to the granule its stores need: whole 32-byte sectors (4 keys) for the per-sector stores,
    // 16 bytes (2 keys) for the bulk copies -- a third of the pads, i.e. ~6 % fewer bytes through both bin levels and the
    // aggregation that reads them.
    constexpr uint32_t PADM = TMA ? 1u : 3u;
    const int bshift = B.shift;                  // <= 62
    const uint64_t lmask = (1ULL << B.hbits) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // bins are scanned in contiguous groups of `per` per thread
    const uint32_t per = (P + SC_THREADS - 1) / SC_THREADS;
    for (uint64_t t0 = (uint64_t)first_tile * SC_TILE; t0 < n; t0 += (uint64_t)tile_stride * SC_TILE) {
        const uint32_t tn = (uint32_t)(n - t0 < SC_TILE ? n - t0 : SC_TILE);
        if (SC_PREFETCH && TMA && threadIdx.x == 0) {   // my next tile -> L2 while this one is sorted
            const uint64_t t1 = t0 + (uint64_t)tile_stride * SC_TILE;
            if (t1 < n && (((uintptr_t)(flat + t1)) & 15) == 0) {
                const uint64_t nk = n - t1 < SC_TILE ? n - t1 : SC_TILE;
                if (nk >= 2) bulk_prefetch_l2(flat + t1, (uint32_t)(nk & ~1ull) * 8u);
            }
        }
        for (uint32_t i = threadIdx.x; i < P; i += SC_THREADS) S.hist[i] = 0;
        __syncthreads();
        // pass A: the tile is read ONCE, 16 bytes per lane and SC_KPT / 2 loads in flight per thread; every key takes its rank
        // inside its bin from the shared histogram and stays in registers until the run offsets are known
        uint64_t key[SC_KPT];
        uint32_t rk[SC_KPT / 2];   // two 16-bit ranks per word
        const uint64_t* src = flat + t0;
        const bool al16 = ((uintptr_t)src & 15) == 0;
#pragma unroll
        for (int j = 0; j < SC_KPT; j += 2) {
            const uint32_t i = (uint32_t)j * SC_THREADS + threadIdx.x * 2;
            key[j] = key[j + 1] = MC_HASH_SENTINEL;
            if (i + 1 < tn) {
                if (al16) {
                    const ulonglong2 v = __ldcs((const ulonglong2*)(src + i));
                    key[j] = v.x;
                    key[j + 1] = v.y;
                } else {   // a caller's own device buffer (mc_table_insert_hashed_device) may be only 8-byte aligned
                    key[j] = __ldcs(src + i);
                    key[j + 1] = __ldcs(src + i + 1);
                }
            } else if (i < tn) {
                key[j] = __ldcs(src + i);
            }
        }
#pragma unroll
        for (int j = 0; j < SC_KPT; j++) {
            const uint64_t h = key[j];
            const bool valid = h != MC_HASH_SENTINEL && (shard_bits == 0 || (uint32_t)(h >> (62 - shard_bits)) == shard_id);
            uint32_t r = 0;
            if (valid) {
                key[j] = h & lmask;
                r = atomicAdd(&S.hist[(uint32_t)(key[j] >> bshift) & (P - 1)], 1u);
            } else {
                key[j] = MC_HASH_SENTINEL;
            }
            if (j & 1) rk[j >> 1] |= r << 16;
            else rk[j >> 1] = r;
        }
        if (TMA) bulk_wait_read();   // my bulk stores of the previous tile have read `sorted` (they overlapped the loads + histogram above)
        __syncthreads();
        // exclusive scan of the padded run lengths; reserve the global runs; write pads and the sector -> bin map
        {
            const uint32_t b0 = threadIdx.x * per;
            uint32_t mine = 0;
            for (uint32_t b = b0; b < b0 + per && b < P; b++) mine += (S.hist[b] + PADM) & ~PADM;
            uint32_t incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0;
            for (int w = 0; w < warp; w++) wbase += s_warp_tot[w];
            uint32_t run = wbase + incl - mine;
#pragma unroll
            for (uint32_t i = 0; i < SC_MAX_PER; i++) {
                const uint32_t b = b0 + i;
                if (i >= per || b >= P) continue;
                const uint32_t c = S.hist[b], pc = (c + PADM) & ~PADM;
                S.off[b] = run;
                const unsigned long long gpos = pc ? atomicAdd(&B.cursors[b], (unsigned long long)pc) : 0ULL;
                S.gbase[b] = gpos;
                for (uint32_t j = c; j < pc; j++) S.sorted[run + j] = MC_HASH_SENTINEL;
                if (!TMA)
                    for (uint32_t j = 0; j < pc; j += 4) S.secbin[(run + j) >> 2] = (uint16_t)b;
                run += pc;
            }
            if (threadIdx.x == SC_THREADS - 1) S.off[P] = run;   // the last thread owns the last bins
        }
        __syncthreads();
        // pass B: keys to their sorted position
#pragma unroll
        for (int j = 0; j < SC_KPT; j++) {
            const uint64_t hl = key[j];
            if (hl == MC_HASH_SENTINEL) continue;
            const uint32_t bin = (uint32_t)(hl >> bshift) & (P - 1);
            const uint32_t r = (j & 1) ? rk[j >> 1] >> 16 : rk[j >> 1] & 0xFFFFu;
            S.sorted[S.off[bin] + r] = hl;
        }
        // one sector of the sorted tile to its bin, or (bin full: heavy hitter or undersized batch) to the shared spill list
        auto store_sector = [&](uint32_t bin, unsigned long long pos, uint32_t at) {
            const ulonglong2 lo2 = *(const ulonglong2*)&S.sorted[at], hi2 = *(const ulonglong2*)&S.sorted[at + 2];
            const uint64_t k0 = lo2.x, k1 = lo2.y, k2 = hi2.x, k3 = hi2.y;
            if (pos + 4 <= B.cap) {
                uint64_t* dst = (D.use ? D.base[bin] : B.keys + (uint64_t)bin * B.cap) + pos;
                asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(k0), "l"(k1), "l"(k2), "l"(k3) : "memory");
            } else {   // applied with direct updates
                const uint64_t kk[4] = {k0, k1, k2, k3};
                uint32_t real = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) real += kk[q] != MC_HASH_SENTINEL;
                unsigned long long sp = atomicAdd(B.spill_cursor, (unsigned long long)real);   // one reservation per sector
                if (sp + real <= B.spill_cap) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (kk[q] != MC_HASH_SENTINEL) B.spill[sp++] = kk[q];
                } else {
                    atomicOr(flags, MC_FLAG_OVERFLOW);   // unreachable: the host keeps spill_free >= keys per scatter
                }
            }
        };
        if (TMA) {
            // pass C, bulk form: the thread that reserved a bin's run hands it to the copy engine in one piece
            fence_async_smem();
            __syncthreads();
            const uint32_t b0 = threadIdx.x * per;
#pragma unroll
            for (uint32_t i = 0; i < SC_MAX_PER; i++) {
                const uint32_t b = b0 + i;
                if (i >= per || b >= P) continue;
                // (off and gbase stay as the scan left them until the next tile's scan, which is behind a barrier)
                const uint32_t run = S.off[b], pc = S.off[b + 1] - run;
                if (pc == 0) continue;
                const unsigned long long pos = S.gbase[b];
                if (pos + pc <= B.cap) {
                    bulk_store_s2g(B.keys + (uint64_t)b * B.cap + pos, &S.sorted[run], pc * 8u);
                } else {
                    // the run crosses the end of its bin (heavy hitter or undersized batch): what fits goes to the bin in one bulk copy
                    // (pos, cap and pc are even: the piece is a whole number of 16-byte granules), the rest takes the spill list with
                    // ONE reservation for the whole remainder.  (A reservation per key -- a dependent global atomic each, issued by one
                    // thread -- made the 3 000 repeat k-mers of cfg4, 8 % of its instances, cost more than everything else together.)
                    const uint32_t fits = pos < B.cap ? (uint32_t)(B.cap - pos) : 0u;   // < pc
                    if (fits) bulk_store_s2g(B.keys + (uint64_t)b * B.cap + pos, &S.sorted[run], fits * 8u);
                    const uint32_t real = pc - (S.sorted[run + pc - 1] == MC_HASH_SENTINEL ? 1u : 0u);   // at most one pad, at the end
                    if (real > fits) {
                        const uint32_t rest = real - fits;
                        const unsigned long long sp = atomicAdd(B.spill_cursor, (unsigned long long)rest);
                        if (sp + rest <= B.spill_cap) {
                            for (uint32_t j = 0; j < rest; j++) B.spill[sp + j] = S.sorted[run + fits + j];
                        } else {
                            atomicOr(flags, MC_FLAG_OVERFLOW);   // unreachable: the host keeps spill_free >= keys per scatter
                        }
                    }
                }
            }
            bulk_commit();
            // no barrier here: the next tile's loads and histogram touch neither `sorted` nor my registers' runs; the wait + barrier
            // in front of the next scan phase orders the overwrite of `sorted` after these copies have read it
        } else {
            __syncthreads();
            // pass C: one aligned 32-byte sector per lane
            const uint32_t n_sec = S.off[P] >> 2;
            for (uint32_t j = threadIdx.x; j < n_sec; j += SC_THREADS) {
                const uint32_t bin = S.secbin[j];
                store_sector(bin, S.gbase[bin] + (j * 4 - S.off[bin]), j * 4);
            }
            __syncthreads();
        }
    }
}

template <bool TMA>
__global__ void __launch_bounds__(SC_THREADS, SC_MINB) k_bin_scatter(const uint64_t* __restrict__ flat, uint64_t n, McBinsDev B, int shard_bits, uint32_t shard_id,
                                                            unsigned int* flags, McBinDst D) {
    scatter_tiles<TMA>(flat, n, B, shard_bits, shard_id, flags, D, blockIdx.x, gridDim.x);
    if (TMA) bulk_wait_read();   // shared memory must outlive the copy engine's reads
}

// Second level: first-level bins b0 .. b0 + G - 1, each split into 2^B2.pbits sub-bins by the hash bits below the
// first-level prefix.  Group slot g owns the sub-bin region and cursor block g of B2; the run lengths of the first level
// are read on the device (no host round trip); the spill list is shared by the whole group.  The (bin, tile) pairs are
// one flat index space walked by a persistent grid, so the CTAs work on the same one or two bins at any time.
template <bool TMA>
__global__ void __launch_bounds__(SC_THREADS, SC_MINB) k_bin_scatter2(McBinsDev B1, uint32_t b0, uint32_t G, uint32_t tiles_per_bin, McBinsDev B2, unsigned int* flags) {
    const uint32_t P2 = 1u << B2.pbits;
    McBinDst none;
    none.use = 0;
    for (uint32_t vt = blockIdx.x; vt < G * tiles_per_bin; vt += gridDim.x) {
        const uint32_t g = vt / tiles_per_bin, tile = vt - g * tiles_per_bin;
        const unsigned long long attempts = B1.cursors[b0 + g];
        const uint64_t n = attempts < B1.cap ? attempts : B1.cap;
        if ((uint64_t)tile * SC_TILE >= n) continue;
        if (SC_PREFETCH && TMA && threadIdx.x == 0) {   // the tile of my next (bin, tile) pair -> L2
            const uint32_t vn = vt + gridDim.x;
            if (vn < G * tiles_per_bin) {
                const uint32_t gn = vn / tiles_per_bin, tl = vn - gn * tiles_per_bin;
                const unsigned long long an = B1.cursors[b0 + gn];
                const uint64_t nn = an < B1.cap ? an : B1.cap;
                const uint64_t at = (uint64_t)tl * SC_TILE;
                if (at < nn) {
                    const uint64_t nk = nn - at < SC_TILE ? nn - at : SC_TILE;
                    if (nk >= 2) bulk_prefetch_l2(B1.keys + (uint64_t)(b0 + gn) * B1.cap + at, (uint32_t)(nk & ~1ull) * 8u);
                }
            }
        }
        McBinsDev S = B2;
        S.keys = B2.keys + (uint64_t)g * P2 * B2.cap;
        S.cursors = B2.cursors + (uint64_t)g * (P2 + 1);
        scatter_tiles<TMA>(B1.keys + (uint64_t)(b0 + g) * B1.cap, n, S, 0, 0, flags, none, tile, 0x40000000u);
    }
    if (TMA) bulk_wait_read();
}

// --------------------------------------------------------------------------------------------
// phase 2
// --------------------------------------------------------------------------------------------
// U = bucket probes a thread keeps in flight, MINB = CTAs per SM the register budget is cut for
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_bin_count(McTableDev t, McBinsDev B, const uint64_t* __restrict__ tile_prefix, uint64_t tile_lo, uint64_t tile_hi,
                                                          uint32_t BIN_TILE, unsigned long long* ticket) {
    const uint32_t P = 1u << B.pbits;
    unsigned long long added = 0;
    __shared__ unsigned long long s_tile;
    for (;;) {
        // Tiles are handed out in order through a global ticket, one per CTA step: the resident CTAs then always work on
        // the lowest unprocessed tiles, so the set of table lines being updated (the sweep window) stays narrow and L2
        // resident.  Without the barrier the warps of a CTA drift tiles apart and the L2 hit rate collapses (measured: 4.5x).
        __syncthreads();
        if (threadIdx.x == 0) s_tile = tile_lo + atomicAdd(ticket, 1ULL);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= tile_hi) break;
        // the bin of this tile: largest b with tile_prefix[b] <= tile (empty bins have no tiles)
        uint32_t lo = 0, hi = P;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(tile_prefix + mid) <= tile) lo = mid;
            else hi = mid;
        }
        const unsigned long long attempts = B.cursors[lo];
        const uint64_t n_bin = attempts < B.cap ? attempts : B.cap;
        const uint64_t start = (tile - __ldg(tile_prefix + lo)) * BIN_TILE;
        const uint64_t* keys = B.keys + (uint64_t)lo * B.cap + start;
        const uint32_t n = (uint32_t)(n_bin - start < BIN_TILE ? n_bin - start : BIN_TILE);
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += 256 * U) {
            uint64_t h[U];
            uint64_t w[U][4];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t idx = base + u * 256 + threadIdx.x;
                h[u] = idx < n ? __ldcs(keys + idx) : MC_HASH_SENTINEL;   // sentinels also pad the sector-aligned runs
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (h[u] != MC_HASH_SENTINEL) mc_load_bucket(t.slots + ((h[u] >> t.tag_bits) & t.nb_mask) * 4, w[u]);
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (h[u] == MC_HASH_SENTINEL) continue;
                // fast path: the key sits in its home bucket (slots fill in order and are never freed, so a match always
                // precedes the first empty slot) -> one RED; everything else takes the general lock-free insert
                const uint64_t want = (h[u] & t.tag_mask) << MC_DISP_BITS;   // tag | disp 0
                uint64_t* bucket = t.slots + ((h[u] >> t.tag_bits) & t.nb_mask) * 4;
                int hit = -1;
                uint64_t wm = 0;
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if ((w[u][s] >> MC_COUNT_BITS) == want && w[u][s] != 0) {
                        hit = s;
                        wm = w[u][s];
                    }
                if (hit >= 0) {
                    if ((wm & MC_COUNT_MASK) < MC_SAT) atomicAdd((unsigned long long*)(bucket + hit), 1ULL);
                } else if (mc_table_add_impl<true>(t, h[u], 1, w[u])) {
                    added++;
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

// --------------------------------------------------------------------------------------------
// phase 2, aggregated: one CTA per sub-bin ("unit").  A unit is small enough (~24 K keys) for its DISTINCT keys to fit a
// shared-memory hash table, so the repeats of a k-mer -- most of a metagenome's k-mer instances -- are summed in shared
// memory and the count table in HBM receives ONE update per distinct key and batch (insert-or-add of the sum) instead of
// one L2 probe + RED per instance.  Whatever does not fit the shared table (probe limit) is applied directly, so the
// kernel is exact for any input; saturating adds commute, so the result does not depend on the grouping.
// --------------------------------------------------------------------------------------------
#define AGG_THREADS 512
#define AGG_SLOTS 8192u      // 8 B key + 4 B count + 2 B list entry = 112 KB per CTA: two CTAs per SM
#ifndef AGG_PROBES
#define AGG_PROBES 12
#endif
#define AGG_U 4       // key loads in flight per thread (double buffered)
#ifndef AGG_FU
#define AGG_FU 1      // bucket probes in flight per thread in the flush (measured: 1 -> 45.3 ms, 2 -> 46.2, 3 -> 50.5 at cfg2)
#endif
#define AGG_PREFETCH_MAX (512u << 10)   // table slices up to this size are pulled into L2 when their unit starts
static size_t agg_smem_bytes() { return (size_t)AGG_SLOTS * 14; }

// COUNT_ONLY: the same aggregation, but the table is only looked up: the kernel reports how many distinct keys of these units are
// NOT in the table yet (counters[MC_CNT_OVF_KEYS]).  Run on the first group -- 1/64 of the hash space -- it tells the host how
// large the table must be for the whole batch before anything is inserted.
// LANE_LOOP: every lane walks its AGG_U keys of a round in ONE loop (a lane that has placed a key goes on to its next key at once),
// so a warp iterates max-over-lanes of the SUM of the lanes' probe counts instead of the sum over the keys of the max: at the
// table's load the probe sequences are short (1.2 probes per key) but their maximum over 32 lanes is 4-5, and the per-key loops
// ran with 8 of 32 lanes active on average (profiles/r2u_extract_agg.ncu-rep, SASS view).
template <bool COUNT_ONLY, bool LANE_LOOP>
__global__ void __launch_bounds__(AGG_THREADS, 2) k_bin_agg(McTableDev t, McBinsDev B2, uint32_t b0, uint32_t n_units, unsigned long long* ticket) {
    extern __shared__ unsigned long long a_mem[];
    unsigned long long* s_key = a_mem;                       // hashed local key + 1, 0 = empty
    uint32_t* s_cnt = (uint32_t*)(a_mem + AGG_SLOTS);
    uint16_t* s_list = (uint16_t*)(s_cnt + AGG_SLOTS);       // claimed slots in claim order: the flush walks (and clears) only these
    __shared__ unsigned long long s_unit;
    __shared__ uint32_t s_used;
    const uint32_t P2 = 1u << B2.pbits;
    const int pbt = B2.hbits - B2.shift;                     // prefix bits shared by the keys of a unit
    unsigned long long added = 0;
    for (uint32_t i = threadIdx.x; i < AGG_SLOTS; i += AGG_THREADS) {
        s_key[i] = 0;
        s_cnt[i] = 0;
    }
    if (threadIdx.x == 0) s_used = 0;
    for (;;) {
        __syncthreads();                                     // the previous unit's flush is complete (table empty again)
        if (threadIdx.x == 0) s_unit = atomicAdd(ticket, 1ULL);
        __syncthreads();
        if (s_unit >= n_units) break;
        const uint32_t unit = (uint32_t)s_unit;
        const uint32_t g = unit >> B2.pbits, sub = unit & (P2 - 1);
        const unsigned long long attempts = B2.cursors[(uint64_t)g * (P2 + 1) + sub];
        const uint32_t n = (uint32_t)(attempts < B2.cap ? attempts : B2.cap);
        const uint64_t* keys = B2.keys + (uint64_t)unit * B2.cap;
        if (n && pbt <= t.q) {
            // the unit's keys all live in one contiguous table slice: start pulling it into L2 now, the flush probes it later
            const uint64_t prefix = ((uint64_t)(b0 + g) << B2.pbits) | sub;
            const uint64_t bytes = (32ull << (t.q - pbt));
            if (bytes <= AGG_PREFETCH_MAX) {
                const char* base = (const char*)(t.slots + ((prefix << (t.q - pbt)) & t.nb_mask) * 4);
                for (uint64_t off = (uint64_t)threadIdx.x * 128; off < bytes; off += (uint64_t)AGG_THREADS * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
            }
        }
        uint64_t h[AGG_U];
#pragma unroll
        for (int u = 0; u < AGG_U; u++) {
            const uint32_t idx = u * AGG_THREADS + threadIdx.x;
            h[u] = idx < n ? __ldcs(keys + idx) : MC_HASH_SENTINEL;
        }
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += AGG_THREADS * AGG_U) {
            uint64_t hn[AGG_U];
#pragma unroll
            for (int u = 0; u < AGG_U; u++) {                // the next round's keys are in flight while this round is aggregated
                const uint32_t idx = base + (AGG_U + u) * AGG_THREADS + threadIdx.x;
                hn[u] = idx < n ? __ldcs(keys + idx) : MC_HASH_SENTINEL;
            }
            if (LANE_LOOP) {
                uint32_t u = 0, p = 0;
                uint64_t key = h[0];
                uint32_t slot = (uint32_t)key & (AGG_SLOTS - 1);
#pragma unroll 1
                while (u < AGG_U) {
                    bool adv = true;
                    if (key != MC_HASH_SENTINEL) {
                        const unsigned long long want = key + 1;
                        unsigned long long cur = *(volatile unsigned long long*)&s_key[slot];
                        if (cur == 0) {
                            cur = atomicCAS(&s_key[slot], 0ULL, want);
                            if (cur == 0) {
                                s_list[atomicAdd(&s_used, 1u)] = (uint16_t)slot;
                                cur = want;
                            }
                        }
                        if (cur == want) {
                            atomicAdd(&s_cnt[slot], 1u);
                        } else if (++p < AGG_PROBES) {
                            slot = (slot + 1) & (AGG_SLOTS - 1);
                            adv = false;
                        } else {
                            if (COUNT_ONLY) added += mc_table_lookup(t, key) < 0;
                            else if (mc_table_add_impl<false>(t, key, 1, nullptr)) added++;
                        }
                    }
                    if (adv) {
                        u++;
                        key = u == 1 ? h[1] : (u == 2 ? h[2] : h[3]);
                        slot = (uint32_t)key & (AGG_SLOTS - 1);
                        p = 0;
                    }
                }
            } else
#pragma unroll
            for (int u = 0; u < AGG_U; u++) {
                if (h[u] == MC_HASH_SENTINEL) continue;
                const unsigned long long want = h[u] + 1;
                uint32_t slot = (uint32_t)h[u] & (AGG_SLOTS - 1);   // low bits of the mixed hash: independent of the bin prefix
                bool done = false;
#pragma unroll 1
                for (int p = 0; p < AGG_PROBES; p++, slot = (slot + 1) & (AGG_SLOTS - 1)) {
                    unsigned long long cur = *(volatile unsigned long long*)&s_key[slot];
                    if (cur == 0) {
                        cur = atomicCAS(&s_key[slot], 0ULL, want);
                        if (cur == 0) {
                            s_list[atomicAdd(&s_used, 1u)] = (uint16_t)slot;
                            cur = want;
                        }
                    }
                    if (cur == want) {
                        atomicAdd(&s_cnt[slot], 1u);
                        done = true;
                        break;
                    }
                }
                if (!done) {
                    if (COUNT_ONLY) added += mc_table_lookup(t, h[u]) < 0;
                    else if (mc_table_add_impl<false>(t, h[u], 1, nullptr)) added++;
                }
            }
#pragma unroll
            for (int u = 0; u < AGG_U; u++) h[u] = hn[u];
        }
        __syncthreads();
        // flush: one insert-or-add per distinct key, AGG_FU bucket probes in flight per thread; the slots are emptied on the way
        const uint32_t used = s_used;
#pragma unroll 1
        for (uint32_t base = 0; base < used; base += AGG_THREADS * AGG_FU) {
            uint64_t hk[AGG_FU];
            uint32_t c[AGG_FU];
            uint64_t w[AGG_FU][4];
#pragma unroll
            for (int u = 0; u < AGG_FU; u++) {
                const uint32_t i = base + u * AGG_THREADS + threadIdx.x;
                hk[u] = MC_HASH_SENTINEL;
                c[u] = 0;
                if (i < used) {
                    const uint32_t slot = s_list[i];
                    hk[u] = s_key[slot] - 1;
                    c[u] = s_cnt[slot] < MC_SAT ? s_cnt[slot] : MC_SAT;   // min(min(a, S) + b, S) == min(a + b, S)
                    s_key[slot] = 0;
                    s_cnt[slot] = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < AGG_FU; u++)
                if (hk[u] != MC_HASH_SENTINEL) mc_load_bucket(t.slots + ((hk[u] >> t.tag_bits) & t.nb_mask) * 4, w[u]);
#pragma unroll
            for (int u = 0; u < AGG_FU; u++) {
                if (hk[u] == MC_HASH_SENTINEL) continue;
                if (COUNT_ONLY) added += mc_table_lookup(t, hk[u]) < 0;
                else if (mc_table_add_impl<true>(t, hk[u], c[u], w[u])) added++;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_used = 0;
    }
    __shared__ unsigned long long s_added;
    if (threadIdx.x == 0) s_added = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&s_added, added);
    __syncthreads();
    if (threadIdx.x == 0 && s_added) atomicAdd(&t.counters[COUNT_ONLY ? MC_CNT_OVF_KEYS : MC_CNT_KEYS], s_added);
}

typedef void (*McBinCountKernel)(McTableDev, McBinsDev, const uint64_t*, uint64_t, uint64_t, uint32_t, unsigned long long*);
static McBinCountKernel bin_count_variant(int v) {
    switch (v) {
        case 1: return k_bin_count<4, 4>;
        case 2: return k_bin_count<2, 6>;
        case 3: return k_bin_count<4, 3>;
        default: return k_bin_count<2, 4>;   // best of the sweep in profiles/r1c_NOTES.md
    }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
// Free device memory + what the context's arena holds for reuse.  cudaMemGetInfo goes to the kernel-mode driver, where it
// queues behind whatever else holds the resource-manager lock (an NVML / nvidia-smi poll takes it for milliseconds): measured on the
// shared boxes as 20 ms "reserve" phases in otherwise identical steps.  The answer only sizes buffers (never more than a fraction
// of it is asked for, and a failed allocation releases the arena and retries), so it is cached and refreshed only after the
// library itself went to the driver for memory.
static uint64_t available_device_bytes(mc_ctx* ctx) {
    if (ctx->free_bytes_cached == 0 || ctx->free_bytes_epoch != ctx->arena_driver_allocs + ctx->scratch_allocs) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return (uint64_t)1 << 30; }
        ctx->free_bytes_cached = free_b ? free_b : 1;
        ctx->free_bytes_epoch = ctx->arena_driver_allocs + ctx->scratch_allocs;
    }
    return (uint64_t)ctx->free_bytes_cached + mc_arena_cached_bytes(ctx);   // blocks cached by the arena are reusable (or released on demand)
}

uint64_t mc_available_device_bytes(mc_ctx* ctx) { return available_device_bytes(ctx); }

// a predicted table size is an estimate: never ask for more than 80 % of what the device can still give.  A table that cannot
// double any more simply runs fuller (4-slot buckets probed over 8-bucket windows stay workable to a load of ~0.8; past that
// the doubling is attempted anyway and fails with MC_ERR_NOMEM if it really does not fit).
static int fit_q(mc_ctx* ctx, const mc_table* t, int q, uint64_t keys_known = 0) {
    const double room = 0.8 * (double)available_device_bytes(ctx);
    while (q > t->dev.q && (double)(32ull << q) > room) q--;
    if (q == t->dev.q && (double)keys_known > 0.8 * 4.0 * (double)t->nb) q = t->dev.q + 1;
    return q;
}
// amortised doubling at load MC_GROW_LOAD, when memory allows
static void grow_if_due(mc_table* t, uint64_t keys_known) {
    extern const double MC_GROW_LOAD;
    if ((double)keys_known <= MC_GROW_LOAD * 4.0 * (double)t->nb) return;
    const int q = fit_q(t->ctx, t, t->dev.q + 1, keys_known);
    if (q > t->dev.q) mc_table_grow_to(t, q);
}

static void batch_free_buffers(mc_ctx* ctx, McBatch* B) {
    mc_dev_free(ctx, B->dev.keys);
    mc_dev_free(ctx, B->dev.cursors);
    mc_dev_free(ctx, B->dev.spill);
    B->dev.keys = nullptr;
    B->dev.cursors = nullptr;
    B->dev.spill = nullptr;
}

static void batch_create(mc_table* t, uint64_t ub, uint64_t total_hint) {
    mc_ctx* ctx = t->ctx;
    McBatch* B = t->batch ? t->batch : (t->batch = new McBatch());
    // capacity: everything the caller expects to add when memory allows, at most half of what is free now
    uint64_t want = total_hint ? std::max(total_hint, ub) : std::max<uint64_t>(ub * 16, (uint64_t)1 << 28);
    if (ctx->tune_batch_keys) want = ctx->tune_batch_keys;
    const uint64_t budget = available_device_bytes(ctx) / 2;
    const uint64_t spill_bytes = (ub + 1024) * 8;
    uint64_t room = budget > spill_bytes ? (budget - spill_bytes) / 10 : 0;   // 8 bytes per key + slack for pads
    uint64_t capacity = std::min(want, std::max<uint64_t>(room, ub / 8 + 1024));
    if (capacity < 1024) capacity = 1024;
    // bins: the table slice of one bin (table_bytes / P) must sit in L2 while it is swept -> <= 64 MiB; and a bin should hold
    // ~2 M keys or more (SC_TILE / P keys per bin and tile are padded to a multiple of 4: 10 % waste at P = 512)
    int pbits = 0;
    while (pbits < 9 && ((uint64_t)1 << (pbits + 21)) < capacity) pbits++;
    while (pbits < 10 && (t->nb * 32) >> pbits > ((uint64_t)64 << 20)) pbits++;
    if (ctx->tune_part_bits >= 0) pbits = (int)std::min<long>(ctx->tune_part_bits, MC_MAX_PBITS);
    const int hbits = 62 - t->dev.shard_bits;
    if (pbits > hbits) pbits = hbits;
    B->P = 1u << pbits;
    B->capacity = capacity;
    // per-bin room: the mean + 3 % + the expected sentinel pads + a constant; whatever exceeds it takes the spill list
    uint64_t cap = capacity / B->P;
    cap = cap + cap / 32 + (capacity / SC_TILE + 1) * 8 / 5 + 4096;   // pads: 1.5 expected per bin and tile
    cap = (cap + 15) & ~15ull;
    B->dev.pbits = pbits;
    B->dev.hbits = hbits;
    B->dev.shift = hbits - pbits;
    B->dev.cap = cap;
    B->dev.spill_cap = ub + 1024;
    B->dev.keys = (uint64_t*)mc_dev_alloc(ctx, (size_t)B->P * cap * 8, "k-mer bins");
    B->dev.spill = (uint64_t*)mc_dev_alloc(ctx, (size_t)B->dev.spill_cap * 8, "k-mer spill list");
    // cursors [P+1] followed by the tile prefix [P+1] used by k_bin_count
    B->dev.cursors = (unsigned long long*)mc_dev_alloc(ctx, (size_t)(B->P + 1) * 16 + 64, "bin cursors");   // + the tile ticket
    B->dev.spill_cursor = B->dev.cursors + B->P;
    MC_CUDA(cudaMemsetAsync(B->dev.cursors, 0, (size_t)(B->P + 1) * 8, ctx->stream));
    B->used = 0;
    B->spill_used = 0;
    B->ub_sum = 0;
}

static uint64_t read_spill_cursor(mc_table* t) {
    McBatch* B = t->batch;
    unsigned long long v = 0;
    MC_CUDA(cudaMemcpyAsync(&v, &B->dev.cursors[B->P], sizeof v, cudaMemcpyDeviceToHost, t->ctx->stream));
    MC_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return v;
}

void mc_batch_reserve(mc_table* t, uint64_t ub, uint64_t total_hint) {
    McBatch* B = t->batch;
    if (B && B->dev.keys) {
        if (B->used) B->spill_used = read_spill_cursor(t);
        // Room in the bins is planned with the density seen so far (+ 5 %) rather than with the chunk's worst case (one k-mer per two
        // input bytes: 30 % above a real FASTQ file's), which ended batches one chunk early and cost a second table sizing + rehash;
        // a bin that does fill up diverts to the spill list, whose room IS checked against the worst case.
        uint64_t expect = ub;
        if (B->ub_sum) expect = std::min<uint64_t>(ub, (uint64_t)((double)ub * ((double)B->used / (double)B->ub_sum) * 1.05) + 1024);
        const bool fits = B->used + expect <= B->capacity && B->spill_used + ub <= B->dev.spill_cap;
        B->last_ub = ub;
        if (fits) return;
        mc_table_flush(t);
        if (ub <= B->dev.spill_cap) return;
        batch_free_buffers(t->ctx, B);   // a larger chunk than the buffers were made for
    }
    batch_create(t, ub, total_hint);
    t->batch->last_ub = ub;
}

static void scatter_attrs(mc_ctx* ctx) {   // per context (= per device): opt in to > 48 KB of dynamic shared memory
    if (!mc_attr_once(ctx, MC_ATTR_SCATTER)) return;
    const int smem = (int)scatter_smem_bytes(1u << MC_MAX_PBITS);
    MC_CUDA(cudaFuncSetAttribute(k_bin_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MC_CUDA(cudaFuncSetAttribute(k_bin_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MC_CUDA(cudaFuncSetAttribute(k_bin_scatter2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MC_CUDA(cudaFuncSetAttribute(k_bin_scatter2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
}

void mc_batch_scatter(mc_table* t, const uint64_t* d_flat, uint64_t n) {
    if (n == 0) return;
    mc_ctx* ctx = t->ctx;
    McBatch* B = t->batch;
    if (!B || !B->dev.keys || B->dev.spill_cap < n) mc_throw(MC_ERR_ARG, "internal: scatter without a reserved batch");
    McStageTimer tm(ctx, MC_ST_SCATTER);
    const size_t smem = scatter_smem_bytes(B->P);
    scatter_attrs(ctx);
    const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 8);
    McBinDst none{};
    if (ctx->tune_scatter_tma) k_bin_scatter<true><<<grid, SC_THREADS, smem, ctx->stream>>>(d_flat, n, B->dev, t->dev.shard_bits, t->shard_id, t->dev.flags, none);
    else k_bin_scatter<false><<<grid, SC_THREADS, smem, ctx->stream>>>(d_flat, n, B->dev, t->dev.shard_bits, t->shard_id, t->dev.flags, none);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    B->used += n;
    B->ub_sum += std::max<uint64_t>(B->last_ub, n);
}

void mc_launch_owner_scatter(mc_ctx* ctx, const uint64_t* d_flat, uint64_t n, const McBinsDev& owners, const McBinDst& dst, unsigned int* flags) {
    if (n == 0) return;
    const uint32_t P = 1u << owners.pbits;
    const size_t smem = scatter_smem_bytes(P);
    scatter_attrs(ctx);
    const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 8);
    k_bin_scatter<false><<<grid, SC_THREADS, smem, ctx->stream>>>(d_flat, n, owners, 0, 0, flags, dst);   // peer-memory destinations: per-sector stores
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) k_insert_hashed(McTableDev t, const uint64_t* hkeys, size_t n);
__global__ void __launch_bounds__(256) k_insert_hashed_n(McTableDev t, const uint64_t* hkeys, const unsigned long long* n_ptr, unsigned long long cap);

static const double MC_FLUSH_TARGET_LOAD = 0.50;
static const uint64_t MC_FLUSH_SEGMENT = 128ull << 20;   // keys per k_bin_count launch (and entries of the overflow list)   // table load aimed at when a flush has to grow the table
extern const double MC_GROW_LOAD;
extern const uint64_t MC_OVF_MAX;

// direct updates from a list of hashed keys (spill lists: heavy hitters that overflowed their bin)
static void insert_list(mc_table* t, const uint64_t* d_keys, uint64_t count) {
    mc_ctx* ctx = t->ctx;
    for (uint64_t off = 0; off < count;) {
        const uint64_t n = std::min<uint64_t>(count - off, MC_OVF_MAX);
        mc_table_prepare_chunk(t, n);
        {
            McStageTimer tm(ctx, MC_ST_COUNT);
            const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
            k_insert_hashed<<<grid, 256, 0, ctx->stream>>>(t->dev, d_keys + off, n);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
        }
        mc_table_finish_chunk(t, false);
        off += n;
    }
}
// the first-level spill list, then the batch is empty again
static void apply_spill_list(mc_table* t, uint64_t sp) {
    McBatch* B = t->batch;
    insert_list(t, B->dev.spill, sp);
    MC_CUDA(cudaMemsetAsync(B->dev.cursors, 0, (size_t)(B->P + 1) * 8, t->ctx->stream));
    B->used = 0;
    B->spill_used = 0;
    B->ub_sum = 0;
    B->inst_equiv = 0;
}

static const uint64_t MC_AGG_MIN_KEYS = 1ull << 25;   // batches below this are swept directly (k_bin_count)
static const uint32_t MC_AGG_UNIT_KEYS = 32768;       // target keys per aggregation unit (sentinel pads included): 4 x AGG_SLOTS

// Groups of first-level bins (<= MC_FLUSH_SEGMENT keys): k_bin_scatter2 splits every bin of the group into 2^pb2 sub-bins
// in a staging buffer, k_bin_agg sums each sub-bin in shared memory and applies the sums.  Table sizing is the same as in
// the direct sweep: the first group is a sample of the hash space, and the overflow list has room for every instance.
static void flush_aggregated(mc_table* t, const std::vector<unsigned long long>& cur, uint64_t batch_keys) {
    McBatch* B = t->batch;
    mc_ctx* ctx = t->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t P1 = B->P;
    const uint64_t cap1 = B->dev.cap;
    // the caller's expectation (mc_table_expect_keys) is in k-mer INSTANCES; keys that came out of de-duplicated super-k-mer records
    // stand for more instances than there are keys
    const uint64_t batch_inst = std::max<uint64_t>(B->inst_equiv, batch_keys);
    uint64_t max_bin = 0;
    for (uint32_t b = 0; b < P1; b++) max_bin = std::max<uint64_t>(max_bin, std::min<uint64_t>(cur[b], cap1));
    if (max_bin == 0) return;
    const uint32_t unit_keys = ctx->tune_agg_unit ? ctx->tune_agg_unit : MC_AGG_UNIT_KEYS;
    int pb2 = 0;
    while (pb2 < MC_MAX_PBITS && pb2 < B->dev.shift && (max_bin >> pb2) > unit_keys) pb2++;
    const uint32_t P2 = 1u << pb2;
    // sub-bin room: the mean of the fullest bin + 50 % + the sentinel pads of its tiles + a constant; the rest spills
    uint64_t cap2 = max_bin / P2;
    cap2 = cap2 + cap2 / 2 + (max_bin / SC_TILE + 1) * 8 / 5 + 1024;   // 50 %: a k-mer seen 1000 x is a tenth of a unit
    cap2 = (cap2 + 15) & ~15ull;
    const uint64_t seg_keys = ctx->tune_segment_keys ? std::min<uint64_t>(ctx->tune_segment_keys, MC_FLUSH_SEGMENT) : MC_FLUSH_SEGMENT;
    const uint32_t g_max = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(P1, seg_keys / max_bin));
    // (sized by the segment, not by g_max * max_bin: the product moves by a few MiB from one batch to the next, and a request just above
    // the cached block of the previous flush would go to the driver -- one cudaMalloc of 1 GiB inside a step)
    const uint64_t spill2_cap = (g_max == P1 ? (uint64_t)g_max * max_bin : std::max<uint64_t>(seg_keys, max_bin)) + 1024;
    // Two staging slots: while k_bin_agg works on the sub-bins of group g (slot g & 1) on the main stream, k_bin_scatter2 fills the
    // slot of group g + 1 on the auxiliary stream.  The two kernels lean on different pipes (the scatter on the shared-memory pipe and
    // its barriers, the aggregation on instruction issue), so run side by side, one CTA per SM each, they finish sooner than one after
    // the other with two CTAs per SM each.  "flush_overlap" = 0 gives the serial schedule (one slot).
    const bool overlap = ctx->tune_flush_overlap != 0 && P1 > 64;
    const int n_slots = overlap ? 2 : 1;
    McBinsDev B2s[2] = {};
    const size_t n_cursors = (size_t)g_max * (P2 + 1) + 2;   // + (unused) + the unit ticket
    struct Staging {   // returned to the arena, and the overflow list detached, also when a step below throws
        mc_ctx* ctx;
        mc_table* t;
        McBinsDev* b;
        int n;
        ~Staging() {
            cudaStreamSynchronize(ctx->aux_stream);
            t->dev.ovf_keys = nullptr;
            t->dev.ovf_cap = 0;
            for (int i = 0; i < n; i++) {
                mc_dev_free(ctx, b[i].keys);
                mc_dev_free(ctx, b[i].spill);
                mc_dev_free(ctx, b[i].cursors);
            }
        }
    } staging{ctx, t, B2s, n_slots};
    for (int i = 0; i < n_slots; i++) {
        McBinsDev& b = B2s[i];
        b.pbits = pb2;
        b.hbits = B->dev.hbits;
        b.shift = B->dev.shift - pb2;
        b.cap = cap2;
        b.spill_cap = spill2_cap;
        b.keys = (uint64_t*)mc_dev_alloc(ctx, (size_t)g_max * P2 * cap2 * 8, "k-mer sub-bins");
        b.spill = (uint64_t*)mc_dev_alloc(ctx, (size_t)spill2_cap * 8, "k-mer sub-bin spill list");
        b.cursors = (unsigned long long*)mc_dev_alloc(ctx, n_cursors * 8, "sub-bin cursors");
        b.spill_cursor = b.cursors + (n_cursors - 2);   // the slot's own spill cursor (copied into the status block after the group)
    }
    scatter_attrs(ctx);
    if (mc_attr_once(ctx, MC_ATTR_AGG)) {
        MC_CUDA(cudaFuncSetAttribute(k_bin_agg<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem_bytes()));
        MC_CUDA(cudaFuncSetAttribute(k_bin_agg<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem_bytes()));
        MC_CUDA(cudaFuncSetAttribute(k_bin_agg<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem_bytes()));
        MC_CUDA(cudaFuncSetAttribute(k_bin_agg<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem_bytes()));
    }
    int agg_per_sm = 0, sc_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&agg_per_sm, k_bin_agg<false, true>, AGG_THREADS, agg_smem_bytes()) != cudaSuccess || agg_per_sm < 1) agg_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sc_per_sm, k_bin_scatter2<true>, SC_THREADS, scatter_smem_bytes(P2)) != cudaSuccess || sc_per_sm < 1) sc_per_sm = 1;
    cudaGetLastError();

    // The host runs ONE group ahead of the device: group g is queued before the status block of group g - 1 is looked at, so
    // the stream never drains between groups and a slow host costs nothing (round 1 synchronised after every group and read
    // three words back one by one; on a slow box that was tens of milliseconds per step).  Everything a group can leave behind
    // is therefore handled on the device or survives one more group: the sub-bin spill list is applied by a kernel that reads
    // its length on the device, and the overflow list has room for every instance of TWO groups.
    const uint64_t ovf_cap = std::max<uint64_t>(2 * (uint64_t)g_max * max_bin, 1024);
    auto attach_overflow_list = [&]() {
        t->dev.ovf_keys = (uint64_t*)mc_scratch(ctx, 3, ovf_cap * sizeof(uint64_t));
        t->dev.ovf_cap = ovf_cap;
        MC_CUDA(cudaMemsetAsync(&t->dev.counters[MC_CNT_OVF], 0, 2 * sizeof(unsigned long long), st));
    };
    attach_overflow_list();
    enum { RING = 2 };
    unsigned long long* h_ring = (unsigned long long*)((uint8_t*)ctx->h_pinned_small + 1024);   // RING status blocks
    struct Pending { bool live = false; cudaEvent_t ev = nullptr; uint64_t keys_done = 0, group_keys = 0; uint32_t b0 = 0, G = 0; } pend[RING];
    cudaEvent_t sc_done[2] = {nullptr, nullptr}, agg_done[2] = {nullptr, nullptr};
    struct EvGuard {
        mc_ctx* c; Pending* p; cudaEvent_t* a; cudaEvent_t* b;
        ~EvGuard() { for (int i = 0; i < RING; i++) mc_event_put(c, p[i].ev); for (int i = 0; i < 2; i++) { mc_event_put(c, a[i]); mc_event_put(c, b[i]); } }
    } ev_guard{ctx, pend, sc_done, agg_done};
    for (int i = 0; i < RING; i++) pend[i].ev = mc_event_get(ctx);
    for (int i = 0; i < 2; i++) { sc_done[i] = mc_event_get(ctx); agg_done[i] = mc_event_get(ctx); }
    const uint64_t keys_before = mc_table_keys(t);
    uint64_t keys_known = keys_before;
    const bool trace = getenv("MC_TRACE") != nullptr;
    // look at the status block a finished group left behind: overflow -> fatal, table growth, re-insertion of overflowed keys
    auto examine = [&](Pending& p, const unsigned long long* h_c) {
        MC_CUDA(cudaEventSynchronize(p.ev));
        p.live = false;
        if ((unsigned int)h_c[MC_STATUS_FLAGS] & MC_FLAG_OVERFLOW) mc_throw(MC_ERR_NOMEM, "k-mer table overflow (stash and overflow list full)");
        const uint64_t spill2 = p.group_keys ? h_c[MC_STATUS_SPILL2] : 0;
        if (trace) fprintf(stderr, "[mc trace] agg group b0=%u G=%u keys=%llu P2=%u cap2=%llu spill2=%llu ovf=%llu table_keys=%llu\n", p.b0, p.G,
                           (unsigned long long)p.group_keys, P2, (unsigned long long)cap2, (unsigned long long)spill2,
                           (unsigned long long)h_c[MC_CNT_OVF], (unsigned long long)h_c[MC_CNT_KEYS]);
        if (spill2 > spill2_cap) mc_throw(MC_ERR_NOMEM, "internal: sub-bin spill list overflow");
        const uint64_t ovf_n = h_c[MC_CNT_OVF];
        const uint64_t now = h_c[MC_CNT_KEYS] + h_c[MC_CNT_OVF_KEYS];   // every push is one distinct key
        const double gained = now > keys_before ? (double)(now - keys_before) : 0.0;
        const double predicted = (double)keys_before + (p.keys_done ? gained * ((double)batch_keys / (double)p.keys_done) : 0.0);
        bool grown = false;
        if (predicted > MC_GROW_LOAD * 4.0 * (double)t->nb) {
            int q = fit_q(ctx, t, mc_q_for_keys((uint64_t)predicted, t->dev.shard_bits, MC_FLUSH_TARGET_LOAD));
            if (q > t->dev.q) {
                mc_table_grow_to(t, q);
                grown = true;
            }
        }
        if (ovf_n) {   // rare: the table was too small; re-insert what overflowed so far (grows until it fits)
            if (t->hook) t->hook->failed = true;   // keys of ranges already handed on come back into the table
            mc_table_finish_chunk(t, grown);
            attach_overflow_list();
            keys_known = mc_table_keys(t);
        } else {
            keys_known = std::max(keys_known, now);
        }
    };
    // the groups of first-level bins: the first one is a sample (1/64 of the hash space) that sizes the table for the rest of the batch
    struct Group { uint32_t b0, G; uint64_t keys; };
    std::vector<Group> groups;
    for (uint32_t b0 = 0; b0 < P1;) {
        uint32_t G = groups.empty() ? std::max<uint32_t>(P1 / 64, 1) : g_max;
        G = std::min(std::min(G, g_max), P1 - b0);
        uint64_t gk = 0;
        for (uint32_t b = b0; b < b0 + G; b++) gk += std::min<uint64_t>(cur[b], cap1);
        groups.push_back(Group{b0, G, gk});
        b0 += G;
    }
    const uint32_t tiles = (uint32_t)((max_bin + SC_TILE - 1) / SC_TILE);
    cudaStream_t s2 = overlap ? ctx->aux_stream : st;
    auto launch_scatter2 = [&](uint32_t gi) {   // sub-bin scatter of group gi into its staging slot, on s2
        const Group& g = groups[gi];
        McBinsDev& S = B2s[gi % n_slots];
        if (overlap && gi >= 2) MC_CUDA(cudaStreamWaitEvent(s2, agg_done[gi & 1], 0));   // agg(gi - 2) has read this slot
        if (g.keys) {
            MC_CUDA(cudaMemsetAsync(S.cursors, 0, n_cursors * 8, s2));
            McStageTimer tm(ctx, MC_ST_SCATTER, s2);
            const uint32_t x = (uint32_t)std::min<uint64_t>((uint64_t)tiles * g.G, (uint64_t)ctx->sm_count * (overlap ? 1 : sc_per_sm));
            if (ctx->tune_scatter_tma) k_bin_scatter2<true><<<x, SC_THREADS, scatter_smem_bytes(P2), s2>>>(B->dev, g.b0, g.G, tiles, S, t->dev.flags);
            else k_bin_scatter2<false><<<x, SC_THREADS, scatter_smem_bytes(P2), s2>>>(B->dev, g.b0, g.G, tiles, S, t->dev.flags);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
        }
        if (overlap) MC_CUDA(cudaEventRecord(sc_done[gi & 1], s2));
    };
    if (overlap) {   // the first-level bins are complete on the main stream before the auxiliary stream reads them
        MC_CUDA(cudaEventRecord(agg_done[0], st));
        MC_CUDA(cudaStreamWaitEvent(s2, agg_done[0], 0));
        launch_scatter2(0);
    }
    uint64_t keys_done = 0;
    for (uint32_t gi = 0; gi < (uint32_t)groups.size(); gi++) {
        const Group& g = groups[gi];
        const bool first = gi == 0;
        McBinsDev& B2 = B2s[gi % n_slots];
        unsigned long long* d_ticket = B2.cursors + (n_cursors - 1);
        grow_if_due(t, keys_known);   // amortised doubling
        Pending& slot = pend[gi % RING];
        unsigned long long* h_c = h_ring + (size_t)(gi % RING) * MC_STATUS_WORDS;
        if (slot.live) examine(slot, h_c);   // (only when RING groups are in flight; normally drained below)
        if (!overlap) launch_scatter2(gi);
        else MC_CUDA(cudaStreamWaitEvent(st, sc_done[gi & 1], 0));
        if (g.keys) {
            const uint32_t n_units = g.G * P2;
            const int grid = (int)std::min<uint64_t>(n_units, (uint64_t)ctx->sm_count * (overlap ? 1 : agg_per_sm));
            if (first) {
                // the sample is only COUNTED first: its distinct keys that the table does not hold yet, scaled to the whole batch,
                // size the table before anything is inserted (no overflow churn, no rehash of a half-filled table)
                {
                    McStageTimer tm(ctx, MC_ST_COUNT);
                    if (ctx->tune_agg_lane_loop) k_bin_agg<true, true><<<grid, AGG_THREADS, agg_smem_bytes(), st>>>(t->dev, B2, g.b0, n_units, d_ticket);
                    else k_bin_agg<true, false><<<grid, AGG_THREADS, agg_smem_bytes(), st>>>(t->dev, B2, g.b0, n_units, d_ticket);
                    MC_LAUNCH(ctx);
                    MC_CUDA(cudaGetLastError());
                }
                if (overlap && groups.size() > 1) launch_scatter2(1);   // the next group's scatter runs under the sample and its read-back
                MC_CUDA(cudaMemcpyAsync(h_c, t->dev.counters, MC_STATUS_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaStreamSynchronize(st));
                // later batches of a stream are assumed to bring as many new keys as this one: an upper bound (the genomic k-mers
                // saturate, the error k-mers keep coming), so the table is grown once instead of doubling batch after batch
                // The first batch of a stream over-predicts most (everything is new): it is trusted for 4 batches only; the
                // second batch already shows the marginal rate and the table is then grown to its final size.
                double batches_left = 1.0;
                if (t->expect_instances > t->seen_instances + batch_inst) batches_left = (double)(t->expect_instances - t->seen_instances) / (double)batch_inst;
                if (t->stream_flushes == 0) batches_left = std::min(batches_left, 4.0);
                const double fresh = (double)h_c[MC_CNT_OVF_KEYS] * ((double)batch_keys / (double)g.keys) * batches_left;
                const double predicted = (double)keys_before + fresh;
                if (trace) fprintf(stderr, "[mc trace] agg sample: %llu new distinct keys in %llu instances -> %.0f predicted for the batch\n",
                                   (unsigned long long)h_c[MC_CNT_OVF_KEYS], (unsigned long long)g.keys, predicted);
                if (predicted > MC_GROW_LOAD * 4.0 * (double)t->nb) {
                    const int q = fit_q(ctx, t, mc_q_for_keys((uint64_t)predicted, t->dev.shard_bits, MC_FLUSH_TARGET_LOAD));
                    if (q > t->dev.q) mc_table_grow_to(t, q);
                }
                attach_overflow_list();
                MC_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned long long), st));
            }
            {
                McStageTimer tm(ctx, MC_ST_COUNT);
                if (ctx->tune_agg_lane_loop) k_bin_agg<false, true><<<grid, AGG_THREADS, agg_smem_bytes(), st>>>(t->dev, B2, g.b0, n_units, d_ticket);
                else k_bin_agg<false, false><<<grid, AGG_THREADS, agg_smem_bytes(), st>>>(t->dev, B2, g.b0, n_units, d_ticket);
                MC_LAUNCH(ctx);
                // a sub-bin that was full (a heavy hitter) spilled: direct updates, length read on the device
                k_insert_hashed_n<<<ctx->sm_count, 256, 0, st>>>(t->dev, B2.spill, B2.spill_cursor, spill2_cap);
                MC_LAUNCH(ctx);
                MC_CUDA(cudaGetLastError());
            }
            if (trace && first) {   // sub-bin fill of the first group
                std::vector<unsigned long long> c2((size_t)g.G * (P2 + 1));
                MC_CUDA(cudaMemcpyAsync(c2.data(), B2.cursors, c2.size() * 8, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaStreamSynchronize(st));
                unsigned long long mx = 0, sum = 0, over = 0, mn = ~0ull;
                for (uint32_t gg = 0; gg < g.G; gg++)
                    for (uint32_t j = 0; j < P2; j++) {
                        unsigned long long v = c2[(size_t)gg * (P2 + 1) + j];
                        mx = std::max(mx, v); mn = std::min(mn, v); sum += v; over += v > cap2;
                    }
                fprintf(stderr, "[mc trace] sub-bin attempts: min %llu mean %.0f max %llu, %llu of %u over cap2; bin0 attempts %llu\n", mn, (double)sum / (g.G * P2), mx, over, g.G * P2, cur[0]);
            }
        }
        if (t->hook && !t->hook->failed) {   // everything whose home bucket lies in this group's hash range is final now
            t->hook->group_done(t->hook->user, t, g.b0, g.G, P1);
            t->hook->groups++;
        }
        keys_done += g.keys;
        slot.live = true;
        slot.keys_done = keys_done;
        slot.group_keys = g.keys;
        slot.b0 = g.b0;
        slot.G = g.G;
        MC_CUDA(cudaMemcpyAsync(h_c, t->dev.counters, MC_STATUS_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (g.keys) MC_CUDA(cudaMemcpyAsync(h_c + MC_STATUS_SPILL2, B2.spill_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaEventRecord(slot.ev, st));
        if (overlap) {
            MC_CUDA(cudaEventRecord(agg_done[gi & 1], st));                      // this slot may be refilled (by the scatter of group gi + 2)
            if (gi >= 1 && gi + 1 < (uint32_t)groups.size()) launch_scatter2(gi + 1);   // (group 1 was queued under the sample)
        }
        // the group queued BEFORE this one has had a whole group's time to finish: look at what it left behind
        if (gi >= 1) {
            Pending& prev = pend[(gi - 1) % RING];
            if (prev.live) examine(prev, h_ring + (size_t)((gi - 1) % RING) * MC_STATUS_WORDS);
        }
    }
    for (int j = 0; j < RING; j++)
        if (pend[j].live) examine(pend[j], h_ring + (size_t)j * MC_STATUS_WORDS);
    if (overlap) MC_CUDA(cudaStreamSynchronize(s2));
    t->seen_instances += batch_inst;
    t->stream_flushes++;
}

void mc_table_flush(mc_table* t) {
    McBatch* B = t->batch;
    if (B && B->sk_nb && !B->sk_flushing && B->sk_kmers) mc_sk_flush(t);   // pending super-k-mer records -> key bins / table (mc_sk.cu)
    if (!B || !B->dev.keys || B->used == 0) return;
    mc_ctx* ctx = t->ctx;
    MC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t P = B->P;
    const uint32_t BIN_TILE = ctx->tune_count_tile ? ctx->tune_count_tile : BIN_TILE_DEFAULT;
    std::vector<unsigned long long> cur(P + 1);
    MC_CUDA(cudaMemcpyAsync(cur.data(), B->dev.cursors, (size_t)(P + 1) * 8, cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    if (cur[P] > B->dev.spill_cap) mc_throw(MC_ERR_NOMEM, "internal: k-mer spill list overflow (%llu > %llu)", cur[P], (unsigned long long)B->dev.spill_cap);
    uint64_t batch_keys = 0;
    for (uint32_t b = 0; b < P; b++) batch_keys += std::min<uint64_t>(cur[b], B->dev.cap);
    // large batches: second-level bins + shared-memory aggregation (one table update per DISTINCT key of a sub-bin)
    const bool aggregate = ctx->tune_agg < 0 ? batch_keys >= MC_AGG_MIN_KEYS : ctx->tune_agg != 0;
    if (aggregate) {
        if (t->hook && cur[P] > 0) t->hook->failed = true;   // the first-level spill list is applied at the very end, to any range
        flush_aggregated(t, cur, batch_keys);
        apply_spill_list(t, cur[P]);
        return;
    }
    if (t->hook) t->hook->failed = true;   // the direct sweep has no per-range completion points
    std::vector<uint64_t> prefix(P + 1);
    prefix[0] = 0;
    for (uint32_t b = 0; b < P; b++) {
        const uint64_t c = std::min<uint64_t>(cur[b], B->dev.cap);
        prefix[b + 1] = prefix[b] + (c + BIN_TILE - 1) / BIN_TILE;
    }
    const uint64_t total_tiles = prefix[P];
    uint64_t* d_prefix = (uint64_t*)(B->dev.cursors + (P + 1));
    MC_CUDA(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, st));

    // One k_bin_count launch per segment of at most seg_keys_max keys.  The overflow list attached to the table has room for a
    // whole segment, so even a table that turns out far too small loses no update; after every segment the key counter
    // tells how many keys the batch will add in total (hashes are uniform: the slices not swept yet gain what the swept
    // ones did) and the table is grown to that size at once, while it is still mostly empty.
    const uint64_t seg_keys_max = ctx->tune_segment_keys ? std::min(std::max<uint64_t>(ctx->tune_segment_keys, BIN_TILE), MC_FLUSH_SEGMENT) : MC_FLUSH_SEGMENT;
    const uint64_t seg_tiles_max = std::max<uint64_t>(seg_keys_max / BIN_TILE, 1);
    const uint64_t ovf_cap = std::max<uint64_t>(std::min<uint64_t>(seg_keys_max, total_tiles * BIN_TILE), 1024);
    auto attach_overflow_list = [&]() {
        t->dev.ovf_keys = (uint64_t*)mc_scratch(ctx, 3, ovf_cap * sizeof(uint64_t));
        t->dev.ovf_cap = ovf_cap;
        MC_CUDA(cudaMemsetAsync(&t->dev.counters[MC_CNT_OVF], 0, 2 * sizeof(unsigned long long), st));
    };
    attach_overflow_list();
    McBinCountKernel kern = bin_count_variant(ctx->tune_count_variant);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    cudaGetLastError();
    unsigned long long* d_ticket = (unsigned long long*)(d_prefix + (P + 1));
    unsigned long long* h_c = (unsigned long long*)((uint8_t*)ctx->h_pinned_small + 1024);   // counters + flags read-back
    unsigned int* h_flags = (unsigned int*)(h_c + MC_CNT_N);
    const uint64_t keys_before = mc_table_keys(t);
    uint64_t keys_known = keys_before;
    uint64_t tile = 0;
    bool first = true;
    while (tile < total_tiles) {
        // the first segment is a sample (1/64 of the hash space) that sizes the table for the rest of the batch
        uint64_t seg = first ? std::max<uint64_t>(total_tiles / 64, 1) : seg_tiles_max;
        seg = std::min(std::min(seg, seg_tiles_max), total_tiles - tile);
        grow_if_due(t, keys_known);   // amortised doubling
        {
            McStageTimer tm(ctx, MC_ST_COUNT);
            // exactly the co-resident CTAs; tiles are handed out in order by the ticket
            const int grid = (int)std::min<uint64_t>(seg, (uint64_t)ctx->sm_count * per_sm);
            MC_CUDA(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned long long), st));
            kern<<<grid, 256, 0, st>>>(t->dev, B->dev, d_prefix, tile, tile + seg, BIN_TILE, d_ticket);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
        }
        tile += seg;
        first = false;
        MC_CUDA(cudaMemcpyAsync(h_c, t->dev.counters, MC_CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaMemcpyAsync(h_flags, t->dev.flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        if (*h_flags & MC_FLAG_OVERFLOW) mc_throw(MC_ERR_NOMEM, "k-mer table overflow (stash and overflow list full)");
        const uint64_t ovf_n = h_c[MC_CNT_OVF];
        const uint64_t now = h_c[MC_CNT_KEYS] + ovf_n;
        const double gained = now > keys_before ? (double)(now - keys_before) : 0.0;
        const double predicted = (double)keys_before + gained * ((double)total_tiles / (double)tile);
        bool grown = false;
        if (predicted > MC_GROW_LOAD * 4.0 * (double)t->nb) {
            int q = fit_q(ctx, t, mc_q_for_keys((uint64_t)predicted, t->dev.shard_bits, MC_FLUSH_TARGET_LOAD));
            if (q > t->dev.q) {
                mc_table_grow_to(t, q);
                grown = true;
            }
        }
        if (ovf_n) {   // rare: the table was too small for this segment; re-insert what overflowed (grows until it fits)
            mc_table_finish_chunk(t, grown);
            attach_overflow_list();
            keys_known = mc_table_keys(t);
        } else {
            keys_known = now;
        }
    }
    t->dev.ovf_keys = nullptr;
    t->dev.ovf_cap = 0;
    apply_spill_list(t, cur[P]);
}

void mc_batch_release(mc_table* t) {
    if (!t->batch) return;
    mc_table_flush(t);
    mc_sk_release(t);
    batch_free_buffers(t->ctx, t->batch);
    delete t->batch;
    t->batch = nullptr;
}

// --------------------------------------------------------------------------------------------
// tuning knobs (tests force tiny batches / many bins; profiles compare against the direct path)
// --------------------------------------------------------------------------------------------
extern "C" int mc_ctx_set_tuning(mc_ctx* ctx, const char* name, long long value) {
    if (!ctx || !name) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    std::string n(name);
    if (n == "part_bits" && value >= -1 && value <= MC_MAX_PBITS) ctx->tune_part_bits = (long)value;
    else if (n == "batch_keys" && value >= 0) ctx->tune_batch_keys = (uint64_t)value;
    else if (n == "segment_keys" && value >= 0) ctx->tune_segment_keys = (uint64_t)value;
    else if (n == "host_batch_chunks" && value >= 0 && value <= 1024) ctx->tune_host_batch_chunks = (uint32_t)value;
    else if (n == "extract_parts" && (value == 0 || value == 1 || value == 4 || value == 32)) ctx->tune_extract_parts = (int)value;
    else if (n == "count_variant" && value >= 0 && value <= 3) ctx->tune_count_variant = (int)value;
    else if (n == "agg" && value >= -1 && value <= 1) ctx->tune_agg = (int)value;
    else if (n == "keep_stage_ms" && value >= 0 && value <= 1) ctx->keep_stage_ms = (int)value;
    else if (n == "scatter_tma" && value >= 0 && value <= 1) ctx->tune_scatter_tma = (int)value;
    else if (n == "agg_lane_loop" && value >= 0 && value <= 1) ctx->tune_agg_lane_loop = (int)value;
    else if (n == "superkmers" && value >= 0 && value <= 1) ctx->tune_superkmers = (int)value;
    else if (n == "sk_bits" && value >= -1 && value <= MC_SK_MAX_BITS) ctx->tune_sk_bits = (int)value;
    else if (n == "sk_cap" && value >= 0 && value <= (1ll << 30)) ctx->tune_sk_cap = (uint64_t)value;
    else if (n == "flush_overlap" && value >= 0 && value <= 1) ctx->tune_flush_overlap = (int)value;
    else if (n == "extract_tma" && value >= 0 && value <= 1) ctx->tune_extract_tma = (int)value;
    else if (n == "defer_flush" && value >= 0 && value <= 1) ctx->tune_defer_flush = (int)value;
    else if (n == "stage_timing" && value >= 0 && value <= 1) ctx->tune_stage_timing = (int)value;
    else if (n == "device_chunk_bytes" && (value == 0 || (value >= 4096 && value <= ((long long)3 << 30)))) ctx->tune_device_chunk_bytes = (size_t)value;
    else if (n == "agg_unit" && value >= 0 && value <= (1 << 24)) ctx->tune_agg_unit = (uint32_t)value;
    else if (n == "count_tile" && value >= 0 && value <= (1 << 20) && value % 256 == 0) ctx->tune_count_tile = (uint32_t)value;
    else {
        mc_set_error("unknown tuning parameter or value out of range: " + n);
        return MC_ERR_ARG;
    }
    return MC_OK;
}
extern "C" int mc_table_expect_keys(mc_table* t, uint64_t n) {
    if (!t) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    t->expect_instances = n;
    t->seen_instances = 0;
    t->stream_flushes = 0;
    return MC_OK;
}
extern "C" int mc_table_flush_pending(mc_table* t) {
    try {
        if (!t) mc_throw(MC_ERR_ARG, "null argument");
        mc_table_flush(t);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
