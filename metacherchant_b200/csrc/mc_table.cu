// The k-mer count table in HBM: creation, growth (rehash), lookups, dump, coverage compaction.
// Replaces ITMO!/structures/map/{BigLong2ShortHashMap,Long2ShortHashMap}.java (semantics, not layout).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mc_common.cuh"

// --------------------------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void mc_set_error(const std::string& msg) { g_last_error = msg; }
void mc_throw(int code, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw McError(code, buf);
}
extern "C" const char* mc_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* mc_version(void) { return "metacherchant_b200 0.1 (sm_100a)"; }

// --------------------------------------------------------------------------------------------
// inverse of mc_mix62
// --------------------------------------------------------------------------------------------
static __host__ __device__ __forceinline__ uint64_t inv_odd(uint64_t c) {   // inverse mod 2^64 (hence mod 2^62)
    uint64_t x = c;
    for (int i = 0; i < 6; i++) x *= 2 - c * x;
    return x;
}
static __host__ __device__ __forceinline__ uint64_t unmix62_impl(uint64_t h) {
    h ^= h >> 32;
    h = (h * inv_odd(MC_MIX_C2)) & MC_M62;
    h ^= (h >> 29) ^ (h >> 58);
    h = (h * inv_odd(MC_MIX_C1)) & MC_M62;
    h ^= h >> 31;
    return h;
}
uint64_t mc_unmix62(uint64_t h) { return unmix62_impl(h); }
__device__ uint64_t mc_unmix62_dev(uint64_t h) { return unmix62_impl(h); }

// --------------------------------------------------------------------------------------------
// allocation / sizing
// --------------------------------------------------------------------------------------------
static const double MC_MAX_LOAD = 0.70;
static const uint32_t MC_STASH_SLOTS = 1u << 16;

int mc_q_for_keys(uint64_t keys, int shard_bits, double load) {
    int q_min = 21 - shard_bits;   // tag must fit 41 bits
    if (q_min < 10) q_min = 10;
    int q = q_min;
    while ((double)keys > load * 4.0 * (double)(1ULL << q)) q++;
    return q;
}
static int q_for_keys(uint64_t keys, int shard_bits) { return mc_q_for_keys(keys, shard_bits, MC_MAX_LOAD); }

// Arena: see mc_ctx.  Requests are rounded (512 B; 2 MiB from 1 MiB up) so that the same logical size always maps to
// the same block size; a cached block is reused when it is at most 1/8 larger than the request.
static size_t arena_round(size_t bytes) {
    if (bytes < 512) bytes = 512;
    size_t g = bytes >= ((size_t)1 << 20) ? ((size_t)2 << 20) : 512;
    // large blocks: 1/32 of the size's power of two, so that requests that drift by a few per cent between calls (bins and staging
    // sized from the data seen so far) keep mapping to the same block size and are served from the cache
    if (bytes >= ((size_t)64 << 20)) {
        int lg = 63;
        while (!(bytes >> lg)) lg--;
        g = (size_t)1 << (lg - 5);
    }
    return (bytes + g - 1) / g * g;
}
static void arena_release_all_locked(mc_ctx* ctx) {
    if (ctx->arena_free.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->arena_free) cudaFree(b.p);
    ctx->arena_driver_allocs++;   // (invalidates the cached free-memory figure)
    ctx->arena_free.clear();
    ctx->arena_cached_bytes = 0;
    cudaGetLastError();
}
void* mc_dev_alloc(mc_ctx* ctx, size_t bytes, const char* what) {
    const size_t need = arena_round(bytes);
    std::lock_guard<std::mutex> lk(ctx->arena_mu);
    size_t best = (size_t)-1, best_bytes = (size_t)-1;
    for (size_t i = 0; i < ctx->arena_free.size(); i++) {
        const size_t b = ctx->arena_free[i].bytes;
        if (b >= need && b < best_bytes) { best = i; best_bytes = b; }
    }
    const size_t slack = need < ((size_t)1 << 20) ? 3 * need : need / 8;
    void* p = nullptr;
    if (best != (size_t)-1 && best_bytes <= need + slack) {
        p = ctx->arena_free[best].p;
        ctx->arena_free[best] = ctx->arena_free.back();
        ctx->arena_free.pop_back();
        ctx->arena_cached_bytes -= best_bytes;
        ctx->arena_live[p] = best_bytes;
        return p;
    }
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {   // give the cache back to the driver and try once more
        cudaGetLastError();
        arena_release_all_locked(ctx);
        e = cudaMalloc(&p, need);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        mc_throw(MC_ERR_NOMEM, "%s of %.2f GiB does not fit in device memory", what, (double)bytes / (1 << 30));
    }
    ctx->arena_driver_allocs++;
    ctx->arena_live[p] = need;
    if (getenv("MC_TRACE")) fprintf(stderr, "[mc trace] arena: cudaMalloc of %.1f MiB for %s (driver call %llu)\n", (double)need / (1 << 20), what, (unsigned long long)ctx->arena_driver_allocs);
    return p;
}
void mc_dev_free(mc_ctx* ctx, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(ctx->arena_mu);
    auto it = ctx->arena_live.find(p);
    if (it == ctx->arena_live.end()) {   // not ours (should not happen): hand it to the driver
        cudaStreamSynchronize(ctx->stream);
        cudaFree(p);
        return;
    }
    ctx->arena_free.push_back({p, it->second});
    ctx->arena_cached_bytes += it->second;
    ctx->arena_live.erase(it);
}
size_t mc_arena_cached_bytes(mc_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->arena_mu);
    return ctx->arena_cached_bytes;
}
void mc_arena_trim(mc_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->arena_mu);
    arena_release_all_locked(ctx);
}
cudaEvent_t mc_event_get(mc_ctx* ctx) {
    {
        std::lock_guard<std::mutex> lk(ctx->arena_mu);
        if (!ctx->event_pool.empty()) {
            cudaEvent_t e = ctx->event_pool.back();
            ctx->event_pool.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void mc_event_put(mc_ctx* ctx, cudaEvent_t e) {
    if (!e) return;
    std::lock_guard<std::mutex> lk(ctx->arena_mu);
    ctx->event_pool.push_back(e);
}

static void alloc_slots(mc_ctx* ctx, McTableDev& d, int q, int shard_bits, cudaStream_t s) {
    d.q = q;
    d.shard_bits = shard_bits;
    d.tag_bits = 62 - shard_bits - q;
    if (d.tag_bits > MC_MAX_TAG_BITS || d.tag_bits < 0) mc_throw(MC_ERR_ARG, "table geometry out of range (q=%d)", q);
    d.tag_mask = d.tag_bits == 0 ? 0 : ((1ULL << d.tag_bits) - 1);
    d.nb_mask = (1ULL << q) - 1;
    size_t bytes = ((size_t)1 << q) * 32;
    d.slots = (uint64_t*)mc_dev_alloc(ctx, bytes, "k-mer table");
    MC_CUDA(cudaMemsetAsync(d.slots, 0, bytes, s));
}

extern "C" int mc_table_create(mc_ctx* ctx, int k, uint64_t capacity_hint, int shard_bits, int shard_id, mc_table** out) {
    try {
        if (!ctx || !out) mc_throw(MC_ERR_ARG, "null argument");
        if (k < 1 || k > 31) mc_throw(k > 31 ? MC_ERR_UNSUPPORTED : MC_ERR_ARG, "k must be in 1..31 (k > 31 is the reference's lossy hashed mode, out of scope)");
        if (shard_bits < 0 || shard_bits > 8 || shard_id < 0 || shard_id >= (1 << shard_bits)) mc_throw(MC_ERR_ARG, "shard_bits must be in 0..8 and shard_id < 2^shard_bits");
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table* t = new mc_table();
        t->ctx = ctx;
        t->k = k;
        t->shard_id = (uint32_t)shard_id;
        alloc_slots(ctx, t->dev, q_for_keys(capacity_hint, shard_bits), shard_bits, ctx->stream);
        t->nb = 1ULL << t->dev.q;
        t->dev.stash_mask = MC_STASH_SLOTS - 1;
        t->dev.ovf_keys = nullptr;
        t->dev.ovf_cap = 0;
        t->dev.stash_keys = (uint64_t*)mc_dev_alloc(ctx, MC_STASH_SLOTS * sizeof(uint64_t), "stash");
        t->dev.stash_counts = (uint32_t*)mc_dev_alloc(ctx, MC_STASH_SLOTS * sizeof(uint32_t), "stash");
        // one status block: counters, the flag word and a spare cursor are read back with ONE copy (MC_STATUS_WORDS)
        t->dev.counters = (unsigned long long*)mc_dev_alloc(ctx, MC_STATUS_WORDS * sizeof(unsigned long long), "counters");
        t->dev.flags = (unsigned int*)(t->dev.counters + MC_STATUS_FLAGS);
        MC_CUDA(cudaMemsetAsync(t->dev.stash_keys, 0, MC_STASH_SLOTS * sizeof(uint64_t), ctx->stream));
        MC_CUDA(cudaMemsetAsync(t->dev.stash_counts, 0, MC_STASH_SLOTS * sizeof(uint32_t), ctx->stream));
        MC_CUDA(cudaMemsetAsync(t->dev.counters, 0, MC_STATUS_WORDS * sizeof(unsigned long long), ctx->stream));
        *out = t;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" void mc_table_free(mc_table* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    if (t->batch) {   // pending updates die with the table
        mc_dev_free(t->ctx, t->batch->dev.keys);
        mc_dev_free(t->ctx, t->batch->dev.cursors);
        mc_dev_free(t->ctx, t->batch->dev.spill);
        delete t->batch;
        t->batch = nullptr;
    }
    // stream-ordered frees: the memory goes back to the context's pool once the work queued so far is done
    mc_dev_free(t->ctx, t->dev.slots);
    mc_dev_free(t->ctx, t->dev.stash_keys);
    mc_dev_free(t->ctx, t->dev.stash_counts);
    mc_dev_free(t->ctx, t->dev.counters);   // the flag word lives in the same block
    delete t;
}

static void read_counters(const mc_table* t, unsigned long long c[MC_CNT_N]) {
    MC_CUDA(cudaMemcpyAsync(c, t->dev.counters, MC_CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->ctx->stream));
    MC_CUDA(cudaStreamSynchronize(t->ctx->stream));
}
uint64_t mc_table_keys(const mc_table* t) {
    unsigned long long c[MC_CNT_N];
    read_counters(t, c);
    return c[MC_CNT_KEYS];
}
static void check_overflow(const mc_table* t) {
    unsigned int f = 0;
    MC_CUDA(cudaMemcpyAsync(&f, t->dev.flags, sizeof f, cudaMemcpyDeviceToHost, t->ctx->stream));
    MC_CUDA(cudaStreamSynchronize(t->ctx->stream));
    if (f & MC_FLAG_OVERFLOW) mc_throw(MC_ERR_NOMEM, "k-mer table overflow (stash full); raise capacity_hint");
}

extern "C" uint64_t mc_table_size(const mc_table* t) {
    try {
        mc_table_flush(const_cast<mc_table*>(t));
        return mc_table_keys(t);
    } catch (const McError& e) {
        mc_set_error(e.what());
        return 0;
    }
}
extern "C" int mc_table_stats(const mc_table* t, uint64_t* reads, uint64_t* runs, uint64_t* bases, uint64_t* kmers) {
    try {
        unsigned long long c[MC_CNT_N];
        mc_table_flush(const_cast<mc_table*>(t));
        read_counters(t, c);
        if (reads) *reads = c[MC_CNT_READS];
        if (runs) *runs = c[MC_CNT_RUNS];
        if (bases) *bases = c[MC_CNT_BASES];
        if (kmers) *kmers = c[MC_CNT_KMERS];
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// --------------------------------------------------------------------------------------------
// rehash: every live slot of `src` re-inserted into `dst` (bigger, or filtered by min_count)
// --------------------------------------------------------------------------------------------
// Re-insertion is a chain of dependent L2 round trips (bucket probe, CAS).  Done lane by lane and slot by slot -- round 1 --
// almost every (bucket, slot) position of a warp has SOME lane with a survivor, so a warp ran ~2.5 serial insert phases per
// 1 KB scanned and the scan crawled at the insert latency (2 TB/s).  Now a warp first scans RH_U x 32 buckets, queues the
// survivors in shared memory, and then inserts them one per lane: one insert phase per 4 KB scanned in the coverage filter.
#ifndef RH_U
#define RH_U 4
#endif
#ifndef RH_MINB
#define RH_MINB 4     // resident CTAs per SM the register budget is cut for: 64 registers, 1024 threads per SM (filter 5.95 -> 4.8 ms at cfg2)
#endif
#define RH_Q 256      // queue entries per warp (what does not fit is inserted on the spot)
__global__ void __launch_bounds__(256, RH_MINB) k_rehash(McTableDev src, McTableDev dst, uint32_t min_count) {
    __shared__ uint64_t q_h[8][RH_Q];
    __shared__ uint32_t q_c[8][RH_Q];
    __shared__ uint32_t q_n[8];
    const uint64_t nb = src.nb_mask + 1;
    const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long added = 0;
    if (lane == 0) q_n[warp] = 0;
    __syncwarp();
    // every lane of a warp runs the same number of steps (the warp-level queue needs the whole warp)
    const uint64_t first = blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u);
    for (uint64_t w0 = first; w0 < nb; w0 += T * RH_U) {
        const uint64_t b0 = w0 + lane;
        uint64_t w4[RH_U][4];
#pragma unroll
        for (int u = 0; u < RH_U; u++) {
            const uint64_t b = b0 + (uint64_t)u * T;
            if (b < nb) mc_load_bucket(src.slots + b * 4, w4[u]);   // one 32-byte bucket per thread and load
            else w4[u][0] = w4[u][1] = w4[u][2] = w4[u][3] = 0;
        }
#pragma unroll
        for (int u = 0; u < RH_U; u++) {
            const uint64_t b = b0 + (uint64_t)u * T;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint64_t w = w4[u][s];
                if (w == 0) continue;
                uint32_t c = (uint32_t)(w & MC_COUNT_MASK);
                if (c > MC_SAT) c = MC_SAT;
                if (c < min_count) continue;
                uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
                uint64_t home = (b - d) & src.nb_mask;
                uint64_t hl = (home << src.tag_bits) | (w >> MC_TAG_SHIFT);
                const uint32_t at = atomicAdd(&q_n[warp], 1u);
                if (at < RH_Q) {
                    q_h[warp][at] = hl;
                    q_c[warp][at] = c;
                } else if (mc_table_add(dst, hl, c)) {
                    added++;
                }
            }
        }
        __syncwarp();
        const uint32_t n = min(q_n[warp], (uint32_t)RH_Q);
        for (uint32_t i = lane; i < n; i += 32)
            if (mc_table_add(dst, q_h[warp][i], q_c[warp][i])) added++;
        __syncwarp();
        if (lane == 0) q_n[warp] = 0;
        __syncwarp();
    }
    // stash entries
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= src.stash_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t kx = src.stash_keys[i];
        if (kx == 0) continue;
        uint32_t c = src.stash_counts[i];
        if (c > MC_SAT) c = MC_SAT;
        if (c < min_count) continue;
        if (mc_table_add(dst, kx - 1, c)) added++;
    }
    // block-aggregate the new-key counter
    __shared__ unsigned long long s_added;
    if (threadIdx.x == 0) s_added = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&s_added, added);
    __syncthreads();
    if (threadIdx.x == 0 && s_added) atomicAdd(&dst.counters[MC_CNT_KEYS], s_added);
}

static int grid_for(const mc_ctx* ctx, uint64_t items, int block, int per_sm = 8) {
    uint64_t need = (items + block - 1) / block;
    uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)std::min<uint64_t>(need, cap);
}

static void grow_to(mc_table* t, int new_q) {
    mc_ctx* ctx = t->ctx;
    McTableDev old = t->dev;
    McTableDev nw = old;
    alloc_slots(ctx, nw, new_q, old.shard_bits, ctx->stream);
    // fresh stash + key counter for the new geometry; the statistics counters carry over
    nw.stash_keys = (uint64_t*)mc_dev_alloc(ctx, MC_STASH_SLOTS * sizeof(uint64_t), "stash");
    nw.stash_counts = (uint32_t*)mc_dev_alloc(ctx, MC_STASH_SLOTS * sizeof(uint32_t), "stash");
    nw.ovf_keys = nullptr;   // a rehash into a table twice the size never needs the overflow list
    nw.ovf_cap = 0;
    MC_CUDA(cudaMemsetAsync(nw.stash_keys, 0, MC_STASH_SLOTS * sizeof(uint64_t), ctx->stream));
    MC_CUDA(cudaMemsetAsync(nw.stash_counts, 0, MC_STASH_SLOTS * sizeof(uint32_t), ctx->stream));
    MC_CUDA(cudaMemsetAsync(&nw.counters[MC_CNT_KEYS], 0, 2 * sizeof(unsigned long long), ctx->stream));
    k_rehash<<<grid_for(ctx, (old.nb_mask + 1 + RH_U - 1) / RH_U, 256, 16), 256, 0, ctx->stream>>>(old, nw, 0);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    mc_dev_free(ctx, old.slots);
    mc_dev_free(ctx, old.stash_keys);
    mc_dev_free(ctx, old.stash_counts);
    nw.ovf_keys = old.ovf_keys;
    nw.ovf_cap = old.ovf_cap;
    t->dev = nw;
    t->nb = 1ULL << new_q;
}

void mc_table_grow_to(mc_table* t, int q) {
    if (t->hook && t->hook->before_grow) t->hook->before_grow(t->hook->user, t);
    if (getenv("MC_TRACE")) {   // (synchronising: trace runs only)
        cudaEvent_t a = mc_event_get(t->ctx), b = mc_event_get(t->ctx);
        const int q0 = t->dev.q;
        cudaEventRecord(a, t->ctx->stream);
        grow_to(t, q);
        cudaEventRecord(b, t->ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        t->ctx->stage_ms[MC_ST_GROW] += ms;
        fprintf(stderr, "[mc trace] table grown from 2^%d to 2^%d buckets (%.1f GiB) in %.2f ms\n", q0, q, (double)(32ull << q) / (1 << 30), ms);
        mc_event_put(t->ctx, a);
        mc_event_put(t->ctx, b);
        return;
    }
    McStageTimer tm(t->ctx, MC_ST_GROW);   // resolved lazily: growing does not stall the host
    grow_to(t, q);
}

void mc_table_reserve(mc_table* t, uint64_t additional) {
    // callers insert without an overflow list: room at load <= 0.55, where the 8-bucket windows practically never fill
    uint64_t keys = mc_table_keys(t);
    int q = mc_q_for_keys(keys + additional, t->dev.shard_bits, 0.55);
    if (q > t->dev.q) mc_table_grow_to(t, q);
}

// Chunk protocol of the counting path.  Growth is amortised (double when the table is more than 55 % full) instead
// of reserving room for the worst case "every k-mer of the chunk is new"; an update that finds neither a free slot in
// its 8-bucket window nor room in the stash lands in the overflow list and is re-inserted after the table has grown.
extern const double MC_GROW_LOAD = 0.55;
extern const uint64_t MC_OVF_MAX = 64ull << 20;

void mc_table_prepare_chunk(mc_table* t, uint64_t ub) {
    mc_ctx* ctx = t->ctx;
    uint64_t keys = mc_table_keys(t);
    uint64_t ovf_cap = std::min<uint64_t>(std::max<uint64_t>(ub, 1024), MC_OVF_MAX);
    int q = t->dev.q;
    // (a) amortised doubling, (b) never start a chunk that could not physically fit in table + overflow list
    while ((double)keys > MC_GROW_LOAD * 4.0 * (double)(1ULL << q) || (double)(keys + ub) > 0.9 * 4.0 * (double)(1ULL << q) + (double)ovf_cap) q++;
    if (q > t->dev.q) mc_table_grow_to(t, q);
    t->dev.ovf_keys = (uint64_t*)mc_scratch(ctx, 3, ovf_cap * sizeof(uint64_t));
    t->dev.ovf_cap = ovf_cap;
    MC_CUDA(cudaMemsetAsync(&t->dev.counters[MC_CNT_OVF], 0, 2 * sizeof(unsigned long long), ctx->stream));
}

__global__ void __launch_bounds__(256) k_insert_hashed(McTableDev t, const uint64_t* hkeys, size_t n);

void mc_table_finish_chunk(mc_table* t, bool grown_already) {
    mc_ctx* ctx = t->ctx;
    while (true) {
        unsigned long long c[MC_CNT_N];
        read_counters(t, c);
        check_overflow(t);
        unsigned long long n = c[MC_CNT_OVF];
        if (n == 0) break;
        if (n > t->dev.ovf_cap) mc_throw(MC_ERR_NOMEM, "k-mer table overflow: pass a capacity hint for this input");
        // move the list aside, double the table, insert the list (which may overflow again, into the now empty list)
        uint64_t* tmp = (uint64_t*)mc_dev_alloc(ctx, n * sizeof(uint64_t), "overflow list");
        MC_CUDA(cudaMemcpyAsync(tmp, t->dev.ovf_keys, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        MC_CUDA(cudaMemsetAsync(&t->dev.counters[MC_CNT_OVF], 0, 2 * sizeof(unsigned long long), ctx->stream));
        if (!grown_already) mc_table_grow_to(t, t->dev.q + 1);
        grown_already = false;
        uint64_t blocks = std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
        k_insert_hashed<<<(unsigned)blocks, 256, 0, ctx->stream>>>(t->dev, tmp, n);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        mc_dev_free(ctx, tmp);
    }
    t->dev.ovf_keys = nullptr;
    t->dev.ovf_cap = 0;
}

// --------------------------------------------------------------------------------------------
// Seam 2: get / dump / filter / add_pairs / insert_hashed
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_get(McTableDev t, const uint64_t* keys, size_t n, int16_t* out, uint32_t shard_id) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t h = mc_mix62(keys[i] & MC_M62);
        int v = -1;
        if (keys[i] <= MC_M62 && (t.shard_bits == 0 || (h >> (62 - t.shard_bits)) == shard_id)) v = mc_table_lookup(t, h);
        out[i] = (int16_t)v;
    }
}

extern "C" int mc_table_get(const mc_table* t, const uint64_t* keys, size_t n, int16_t* out) {
    try {
        if (!t || (n && (!keys || !out))) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(t));
        uint64_t* d_keys;
        int16_t* d_out;
        d_keys = (decltype(d_keys))mc_dev_alloc(ctx, n * sizeof(uint64_t), "scratch");
        d_out = (decltype(d_out))mc_dev_alloc(ctx, n * sizeof(int16_t), "scratch");
        MC_CUDA(cudaMemcpyAsync(d_keys, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        k_get<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t->dev, d_keys, n, d_out, t->shard_id);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaMemcpyAsync(out, d_out, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        mc_dev_free(ctx, d_keys);
        mc_dev_free(ctx, d_out);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// pass 1: count matching entries; pass 2: write (key, count) with warp-aggregated appends.  One 32-byte bucket per thread.
template <bool WRITE>
__global__ void __launch_bounds__(256) k_dump(McTableDev t, uint32_t min_count, unsigned long long* cursor, uint64_t* keys, int16_t* counts,
                                             uint64_t shard_prefix) {
    const uint64_t nb = t.nb_mask + 1;
    const uint64_t stash_units = ((uint64_t)t.stash_mask + 1 + 3) / 4;
    const uint64_t total = nb + stash_units;   // work units of 4 entries
    unsigned long long local = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // every lane of a warp runs the same number of iterations so the ballots below are convergent
    const uint64_t iters = (total + stride - 1) / stride;
    uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (uint64_t it = 0; it < iters; it++, u += stride) {
        uint64_t hl[4];
        uint32_t c[4];
        bool hit[4] = {false, false, false, false};
        if (u < nb) {
            uint64_t w4[4];
            mc_load_bucket(t.slots + u * 4, w4);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint64_t w = w4[s];
                c[s] = (uint32_t)(w & MC_COUNT_MASK);
                uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
                hl[s] = (((u - d) & t.nb_mask) << t.tag_bits) | (w >> MC_TAG_SHIFT);
                hit[s] = w != 0;
            }
        } else if (u < total) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint64_t i = (u - nb) * 4 + s;
                uint64_t kx = i <= t.stash_mask ? t.stash_keys[i] : 0;
                hit[s] = kx != 0;
                hl[s] = kx - 1;
                c[s] = hit[s] ? t.stash_counts[i] : 0;
            }
        }
        uint32_t n_hit = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (c[s] > MC_SAT) c[s] = MC_SAT;
            hit[s] = hit[s] && c[s] >= min_count;
            n_hit += hit[s];
        }
        if (!WRITE) {
            local += n_hit;
        } else {
            // warp-aggregated reservation: exclusive scan of n_hit over the warp
            uint32_t incl = n_hit;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long base = 0;
            if (lane == 0 && warp_total) base = atomicAdd(cursor, (unsigned long long)warp_total);
            base = __shfl_sync(0xffffffffu, base, 0);
            unsigned long long pos = base + incl - n_hit;
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (hit[s]) {
                    keys[pos] = mc_unmix62_dev(shard_prefix | hl[s]);
                    counts[pos] = (int16_t)c[s];
                    pos++;
                }
        }
    }
    if (!WRITE) {
        for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if (lane == 0 && local) atomicAdd(cursor, local);
    }
}

static uint64_t count_matching(const mc_table* t, int min_count, unsigned long long* d_cursor) {
    mc_ctx* ctx = t->ctx;
    MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
    uint32_t mc = min_count < 0 ? 0 : (uint32_t)min_count;
    k_dump<false><<<grid_for(ctx, t->nb, 256, 16), 256, 0, ctx->stream>>>(t->dev, mc, d_cursor, nullptr, nullptr, 0);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    MC_CUDA(cudaMemcpyAsync(&n, d_cursor, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    return n;
}

// Order-independent digest of the entries (mc_table_checksum): one 32-byte bucket per thread, like k_dump
__global__ void __launch_bounds__(256) k_checksum(McTableDev t, uint32_t min_count, unsigned long long* acc, uint64_t shard_prefix) {
    const uint64_t nb = t.nb_mask + 1;
    const uint64_t stash_units = ((uint64_t)t.stash_mask + 1 + 3) / 4;
    const uint64_t total = nb + stash_units;
    unsigned long long n = 0, sum = 0, dig = 0, mx = 0;
    for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t hl[4];
        uint32_t c[4];
        if (u < nb) {
            uint64_t w4[4];
            mc_load_bucket(t.slots + u * 4, w4);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint64_t w = w4[s];
                c[s] = (uint32_t)(w & MC_COUNT_MASK);
                const uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
                hl[s] = (((u - d) & t.nb_mask) << t.tag_bits) | (w >> MC_TAG_SHIFT);
                if (w == 0) c[s] = 0;
            }
        } else {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint64_t i = (u - nb) * 4 + s;
                const uint64_t kx = i <= t.stash_mask ? t.stash_keys[i] : 0;
                hl[s] = kx - 1;
                c[s] = kx ? t.stash_counts[i] : 0;
            }
        }
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (c[s] > MC_SAT) c[s] = MC_SAT;
            if (c[s] == 0 || c[s] < min_count) continue;
            const uint64_t key = mc_unmix62_dev(shard_prefix | hl[s]);
            n++;
            sum += c[s];
            dig += ((key ^ (key >> 29)) * 0xBF58476D1CE4E5B9ULL + 1ULL) * (unsigned long long)c[s];
            mx = mx > c[s] ? mx : c[s];
        }
    }
    for (int o = 16; o; o >>= 1) {
        n += __shfl_down_sync(0xffffffffu, n, o);
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        dig += __shfl_down_sync(0xffffffffu, dig, o);
        const unsigned long long m2 = __shfl_down_sync(0xffffffffu, mx, o);
        mx = mx > m2 ? mx : m2;
    }
    if ((threadIdx.x & 31) == 0) {
        if (n) atomicAdd(&acc[0], n);
        if (sum) atomicAdd(&acc[1], sum);
        if (dig) atomicAdd(&acc[2], dig);
        if (mx) atomicMax(&acc[3], mx);
    }
}

extern "C" int mc_table_checksum(const mc_table* t, int min_count, uint64_t out[4]) {
    try {
        if (!t || !out) mc_throw(MC_ERR_ARG, "null argument");
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(t));
        check_overflow(t);
        unsigned long long* d_acc = (unsigned long long*)mc_dev_alloc(ctx, 4 * sizeof(unsigned long long), "scratch");
        MC_CUDA(cudaMemsetAsync(d_acc, 0, 4 * sizeof(unsigned long long), ctx->stream));
        const uint64_t prefix = t->dev.shard_bits ? ((uint64_t)t->shard_id << (62 - t->dev.shard_bits)) : 0;
        k_checksum<<<grid_for(ctx, t->nb, 256, 16), 256, 0, ctx->stream>>>(t->dev, min_count < 0 ? 0u : (uint32_t)min_count, d_acc, prefix);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned long long h[4];
        MC_CUDA(cudaMemcpyAsync(h, d_acc, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        mc_dev_free(ctx, d_acc);
        for (int i = 0; i < 4; i++) out[i] = h[i];
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_table_dump(const mc_table* t, int min_count, uint64_t* keys, int16_t* counts, size_t cap, size_t* n_out) {
    try {
        if (!t || !n_out) mc_throw(MC_ERR_ARG, "null argument");
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(t));
        check_overflow(t);
        unsigned long long* d_cursor;
        d_cursor = (decltype(d_cursor))mc_dev_alloc(ctx, sizeof(unsigned long long), "scratch");
        uint64_t n = count_matching(t, min_count, d_cursor);
        *n_out = n;
        if (cap == 0 || n == 0) {
            mc_dev_free(ctx, d_cursor);
            return MC_OK;
        }
        if (n > 0xFFFFFFFFull) {
            mc_dev_free(ctx, d_cursor);
            mc_throw(MC_ERR_UNSUPPORTED, "dump of more than 2^32 entries");
        }
        if (cap < n) {
            mc_dev_free(ctx, d_cursor);
            mc_throw(MC_ERR_CAPACITY, "dump needs room for %llu entries", (unsigned long long)n);
        }
        uint64_t* d_keys;
        int16_t* d_counts;
        d_keys = (decltype(d_keys))mc_dev_alloc(ctx, n * sizeof(uint64_t), "scratch");
        d_counts = (decltype(d_counts))mc_dev_alloc(ctx, n * sizeof(int16_t), "scratch");
        MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
        uint64_t prefix = t->dev.shard_bits ? ((uint64_t)t->shard_id << (62 - t->dev.shard_bits)) : 0;
        uint32_t mc = min_count < 0 ? 0 : (uint32_t)min_count;
        k_dump<true><<<grid_for(ctx, t->nb, 256, 16), 256, 0, ctx->stream>>>(t->dev, mc, d_cursor, d_keys, d_counts, prefix);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        std::vector<uint64_t> hk(n);
        std::vector<int16_t> hc(n);
        MC_CUDA(cudaMemcpyAsync(hk.data(), d_keys, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaMemcpyAsync(hc.data(), d_counts, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        mc_dev_free(ctx, d_keys);
        mc_dev_free(ctx, d_counts);
        mc_dev_free(ctx, d_cursor);
        // sorted by key for parity checks (the reference iterates in map order; consumers treat it as a set)
        std::vector<uint32_t> idx(n);
        for (uint64_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return hk[a] < hk[b]; });
        for (uint64_t i = 0; i < n; i++) {
            keys[i] = hk[idx[i]];
            counts[i] = hc[idx[i]];
        }
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// Device-buffer forms of dump / add_pairs: the multi-GPU gather of the solid (key, count16) set moves the
// pairs GPU to GPU and must not bounce them through the host.  Entries leave in table order (unsorted).
extern "C" int mc_table_dump_device(const mc_table* t, int min_count, uint64_t* d_keys, int16_t* d_counts, size_t cap, size_t* n_out) {
    try {
        if (!t || !n_out) mc_throw(MC_ERR_ARG, "null argument");
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(t));
        check_overflow(t);
        unsigned long long* d_cursor = (unsigned long long*)mc_dev_alloc(ctx, sizeof(unsigned long long), "scratch");
        struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } fr{ctx, d_cursor};
        const uint64_t n = count_matching(t, min_count, d_cursor);
        *n_out = n;
        if (cap == 0 || n == 0) return MC_OK;
        if (cap < n) mc_throw(MC_ERR_CAPACITY, "dump needs room for %llu entries", (unsigned long long)n);
        if (!d_keys || !d_counts) mc_throw(MC_ERR_ARG, "null argument");
        MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
        const uint64_t prefix = t->dev.shard_bits ? ((uint64_t)t->shard_id << (62 - t->dev.shard_bits)) : 0;
        k_dump<true><<<grid_for(ctx, t->nb, 256, 16), 256, 0, ctx->stream>>>(t->dev, min_count < 0 ? 0u : (uint32_t)min_count, d_cursor, d_keys, d_counts, prefix);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

__global__ void __launch_bounds__(256) k_add_pairs(McTableDev t, const uint64_t* keys, const int16_t* counts, size_t n, uint32_t shard_id, unsigned int* bad);
extern "C" int mc_table_add_pairs_device(mc_table* t, const uint64_t* d_keys, const int16_t* d_counts, size_t n) {
    try {
        if (!t || (n && (!d_keys || !d_counts))) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(t);
        mc_table_reserve(t, n);
        unsigned int* d_bad = (unsigned int*)mc_dev_alloc(ctx, sizeof(unsigned int), "scratch");
        struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } fr{ctx, d_bad};
        MC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
        k_add_pairs<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t->dev, d_keys, d_counts, n, t->shard_id, d_bad);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned int bad = 0;
        MC_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        check_overflow(t);
        if (bad) mc_throw(MC_ERR_ARG, "a key above 2^62 is not a packed k-mer of k <= 31");
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_table_filter(mc_ctx* ctx, const mc_table* t, int min_count, mc_table** solid) {
    try {
        if (!ctx || !t || !solid) mc_throw(MC_ERR_ARG, "null argument");
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(t));
        check_overflow(t);
        McHostTimer ht(ctx, MC_ST_FILTER);
        McStageTimer tm(ctx, MC_ST_FILTER);
        unsigned long long* d_cursor;
        d_cursor = (decltype(d_cursor))mc_dev_alloc(ctx, sizeof(unsigned long long), "scratch");
        struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } fr{ctx, d_cursor};
        // Size of the solid table: a large table is only SAMPLED -- the survivors of its first 1/64 (the hash spreads the keys
        // evenly over the buckets) scaled up, plus a margin -- so the table is scanned once, by the rehash, instead of twice.
        // The rehash counts exactly; if the estimate was too low the compaction is simply redone with the exact figure.
        const bool sampled = t->nb >= ((uint64_t)1 << 22);
        uint64_t n;
        if (sampled) {
            McTableDev sub = t->dev;
            sub.nb_mask = (t->nb >> 6) - 1;
            MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
            k_dump<false><<<grid_for(ctx, sub.nb_mask + 1, 256, 16), 256, 0, ctx->stream>>>(sub, min_count < 0 ? 0u : (uint32_t)min_count, d_cursor, nullptr, nullptr, 0);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
            unsigned long long v = 0;
            MC_CUDA(cudaMemcpyAsync(&v, d_cursor, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
            MC_CUDA(cudaStreamSynchronize(ctx->stream));
            n = v * 64 + v * 64 / 16 + 65536;   // + 6 % and a constant: far outside the sampling spread of >= 64 Ki buckets
        } else {
            n = count_matching(t, min_count, d_cursor);
        }
        mc_table* s = nullptr;
        uint32_t mc = min_count < 1 ? 1 : (uint32_t)min_count;
        for (int attempt = 0;; attempt++) {
            // dense but comfortably probe-able: load <= 0.35 keeps almost every lookup inside one sector
            int rc = mc_table_create(ctx, t->k, n * 2, t->dev.shard_bits, (int)t->shard_id, &s);
            if (rc != MC_OK) return rc;
            k_rehash<<<grid_for(ctx, (t->nb + RH_U - 1) / RH_U, 256, 16), 256, 0, ctx->stream>>>(t->dev, s->dev, mc);
            MC_LAUNCH(ctx);
            MC_CUDA(cudaGetLastError());
            if (!sampled || attempt > 0) break;
            const uint64_t exact = mc_table_keys(s);
            if ((double)exact <= 0.40 * (double)s->capacity_slots()) break;
            mc_table_free(s);   // the sample under-estimated (cannot happen for hashed keys; kept for safety): redo with the exact count
            s = nullptr;
            n = exact;
        }
        // statistics travel with the filtered table
        MC_CUDA(cudaMemcpyAsync(&s->dev.counters[MC_CNT_READS], &t->dev.counters[MC_CNT_READS], 4 * sizeof(unsigned long long),
                                cudaMemcpyDeviceToDevice, ctx->stream));
        // no overflow list is attached to `s`: an entry that found neither its window nor the stash free would be LOST and a solid
        // k-mer would silently vanish from every environment.  Improbable at load 0.35, but checked.
        try {
            check_overflow(s);
        } catch (...) {
            mc_table_free(s);
            throw;
        }
        *solid = s;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// pairs whose key this shard does not own are skipped (the other shards' add_pairs take them); a key above 2^62 is not a
// k-mer of k <= 31 and is reported through `bad`
__global__ void __launch_bounds__(256) k_add_pairs(McTableDev t, const uint64_t* keys, const int16_t* counts, size_t n, uint32_t shard_id, unsigned int* bad) {
    unsigned long long added = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int c = counts[i];
        if (c <= 0) continue;   // Kmers2HMWorker only adds freq > threshold >= 0
        if (keys[i] > MC_M62) { atomicOr(bad, 1u); continue; }
        const uint64_t h = mc_mix62(keys[i]);
        if (t.shard_bits != 0 && (uint32_t)(h >> (62 - t.shard_bits)) != shard_id) continue;
        if (mc_table_add(t, h, (uint32_t)c)) added++;
    }
    if (added) atomicAdd(&t.counters[MC_CNT_KEYS], added);
}

extern "C" int mc_table_add_pairs(mc_table* t, const uint64_t* keys, const int16_t* counts, size_t n) {
    try {
        if (!t || (n && (!keys || !counts))) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(t);
        mc_table_reserve(t, n);
        uint64_t* d_keys;
        int16_t* d_counts;
        d_keys = (decltype(d_keys))mc_dev_alloc(ctx, n * sizeof(uint64_t), "scratch");
        d_counts = (decltype(d_counts))mc_dev_alloc(ctx, n * sizeof(int16_t), "scratch");
        unsigned int* d_bad = (unsigned int*)mc_dev_alloc(ctx, sizeof(unsigned int), "scratch");
        MC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
        MC_CUDA(cudaMemcpyAsync(d_keys, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        MC_CUDA(cudaMemcpyAsync(d_counts, counts, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        k_add_pairs<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t->dev, d_keys, d_counts, n, t->shard_id, d_bad);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        unsigned int bad = 0;
        MC_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        mc_dev_free(ctx, d_keys);
        mc_dev_free(ctx, d_counts);
        mc_dev_free(ctx, d_bad);
        check_overflow(t);
        if (bad) mc_throw(MC_ERR_ARG, "a key above 2^62 is not a packed k-mer of k <= 31");
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

__global__ void __launch_bounds__(256) k_insert_hashed(McTableDev t, const uint64_t* hkeys, size_t n) {
    unsigned long long added = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t h = hkeys[i];
        if (h == MC_HASH_SENTINEL) continue;
        if (mc_table_add(t, h, 1)) added++;
    }
    __shared__ unsigned long long s_added;
    if (threadIdx.x == 0) s_added = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&s_added, added);
    __syncthreads();
    if (threadIdx.x == 0 && s_added) atomicAdd(&t.counters[MC_CNT_KEYS], s_added);
}

// the same with the list length read on the device (sub-bin spill lists of the aggregated flush: no host round trip)
__global__ void __launch_bounds__(256) k_insert_hashed_n(McTableDev t, const uint64_t* hkeys, const unsigned long long* n_ptr, unsigned long long cap) {
    const unsigned long long n = *n_ptr < cap ? *n_ptr : cap;
    if (n == 0) return;
    unsigned long long added = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t h = hkeys[i];
        if (h == MC_HASH_SENTINEL) continue;
        if (mc_table_add(t, h, 1)) added++;
    }
    for (int o = 16; o; o >>= 1) added += __shfl_down_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&t.counters[MC_CNT_KEYS], added);
}

extern "C" int mc_table_insert_hashed_device(mc_table* t, const uint64_t* d_hkeys, size_t n) {
    try {
        if (!t || (n && !d_hkeys)) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        // buffered: the keys are binned now and applied by the next flush (any read of the table flushes)
        mc_batch_reserve(t, n, 0);
        mc_batch_scatter(t, d_hkeys, n);
        MC_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller may reuse d_hkeys
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
