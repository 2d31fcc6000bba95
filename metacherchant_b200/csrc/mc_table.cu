// The k-mer count table in HBM: creation, growth (rehash), lookups, dump, coverage compaction.
// Replaces ITMO!/structures/map/{BigLong2ShortHashMap,Long2ShortHashMap}.java (semantics, not layout).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
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

static int q_for_keys(uint64_t keys, int shard_bits) {
    int q_min = 21 - shard_bits;   // tag must fit 41 bits
    if (q_min < 10) q_min = 10;
    int q = q_min;
    while ((double)keys > MC_MAX_LOAD * 4.0 * (double)(1ULL << q)) q++;
    return q;
}

static void alloc_slots(McTableDev& d, int q, int shard_bits, cudaStream_t s) {
    d.q = q;
    d.shard_bits = shard_bits;
    d.tag_bits = 62 - shard_bits - q;
    if (d.tag_bits > MC_MAX_TAG_BITS || d.tag_bits < 0) mc_throw(MC_ERR_ARG, "table geometry out of range (q=%d)", q);
    d.tag_mask = d.tag_bits == 0 ? 0 : ((1ULL << d.tag_bits) - 1);
    d.nb_mask = (1ULL << q) - 1;
    size_t bytes = ((size_t)1 << q) * 32;
    cudaError_t e = cudaMalloc(&d.slots, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        mc_throw(MC_ERR_NOMEM, "k-mer table of %.1f GiB does not fit in device memory", (double)bytes / (1 << 30));
    }
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
        alloc_slots(t->dev, q_for_keys(capacity_hint, shard_bits), shard_bits, ctx->stream);
        t->nb = 1ULL << t->dev.q;
        t->dev.stash_mask = MC_STASH_SLOTS - 1;
        MC_CUDA(cudaMalloc(&t->dev.stash_keys, MC_STASH_SLOTS * sizeof(uint64_t)));
        MC_CUDA(cudaMalloc(&t->dev.stash_counts, MC_STASH_SLOTS * sizeof(uint32_t)));
        MC_CUDA(cudaMalloc(&t->dev.counters, MC_CNT_N * sizeof(unsigned long long)));
        MC_CUDA(cudaMalloc(&t->dev.flags, sizeof(unsigned int)));
        MC_CUDA(cudaMemsetAsync(t->dev.stash_keys, 0, MC_STASH_SLOTS * sizeof(uint64_t), ctx->stream));
        MC_CUDA(cudaMemsetAsync(t->dev.stash_counts, 0, MC_STASH_SLOTS * sizeof(uint32_t), ctx->stream));
        MC_CUDA(cudaMemsetAsync(t->dev.counters, 0, MC_CNT_N * sizeof(unsigned long long), ctx->stream));
        MC_CUDA(cudaMemsetAsync(t->dev.flags, 0, sizeof(unsigned int), ctx->stream));
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
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->dev.slots);
    cudaFree(t->dev.stash_keys);
    cudaFree(t->dev.stash_counts);
    cudaFree(t->dev.counters);
    cudaFree(t->dev.flags);
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
        return mc_table_keys(t);
    } catch (const McError& e) {
        mc_set_error(e.what());
        return 0;
    }
}
extern "C" int mc_table_stats(const mc_table* t, uint64_t* reads, uint64_t* runs, uint64_t* bases, uint64_t* kmers) {
    try {
        unsigned long long c[MC_CNT_N];
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
__global__ void __launch_bounds__(256) k_rehash(McTableDev src, McTableDev dst, uint32_t min_count) {
    const uint64_t n_slots = (src.nb_mask + 1) * 4;
    unsigned long long added = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t w = src.slots[i];
        if (w == 0) continue;
        uint32_t c = (uint32_t)(w & MC_COUNT_MASK);
        if (c > MC_SAT) c = MC_SAT;
        if (c < min_count) continue;
        uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
        uint64_t home = ((i >> 2) - d) & src.nb_mask;
        uint64_t hl = (home << src.tag_bits) | (w >> MC_TAG_SHIFT);
        if (mc_table_add(dst, hl, c)) added++;
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
    alloc_slots(nw, new_q, old.shard_bits, ctx->stream);
    // fresh stash + key counter for the new geometry; the statistics counters carry over
    MC_CUDA(cudaMalloc(&nw.stash_keys, MC_STASH_SLOTS * sizeof(uint64_t)));
    MC_CUDA(cudaMalloc(&nw.stash_counts, MC_STASH_SLOTS * sizeof(uint32_t)));
    MC_CUDA(cudaMemsetAsync(nw.stash_keys, 0, MC_STASH_SLOTS * sizeof(uint64_t), ctx->stream));
    MC_CUDA(cudaMemsetAsync(nw.stash_counts, 0, MC_STASH_SLOTS * sizeof(uint32_t), ctx->stream));
    MC_CUDA(cudaMemsetAsync(&nw.counters[MC_CNT_KEYS], 0, 2 * sizeof(unsigned long long), ctx->stream));
    k_rehash<<<grid_for(ctx, (old.nb_mask + 1) * 4, 256, 16), 256, 0, ctx->stream>>>(old, nw, 0);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(old.slots);
    cudaFree(old.stash_keys);
    cudaFree(old.stash_counts);
    t->dev = nw;
    t->nb = 1ULL << new_q;
}

void mc_table_reserve(mc_table* t, uint64_t additional) {
    uint64_t keys = mc_table_keys(t);
    int q = q_for_keys(keys + additional, t->dev.shard_bits);
    if (q > t->dev.q) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, t->ctx->stream);
        grow_to(t, q);
        cudaEventRecord(b, t->ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        t->ctx->stage_ms[2] += ms;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
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
        uint64_t* d_keys;
        int16_t* d_out;
        MC_CUDA(cudaMalloc(&d_keys, n * sizeof(uint64_t)));
        MC_CUDA(cudaMalloc(&d_out, n * sizeof(int16_t)));
        MC_CUDA(cudaMemcpyAsync(d_keys, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        k_get<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t->dev, d_keys, n, d_out, t->shard_id);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaMemcpyAsync(out, d_out, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_keys);
        cudaFree(d_out);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// pass 1: count matching entries; pass 2: write (key, count) with warp-aggregated appends
template <bool WRITE>
__global__ void __launch_bounds__(256) k_dump(McTableDev t, uint32_t min_count, unsigned long long* cursor, uint64_t* keys, int16_t* counts,
                                             uint64_t shard_prefix) {
    const uint64_t n_slots = (t.nb_mask + 1) * 4;
    const uint64_t total = n_slots + t.stash_mask + 1;
    unsigned long long local = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // every lane of a warp runs the same number of iterations so the ballots below are convergent
    const uint64_t iters = (total + stride - 1) / stride;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < iters; it++, i += stride) {
        bool hit = false;
        uint64_t hl = 0;
        uint32_t c = 0;
        if (i < n_slots) {
            uint64_t w = t.slots[i];
            if (w != 0) {
                c = (uint32_t)(w & MC_COUNT_MASK);
                uint64_t d = (w >> MC_COUNT_BITS) & (MC_MAX_DISP - 1);
                uint64_t home = ((i >> 2) - d) & t.nb_mask;
                hl = (home << t.tag_bits) | (w >> MC_TAG_SHIFT);
                hit = true;
            }
        } else if (i < total) {
            uint64_t kx = t.stash_keys[i - n_slots];
            if (kx != 0) {
                hl = kx - 1;
                c = t.stash_counts[i - n_slots];
                hit = true;
            }
        }
        if (c > MC_SAT) c = MC_SAT;
        hit = hit && c >= min_count;
        if (!WRITE) {
            local += hit;
        } else {
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                int lane = threadIdx.x & 31;
                int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (hit) {
                    unsigned long long pos = base + __popc(m & ((1u << lane) - 1));
                    keys[pos] = mc_unmix62_dev(shard_prefix | hl);
                    counts[pos] = (int16_t)c;
                }
            }
        }
    }
    if (!WRITE) {
        for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(cursor, local);
    }
}

static uint64_t count_matching(const mc_table* t, int min_count, unsigned long long* d_cursor) {
    mc_ctx* ctx = t->ctx;
    MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
    uint32_t mc = min_count < 0 ? 0 : (uint32_t)min_count;
    k_dump<false><<<grid_for(ctx, t->capacity_slots(), 256, 16), 256, 0, ctx->stream>>>(t->dev, mc, d_cursor, nullptr, nullptr, 0);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    MC_CUDA(cudaMemcpyAsync(&n, d_cursor, sizeof n, cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    return n;
}

extern "C" int mc_table_dump(const mc_table* t, int min_count, uint64_t* keys, int16_t* counts, size_t cap, size_t* n_out) {
    try {
        if (!t || !n_out) mc_throw(MC_ERR_ARG, "null argument");
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        check_overflow(t);
        unsigned long long* d_cursor;
        MC_CUDA(cudaMalloc(&d_cursor, sizeof(unsigned long long)));
        uint64_t n = count_matching(t, min_count, d_cursor);
        *n_out = n;
        if (cap == 0 || n == 0) {
            cudaFree(d_cursor);
            return MC_OK;
        }
        if (cap < n) {
            cudaFree(d_cursor);
            mc_throw(MC_ERR_CAPACITY, "dump needs room for %llu entries", (unsigned long long)n);
        }
        uint64_t* d_keys;
        int16_t* d_counts;
        MC_CUDA(cudaMalloc(&d_keys, n * sizeof(uint64_t)));
        MC_CUDA(cudaMalloc(&d_counts, n * sizeof(int16_t)));
        MC_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), ctx->stream));
        uint64_t prefix = t->dev.shard_bits ? ((uint64_t)t->shard_id << (62 - t->dev.shard_bits)) : 0;
        uint32_t mc = min_count < 0 ? 0 : (uint32_t)min_count;
        k_dump<true><<<grid_for(ctx, t->capacity_slots(), 256, 16), 256, 0, ctx->stream>>>(t->dev, mc, d_cursor, d_keys, d_counts, prefix);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        std::vector<uint64_t> hk(n);
        std::vector<int16_t> hc(n);
        MC_CUDA(cudaMemcpyAsync(hk.data(), d_keys, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaMemcpyAsync(hc.data(), d_counts, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_keys);
        cudaFree(d_counts);
        cudaFree(d_cursor);
        // sorted by key for parity checks (the reference iterates in map order; consumers treat it as a set)
        std::vector<uint32_t> idx(n);
        for (uint64_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
        if (n > 0xFFFFFFFFull) mc_throw(MC_ERR_UNSUPPORTED, "dump of more than 2^32 entries");
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

extern "C" int mc_table_filter(mc_ctx* ctx, const mc_table* t, int min_count, mc_table** solid) {
    try {
        if (!ctx || !t || !solid) mc_throw(MC_ERR_ARG, "null argument");
        MC_CUDA(cudaSetDevice(ctx->device));
        check_overflow(t);
        unsigned long long* d_cursor;
        MC_CUDA(cudaMalloc(&d_cursor, sizeof(unsigned long long)));
        uint64_t n = count_matching(t, min_count, d_cursor);
        cudaFree(d_cursor);
        mc_table* s = nullptr;
        // dense but comfortably probe-able: load <= 0.35 keeps almost every lookup inside one sector
        int rc = mc_table_create(ctx, t->k, n * 2, t->dev.shard_bits, (int)t->shard_id, &s);
        if (rc != MC_OK) return rc;
        uint32_t mc = min_count < 1 ? 1 : (uint32_t)min_count;
        k_rehash<<<grid_for(ctx, t->capacity_slots(), 256, 16), 256, 0, ctx->stream>>>(t->dev, s->dev, mc);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        // statistics travel with the filtered table
        MC_CUDA(cudaMemcpyAsync(&s->dev.counters[MC_CNT_READS], &t->dev.counters[MC_CNT_READS], 4 * sizeof(unsigned long long),
                                cudaMemcpyDeviceToDevice, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        *solid = s;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

__global__ void __launch_bounds__(256) k_add_pairs(McTableDev t, const uint64_t* keys, const int16_t* counts, size_t n) {
    unsigned long long added = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int c = counts[i];
        if (c <= 0) continue;   // Kmers2HMWorker only adds freq > threshold >= 0
        if (mc_table_add(t, mc_mix62(keys[i] & MC_M62), (uint32_t)c)) added++;
    }
    if (added) atomicAdd(&t.counters[MC_CNT_KEYS], added);
}

extern "C" int mc_table_add_pairs(mc_table* t, const uint64_t* keys, const int16_t* counts, size_t n) {
    try {
        if (!t || (n && (!keys || !counts))) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_reserve(t, n);
        uint64_t* d_keys;
        int16_t* d_counts;
        MC_CUDA(cudaMalloc(&d_keys, n * sizeof(uint64_t)));
        MC_CUDA(cudaMalloc(&d_counts, n * sizeof(int16_t)));
        MC_CUDA(cudaMemcpyAsync(d_keys, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        MC_CUDA(cudaMemcpyAsync(d_counts, counts, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        k_add_pairs<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(t->dev, d_keys, d_counts, n);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_keys);
        cudaFree(d_counts);
        check_overflow(t);
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

extern "C" int mc_table_insert_hashed_device(mc_table* t, const uint64_t* d_hkeys, size_t n) {
    try {
        if (!t || (n && !d_hkeys)) mc_throw(MC_ERR_ARG, "null argument");
        if (n == 0) return MC_OK;
        mc_ctx* ctx = t->ctx;
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_reserve(t, n);
        k_insert_hashed<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(t->dev, d_hkeys, n);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
