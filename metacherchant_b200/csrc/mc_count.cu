// K2 + K3: 2-bit canonical k-mer extraction fused with the count-table update, and the host-side
// chunk pipeline behind Seam 1 (IOUtils.loadReads, src/io/IOUtils.java:217-248).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mc_parse.cuh"

// --------------------------------------------------------------------------------------------
// K2/K3 kernel.  Four threads (a "quad") share one record: the quad first validates the whole record
// (each lane a quarter of its bases: nucleotide / quality checks, N detection, reader statistics), then
// lane p rolls the forward and reverse-complement words over k-mer start positions [p*S, (p+1)*S),
// S = ceil((len-k+1)/4), re-reading k-1 bases of context, and feeds every valid k-mer to the table.
//
// Reference semantics restated (all integer):
//   FASTQ (FastqReader.java:72-80, FastaReaderFromXQSourceTrunc.java:61-94, DnaQ.java:147-150,176-186):
//     base is a split point  <=>  char in {N,n,.}  or  ((qual - offset) & 63) == 0;
//     the quality char of a non-N base must lie in [offset, 126] (Illumina.java:7-12, Sanger.java:7-12);
//     k-mers never span a split point.
//   FASTA (FastaReader.java:54-64): a record containing N or n is dropped whole; any other non-ACGT char throws.
//   k-mers (ShortKmer.java:68-71,122-149): fw=((fw<<2)|b)&mask, rc=(rc>>2)|((3-b)<<(2k-2)), key=min(fw,rc).
// --------------------------------------------------------------------------------------------
extern "C" uint64_t mc_max_kmers(size_t len, int fmt, int k);

struct McExtractArgs {
    const uint8_t* base;
    const McRec* recs;
    uint32_t n_records;
    uint32_t n_bytes;         // readable bytes at `base` (the chunk, or the compacted FASTA sequence buffer)
    int k;
    int qoff;                 // 33 / 64 (FASTQ) ; ignored for FASTA
    McParseHeader* header;    // errors
    McTableDev table;
    uint32_t shard_id;        // only keys whose top shard_bits hash bits equal shard_id are kept (shard_bits may be 0)
    uint64_t* out_keys;       // MODE 1: flat hashed keys (sentinel-padded)
    unsigned long long* out_cursor;
};

__device__ __forceinline__ bool is_iupac(uint8_t c) {
    switch (c | 0x20) {
        case 'r': case 'y': case 'm': case 'k': case 's': case 'w': case 'h': case 'b': case 'v': case 'd': return true;
        default: return false;
    }
}

// The bytes of a warp's 8 consecutive records form one contiguous span of the chunk (2.5 KB for 150 bp FASTQ).  Reading
// them lane by lane and byte by byte costs ~10 L1 wavefronts per load instruction; instead the warp copies the span into
// shared memory with coalesced 16-byte loads and the per-base loop reads from there.  Spans that do not fit (long reads)
// keep the direct global loads.
#define EX_STAGE_BYTES 4096
__device__ __forceinline__ uint4 ld_stream128(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <bool FASTA, int MODE>
__global__ void __launch_bounds__(256) k_extract(McExtractArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint8_t(*s_stage)[EX_STAGE_BYTES] = reinterpret_cast<uint8_t(*)[EX_STAGE_BYTES]>(s_dyn);
    const int lane = threadIdx.x & 31;
    const int part = lane & 3;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int k = a.k;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int rc_shift = 2 * k - 2;

    unsigned long long st_reads = 0, st_runs = 0, st_bases = 0, st_kmers = 0, st_new = 0;

    // each warp takes 8 consecutive records per step; all lanes run the same number of steps
    for (uint32_t g = warp_global; (uint64_t)g * 8 < a.n_records; g += warps_total) {
        const uint32_t r = g * 8 + (lane >> 2);
        const bool live = r < a.n_records;
        McRec rec = live ? a.recs[r] : McRec{0, 0, 0, 0};
        const uint8_t* seq = a.base + rec.seq_off;
        const uint8_t* qual = a.base + rec.qual_off;
        {
            uint32_t lo = live ? rec.seq_off : 0xFFFFFFFFu;
            uint32_t hi = !live ? 0u : FASTA ? rec.seq_off + rec.seq_len : max(rec.seq_off + rec.seq_len, rec.qual_off + rec.qual_len);
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (hi > lo && hi <= a.n_bytes) {
                const uint8_t* g = a.base + lo;
                const uint32_t al = (uint32_t)((uintptr_t)g & 15);
                const uint32_t span = hi - lo + al;
                if (span <= EX_STAGE_BYTES) {
                    const uint8_t* g0 = g - al;
                    const uint8_t* end = a.base + a.n_bytes;
                    uint8_t* sw = s_stage[threadIdx.x >> 5];
                    __syncwarp();   // the previous group's readers are done
                    for (uint32_t off = lane * 16; off < span; off += 512) {
                        const uint8_t* p = g0 + off;
                        if (p >= a.base && p + 16 <= end) {
                            *reinterpret_cast<uint4*>(sw + off) = ld_stream128(p);
                        } else {
                            for (int j = 0; j < 16; j++)
                                if (p + j >= a.base && p + j < end) sw[off + j] = p[j];
                        }
                    }
                    __syncwarp();
                    seq = sw + al + (rec.seq_off - lo);
                    qual = sw + al + (rec.qual_off - lo);
                }
            }
        }
        uint32_t len = rec.seq_len;
        bool skip = !live;
        if (!FASTA && live && rec.qual_len != rec.seq_len) {
            if (part == 0) mc_report_error(a.header, MC_PE_FQ_LEN, rec.seq_off);
            skip = true;
        }
        // ---- validation + statistics over my quarter of the record ---------------------------
        uint32_t q_lo = (uint32_t)(((uint64_t)len * part) >> 2), q_hi = (uint32_t)(((uint64_t)len * (part + 1)) >> 2);
        bool has_n = false;
        uint32_t good = 0, splits = 0, bad_pos = 0xFFFFFFFFu, bad_code = 0;
        if (!skip) {
            for (uint32_t j = q_lo; j < q_hi; j++) {
                uint8_t c = seq[j];
                int code = mc_nuc_code(c);
                if (FASTA) {
                    if (code == 4) has_n = true;
                    else if (code == 5 && bad_pos == 0xFFFFFFFFu) { bad_pos = rec.seq_off + j; bad_code = is_iupac(c) ? MC_PE_IUPAC : MC_PE_NUC; }
                } else {
                    bool split = (code == 4) || (c == '.');
                    if (!split) {
                        if (code == 5) mc_report_error(a.header, is_iupac(c) ? MC_PE_IUPAC : MC_PE_NUC, rec.seq_off + j);
                        uint8_t q = qual[j];
                        if (q < a.qoff || q > 126) mc_report_error(a.header, MC_PE_QUAL, rec.qual_off + j);
                        split = (((uint32_t)q - (uint32_t)a.qoff) & 63u) == 0;
                    }
                    if (split) { if (j + 1 != len) splits++; }
                    else good++;
                }
            }
        }
        if (FASTA) {
            // the whole quad must agree on has_n before anything is reported or counted
            unsigned m = __ballot_sync(0xffffffffu, has_n);
            bool rec_has_n = (m >> (lane & ~3)) & 0xF;
            if (!skip && !rec_has_n && bad_pos != 0xFFFFFFFFu) mc_report_error(a.header, bad_code, bad_pos);
            if (rec_has_n || len == 0) skip = true;
            if (!skip && part == 0) { st_reads++; st_runs++; st_bases += len; }
        } else if (!skip) {
            if (part == 0) { st_reads++; st_runs++; }
            st_runs += splits;
            st_bases += good;
        }
        // ---- rolling extraction over my share of k-mer start positions ------------------------
        uint32_t nk = (!skip && len >= (uint32_t)k) ? len - k + 1 : 0;
        uint32_t S = (nk + 3) >> 2;
        uint32_t s_lo = min(nk, part * S), s_hi = min(nk, s_lo + S);
        uint32_t reserve = s_hi - s_lo;
        unsigned long long out_pos = 0;
        if (MODE == 1) {
            // warp-aggregated reservation of `reserve` slots per lane in the flat key buffer
            uint32_t incl = reserve;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long base = 0;
            if (lane == 0 && warp_total) base = atomicAdd(a.out_cursor, (unsigned long long)warp_total);
            base = __shfl_sync(0xffffffffu, base, 0);
            out_pos = base + incl - reserve;
        }
        // (collecting a warp's keys in shared memory for coalesced stores was measured slower: the extra 64 KB per CTA
        // halves the occupancy this ALU/latency-bound loop needs, profiles/r1c_NOTES.md)
        uint64_t* const my_out = (MODE == 1) ? a.out_keys + out_pos : nullptr;
        uint32_t emitted = 0;
        if (reserve) {
            uint64_t fw = 0, rc = 0;
            uint32_t run = 0;
            const uint32_t j_end = s_hi + k - 1;   // exclusive; <= len
            for (uint32_t j = s_lo; j < j_end; j++) {
                uint8_t c = seq[j];
                int code = mc_nuc_code(c);
                bool bad = code >= 4;
                if (!FASTA && !bad) {
                    uint8_t q = qual[j];
                    bad = (((uint32_t)q - (uint32_t)a.qoff) & 63u) == 0;
                }
                uint64_t b = (uint64_t)(code & 3);
                fw = ((fw << 2) | b) & kmask;
                rc = (rc >> 2) | ((3ULL - b) << rc_shift);
                run = bad ? 0 : run + 1;
                if (run >= (uint32_t)k) {
                    uint64_t key = fw < rc ? fw : rc;
                    uint64_t h = mc_mix62(key);
                    st_kmers++;
                    if (MODE == 0) {
                        if (a.table.shard_bits == 0 || (h >> (62 - a.table.shard_bits)) == a.shard_id)
                            if (mc_table_add(a.table, h, 1)) st_new++;
                    } else {
                        my_out[emitted] = h;
                        emitted++;
                    }
                }
            }
        }
        if (MODE == 1) {
            for (uint32_t e = emitted; e < reserve; e++) my_out[e] = MC_HASH_SENTINEL;
        }
    }

    // ---- block-aggregated statistics ---------------------------------------------------------
    __shared__ unsigned long long s_acc[5];
    if (threadIdx.x < 5) s_acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long v[5] = {st_reads, st_runs, st_bases, st_kmers, st_new};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long x = v[i];
        for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&s_acc[i], x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_acc[0]) atomicAdd(&a.table.counters[MC_CNT_READS], s_acc[0]);
        if (s_acc[1]) atomicAdd(&a.table.counters[MC_CNT_RUNS], s_acc[1]);
        if (s_acc[2]) atomicAdd(&a.table.counters[MC_CNT_BASES], s_acc[2]);
        if (s_acc[3]) atomicAdd(&a.table.counters[MC_CNT_KMERS], s_acc[3]);
        if (s_acc[4]) atomicAdd(&a.table.counters[MC_CNT_KEYS], s_acc[4]);
    }
}

// --------------------------------------------------------------------------------------------
// host side: quality-offset sniffing (ITMO!/io/ReadersUtils.java:63-77) on the head of the input
// --------------------------------------------------------------------------------------------
// Restates "read the first 1000 records with Phred+64; IllegalQualityValueException => Phred+33".
// Returns 33 or 64; -1 if `p` ends before 1000 records were seen and more input exists (caller retries with more).
static int sniff_quality_offset(const uint8_t* p, size_t n, bool complete) {
    size_t pos = 0;
    auto read_line = [&](size_t& s, size_t& e) -> bool {   // BufferedReader.readLine
        if (pos >= n) return false;
        s = pos;
        while (pos < n && p[pos] != '\n' && p[pos] != '\r') pos++;
        e = pos;
        if (pos >= n && !complete) return false;   // partial line
        if (pos < n) {
            if (p[pos] == '\r' && pos + 1 < n && p[pos + 1] == '\n') pos += 2;
            else pos += 1;
        }
        return true;
    };
    auto next_data_line = [&](size_t& s, size_t& e) -> int {   // 1 ok, 0 eof/partial, -1 structural error (not an IllegalQualityValue)
        size_t ls, le;
        do {
            if (!read_line(ls, le)) return 0;
        } while (le == ls);
        if (p[ls] != '@' && p[ls] != '+') return -1;
        if (!read_line(s, e)) return 0;
        return 1;
    };
    for (int rec = 0; rec < 1000; rec++) {
        size_t ss, se, qs, qe;
        int a = next_data_line(ss, se);
        if (a == 0) return complete ? 64 : -1;
        if (a < 0) return 64;   // the structural error resurfaces, with its own message, when the file is parsed
        int b = next_data_line(qs, qe);
        if (b == 0) return complete ? 64 : -1;
        if (b < 0 || se - ss != qe - qs) return 64;
        for (size_t i = 0; i < se - ss; i++) {
            uint8_t c = p[ss + i];
            if (c == 'N' || c == 'n' || c == '.') continue;
            if (mc_nuc_code(c) >= 4) return 64;   // IllegalArgumentException propagates; reported by the parser later
            uint8_t q = p[qs + i];
            if (q < 64 || q > 126) return 33;
        }
    }
    return 64;
}

static int resolve_qoff_host(const uint8_t* buf, size_t len, int fmt) {
    if (fmt == MC_FMT_FASTQ33) return 33;
    if (fmt == MC_FMT_FASTQ64) return 64;
    size_t sample = std::min<size_t>(len, (size_t)1 << 20);
    while (true) {
        int r = sniff_quality_offset(buf, sample, sample == len);
        if (r > 0) return r;
        if (sample == len) return 64;
        sample = std::min<size_t>(len, sample * 4);
    }
}

static int resolve_qoff_device(mc_ctx* ctx, const uint8_t* dbuf, size_t len, int fmt) {
    if (fmt == MC_FMT_FASTQ33) return 33;
    if (fmt == MC_FMT_FASTQ64) return 64;
    size_t sample = std::min<size_t>(len, (size_t)1 << 20);
    std::vector<uint8_t> h;
    while (true) {
        h.resize(sample);
        MC_CUDA(cudaMemcpyAsync(h.data(), dbuf, sample, cudaMemcpyDeviceToHost, ctx->stream));
        MC_CUDA(cudaStreamSynchronize(ctx->stream));
        int r = sniff_quality_offset(h.data(), sample, sample == len);
        if (r > 0) return r;
        if (sample == len) return 64;
        sample = std::min<size_t>(len, sample * 4);
    }
}

// --------------------------------------------------------------------------------------------
// one device chunk: parse -> reserve table room -> extract+count.  Returns bytes consumed.
// --------------------------------------------------------------------------------------------
// reads back K2's error word (and, when given, the flat-key cursor) with one synchronisation
static unsigned long long check_k2_errors(mc_ctx* ctx, McParseHeader* d_header, const uint8_t* d_buf, uint32_t n, const unsigned long long* d_cursor = nullptr) {
    McParseHeader* hp = (McParseHeader*)ctx->h_pinned_small;
    unsigned long long* hc = (unsigned long long*)((uint8_t*)ctx->h_pinned_small + 512);
    *hc = 0;
    MC_CUDA(cudaMemcpyAsync(hp, d_header, sizeof(McParseHeader), cudaMemcpyDeviceToHost, ctx->stream));
    if (d_cursor) MC_CUDA(cudaMemcpyAsync(hc, d_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hp->err_packed != MC_NO_ERROR) mc_throw_parse_error(hp->err_packed, d_buf, n, ctx);
    return *hc;
}

template <bool FASTA, int MODE>
static void launch_extract(mc_ctx* ctx, McExtractArgs& a) {
    const size_t smem = 8 * EX_STAGE_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA(cudaFuncSetAttribute(k_extract<FASTA, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    uint64_t warps = ((uint64_t)a.n_records + 7) / 8;
    uint64_t blocks = (warps + 7) / 8;                           // 8 warps per 256-thread CTA
    uint64_t cap = (uint64_t)ctx->sm_count * 6 * 4;   // resident CTAs per SM x 4 waves
    if (blocks < 1) blocks = 1;
    k_extract<FASTA, MODE><<<(unsigned)std::min(blocks, cap), 256, smem, ctx->stream>>>(a);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
}


#define MC_FLAT_HEAD 8192   // cursor + per-owner counters and cursors in front of the flat keys
// Launches k_extract in its flat-key mode into scratch slot 2; returns the device buffer, *n_flat = keys written (sentinels included)
static uint64_t* extract_flat(mc_ctx* ctx, McExtractArgs& a, bool fasta, uint64_t ub, const uint8_t* d_buf, uint32_t n, uint64_t* n_flat) {
    cudaStream_t st = ctx->stream;
    uint8_t* s2 = (uint8_t*)mc_scratch(ctx, 2, MC_FLAT_HEAD + ub * sizeof(uint64_t));
    unsigned long long* d_cursor = (unsigned long long*)s2;
    uint64_t* d_flat = (uint64_t*)(s2 + MC_FLAT_HEAD);
    MC_CUDA(cudaMemsetAsync(s2, 0, MC_FLAT_HEAD, st));
    a.out_keys = d_flat;
    a.out_cursor = d_cursor;
    {
        McStageTimer tm(ctx, MC_ST_EXTRACT);
        if (fasta) launch_extract<true, 1>(ctx, a);
        else launch_extract<false, 1>(ctx, a);
    }
    *n_flat = check_k2_errors(ctx, a.header, d_buf, n, d_cursor);
    if (*n_flat > ub) mc_throw(MC_ERR_CUDA, "internal: %llu k-mers extracted from a chunk bounded by %llu", (unsigned long long)*n_flat, (unsigned long long)ub);
    return d_flat;
}

static uint32_t count_chunk(mc_ctx* ctx, mc_table* t, const uint8_t* d_buf, uint32_t n, bool fasta, int qoff, bool final_chunk, uint64_t total_hint) {
    McParsed P;
    {
        McStageTimer tm(ctx, MC_ST_PARSE);
        mc_parse_chunk(ctx, d_buf, n, fasta, final_chunk, &P);
    }
    if (P.h.err_packed != MC_NO_ERROR) mc_throw_parse_error(P.h.err_packed, d_buf, n, ctx);
    if (P.h.n_records > 0) {
        uint64_t ub = fasta ? P.h.total_seq_bytes : (uint64_t)n / 2;
        McExtractArgs a{};
        a.base = P.base;
        a.recs = P.recs;
        a.n_records = P.h.n_records;
        a.n_bytes = fasta ? P.h.total_seq_bytes : n;
        a.k = t->k;
        a.qoff = qoff;
        a.header = P.d_header;
        a.table = t->dev;
        a.shard_id = t->shard_id;
        if (ctx->tune_direct) {
            // the round-1a design, kept for A/B measurements: extraction fused with direct probing of the HBM-sized table
            mc_table_prepare_chunk(t, ub);
            a.table = t->dev;
            {
                McStageTimer tm(ctx, MC_ST_EXTRACT);
                if (fasta) launch_extract<true, 0>(ctx, a);
                else launch_extract<false, 0>(ctx, a);
            }
            check_k2_errors(ctx, P.d_header, d_buf, n);
            mc_table_finish_chunk(t);
        } else {
            // extract -> flat hashed keys -> bins; the table itself is touched only when the batch is flushed
            mc_batch_reserve(t, ub + 1, total_hint);
            a.table = t->dev;   // a flush may have grown the table (the extractor only uses its statistics counters)
            uint64_t n_flat = 0;
            uint64_t* d_flat = extract_flat(ctx, a, fasta, ub + 1, d_buf, n, &n_flat);
            mc_batch_scatter(t, d_flat, n_flat);
        }
    }
    uint32_t consumed = P.h.consumed;
    if (!final_chunk && consumed == 0 && n >= ctx->chunk_bytes / 2) mc_throw_parse_error(((unsigned long long)0 << 8) | MC_PE_TOO_LONG, nullptr, 0, ctx);
    return consumed;
}

static void finish_table(mc_table* t) {
    mc_table_flush(t);
    mc_resolve_timers(t->ctx);
    unsigned int f = 0;
    MC_CUDA(cudaMemcpyAsync(&f, t->dev.flags, sizeof f, cudaMemcpyDeviceToHost, t->ctx->stream));
    MC_CUDA(cudaStreamSynchronize(t->ctx->stream));
    if (f & MC_FLAG_OVERFLOW) mc_throw(MC_ERR_NOMEM, "k-mer table overflow (stash full); raise capacity_hint");
}

static void validate_fmt(int fmt) {
    if (fmt != MC_FMT_FASTA && fmt != MC_FMT_FASTQ && fmt != MC_FMT_FASTQ33 && fmt != MC_FMT_FASTQ64) mc_throw(MC_ERR_ARG, "unknown input format %d", fmt);
}

extern "C" int mc_count_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, mc_table* t) {
    try {
        if (!ctx || !t || (len && !dbuf)) mc_throw(MC_ERR_ARG, "null argument");
        validate_fmt(fmt);
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_resolve_timers(ctx);
        for (auto& v : ctx->stage_ms) v = 0;
        const uint8_t* p = (const uint8_t*)dbuf;
        bool fasta = fmt == MC_FMT_FASTA;
        int qoff = fasta ? 0 : resolve_qoff_device(ctx, p, len, fmt);
        size_t pos = 0;
        while (pos < len) {
            size_t n = std::min(len - pos, ctx->chunk_bytes);
            bool final_chunk = pos + n == len;
            uint32_t consumed = count_chunk(ctx, t, p + pos, (uint32_t)n, fasta, qoff, final_chunk, mc_max_kmers(len, fmt, t->k));
            if (final_chunk) break;
            if (consumed == 0) mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk (%zu bytes); raise mc_ctx_set_chunk_bytes", ctx->chunk_bytes);
            pos += consumed;
        }
        finish_table(t);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// Host buffer (pinned memory gives full PCIe speed, pageable works too).  Two device buffers of 2*chunk bytes:
// the slice of chunk i+1 is uploaded on the copy stream to offset `chunk` of the idle buffer while chunk i is parsed
// and counted on the compute stream; the unconsumed tail of chunk i (an incomplete record, < chunk bytes) is then
// moved device-to-device right in front of that slice, so every chunk the kernels see starts at a record boundary.
extern "C" int mc_count_bytes(mc_ctx* ctx, const void* buf, size_t len, int fmt, mc_table* t) {
    cudaEvent_t copied[2] = {nullptr, nullptr};
    cudaEvent_t consumed_ev = nullptr;   // the compute stream is done reading the buffer the next prefetch overwrites
    try {
        if (!ctx || !t || (len && !buf)) mc_throw(MC_ERR_ARG, "null argument");
        validate_fmt(fmt);
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_resolve_timers(ctx);
        for (auto& v : ctx->stage_ms) v = 0;
        const uint8_t* p = (const uint8_t*)buf;
        const bool fasta = fmt == MC_FMT_FASTA;
        const int qoff = fasta ? 0 : resolve_qoff_host(p, len, fmt);
        const size_t chunk = ctx->chunk_bytes;
        const size_t cap = 2 * chunk + 256;
        for (int i = 0; i < 2; i++)
            if (ctx->d_chunk_cap[i] < cap) {
                MC_CUDA(cudaStreamSynchronize(ctx->stream));
                cudaFree(ctx->d_chunk[i]);
                ctx->d_chunk[i] = nullptr;
                ctx->d_chunk_cap[i] = 0;
                cudaError_t e = cudaMalloc(&ctx->d_chunk[i], cap);
                if (e != cudaSuccess) { cudaGetLastError(); mc_throw(MC_ERR_NOMEM, "cannot allocate the %zu-byte device chunk buffers", cap); }
                ctx->d_chunk_cap[i] = cap;
            }
        for (int i = 0; i < 2; i++) MC_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        MC_CUDA(cudaEventCreateWithFlags(&consumed_ev, cudaEventDisableTiming));
        size_t pos = 0;      // next host byte to upload
        size_t carry = 0;    // bytes of the previous chunk sitting right in front of offset `chunk`
        int cur = 0;
        size_t n_slice = std::min(len, chunk);
        if (n_slice) MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[cur] + chunk, p, n_slice, cudaMemcpyHostToDevice, ctx->copy_stream));
        MC_CUDA(cudaEventRecord(copied[cur], ctx->copy_stream));
        pos = n_slice;
        while (true) {
            const bool final_chunk = pos == len;
            const int nxt = cur ^ 1;
            size_t n_next = 0;
            if (!final_chunk) {   // prefetch the next slice; the idle buffer's kernels finished (count_chunk synchronises)
                n_next = std::min(len - pos, chunk);
                // buffer `nxt` held the previous chunk: its kernels and the carry copy out of it must have finished
                MC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, consumed_ev, 0));
                MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[nxt] + chunk, p + pos, n_next, cudaMemcpyHostToDevice, ctx->copy_stream));
                MC_CUDA(cudaEventRecord(copied[nxt], ctx->copy_stream));
            }
            MC_CUDA(cudaStreamWaitEvent(ctx->stream, copied[cur], 0));
            const uint8_t* d_start = (const uint8_t*)ctx->d_chunk[cur] + chunk - carry;
            const size_t n = carry + n_slice;
            uint32_t consumed = count_chunk(ctx, t, d_start, (uint32_t)n, fasta, qoff, final_chunk, mc_max_kmers(len, fmt, t->k));
            if (final_chunk) break;
            size_t left = n - consumed;
            if (left > chunk) mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk (%zu bytes); raise mc_ctx_set_chunk_bytes", chunk);
            if (left) MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[nxt] + chunk - left, d_start + consumed, left, cudaMemcpyDeviceToDevice, ctx->stream));
            MC_CUDA(cudaEventRecord(consumed_ev, ctx->stream));
            carry = left;
            pos += n_next;
            n_slice = n_next;
            cur = nxt;
        }
        for (int i = 0; i < 2; i++) cudaEventDestroy(copied[i]);
        cudaEventDestroy(consumed_ev);
        finish_table(t);
        return MC_OK;
    } catch (const McError& e) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 2; i++) if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed_ev) cudaEventDestroy(consumed_ev);
        mc_set_error(e.what());
        return e.code;
    }
}

// --------------------------------------------------------------------------------------------
// files: format by suffix (ITMO!/io/ReadersUtils.java:27-54), one file after the other (src/io/IOUtils.java:283-316)
// --------------------------------------------------------------------------------------------
static int detect_file_format(const std::string& path) {
    size_t slash = path.find_last_of('/');
    std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    for (auto& c : name) c = (char)tolower((unsigned char)c);
    auto ends = [&](const char* suf) { size_t m = strlen(suf); return name.size() >= m && name.compare(name.size() - m, m, suf) == 0; };
    if (ends(".gz") || ends(".bz2")) mc_throw(MC_ERR_UNSUPPORTED, "Failed to read from file %s (compressed input is not supported by the GPU path)", path.c_str());
    if (ends(".binq")) mc_throw(MC_ERR_UNSUPPORTED, "Failed to read from file %s (binq input is not supported by the GPU path)", path.c_str());
    if (ends(".fastq") || ends(".fq")) return MC_FMT_FASTQ;
    if (ends(".fasta") || ends(".fa") || ends(".fn") || ends(".fna")) return MC_FMT_FASTA;
    mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());   // "Can't detect file format" wrapped by IOUtils.java:292
}

extern "C" int mc_count_file(mc_ctx* ctx, const char* path, mc_table* t) {
    try {
        if (!ctx || !path || !t) mc_throw(MC_ERR_ARG, "null argument");
        int fmt = detect_file_format(path);
        int fd = open(path, O_RDONLY);
        if (fd < 0) mc_throw(MC_ERR_IO, "Failed to read from file %s", path);
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); mc_throw(MC_ERR_IO, "Failed to read from file %s", path); }
        size_t len = (size_t)sb.st_size;
        void* map = len ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
        if (len && map == MAP_FAILED) { close(fd); mc_throw(MC_ERR_IO, "Failed to read from file %s", path); }
        if (len) madvise(map, len, MADV_SEQUENTIAL);
        int rc = mc_count_bytes(ctx, map, len, fmt, t);
        if (len) munmap(map, len);
        close(fd);
        return rc;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_count_files(mc_ctx* ctx, const char* const* paths, int n_paths, int k, mc_table** out) {
    mc_table* t = nullptr;
    try {
        if (!ctx || !out || (n_paths && !paths)) mc_throw(MC_ERR_ARG, "null argument");
        int rc = mc_table_create(ctx, k, 0, 0, 0, &t);
        if (rc != MC_OK) return rc;
        for (int i = 0; i < n_paths; i++) {
            rc = mc_count_file(ctx, paths[i], t);
            if (rc != MC_OK) { mc_table_free(t); return rc; }
        }
        *out = t;
        return MC_OK;
    } catch (const McError& e) {
        if (t) mc_table_free(t);
        mc_set_error(e.what());
        return e.code;
    }
}

// --------------------------------------------------------------------------------------------
// multi-GPU building block: extract the hashed keys of a device-resident batch and group them by owner
// --------------------------------------------------------------------------------------------
extern "C" uint64_t mc_max_kmers(size_t len, int fmt, int k) {
    (void)k;
    return fmt == MC_FMT_FASTA ? (uint64_t)len : (uint64_t)len / 2 + 1;
}

#define MC_MAX_OWNERS 256
__global__ void __launch_bounds__(256) k_owner_count(const uint64_t* __restrict__ keys, uint64_t n, int shard_bits, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_hist[MC_MAX_OWNERS];
    for (int i = threadIdx.x; i < MC_MAX_OWNERS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = keys[i];
        if (h == MC_HASH_SENTINEL) continue;
        atomicAdd(&s_hist[shard_bits ? (uint32_t)(h >> (62 - shard_bits)) : 0u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (1 << shard_bits); i += blockDim.x)
        if (s_hist[i]) atomicAdd(&counts[i], (unsigned long long)s_hist[i]);
}
// each CTA works on a contiguous tile so that it can reserve one run per owner and fill it with shared-memory cursors
__global__ void __launch_bounds__(256) k_owner_scatter(const uint64_t* __restrict__ keys, uint64_t n, int shard_bits,
                                                       unsigned long long* __restrict__ cursors, uint64_t* __restrict__ out, uint64_t tile) {
    __shared__ unsigned int s_hist[MC_MAX_OWNERS];
    __shared__ unsigned long long s_base[MC_MAX_OWNERS];
    const int owners = 1 << shard_bits;
    for (uint64_t t0 = blockIdx.x * tile; t0 < n; t0 += (uint64_t)gridDim.x * tile) {
        uint64_t t1 = min(n, t0 + tile);
        for (int i = threadIdx.x; i < owners; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (uint64_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
            uint64_t h = keys[i];
            if (h != MC_HASH_SENTINEL) atomicAdd(&s_hist[shard_bits ? (uint32_t)(h >> (62 - shard_bits)) : 0u], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < owners; i += blockDim.x) {
            s_base[i] = s_hist[i] ? atomicAdd(&cursors[i], (unsigned long long)s_hist[i]) : 0ULL;
            s_hist[i] = 0;
        }
        __syncthreads();
        for (uint64_t i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
            uint64_t h = keys[i];
            if (h == MC_HASH_SENTINEL) continue;
            uint32_t o = shard_bits ? (uint32_t)(h >> (62 - shard_bits)) : 0u;
            out[s_base[o] + atomicAdd(&s_hist[o], 1u)] = h;
        }
        __syncthreads();
    }
}

extern "C" int mc_extract_partition_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, int k, int shard_bits, uint64_t* d_keys,
                                           uint64_t* counts, mc_table* stats_sink) {
    try {
        if (!ctx || !d_keys || !counts || !stats_sink || (len && !dbuf)) mc_throw(MC_ERR_ARG, "null argument");
        validate_fmt(fmt);
        if (shard_bits < 0 || shard_bits > 8) mc_throw(MC_ERR_ARG, "shard_bits must be in 0..8");
        if (len > ((size_t)3 << 30)) mc_throw(MC_ERR_ARG, "batch larger than 3 GiB");
        if (k != stats_sink->k) mc_throw(MC_ERR_ARG, "k differs from the table's");
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const int owners = 1 << shard_bits;
        for (int i = 0; i < owners; i++) counts[i] = 0;
        const uint8_t* p = (const uint8_t*)dbuf;
        const bool fasta = fmt == MC_FMT_FASTA;
        const int qoff = fasta ? 0 : resolve_qoff_device(ctx, p, len, fmt);
        McParsed P;
        mc_parse_chunk(ctx, p, (uint32_t)len, fasta, true, &P);
        if (P.h.err_packed != MC_NO_ERROR) mc_throw_parse_error(P.h.err_packed, p, (uint32_t)len, ctx);
        if (P.h.n_records == 0) return MC_OK;
        const uint64_t ub = mc_max_kmers(len, fmt, k);
        McExtractArgs a{};
        a.base = P.base;
        a.recs = P.recs;
        a.n_records = P.h.n_records;
        a.n_bytes = fasta ? P.h.total_seq_bytes : (uint32_t)len;
        a.k = k;
        a.qoff = qoff;
        a.header = P.d_header;
        a.table = stats_sink->dev;
        uint64_t n_flat = 0;
        uint64_t* d_flat = extract_flat(ctx, a, fasta, ub, p, (uint32_t)len, &n_flat);
        // the head of slot 2 (8 KiB, zeroed by extract_flat): [0] flat cursor, [8..) per-owner counters, then cursors
        unsigned long long* d_counts = (unsigned long long*)((uint8_t*)d_flat - MC_FLAT_HEAD) + 8;
        unsigned long long* d_cursors = d_counts + MC_MAX_OWNERS;
        if (n_flat == 0) return MC_OK;
        int grid = (int)std::min<uint64_t>((n_flat + 4095) / 4096, (uint64_t)ctx->sm_count * 8);
        k_owner_count<<<grid, 256, 0, st>>>(d_flat, n_flat, shard_bits, d_counts);
        MC_LAUNCH(ctx);
        unsigned long long h_counts[MC_MAX_OWNERS];
        MC_CUDA(cudaMemcpyAsync(h_counts, d_counts, owners * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        unsigned long long h_bases[MC_MAX_OWNERS], acc = 0;
        for (int i = 0; i < owners; i++) { h_bases[i] = acc; acc += h_counts[i]; counts[i] = h_counts[i]; }
        MC_CUDA(cudaMemcpyAsync(d_cursors, h_bases, owners * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        k_owner_scatter<<<grid, 256, 0, st>>>(d_flat, n_flat, shard_bits, d_cursors, d_keys, 4096);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaStreamSynchronize(st));
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
