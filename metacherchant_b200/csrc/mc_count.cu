// K2 + K3: 2-bit canonical k-mer extraction fused with the count-table update, and the host-side
// chunk pipeline behind Seam 1 (IOUtils.loadReads, src/io/IOUtils.java:217-248).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
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
    uint32_t shard_id;        // (kept for the statistics sink; keys are filtered by shard when they are binned)
    uint64_t* out_keys;       // flat hashed keys (sentinel-padded)
    unsigned long long* out_cursor;
    int use_tma;              // stage the warps' record spans with bulk copies + mbarriers (0: per-lane 16-byte loads)
    McSkDev sk;               // super-k-mer mode (k_extract<.., W > 0>): records go to these minimizer bins, out_keys is unused
};

__device__ __forceinline__ bool is_iupac(uint8_t c) {
    switch (c | 0x20) {
        case 'r': case 'y': case 'm': case 'k': case 's': case 'w': case 'h': case 'b': case 'v': case 'd': return true;
        default: return false;
    }
}

// --------------------------------------------------------------------------------------------
// K2 kernel, second design (profiles/r1c_NOTES.md).  The first one (a quad per record, a validation pass followed by a
// rolling pass, byte loads from global memory, nested branches) executed ~210 thread instructions per k-mer and was
// issue-bound.  Now:
//   * the byte span of a warp's records is copied to shared memory with coalesced 16-byte loads and the loops read from there;
//   * PARTS lanes share one record; PARTS = 1 for short reads (one lane rolls over the whole read: no k-1 bases of
//     context are re-read), 4 or 32 for longer ones.  Lane p owns the k-mer starts [p*S, (p+1)*S), S = ceil(nk/PARTS),
//     and the reader statistics of the bases [p*S, (p+1)*S) (the last lane: up to the end of the read);
//   * one pass, branch-free per base: a 256-entry class table (2-bit code | split | invalid) in shared memory, error
//     conditions are only OR-ed into a flag and the exact first error is located by a rare re-scan;
//   * keys always leave as flat hashed keys (the binned table update consumes them; there is no direct-probing mode).
// --------------------------------------------------------------------------------------------
#define EX_THREADS 128
#define EX_WARPS (EX_THREADS / 32)
#ifndef EX_STAGE_BYTES
#define EX_STAGE_BYTES 12288   // per warp: 32 x 150 bp FASTQ records are ~10 KB
#endif
#define EX_CLS_SPLIT 4u        // no k-mer may span this base: N, n, '.' in FASTQ, and every invalid character
#define EX_CLS_BAD 8u          // not a nucleotide character (an error unless the FASTA record is dropped for an N)
#define EX_CLS_N 16u           // N or n
#define EX_FLAG_QUAL 32u       // quality character outside [offset, 126] under a real base
__device__ __forceinline__ uint4 ld_stream128(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint8_t ex_class(uint32_t c, bool fasta) {
    int code = mc_nuc_code((uint8_t)c);
    if (code < 4) return (uint8_t)code;
    if (code == 4) return EX_CLS_SPLIT | EX_CLS_N;
    if (!fasta && c == '.') return EX_CLS_SPLIT;
    return EX_CLS_SPLIT | EX_CLS_BAD;
}

__device__ __forceinline__ uint32_t ex_lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// ---- bulk asynchronous copy global -> shared with an mbarrier (TMA unit; SASS UBLKCP.S.G + SYNCS) -----------------------------------
// The byte span of a warp's records is ONE contiguous, 16-byte aligned piece of the chunk: lane 0 arms the warp's mbarrier with the
// byte count and hands the whole span to the copy engine; the lanes wait on the barrier's phase instead of moving 16 bytes each
// through registers (LDG.128 -> STS.128).  Spans that touch the edges of the buffer keep the per-lane loop.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t phase) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load_g2s(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

struct McRollOut {
    uint32_t emitted, good, splits, flags;
};
// Where a lane's e-th key goes inside its warp's block of the flat key buffer.  Lanes emit in lockstep, so the first `rows`
// keys of every active lane are interleaved (row e = `stride` consecutive slots, one per active lane: one coalesced store
// per emission step instead of 32 scattered 8-byte stores); what a lane has beyond `rows` follows in a private tail.
struct McOutSlots {
    uint64_t* row;    // block + my rank among the active lanes
    uint64_t* tail;   // my private tail
    uint32_t rows, stride;
    __device__ __forceinline__ uint64_t* at(uint32_t e) const { return e < rows ? row + (uint64_t)e * stride : tail + (e - rows); }
};
// One pass over bases [s_lo, j_end) of a read: reader statistics for [s_lo, own_hi), every k-mer that lies inside a run to `out`.
// STAGED: seq/qual are shared-memory addresses (the warp's staged span), else global pointers.
template <bool FASTA, bool STAGED>
__device__ __forceinline__ McRollOut ex_roll(uint64_t seq, uint64_t qual, uint32_t cls_saddr, uint32_t s_lo, uint32_t j_end, uint32_t own_hi, uint32_t len,
                                             int k, uint64_t kmask, int rc_shift, uint32_t qoff, uint32_t qspan, const McOutSlots& out) {
    McRollOut o{0, 0, 0, 0};
    uint64_t fw = 0, rc = 0;
    uint32_t run = 0;
#pragma unroll 2
    for (uint32_t j = s_lo; j < j_end; j++) {
        const uint32_t c = STAGED ? ex_lds_u8((uint32_t)seq + j) : (uint32_t) __ldg((const uint8_t*)seq + j);
        const uint32_t cls = ex_lds_u8(cls_saddr + c);
        uint32_t split = cls & EX_CLS_SPLIT;
        if (!FASTA) {
            const uint32_t q = STAGED ? ex_lds_u8((uint32_t)qual + j) : (uint32_t) __ldg((const uint8_t*)qual + j);
            const uint32_t qd = q - qoff;
            o.flags |= (cls & EX_CLS_BAD) | ((split == 0 && qd > qspan) ? EX_FLAG_QUAL : 0u);
            split |= ((qd & 63u) == 0) ? 1u : 0u;
            if (j < own_hi) {
                o.good += split ? 0u : 1u;
                o.splits += (split && j + 1 != len) ? 1u : 0u;
            }
        } else {
            o.flags |= cls;   // EX_CLS_N: the record holds an N, EX_CLS_BAD: an invalid character
        }
        const uint64_t b = cls & 3u;
        fw = ((fw << 2) | b) & kmask;
        rc = (rc >> 2) | ((3ULL - b) << rc_shift);
        run = split ? 0u : run + 1u;
        if (run >= (uint32_t)k) {
            *out.at(o.emitted) = mc_mix62(fw < rc ? fw : rc);
            o.emitted++;
        }
    }
    return o;
}

// ---- super-k-mer mode ------------------------------------------------------------------------------------------------------
// One finished super-k-mer -> its minimizer bin (or the spill list when the bin is full).  r0 = the first k-mer (forward strand),
// ext = the n - 1 bases that follow it (first one highest).  Not inlined: lanes finish records at different steps, the call sits
// behind a divergent branch and should stay out of the rolling loop's instruction footprint.
struct McSkSink {   // what sk_emit needs of McSkDev, passed by value (registers): a reference would put the struct on the local stack
    uint4* recs;
    unsigned long long* cursors;
    uint4* spill;
    unsigned int* flags;
    uint32_t cap, spill_cap, nb_mask;
};
__device__ __noinline__ void sk_emit(McSkSink o, uint32_t bin, uint64_t r0, uint64_t ext, uint32_t n, int k) {
    const int s0 = 122 - 2 * k;                      // 58 .. 78 for k in [22, 32]
    uint64_t hi = s0 >= 64 ? r0 << (s0 - 64) : r0 >> (64 - s0);
    uint64_t lo = s0 >= 64 ? 0ULL : r0 << s0;
    const uint32_t e = n - 1;
    if (e) {
        const int s1 = 2 * (61 - k - (int)e);        // >= 0: offset of the last base
        hi |= s1 >= 64 ? ext << (s1 - 64) : (s1 ? ext >> (64 - s1) : 0ULL);
        lo |= s1 >= 64 ? 0ULL : ext << s1;
    }
    hi |= (uint64_t)n << 58;
    const uint4 rec = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    const unsigned long long at = atomicAdd(&o.cursors[bin], ((unsigned long long)n << 32) | 1ULL);
    const uint32_t pos = (uint32_t)at;
    if (pos < o.cap) {
        o.recs[(uint64_t)bin * o.cap + pos] = rec;
    } else {
        const uint32_t nb = o.nb_mask + 1;
        const unsigned long long sp = atomicAdd(&o.cursors[nb], 1ULL);
        atomicAdd(&o.cursors[nb + 1], (unsigned long long)n);
        if (sp < o.spill_cap) o.spill[sp] = rec;
        else atomicOr(o.flags, MC_FLAG_OVERFLOW);   // the host withdraws this chunk and redoes it with flat keys
    }
}
// 32-bit order-and-bin hash of a canonical m-mer (m <= 16: the m-mer itself fits 32 bits)
__device__ __forceinline__ uint32_t sk_mhash(uint32_t cm) {
    uint32_t x = cm * 0x9E3779B1u;
    return x ^ (x >> 15);
}
__device__ __forceinline__ uint32_t sk_bin_of(uint32_t mn, uint32_t nb_mask) { return ((mn * 0x85EBCA6Bu) >> 13) & nb_mask; }

// The rolling pass of ex_roll with the k-mer hash replaced by the minimizer machinery.  The loop is unrolled W times so that every
// access to the register ring below has a static index; the iteration count is the same for all lanes of a warp (each starts at its
// own s_lo), so the unrolled position is warp-uniform.
//
// Window minimum in ~3 operations per base (van Herk / Gil-Werman): positions are cut into blocks of W.  While block B is being
// read, ring[0..u] holds its raw m-mer hashes and ring[u+1..W-1] the SUFFIX minima of block B - 1; the minimum over the last W
// m-mers is min(prefix minimum of block B up to u, ring[u + 1]).  At the end of a block one backward pass turns its raw values into
// suffix minima.  An m-mer that is not inside a run enters as 0xFFFFFFFF; a k-mer is complete when the run is >= k = m + W - 1 long,
// and by then the last W entries all belong to it.
template <bool FASTA, bool STAGED, int W>
__device__ __forceinline__ McRollOut ex_roll_sk(uint64_t seq, uint64_t qual, uint32_t cls_saddr, uint32_t s_lo, uint32_t j_end, uint32_t own_hi, uint32_t len,
                                                int k, uint64_t kmask, uint32_t qoff, uint32_t qspan, const McSkDev& skd) {
    McRollOut o{0, 0, 0, 0};
    const McSkSink sk{skd.recs, skd.cursors, skd.spill, skd.flags, (uint32_t)skd.cap, (uint32_t)min(skd.spill_cap, (uint64_t)0xFFFFFFFFu), skd.nb_mask};
    const uint32_t n_max = skd.n_max;
    uint64_t fw = 0;
    uint32_t run = 0;
    const int m = k - W + 1;                         // 11 .. 16
    const uint32_t mmask = m == 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    const int mshift = 2 * m - 2;
    uint32_t mf = 0, mr = 0;
    uint32_t ring[W];
#pragma unroll
    for (int u = 0; u < W; u++) ring[u] = 0xFFFFFFFFu;
    uint32_t prefix = 0xFFFFFFFFu;
    uint64_t r0 = 0, ext = 0;
    uint32_t cur_n = 0, cur_bin = 0;
    uint32_t j = s_lo;
    while (j < j_end) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            if (j < j_end) {
                const uint32_t c = STAGED ? ex_lds_u8((uint32_t)seq + j) : (uint32_t) __ldg((const uint8_t*)seq + j);
                const uint32_t cls = ex_lds_u8(cls_saddr + c);
                uint32_t split = cls & EX_CLS_SPLIT;
                if (!FASTA) {
                    const uint32_t q = STAGED ? ex_lds_u8((uint32_t)qual + j) : (uint32_t) __ldg((const uint8_t*)qual + j);
                    const uint32_t qd = q - qoff;
                    o.flags |= (cls & EX_CLS_BAD) | ((split == 0 && qd > qspan) ? EX_FLAG_QUAL : 0u);
                    split |= ((qd & 63u) == 0) ? 1u : 0u;
                    if (j < own_hi) {
                        o.good += split ? 0u : 1u;
                        o.splits += (split && j + 1 != len) ? 1u : 0u;
                    }
                } else {
                    o.flags |= cls;
                }
                const uint32_t b = cls & 3u;
                fw = ((fw << 2) | b) & kmask;
                mf = ((mf << 2) | b) & mmask;
                mr = (mr >> 2) | ((3u - b) << mshift);
                run = split ? 0u : run + 1u;
                const uint32_t mh = run >= (uint32_t)m ? sk_mhash(min(mf, mr)) : 0xFFFFFFFFu;
                prefix = u == 0 ? mh : min(prefix, mh);
                const uint32_t mn = u == W - 1 ? prefix : min(prefix, ring[u == W - 1 ? 0 : u + 1]);
                ring[u] = mh;
                if (u == W - 1) {
#pragma unroll
                    for (int x = W - 2; x >= 0; x--) ring[x] = min(ring[x], ring[x + 1]);
                }
                if (run >= (uint32_t)k) {
                    const uint32_t bin = sk_bin_of(mn, sk.nb_mask);
                    if (cur_n != 0 && bin == cur_bin && cur_n < n_max) {
                        ext = (ext << 2) | b;
                        cur_n++;
                    } else {
                        if (cur_n) sk_emit(sk, cur_bin, r0, ext, cur_n, k);
                        r0 = fw;
                        ext = 0;
                        cur_n = 1;
                        cur_bin = bin;
                    }
                    o.emitted++;
                } else if (cur_n) {   // the run ended
                    sk_emit(sk, cur_bin, r0, ext, cur_n, k);
                    cur_n = 0;
                }
                j++;
            }
        }
    }
    if (cur_n) sk_emit(sk, cur_bin, r0, ext, cur_n, k);
    return o;
}

template <bool FASTA, int PARTS, int W>
__global__ void __launch_bounds__(EX_THREADS) k_extract(McExtractArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    constexpr int RPW = 32 / PARTS;   // records per warp and step
    uint8_t* const s_cls = s_dyn + EX_WARPS * EX_STAGE_BYTES;
    __shared__ __align__(8) unsigned long long s_bar[EX_WARPS];   // one mbarrier per warp (its staging window)
    for (int i = threadIdx.x; i < 256; i += EX_THREADS) s_cls[i] = ex_class((uint32_t)i, FASTA);
    if ((threadIdx.x & 31) == 0) mbar_init((uint32_t)__cvta_generic_to_shared(&s_bar[threadIdx.x >> 5]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t my_bar = (uint32_t)__cvta_generic_to_shared(&s_bar[threadIdx.x >> 5]);
    uint32_t bar_phase = 0;
    const int lane = threadIdx.x & 31;
    const int part = lane & (PARTS - 1);
    const uint32_t group_mask = PARTS == 32 ? 0xffffffffu : (((1u << PARTS) - 1u) << (lane & ~(PARTS - 1)));
    uint8_t* const sw = s_dyn + (threadIdx.x >> 5) * EX_STAGE_BYTES;
    const uint32_t warps_total = (gridDim.x * EX_THREADS) >> 5;
    const uint32_t warp_global = (blockIdx.x * EX_THREADS + threadIdx.x) >> 5;
    const int k = a.k;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int rc_shift = 2 * k - 2;
    const uint32_t qoff = (uint32_t)a.qoff, qspan = 126u - qoff;

    unsigned long long st_reads = 0, st_runs = 0, st_bases = 0, st_kmers = 0;

    for (uint32_t g = warp_global; (uint64_t)g * RPW < a.n_records; g += warps_total) {
        const uint32_t r = g * RPW + lane / PARTS;
        const bool live = r < a.n_records;
        const McRec rec = live ? a.recs[r] : McRec{0, 0, 0, 0};
        const uint8_t* seq = a.base + rec.seq_off;
        const uint8_t* qual = a.base + rec.qual_off;
        bool staged = false;
        // ---- the warp's byte span -> shared memory ------------------------------------------------
        {
            uint32_t lo = live ? rec.seq_off : 0xFFFFFFFFu;
            uint32_t hi = !live ? 0u : FASTA ? rec.seq_off + rec.seq_len : max(rec.seq_off + rec.seq_len, rec.qual_off + rec.qual_len);
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            __syncwarp();   // the previous step's readers are done with the buffer
            if (hi > lo && hi <= a.n_bytes) {
                const uint8_t* gp = a.base + lo;
                const uint32_t al = (uint32_t)((uintptr_t)gp & 15);
                const uint32_t span = hi - lo + al;
                if (span <= EX_STAGE_BYTES) {
                    const uint8_t* g0 = gp - al;
                    const uint8_t* end = a.base + a.n_bytes;
                    const uint32_t span16 = (span + 15u) & ~15u;
                    if (a.use_tma && g0 >= a.base && g0 + span16 <= end && span16 <= EX_STAGE_BYTES) {
                        if (lane == 0) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the window's earlier (generic) readers before the engine's writes
                            mbar_expect_tx(my_bar, span16);
                            bulk_load_g2s((uint32_t)__cvta_generic_to_shared(sw), g0, span16, my_bar);
                        }
                        while (!mbar_try_wait(my_bar, bar_phase)) {}
                        bar_phase ^= 1u;
                    } else
                    for (uint32_t off = lane * 16; off < span; off += 512) {
                        const uint8_t* p = g0 + off;
                        if (p >= a.base && p + 16 <= end) {
                            *reinterpret_cast<uint4*>(sw + off) = ld_stream128(p);
                        } else {
                            for (int j = 0; j < 16; j++)
                                if (p + j >= a.base && p + j < end) sw[off + j] = p[j];
                        }
                    }
                    seq = sw + al + (rec.seq_off - lo);
                    qual = sw + al + (rec.qual_off - lo);
                    staged = true;
                }
            }
            __syncwarp();
        }
        const uint32_t len = rec.seq_len;
        bool skip = !live;
        if (!FASTA && live && rec.qual_len != rec.seq_len) {
            if (part == 0) mc_report_error(a.header, MC_PE_FQ_LEN, rec.seq_off);
            skip = true;
        }
        const uint32_t cls_saddr = (uint32_t)__cvta_generic_to_shared(s_cls);
        bool predrop = false;
        if (W > 0 && FASTA) {
            // records cannot be withdrawn from their bins afterwards: a FASTA record holding an N (dropped whole, FastaReader.java:58-59)
            // is recognised before anything is emitted
            uint32_t f = 0;
            if (!skip) {
                const uint32_t S0 = (len + PARTS - 1) / PARTS;
                const uint32_t lo0 = min(len, part * S0), hi0 = min(len, lo0 + S0);
                for (uint32_t j = lo0; j < hi0; j++) f |= s_cls[seq[j]];
            }
            predrop = (__ballot_sync(0xffffffffu, (f & EX_CLS_N) != 0) & group_mask) != 0;
        }
        // ---- my share of the record -------------------------------------------------------------
        const uint32_t nk = (!skip && !predrop && len >= (uint32_t)k) ? len - k + 1 : 0;
        const uint32_t S = (nk + PARTS - 1) / PARTS;
        const uint32_t s_lo = min(nk, part * S), s_hi = min(nk, s_lo + S);
        const uint32_t reserve = s_hi - s_lo;
        const bool idle = skip || predrop;
        const uint32_t own_hi = idle ? s_lo : (part == PARTS - 1 ? len : s_hi);          // statistics: bases [s_lo, own_hi)
        const uint32_t j_end = idle ? s_lo : (part == PARTS - 1 ? len : (reserve ? s_hi + k - 1 : s_lo));
        McRollOut ro;
        if (W == 0) {
            // warp-aggregated reservation of the warp's flat-key block (sum of `reserve`), laid out as McOutSlots describes
            McOutSlots out;
            {
                const uint32_t act = __ballot_sync(0xffffffffu, reserve > 0);
                out.stride = __popc(act);
                out.rows = act ? __reduce_min_sync(0xffffffffu, reserve ? reserve : 0xFFFFFFFFu) : 0u;
                const uint32_t rest = reserve > out.rows ? reserve - out.rows : 0u;
                uint32_t incl = rest;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t warp_total = out.rows * out.stride + __shfl_sync(0xffffffffu, incl, 31);
                unsigned long long base = 0;
                if (lane == 0 && warp_total) base = atomicAdd(a.out_cursor, (unsigned long long)warp_total);
                base = __shfl_sync(0xffffffffu, base, 0);
                out.row = a.out_keys + base + __popc(act & ((1u << lane) - 1u));
                out.tail = a.out_keys + base + (uint64_t)out.rows * out.stride + (incl - rest);
            }
            if (staged)
                ro = ex_roll<FASTA, true>((uint64_t)__cvta_generic_to_shared(seq), (uint64_t)__cvta_generic_to_shared(qual), cls_saddr, s_lo, j_end, own_hi, len, k,
                                          kmask, rc_shift, qoff, qspan, out);
            else
                ro = ex_roll<FASTA, false>((uint64_t)seq, (uint64_t)qual, cls_saddr, s_lo, j_end, own_hi, len, k, kmask, rc_shift, qoff, qspan, out);
            // (a FASTA record with an N: its k-mers are withdrawn below) -- unused slots of the reservation become sentinels
            const bool dropped = FASTA && (__ballot_sync(0xffffffffu, (ro.flags & EX_CLS_N) != 0) & group_mask) != 0;
            for (uint32_t e = dropped ? 0u : ro.emitted; e < reserve; e++) *out.at(e) = MC_HASH_SENTINEL;
        } else {
            constexpr int WW = W > 0 ? W : 12;
            if (staged)
                ro = ex_roll_sk<FASTA, true, WW>((uint64_t)__cvta_generic_to_shared(seq), (uint64_t)__cvta_generic_to_shared(qual), cls_saddr, s_lo, j_end, own_hi, len,
                                                 k, kmask, qoff, qspan, a.sk);
            else
                ro = ex_roll_sk<FASTA, false, WW>((uint64_t)seq, (uint64_t)qual, cls_saddr, s_lo, j_end, own_hi, len, k, kmask, qoff, qspan, a.sk);
            if (predrop) ro.flags |= EX_CLS_N;
        }
        uint32_t emitted = ro.emitted;
        const uint32_t good = ro.good, splits = ro.splits, flags = ro.flags;
        // ---- record-level outcome ---------------------------------------------------------------
        bool drop = false;
        if (FASTA) {
            // FastaReader.java:58-59: a record holding N or n is skipped whole; invalid characters only matter otherwise
            drop = (__ballot_sync(0xffffffffu, (flags & EX_CLS_N) != 0) & group_mask) != 0;
            if (!skip && !drop && (flags & EX_CLS_BAD)) {
                for (uint32_t j = s_lo; j < own_hi; j++) {
                    const uint8_t c = seq[j];
                    if (s_cls[c] & EX_CLS_BAD) { mc_report_error(a.header, is_iupac(c) ? MC_PE_IUPAC : MC_PE_NUC, rec.seq_off + j); break; }
                }
            }
            if (drop || len == 0) skip = true;
            if (!skip && part == 0) { st_reads++; st_runs++; st_bases += len; }
        } else if (!skip) {
            if (flags & (EX_CLS_BAD | EX_FLAG_QUAL)) {   // rare: locate the first offending base of my share, in the reader's order of checks
                for (uint32_t j = s_lo; j < own_hi; j++) {
                    const uint8_t c = seq[j];
                    const uint32_t cls = s_cls[c];
                    if (cls & EX_CLS_SPLIT) {
                        if (cls & EX_CLS_BAD) { mc_report_error(a.header, is_iupac(c) ? MC_PE_IUPAC : MC_PE_NUC, rec.seq_off + j); break; }
                        continue;
                    }
                    if ((uint32_t)qual[j] - qoff > qspan) { mc_report_error(a.header, MC_PE_QUAL, rec.qual_off + j); break; }
                }
            }
            if (part == 0) { st_reads++; st_runs++; }
            st_runs += splits;
            st_bases += good;
        }
        if (drop) emitted = 0;   // FASTA record with an N: its k-mers are withdrawn
        st_kmers += emitted;
    }

    // ---- block-aggregated statistics ---------------------------------------------------------
    __shared__ unsigned long long s_acc[4];
    if (threadIdx.x < 4) s_acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long v[4] = {st_reads, st_runs, st_bases, st_kmers};
#pragma unroll
    for (int i = 0; i < 4; i++) {
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
        if (W > 0 && s_acc[3]) atomicAdd(&a.sk.cursors[a.sk.nb_mask + 3], s_acc[3]);   // k-mers of this launch (the host's chunk accounting)
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

// super-k-mer mode: the extractor's error word, the k-mers it turned into records and the overflow flag of the bins, one synchronisation
static unsigned long long sk_finish_chunk(mc_ctx* ctx, mc_table* t, McParseHeader* d_header, const uint8_t* d_buf, uint32_t n, bool* overflow) {
    const McSkDev& sk = t->batch->sk;
    const uint32_t nb = sk.nb_mask + 1;
    McParseHeader* hp = (McParseHeader*)ctx->h_pinned_small;
    unsigned long long* hc = (unsigned long long*)((uint8_t*)ctx->h_pinned_small + 512);   // [0] k-mers of this launch, [1] flag word
    hc[0] = hc[1] = 0;
    MC_CUDA(cudaMemcpyAsync(hp, d_header, sizeof(McParseHeader), cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaMemcpyAsync(hc, &sk.cursors[nb + 2], 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    MC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hp->err_packed != MC_NO_ERROR) mc_throw_parse_error(hp->err_packed, d_buf, n, ctx);
    *overflow = ((unsigned int)hc[1] & MC_FLAG_OVERFLOW) != 0;
    return hc[0];
}

template <bool FASTA, int PARTS, int W>
static void launch_extract_p(mc_ctx* ctx, McExtractArgs& a) {
    const size_t smem = EX_WARPS * EX_STAGE_BYTES + 256;
    // function attributes are per device: one opt-in per context and instantiation (a process may hold contexts on several GPUs)
    const int variant = (FASTA ? 3 : 0) + (PARTS == 1 ? 0 : PARTS == 4 ? 1 : 2);
    if (mc_attr_once(ctx, W == 0 ? MC_ATTR_EXTRACT + variant : MC_ATTR_EXTRACT_SK + (W == 20 ? 6 : 0) + variant))
        MC_CUDA(cudaFuncSetAttribute(k_extract<FASTA, PARTS, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t per_cta = (uint64_t)(32 / PARTS) * EX_WARPS;          // records per CTA step
    uint64_t blocks = ((uint64_t)a.n_records + per_cta - 1) / per_cta;
    const uint64_t cap = (uint64_t)ctx->sm_count * 4 * 4;                 // resident CTAs per SM x 4 waves
    if (blocks < 1) blocks = 1;
    k_extract<FASTA, PARTS, W><<<(unsigned)std::min(blocks, cap), EX_THREADS, smem, ctx->stream>>>(a);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
}
// lanes per record: one for short reads (no context bases re-read), more when a warp's records would not fit its shared-memory
// window or a single lane would serialise a long read
template <bool FASTA, int W>
static void launch_extract_w(mc_ctx* ctx, McExtractArgs& a, uint32_t max_seq_len) {
    // one lane per record while the span of a warp's 32 records fits its staging window (FASTQ: 2 L + ~24 bytes per record, FASTA: L + ~10)
#ifndef EX_ONE_FASTQ
#define EX_ONE_FASTQ ((EX_STAGE_BYTES / 32 - 44) / 2)
#endif
    const uint32_t one = FASTA ? (uint32_t)(EX_STAGE_BYTES / 32 - 4) : (uint32_t)EX_ONE_FASTQ;
    if (ctx->tune_extract_parts == 1 || (ctx->tune_extract_parts == 0 && max_seq_len <= one)) launch_extract_p<FASTA, 1, W>(ctx, a);
    else if (ctx->tune_extract_parts == 4 || (ctx->tune_extract_parts == 0 && max_seq_len <= 4 * one)) launch_extract_p<FASTA, 4, W>(ctx, a);
    else launch_extract_p<FASTA, 32, W>(ctx, a);
}
// sk_w: 0 = flat hashed keys, else the minimizer window of the super-k-mer mode (mc_sk_window(k))
template <bool FASTA>
static void launch_extract(mc_ctx* ctx, McExtractArgs& a, uint32_t max_seq_len, int sk_w = 0) {
    if (sk_w == 20) launch_extract_w<FASTA, 20>(ctx, a, max_seq_len);
    else if (sk_w == 12) launch_extract_w<FASTA, 12>(ctx, a, max_seq_len);
    else launch_extract_w<FASTA, 0>(ctx, a, max_seq_len);
}

#define MC_FLAT_HEAD 8192   // cursor + per-owner counters and cursors in front of the flat keys
// Launches k_extract in its flat-key mode into scratch slot 2; returns the device buffer, *n_flat = keys written (sentinels included)
static uint64_t* extract_flat(mc_ctx* ctx, McExtractArgs& a, bool fasta, uint32_t max_seq_len, uint64_t ub, const uint8_t* d_buf, uint32_t n, uint64_t* n_flat) {
    cudaStream_t st = ctx->stream;
    uint8_t* s2 = (uint8_t*)mc_scratch(ctx, 2, MC_FLAT_HEAD + ub * sizeof(uint64_t));
    unsigned long long* d_cursor = (unsigned long long*)s2;
    uint64_t* d_flat = (uint64_t*)(s2 + MC_FLAT_HEAD);
    MC_CUDA(cudaMemsetAsync(s2, 0, MC_FLAT_HEAD, st));
    a.out_keys = d_flat;
    a.out_cursor = d_cursor;
    a.use_tma = ctx->tune_extract_tma;
    {
        McStageTimer tm(ctx, MC_ST_EXTRACT);
        if (fasta) launch_extract<true>(ctx, a, max_seq_len);
        else launch_extract<false>(ctx, a, max_seq_len);
    }
    *n_flat = check_k2_errors(ctx, a.header, d_buf, n, d_cursor);
    if (*n_flat > ub) mc_throw(MC_ERR_CUDA, "internal: %llu k-mers extracted from a chunk bounded by %llu", (unsigned long long)*n_flat, (unsigned long long)ub);
    return d_flat;
}

static uint32_t count_chunk(mc_ctx* ctx, mc_table* t, const uint8_t* d_buf, uint32_t n, bool fasta, int qoff, bool final_chunk, uint64_t total_hint,
                            uint64_t stream_bytes = 0) {
    McParsed P;
    {
        McHostTimer ht(ctx, MC_ST_PARSE);
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
        bool done = false;
        if (ctx->tune_superkmers && mc_sk_supported(t->k)) {
            // super-k-mer records -> minimizer bins (mc_sk.cu); unpacked into hashed keys when the table is flushed
            bool room;
            {
                McHostTimer ht(ctx, MC_ST_GROW);
                room = mc_sk_reserve(t, ub + 1, total_hint);
            }
            if (room) {
                McHostTimer ht(ctx, MC_ST_EXTRACT);
                McBatch* B = t->batch;
                a.table = t->dev;
                a.sk = B->sk;
                a.use_tma = ctx->tune_extract_tma;
                a.out_keys = nullptr;
                a.out_cursor = nullptr;
                mc_sk_snapshot(t);
                {
                    McStageTimer tm(ctx, MC_ST_EXTRACT);
                    if (fasta) launch_extract<true>(ctx, a, P.h.max_seq_len, mc_sk_window(t->k));
                    else launch_extract<false>(ctx, a, P.h.max_seq_len, mc_sk_window(t->k));
                }
                // one synchronisation: the extractor's error word, the k-mers it found (counter delta) and the overflow flag of the bins
                unsigned long long n_kmers = 0;
                bool overflow = false;
                try {
                    n_kmers = sk_finish_chunk(ctx, t, a.header, d_buf, n, &overflow);
                } catch (...) {
                    mc_sk_rollback(t);   // the chunk's records leave the bins again: an input error applies nothing of its chunk
                    throw;
                }
                if (!overflow) {
                    B->sk_kmers += n_kmers;
                    B->sk_records_est += (uint64_t)((double)n_kmers * 2.0 / (mc_sk_window(t->k) + 1)) + P.h.n_records;
                    if (stream_bytes && t->expect_instances == 0 && P.h.consumed)
                        mc_table_expect_keys(t, (uint64_t)((double)n_kmers * ((double)stream_bytes / (double)P.h.consumed)));
                    done = true;
                } else {
                    // the bins AND their spill list filled up (input far from what the bins were sized for): withdraw the chunk, unpack
                    // what the bins hold, and take this chunk through the flat-key path
                    mc_sk_rollback(t);
                    mc_sk_flush(t);
                }
            }
        }
        if (!done) {
            // extract -> flat hashed keys -> bins; the table itself is touched only when the batch is flushed
            {
                McHostTimer ht(ctx, MC_ST_GROW);   // (host view: reserving / creating the batch of bins)
                mc_batch_reserve(t, ub + 1, total_hint);
            }
            a.table = t->dev;   // a flush may have grown the table (the extractor only uses its statistics counters)
            uint64_t n_flat = 0;
            uint64_t* d_flat;
            {
                McHostTimer ht(ctx, MC_ST_EXTRACT);
                d_flat = extract_flat(ctx, a, fasta, P.h.max_seq_len, ub + 1, d_buf, n, &n_flat);
            }
            // a stream of chunks (host input): the first chunk's k-mers per byte, scaled to the whole input, tells the flush how
            // many more instances to expect
            if (stream_bytes && t->expect_instances == 0 && P.h.consumed)
                mc_table_expect_keys(t, (uint64_t)((double)n_flat * ((double)stream_bytes / (double)P.h.consumed)));
            McHostTimer ht(ctx, MC_ST_SCATTER);
            mc_batch_scatter(t, d_flat, n_flat);
        }
    }
    uint32_t consumed = P.h.consumed;
    if (!final_chunk && consumed == 0 && n >= ctx->chunk_bytes / 2) mc_throw_parse_error(((unsigned long long)0 << 8) | MC_PE_TOO_LONG, nullptr, 0, ctx);
    return consumed;
}

static void finish_table(mc_table* t) {
    if (!t->ctx->tune_defer_flush) {
        McHostTimer ht(t->ctx, MC_ST_COUNT);   // host view of the whole flush
        mc_table_flush(t);
    }
    McHostTimer ht2(t->ctx, 6);
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
        if (!ctx->keep_stage_ms)
            for (auto& v : ctx->stage_ms) v = 0;
        const uint8_t* p = (const uint8_t*)dbuf;
        bool fasta = fmt == MC_FMT_FASTA;
        int qoff = fasta ? 0 : resolve_qoff_device(ctx, p, len, fmt);
        // Device-resident input needs no upload pipelining, so it is cut into the largest pieces the 32-bit record offsets allow
        // (3 GiB) rather than into upload-sized chunks: every chunk costs four host round trips (line count, record count, key
        // count, spill cursor), and on a slow host those gaps, not the kernels, were a tenth of the step.
        const size_t dev_chunk = ctx->tune_device_chunk_bytes ? ctx->tune_device_chunk_bytes : std::max(ctx->chunk_bytes, ((size_t)3 << 30) - ((size_t)1 << 20));
        size_t pos = 0;
        while (pos < len) {
            size_t n = std::min(len - pos, dev_chunk);
            bool final_chunk = pos + n == len;
            uint32_t consumed = count_chunk(ctx, t, p + pos, (uint32_t)n, fasta, qoff, final_chunk, mc_max_kmers(len, fmt, t->k));
            if (final_chunk) break;
            if (consumed == 0) mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk (%zu bytes); raise mc_ctx_set_chunk_bytes", dev_chunk);
            pos += consumed;
        }
        finish_table(t);
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}

// --------------------------------------------------------------------------------------------
// host input as a sequence of slices: a caller's buffer, or a decompressor's output produced slice by slice (a .gz / .bz2 file is
// never inflated whole: two slices of chunk_bytes are all the host memory it needs)
// --------------------------------------------------------------------------------------------
struct McSliceSource {
    virtual ~McSliceSource() {}
    // the first (up to) `want` bytes of the input, without consuming them; *complete = that is the whole input
    virtual size_t peek(size_t want, const uint8_t** p, bool* complete) = 0;
    // the next slice of at most `max` bytes (0: the input is over); the memory stays valid until the call after the next one
    virtual size_t next(size_t max, const uint8_t** p) = 0;
};
struct McMemSource : McSliceSource {
    const uint8_t* base;
    size_t len, pos = 0;
    McMemSource(const void* b, size_t n) : base((const uint8_t*)b), len(n) {}
    size_t peek(size_t want, const uint8_t** p, bool* complete) override {
        *p = base;
        *complete = want >= len;
        return std::min(want, len);
    }
    size_t next(size_t max, const uint8_t** p) override {
        const size_t n = std::min(max, len - pos);
        *p = base + pos;
        pos += n;
        return n;
    }
};
// a decompressor: pull() delivers the inflated bytes in order, 0 at the end of the input
struct McPullSource : McSliceSource {
    std::vector<uint8_t> head;
    std::unique_ptr<uint8_t[]> buf[2];   // (not vectors: a resize would zero-fill a GiB per slice buffer)
    size_t buf_cap[2] = {0, 0};
    size_t head_pos = 0;
    int cur = 0;
    bool eof = false;
    virtual size_t pull(uint8_t* dst, size_t max) = 0;
    size_t peek(size_t want, const uint8_t** p, bool* complete) override {
        while (head.size() < want && !eof) {
            const size_t old = head.size();
            head.resize(want);
            const size_t got = pull(head.data() + old, want - old);
            head.resize(old + got);
            if (!got) eof = true;
        }
        *p = head.data();
        *complete = eof;
        return head.size();
    }
    size_t next(size_t max, const uint8_t** p) override {
        if (buf_cap[cur] < max) {
            buf[cur].reset(new uint8_t[max]);
            buf_cap[cur] = max;
        }
        uint8_t* const b = buf[cur].get();
        cur ^= 1;
        size_t got = 0;
        while (got < max) {
            if (head_pos < head.size()) {
                const size_t n = std::min(max - got, head.size() - head_pos);
                memcpy(b + got, head.data() + head_pos, n);
                head_pos += n;
                got += n;
                if (head_pos == head.size()) { std::vector<uint8_t>().swap(head); head_pos = 0; }
            } else if (eof) {
                break;
            } else {
                const size_t n = pull(b + got, max - got);
                if (!n) eof = true;
                got += n;
            }
        }
        *p = b;
        return got;
    }
};

// ReadersUtils.determineQualityFormat on the head of the input (ITMO!/io/ReadersUtils.java:63-77): the first 1000 records
static int resolve_qoff_source(McSliceSource& src, int fmt) {
    if (fmt == MC_FMT_FASTQ33) return 33;
    if (fmt == MC_FMT_FASTQ64) return 64;
    size_t sample = (size_t)1 << 20;
    while (true) {
        const uint8_t* p = nullptr;
        bool complete = false;
        const size_t n = src.peek(sample, &p, &complete);
        const int r = sniff_quality_offset(p, n, complete);
        if (r > 0) return r;
        if (complete) return 64;
        sample *= 4;
    }
}

// Host input (pinned memory gives full PCIe speed, pageable works too).  Two device buffers of 2*chunk bytes:
// the slice of chunk i+1 is uploaded on the copy stream to offset `chunk` of the idle buffer while chunk i is parsed
// and counted on the compute stream; the unconsumed tail of chunk i (an incomplete record, < chunk bytes) is then
// moved device-to-device right in front of that slice, so every chunk the kernels see starts at a record boundary.
// total_len: the input's size when known (table sizing from the first chunk's k-mer density), 0 for a stream.
static int count_source(mc_ctx* ctx, McSliceSource& src, size_t total_len, int fmt, mc_table* t) {
    cudaEvent_t copied[2] = {nullptr, nullptr};
    cudaEvent_t consumed_ev = nullptr;   // the compute stream is done reading the buffer the next prefetch overwrites
    try {
        validate_fmt(fmt);
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_resolve_timers(ctx);
        if (!ctx->keep_stage_ms)
            for (auto& v : ctx->stage_ms) v = 0;
        const bool fasta = fmt == MC_FMT_FASTA;
        const int qoff = fasta ? 0 : resolve_qoff_source(src, fmt);
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
                ctx->scratch_allocs++;
            }
        for (int i = 0; i < 2; i++) MC_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        MC_CUDA(cudaEventCreateWithFlags(&consumed_ev, cudaEventDisableTiming));
        // Host input is PCIe-bound (~20 ms per GiB): the bins are applied after every chunk (measured best: profiles/r1c_NOTES.md), so that the table sweep runs in the
        // shadow of the uploads instead of after the last one (device-resident input is binned whole, one sweep)
        const uint64_t host_chunks = ctx->tune_host_batch_chunks ? ctx->tune_host_batch_chunks : 1;
        uint64_t batch_hint = host_chunks * mc_max_kmers(chunk, fmt, t->k);
        if (total_len) batch_hint = std::min<uint64_t>(batch_hint, mc_max_kmers(total_len, fmt, t->k));
        mc_table_expect_keys(t, 0);   // set from the first chunk's k-mers per byte (count_chunk)
        size_t carry = 0;    // bytes of the previous chunk sitting right in front of offset `chunk`
        int cur = 0;
        const uint8_t* ps = nullptr;
        size_t n_slice = src.next(chunk, &ps);
        if (n_slice) MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[cur] + chunk, ps, n_slice, cudaMemcpyHostToDevice, ctx->copy_stream));
        MC_CUDA(cudaEventRecord(copied[cur], ctx->copy_stream));
        while (true) {
            const int nxt = cur ^ 1;
            const uint8_t* pn = nullptr;
            const size_t n_next = n_slice ? src.next(chunk, &pn) : 0;   // (a decompressor produces the slice here, on the host)
            const bool final_chunk = n_next == 0;
            if (!final_chunk) {   // prefetch the next slice; the idle buffer's kernels finished (count_chunk synchronises)
                // buffer `nxt` held the previous chunk: its kernels and the carry copy out of it must have finished
                MC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, consumed_ev, 0));
                MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[nxt] + chunk, pn, n_next, cudaMemcpyHostToDevice, ctx->copy_stream));
                MC_CUDA(cudaEventRecord(copied[nxt], ctx->copy_stream));
            }
            MC_CUDA(cudaStreamWaitEvent(ctx->stream, copied[cur], 0));
            const uint8_t* d_start = (const uint8_t*)ctx->d_chunk[cur] + chunk - carry;
            const size_t n = carry + n_slice;
            uint32_t consumed = count_chunk(ctx, t, d_start, (uint32_t)n, fasta, qoff, final_chunk, batch_hint, total_len);
            if (final_chunk) break;
            size_t left = n - consumed;
            if (left > chunk) mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk (%zu bytes); raise mc_ctx_set_chunk_bytes", chunk);
            if (left) MC_CUDA(cudaMemcpyAsync((uint8_t*)ctx->d_chunk[nxt] + chunk - left, d_start + consumed, left, cudaMemcpyDeviceToDevice, ctx->stream));
            MC_CUDA(cudaEventRecord(consumed_ev, ctx->stream));
            carry = left;
            n_slice = n_next;
            cur = nxt;
        }
        for (int i = 0; i < 2; i++) cudaEventDestroy(copied[i]);
        cudaEventDestroy(consumed_ev);
        finish_table(t);
        mc_table_expect_keys(t, 0);
        return MC_OK;
    } catch (const McError& e) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 2; i++) if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed_ev) cudaEventDestroy(consumed_ev);
        if (t) mc_table_expect_keys(t, 0);
        mc_set_error(e.what());
        return e.code;
    }
}

extern "C" int mc_count_bytes(mc_ctx* ctx, const void* buf, size_t len, int fmt, mc_table* t) {
    if (!ctx || !t || (len && !buf)) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    McMemSource src(buf, len);
    return count_source(ctx, src, len, fmt, t);
}

// --------------------------------------------------------------------------------------------
// files: format by suffix (ITMO!/io/ReadersUtils.java:27-54), one file after the other (src/io/IOUtils.java:283-316)
// --------------------------------------------------------------------------------------------
// libbz2 through dlopen: the image ships libbz2.so.1.0 but no bzlib.h; the structure and the three entry points used are the
// library's stable public ABI (bzlib.h 1.0.x).  Concatenated streams are read one after the other, like
// org.apache.hadoop.io.compress.BZip2Codec (ITMO!/io/readers/FastqBZ2Reader.java:26-34, FastaBZ2Reader.java).
struct McBzStream {
    char* next_in;
    unsigned int avail_in, total_in_lo32, total_in_hi32;
    char* next_out;
    unsigned int avail_out, total_out_lo32, total_out_hi32;
    void* state;
    void* (*bzalloc)(void*, int, int);
    void (*bzfree)(void*, void*);
    void* opaque;
};
struct McBz2Source : McPullSource {
    typedef int (*init_t)(McBzStream*, int, int);
    typedef int (*step_t)(McBzStream*);
    std::string path;
    const uint8_t* in;
    size_t n, pos = 0;
    init_t init = nullptr;
    step_t step = nullptr, end = nullptr;
    McBzStream z{};
    bool active = false;
    McBz2Source(const char* path_, const uint8_t* in_, size_t n_) : path(path_), in(in_), n(n_) {
        static void* lib = nullptr;
        if (!lib) lib = dlopen("libbz2.so.1.0", RTLD_NOW);
        if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW);
        if (!lib) mc_throw(MC_ERR_UNSUPPORTED, "Failed to read from file %s (libbz2 is not installed)", path.c_str());
        init = (init_t)dlsym(lib, "BZ2_bzDecompressInit");
        step = (step_t)dlsym(lib, "BZ2_bzDecompress");
        end = (step_t)dlsym(lib, "BZ2_bzDecompressEnd");
        if (!init || !step || !end) mc_throw(MC_ERR_UNSUPPORTED, "Failed to read from file %s (libbz2 lacks the decompression entry points)", path.c_str());
    }
    ~McBz2Source() override {
        if (active) end(&z);
    }
    size_t pull(uint8_t* dst, size_t max) override {
        size_t got = 0;
        while (got < max) {
            if (!active) {
                if (pos >= n) break;   // no further stream
                z = McBzStream{};
                if (init(&z, 0, 0) != 0) mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());
                active = true;
                z.next_in = (char*)in + pos;
                z.avail_in = (unsigned)std::min<size_t>(n - pos, 1u << 30);
            }
            z.next_out = (char*)dst + got;
            z.avail_out = (unsigned)std::min<size_t>(max - got, 1u << 30);
            const unsigned in_before = z.avail_in, out_before = z.avail_out;
            const int rc = step(&z);
            if (rc != 0 && rc != 4) mc_throw(MC_ERR_IO, "Failed to read from file %s (corrupt bzip2 stream)", path.c_str());
            got += out_before - z.avail_out;
            pos += in_before - z.avail_in;
            if (rc == 4) {   // BZ_STREAM_END: a concatenated stream may follow
                end(&z);
                active = false;
            } else if (z.avail_in == 0) {
                if (pos >= n) mc_throw(MC_ERR_IO, "Failed to read from file %s (truncated bzip2 stream)", path.c_str());
                z.next_in = (char*)in + pos;
                z.avail_in = (unsigned)std::min<size_t>(n - pos, 1u << 30);
            }
        }
        return got;
    }
};
// zlib's gzread: concatenated members like java.util.zip.GZIPInputStream (ITMO!/io/readers/FastqGZReader.java:26-34)
struct McGzSource : McPullSource {
    std::string path;
    gzFile g;
    McGzSource(const char* path_) : path(path_), g(gzopen(path_, "rb")) {
        if (!g) mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());
        gzbuffer(g, 1 << 20);
    }
    ~McGzSource() override { gzclose(g); }
    size_t pull(uint8_t* dst, size_t max) override {
        size_t got = 0;
        while (got < max) {
            const int r = gzread(g, dst + got, (unsigned)std::min<size_t>(max - got, 1u << 30));
            if (r < 0) mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());
            if (r == 0) break;
            got += (size_t)r;
        }
        return got;
    }
};

// *gz: 0 plain, 1 ".gz", 2 ".bz2" (ReadersUtils strips ".gz", then ".bz2": ITMO!/io/ReadersUtils.java:36-43)
static int detect_file_format(const std::string& path, int* gz) {
    size_t slash = path.find_last_of('/');
    std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    for (auto& c : name) c = (char)tolower((unsigned char)c);
    auto ends = [&](const char* suf) { size_t m = strlen(suf); return name.size() >= m && name.compare(name.size() - m, m, suf) == 0; };
    *gz = 0;
    if (ends(".gz")) {   // "fastq.gz" / "fasta.gz": the same readers over a GZIPInputStream (ReadersUtils.java:36-39, FastqGZReader.java:26-34)
        *gz = 1;
        name.resize(name.size() - 3);
    }
    if (ends(".bz2")) {  // "fastq.bz2" / "fasta.bz2": the same readers over Hadoop's BZip2Codec stream
        *gz = 2;
        name.resize(name.size() - 4);
    }
    if (ends(".gz") || ends(".bz2")) mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());   // "Can't detect file format"
    if (ends(".binq")) mc_throw(MC_ERR_UNSUPPORTED, "Failed to read from file %s (binq input is not supported by the GPU path)", path.c_str());
    if (ends(".fastq") || ends(".fq")) return MC_FMT_FASTQ;
    if (ends(".fasta") || ends(".fa") || ends(".fn") || ends(".fna")) return MC_FMT_FASTA;
    mc_throw(MC_ERR_IO, "Failed to read from file %s", path.c_str());   // "Can't detect file format" wrapped by IOUtils.java:292
}

extern "C" int mc_detect_file_format(const char* path, int* gz) {
    try {
        if (!path) mc_throw(MC_ERR_ARG, "null argument");
        int z = 0;
        const int fmt = detect_file_format(path, &z);
        if (gz) *gz = z;
        return fmt;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
extern "C" int mc_sniff_quality_offset(const void* buf, size_t len) {
    if (len && !buf) {
        mc_set_error("null argument");
        return MC_ERR_ARG;
    }
    return resolve_qoff_host((const uint8_t*)buf, len, MC_FMT_FASTQ);
}

extern "C" int mc_count_file(mc_ctx* ctx, const char* path, mc_table* t) {
    try {
        if (!ctx || !path || !t) mc_throw(MC_ERR_ARG, "null argument");
        int gz = 0;
        int fmt = detect_file_format(path, &gz);
        // .gz / .bz2: inflated on the host SLICE BY SLICE (zlib; libbz2 through dlopen) into two alternating buffers of chunk_bytes that
        // feed the usual chunked upload: host memory stays at two slices whatever the file's size, and the GPU counts slice i while the
        // decompressor -- which bounds such an input either way -- produces slice i + 1
        if (gz == 2) {
            int fd = open(path, O_RDONLY);
            if (fd < 0) mc_throw(MC_ERR_IO, "Failed to read from file %s", path);
            struct stat sb;
            if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); mc_throw(MC_ERR_IO, "Failed to read from file %s", path); }
            const size_t len = (size_t)sb.st_size;
            void* map = len ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
            if (len && map == MAP_FAILED) { close(fd); mc_throw(MC_ERR_IO, "Failed to read from file %s", path); }
            int rc;
            try {
                McBz2Source src(path, (const uint8_t*)map, len);
                rc = count_source(ctx, src, 0, fmt, t);
            } catch (...) {
                if (len) munmap(map, len);
                close(fd);
                throw;
            }
            if (len) munmap(map, len);
            close(fd);
            return rc;
        }
        if (gz == 1) {
            McGzSource src(path);
            return count_source(ctx, src, 0, fmt, t);
        }
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

// parse + extract one device-resident, record-aligned batch; the flat hashed keys stay in scratch slot 2 (valid until the next
// extraction on this context).  Shared by the NCCL path below and the peer-memory exchange (mc_xchg.cu).
uint64_t* mc_extract_flat_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, int k, mc_table* stats_sink, uint64_t* n_flat) {
    if (len > ((size_t)3 << 30)) mc_throw(MC_ERR_ARG, "batch larger than 3 GiB");
    if (k != stats_sink->k) mc_throw(MC_ERR_ARG, "k differs from the table's");
    *n_flat = 0;
    const uint8_t* p = (const uint8_t*)dbuf;
    const bool fasta = fmt == MC_FMT_FASTA;
    const int qoff = fasta ? 0 : resolve_qoff_device(ctx, p, len, fmt);
    McParsed P;
    mc_parse_chunk(ctx, p, (uint32_t)len, fasta, true, &P);
    if (P.h.err_packed != MC_NO_ERROR) mc_throw_parse_error(P.h.err_packed, p, (uint32_t)len, ctx);
    if (P.h.n_records == 0) return nullptr;
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
    return extract_flat(ctx, a, fasta, P.h.max_seq_len, ub, p, (uint32_t)len, n_flat);
}

extern "C" int mc_extract_partition_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, int k, int shard_bits, uint64_t* d_keys,
                                           uint64_t* counts, mc_table* stats_sink) {
    try {
        if (!ctx || !d_keys || !counts || !stats_sink || (len && !dbuf)) mc_throw(MC_ERR_ARG, "null argument");
        validate_fmt(fmt);
        if (shard_bits < 0 || shard_bits > 8) mc_throw(MC_ERR_ARG, "shard_bits must be in 0..8");
        MC_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const int owners = 1 << shard_bits;
        for (int i = 0; i < owners; i++) counts[i] = 0;
        uint64_t n_flat = 0;
        uint64_t* d_flat = mc_extract_flat_device(ctx, dbuf, len, fmt, k, stats_sink, &n_flat);
        if (!d_flat) return MC_OK;
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
