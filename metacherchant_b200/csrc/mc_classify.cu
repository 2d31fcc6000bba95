// "next" row f-4b: the core of the reference's reads-classifier on the device.
//
// Restates ReadsFinderInGraph.findRead / getCoverage / getTheoryWidth / delta / findReadWithCorrection
// (src/algo/ReadsFinderInGraph.java:37-123) for every read of a FASTQ buffer, one warp per read:
//   * the read is taken as LazyDnaQReaderTool yields it (ITMO!/io/readers/FastqReader.java:53-82, NO quality truncation):
//     N, n and '.' are nucleotide 0 with phred 0, every other base keeps (quality char - offset) & 63 as its phred;
//   * coverage of every k-mer = graph.getWithZero(min(fw, rc)) (0 when absent): the lanes split the k-mer start positions
//     into 32 contiguous pieces and roll the forward / reverse-complement words over their piece;
//   * breadth = (#covered k-mers + [last covered] * (k-1)) / length, depth = (sum + last * (k-1)) / length,
//     found = !(breadth < threshold) && (breadth == 1 || (breadth != 0 && |breadth - (1 - e^-depth)| <= z * sqrt(e^-depth (1 - e^-depth) / length)));
//   * --correction: a read with exactly ONE base of phred < 10 is tried with each of the four nucleotides at that position,
//     with the threshold 0.9 the reference hard-codes there (:113).
// The sums are integers and exact; the final comparison is IEEE double arithmetic with exp and sqrt from the CUDA math
// library (<= 1 ulp, like java.lang.Math): a class can differ from the JVM's only for a read whose breadth sits within an ulp
// of the interval's edge.  Tests compare classes with the oracle and skip reads the oracle reports to be that close.
#include <algorithm>
#include <cmath>
#include <vector>

#include "mc_parse.cuh"

#define CL_THREADS 256

struct McClassifyArgs {
    const uint8_t* base;
    const McRec* recs;
    uint32_t n_records;
    int k, qoff, correction;
    double z, threshold;
    McTableDev table;
    uint32_t shard_id;
    uint8_t* found;
    McParseHeader* header;
};

__device__ __forceinline__ bool cl_is_iupac(uint8_t c) {
    switch (c | 0x20) {
        case 'r': case 'y': case 'm': case 'k': case 's': case 'w': case 'h': case 'b': case 'v': case 'd': return true;
        default: return false;
    }
}
// nucleotide of a base character as the reader stores it: N, n, '.' -> 0 (unsafeAppendUnknown)
__device__ __forceinline__ uint32_t cl_nuc(uint8_t c) {
    const int code = mc_nuc_code(c);
    return code < 4 ? (uint32_t)code : 0u;
}

struct McCover { int sum, nonzero, last; };
// coverage sums of one read with (optionally) the base at `opos` replaced by `onuc`; every lane returns the warp totals
__device__ __forceinline__ McCover cl_cover(const McClassifyArgs& a, const uint8_t* seq, uint32_t len, int opos, uint32_t onuc, int lane) {
    const int k = a.k;
    const uint32_t nk = len - (uint32_t)k + 1;
    const uint32_t S = (nk + 31) / 32;
    const uint32_t lo = min(nk, (uint32_t)lane * S), hi = min(nk, lo + S);
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    int sum = 0, nz = 0, last = 0;
    if (lo < hi) {
        uint64_t fw = 0, rc = 0;
        for (uint32_t j = lo; j < hi + (uint32_t)k - 1; j++) {
            uint64_t b = cl_nuc(__ldg(seq + j));
            if ((int)j == opos) b = onuc;
            fw = ((fw << 2) | b) & kmask;
            rc = (rc >> 2) | ((3ULL - b) << (2 * k - 2));
            if (j + 1 >= lo + (uint32_t)k) {
                int c = mc_table_lookup(a.table, mc_mix62(fw < rc ? fw : rc));
                if (c < 0) c = 0;   // getWithZero
                sum += c;
                nz += c > 0;
                if (j + 1 == len) last = c;
            }
        }
    }
    McCover r;
    r.sum = __reduce_add_sync(0xffffffffu, sum);
    r.nonzero = __reduce_add_sync(0xffffffffu, nz);
    r.last = __reduce_add_sync(0xffffffffu, last);
    return r;
}
__device__ __forceinline__ bool cl_width_test(const McCover& c, int k, uint32_t len, double z, double threshold) {
    const double depth = (double)(c.sum + c.last * (k - 1)) / (double)len;
    const double width = (double)(c.nonzero + (c.last > 0 ? 1 : 0) * (k - 1)) / (double)len;
    const double e = exp(-depth);
    const double theory = 1.0 - e;
    const double sd = z * sqrt(e * (1 - e) / (double)len);
    const double d = width - theory;
    const bool delta = width == 1 || (width != 0 && -sd <= d && d <= sd);
    return !(width < threshold) && delta;
}

__global__ void __launch_bounds__(CL_THREADS) k_classify(McClassifyArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * CL_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * CL_THREADS) >> 5;
    for (uint32_t r = warp; r < a.n_records; r += n_warps) {
        const McRec rec = a.recs[r];
        const uint8_t* seq = a.base + rec.seq_off;
        const uint8_t* qual = a.base + rec.qual_off;
        const uint32_t len = rec.seq_len;
        if (rec.qual_len != rec.seq_len) {
            if (lane == 0) { mc_report_error(a.header, MC_PE_FQ_LEN, rec.seq_off); a.found[r] = 0; }
            continue;
        }
        // validation + the positions of phred < 10 (FastqReader.java:72-80, ReadsFinderInGraph.java:94-99)
        uint32_t n_bad = 0, first_bad = 0xFFFFFFFFu;
        for (uint32_t j = lane; j < len; j += 32) {
            const uint8_t c = __ldg(seq + j), q = __ldg(qual + j);
            uint32_t phred = 0;
            if (!(c == 'N' || c == 'n' || c == '.')) {
                if (mc_nuc_code(c) >= 4) mc_report_error(a.header, cl_is_iupac(c) ? MC_PE_IUPAC : MC_PE_NUC, rec.seq_off + j);
                if (q < (uint32_t)a.qoff || q > 126u) mc_report_error(a.header, MC_PE_QUAL, rec.qual_off + j);
                phred = ((uint32_t)q - (uint32_t)a.qoff) & 63u;
            }
            if (phred < 10) { n_bad++; first_bad = min(first_bad, j); }
        }
        n_bad = __reduce_add_sync(0xffffffffu, n_bad);
        first_bad = __reduce_min_sync(0xffffffffu, first_bad);
        bool found = false;
        if (len >= (uint32_t)a.k) {
            if (a.correction && n_bad == 1) {
                for (uint32_t nuc = 0; nuc < 4 && !found; nuc++) found = cl_width_test(cl_cover(a, seq, len, (int)first_bad, nuc, lane), a.k, len, a.z, 0.9);
            } else {
                found = cl_width_test(cl_cover(a, seq, len, -1, 0, lane), a.k, len, a.z, a.threshold);
            }
        }
        if (lane == 0) a.found[r] = found ? 1 : 0;
    }
}

// Host buffer of whole FASTQ records -> found[] in read order.  The chunk loop is the one of mc_count_bytes without the
// upload pipelining: classification is a one-off pass whose cost is the table probes, not the copy.
extern "C" int mc_classify_bytes(mc_ctx* ctx, const mc_table* graph, const void* buf, size_t len, int fmt, double z, double found_threshold, int correction,
                                 uint8_t* found, size_t cap, size_t* n_reads) {
    try {
        if (!ctx || !graph || !n_reads || (len && !buf)) mc_throw(MC_ERR_ARG, "null argument");
        if (fmt != MC_FMT_FASTQ && fmt != MC_FMT_FASTQ33 && fmt != MC_FMT_FASTQ64) mc_throw(MC_ERR_UNSUPPORTED, "reads-classifier: FASTQ input only");
        if (graph->dev.shard_bits != 0) mc_throw(MC_ERR_UNSUPPORTED, "reads-classifier needs a complete table, not one shard");
        MC_CUDA(cudaSetDevice(ctx->device));
        mc_table_flush(const_cast<mc_table*>(graph));
        cudaStream_t st = ctx->stream;
        const uint8_t* p = (const uint8_t*)buf;
        const int qoff = fmt == MC_FMT_FASTQ33 ? 33 : fmt == MC_FMT_FASTQ64 ? 64 : mc_sniff_quality_offset(buf, len);
        const size_t chunk = ctx->chunk_bytes;
        uint8_t* d_buf = (uint8_t*)mc_dev_alloc(ctx, 2 * chunk + 256, "classifier chunk");
        struct Free { mc_ctx* c; void* p; ~Free() { mc_dev_free(c, p); } } free_buf{ctx, d_buf};
        size_t pos = 0, carry = 0, total = 0;
        std::vector<uint8_t> tail;   // bytes of an incomplete record carried into the next chunk
        while (true) {
            const size_t n_slice = std::min(len - pos, chunk);
            const bool final_chunk = pos + n_slice == len;
            if (carry) MC_CUDA(cudaMemcpyAsync(d_buf, tail.data(), carry, cudaMemcpyHostToDevice, st));
            if (n_slice) MC_CUDA(cudaMemcpyAsync(d_buf + carry, p + pos, n_slice, cudaMemcpyHostToDevice, st));
            const size_t n = carry + n_slice;
            McParsed P;
            mc_parse_chunk(ctx, d_buf, (uint32_t)n, false, final_chunk, &P);
            if (P.h.err_packed != MC_NO_ERROR) mc_throw_parse_error(P.h.err_packed, d_buf, (uint32_t)n, ctx);
            if (P.h.n_records) {
                if (total + P.h.n_records > cap) {
                    if (!found || cap == 0) {   // counting call: only the number of reads is wanted
                    } else mc_throw(MC_ERR_CAPACITY, "room for %zu reads, the input holds more", cap);
                }
                uint8_t* d_found = (uint8_t*)mc_dev_alloc(ctx, P.h.n_records, "classifier flags");
                Free free_found{ctx, d_found};
                McClassifyArgs a{};
                a.base = P.base;
                a.recs = P.recs;
                a.n_records = P.h.n_records;
                a.k = graph->k;
                a.qoff = qoff;
                a.correction = correction ? 1 : 0;
                a.z = z;
                a.threshold = found_threshold;
                a.table = graph->dev;
                a.shard_id = graph->shard_id;
                a.found = d_found;
                a.header = P.d_header;
                const uint64_t warps = P.h.n_records;
                const int grid = (int)std::min<uint64_t>((warps * 32 + CL_THREADS - 1) / CL_THREADS, (uint64_t)ctx->sm_count * 16);
                k_classify<<<grid, CL_THREADS, 0, st>>>(a);
                MC_LAUNCH(ctx);
                MC_CUDA(cudaGetLastError());
                McParseHeader* hp = (McParseHeader*)ctx->h_pinned_small;
                MC_CUDA(cudaMemcpyAsync(hp, P.d_header, sizeof(McParseHeader), cudaMemcpyDeviceToHost, st));
                if (found && cap) MC_CUDA(cudaMemcpyAsync(found + total, d_found, P.h.n_records, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaStreamSynchronize(st));
                if (hp->err_packed != MC_NO_ERROR) mc_throw_parse_error(hp->err_packed, d_buf, (uint32_t)n, ctx);
                total += P.h.n_records;
            }
            if (final_chunk) break;
            const size_t consumed = P.h.consumed;
            const size_t left = n - consumed;
            if (consumed == 0 || left > chunk) mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk (%zu bytes); raise mc_ctx_set_chunk_bytes", chunk);
            // the unconsumed tail sits in host memory too: [pos + n_slice - left, pos + n_slice) unless it reaches back into the previous carry
            tail.resize(left);
            if (left <= n_slice) {
                memcpy(tail.data(), p + pos + n_slice - left, left);
            } else {
                MC_CUDA(cudaMemcpyAsync(tail.data(), d_buf + consumed, left, cudaMemcpyDeviceToHost, st));
                MC_CUDA(cudaStreamSynchronize(st));
            }
            carry = left;
            pos += n_slice;
        }
        *n_reads = total;
        return MC_OK;
    } catch (const McError& e) {
        mc_set_error(e.what());
        return e.code;
    }
}
