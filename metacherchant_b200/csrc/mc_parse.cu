// K1: record scan of a FASTA/FASTQ chunk, entirely on the device.
//
// Restates the *line protocol* of the reference's readers as data-parallel scans:
//   java.io.BufferedReader.readLine        line ends at "\n", "\r" or "\r\n"
//   FastqReader.readNextDataLine           (ITMO!/io/readers/FastqReader.java:84-110) two-state machine:
//       state 0: skip blank lines; a non-blank line must start with '@' or '+'  -> state 1
//       state 1: the next line, whatever it holds, is a data line                -> state 0
//     consecutive data lines pair up as (sequence, quality) (FastqReader.java:53-70)
//   FastaReader.readNextDataLine           (ITMO!/io/readers/FastaReader.java:78-104): a record is the concatenation
//     of the non-comment lines between two comment lines ('>' or ';'); empty groups vanish.
// The state machine is evaluated with a prefix scan over function composition (each line is a map
// {0,1}->{0,1} plus the number of data lines it emits from either start state); FASTA uses the same scan
// with identity maps, i.e. two plain prefix sums (sequence bytes, record starts).
#include <cstdio>
#include <vector>

#include "mc_parse.cuh"

#define TILE_WPT 4u               // 32-byte words per thread and tile
#define TILE_WORDS (256u * TILE_WPT)
#define TILE_BYTES (32u * TILE_WORDS)   // one CTA pass: 256 threads x 4 words of 32 bytes = 32 KiB (131 K tiles per GiB made the
                                  // single-CTA scan of the tile counts 3 ms per step)
#define LINES_PER_THREAD 4
#define LINES_PER_BLOCK (256 * LINES_PER_THREAD)

// --------------------------------------------------------------------------------------------
// line terminators
// --------------------------------------------------------------------------------------------
// The chunk is addressed through a 16-byte aligned pointer `abuf` plus the misalignment `mis` of its first byte,
// so chunks may start anywhere inside a device-resident buffer and still be scanned with 128-bit loads.
// `base` is an offset in aligned coordinates (multiple of 32); bit j of the result = logical byte base+j-mis is a line
// terminator.  Bytes before the chunk read as 0; the byte after the chunk reads as `pad` ('\n' for a non-final chunk,
// so a trailing '\r' stays part of the unfinished line and is carried over; 0 at end of input).
// 0x80 in every byte of w that equals the (replicated) byte pattern; exact per byte (no carries between bytes)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t pat) {
    const uint32_t x = w ^ pat;
    const uint32_t t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;
    return ~t & 0x80808080u;
}
// Terminator mask of the 32 bytes at `base`: a byte is a line terminator when it is '\n', or '\r' not followed by '\n'.
// The mask is STORED TRANSPOSED: byte 4i + b of the word (b-th byte of its i-th 32-bit piece) is bit 8b + i.  The SWAR compare
// leaves one flag per byte at bit 7 of that byte; shifting piece i's flags right by 7 - i and OR-ing the eight pieces builds the
// transposed mask with two instructions per piece, where packing every piece's four flags into adjacent bits cost six.  Pass 2
// turns a non-empty mask back into byte order (transposed_to_mask); popc is the same either way.
// What else the later passes need to know about a line -- does its terminator belong to a "\r\n" pair, does the NEXT line open
// with a header / comment mark -- is read by pass 2 from the two bytes around every terminator (one sector per line), and kept
// as a flag byte per line: passes 3 and 4 never touch the input (round 1 re-read the whole chunk twice for one byte per line).
// (Measured alternatives, profiles/r2_NOTES.md: three SWAR mask planes from pass 1, 8.0 ms; marks / CR-LF looked up per terminator
// inside pass 1, 9.2 ms -- the divergent look-up loop costs more than it saves; this form: pass 1 at its round-1 cost.)
__device__ __forceinline__ uint32_t mask_to_transposed(uint32_t m) {
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t x = (m >> b) & 0x11111111u;        // bytes 4i + b, i = 0..7, at bits 4i
        x = (x | (x >> 3)) & 0x03030303u;
        x = (x | (x >> 6)) & 0x000F000Fu;
        x = (x | (x >> 12)) & 0xFFu;
        r |= x << (8 * b);
    }
    return r;
}
__device__ __forceinline__ uint32_t transposed_to_mask(uint32_t t) {
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t x = (t >> (8 * b)) & 0xFFu;
        x = (x | (x << 12)) & 0x000F000Fu;
        x = (x | (x << 6)) & 0x03030303u;
        x = (x | (x << 3)) & 0x11111111u;
        r |= x << b;
    }
    return r;
}
__device__ __forceinline__ uint32_t term_mask32(const uint8_t* __restrict__ abuf, uint32_t mis, uint32_t n, uint32_t base, uint8_t pad) {
    const uint32_t lo = mis, hi = mis + n;   // valid aligned-coordinate range [lo, hi)
    if (base >= lo && base + 33 <= hi) {
        // interior: eight 32-bit words, four bytes per step (SWAR)
        const uint4* p = reinterpret_cast<const uint4*>(abuf + base);
        const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t nl[9];
#pragma unroll
        for (int i = 0; i < 8; i++) nl[i] = eq_bytes(w[i], 0x0A0A0A0Au);
        nl[8] = __ldg(abuf + base + 32) == '\n' ? 0x80u : 0u;
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t cr = eq_bytes(w[i], 0x0D0D0D0Du);
            const uint32_t nl_next = (nl[i] >> 8) | (nl[i + 1] << 24);
            m |= (nl[i] | (cr & ~nl_next)) >> (7 - i);   // flags at bits 7, 15, 23, 31 -> bits i, 8 + i, 16 + i, 24 + i
        }
        return m;
    }
    uint8_t b[34];
#pragma unroll
    for (int j = 0; j < 33; j++) {
        const uint32_t a = base + j;
        b[j] = (a >= lo && a < hi) ? __ldg(abuf + a) : (a == hi ? pad : (uint8_t)0);
    }
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint8_t c = b[j];
        m |= (uint32_t)((c == '\n') || (c == '\r' && b[j + 1] != '\n')) << j;
    }
    // mask off positions outside the chunk
    if (base < lo) m &= (lo - base >= 32) ? 0u : (0xFFFFFFFFu << (lo - base));
    if (base + 32 > hi) m &= (hi <= base) ? 0u : (0xFFFFFFFFu >> (base + 32 - hi));
    return mask_to_transposed(m);
}

// pass 1 of the line index: the terminator masks (kept: pass 2 reads them instead of scanning the bytes again) and the number of
// terminators per tile
__global__ void __launch_bounds__(256) k_count_terms(const uint8_t* __restrict__ abuf, uint32_t mis, uint32_t n, uint8_t pad,
                                                     uint32_t* __restrict__ tile_counts, uint32_t* __restrict__ masks) {
    uint32_t c = 0;
#pragma unroll
    for (uint32_t j = 0; j < TILE_WPT; j++) {
        const uint32_t wi = blockIdx.x * TILE_WORDS + j * 256u + threadIdx.x;
        const uint32_t base = wi * 32u;
        const uint32_t m = base < mis + n ? term_mask32(abuf, mis, n, base, pad) : 0u;
        masks[wi] = m;
        c += __popc(m);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s[i];
        tile_counts[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of tile_counts in place, total (+ virtual last line) into the header
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* tile_counts, uint32_t n_tiles, const uint8_t* buf, uint32_t n, McParseHeader* h,
                                                     int final_chunk, int fasta) {
    __shared__ uint32_t s[1024];
    uint32_t per = (n_tiles + 1023) / 1024;
    uint32_t lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += tile_counts[i];
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        uint32_t v = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; i++) {
        uint32_t c = tile_counts[i];
        tile_counts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) {
        uint32_t total = s[1023];
        // end of input: a last line without terminator still counts (BufferedReader.readLine); in a non-final chunk the
        // unterminated tail is simply carried over to the next chunk
        bool virt = final_chunk && n > 0 && buf[n - 1] != '\n' && buf[n - 1] != '\r';
        h->err_packed = MC_NO_ERROR;
        h->n_lines = total + (virt ? 1 : 0);
        h->n_records = 0;
        h->consumed = n;
        h->data_lines = 0;
        h->final_state = 0;
        h->total_seq_bytes = 0;
        h->max_seq_len = 0;
        const uint8_t c0 = n ? buf[0] : 0;
        h->first_mark = fasta ? (c0 == '>' || c0 == ';') : (c0 == '@' || c0 == '+');
        h->pad_ = 0;
    }
}

// pass 2: line_end[i] = position of line i's terminator, and one flag byte per line:
//   bit 0 = the terminator is the '\n' of "\r\n" (content ends one byte earlier),
//   bit 1 = the first byte of the NEXT line is a header / comment mark ('@' '+' for FASTQ, '>' ';' for FASTA; line 0's own mark is
//           McParseHeader.first_mark).
template <bool FASTA>
__global__ void __launch_bounds__(256) k_write_terms(const uint8_t* __restrict__ abuf, uint32_t mis, uint32_t n, const uint32_t* __restrict__ masks,
                                                     const uint32_t* __restrict__ tile_offsets, uint32_t* __restrict__ line_end, uint8_t* __restrict__ line_flags,
                                                     uint32_t n_lines, int final_chunk) {
    const uint8_t* buf = abuf + mis;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s[8];
    uint32_t tile_pos = tile_offsets[blockIdx.x];
    for (uint32_t j = 0; j < TILE_WPT; j++) {   // the tile's words in position order: row j = words j * 256 .. j * 256 + 255
        const uint32_t wi = blockIdx.x * TILE_WORDS + j * 256u + threadIdx.x;
        const uint32_t base = wi * 32u;
        uint32_t m = masks[wi];
        const uint32_t c = __popc(m);
        if (m) m = transposed_to_mask(m);   // back to byte order (one word in three has a terminator)
        uint32_t incl = c;   // block exclusive scan of c
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();     // the previous row's readers are done with s[]
        if (lane == 31) s[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, row_total = 0;
        for (int i = 0; i < 8; i++) {
            if (i < warp) woff += s[i];
            row_total += s[i];
        }
        uint32_t pos = tile_pos + woff + incl - c;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t p = base + b - mis;   // chunk coordinate of the terminator
            const bool crlf = p > 0 && __ldg(buf + p) == '\n' && __ldg(buf + p - 1) == '\r';
            bool mark = false;
            if (p + 1 < n) {
                const uint8_t nx = __ldg(buf + p + 1);
                mark = FASTA ? (nx == '>' || nx == ';') : (nx == '@' || nx == '+');
            }
            line_flags[pos] = (uint8_t)((crlf ? 1u : 0u) | (mark ? 2u : 0u));
            line_end[pos++] = p;
        }
        tile_pos += row_total;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && final_chunk && n > 0 && buf[n - 1] != '\n' && buf[n - 1] != '\r') {
        line_end[n_lines - 1] = n;      // the unterminated last line of the input (BufferedReader.readLine)
        line_flags[n_lines - 1] = 0;
    }
}

// content span of line i: [start, end), the '\r' of a "\r\n" pair excluded; `marked` = its first byte opens a header / comment
// line.  Everything comes from the line index: no input byte is read.
struct McLineIdx {
    const uint32_t* line_end;
    const uint8_t* line_flags;
    uint32_t first_mark;     // line 0 starts with a mark character
};
__device__ __forceinline__ void line_span(const McLineIdx& L, uint32_t i, uint32_t& start, uint32_t& end, bool& marked) {
    start = i == 0 ? 0 : L.line_end[i - 1] + 1;
    end = L.line_end[i] - (L.line_flags[i] & 1u);
    marked = (i == 0 ? L.first_mark != 0 : (L.line_flags[i - 1] & 2u) != 0) && end > start;
}

// --------------------------------------------------------------------------------------------
// scan over function composition
// --------------------------------------------------------------------------------------------
struct Comp {
    uint32_t c0, c1;   // items emitted when entered in state 0 / state 1
    uint32_t f;        // bit0 = next state from state 0, bit1 = next state from state 1
};
__device__ __forceinline__ Comp comp_identity() { return Comp{0, 0, 2}; }
__device__ __forceinline__ Comp compose(const Comp& a, const Comp& b) {   // a first, then b
    uint32_t a0 = a.f & 1, a1 = (a.f >> 1) & 1;
    Comp r;
    r.c0 = a.c0 + (a0 ? b.c1 : b.c0);
    r.c1 = a.c1 + (a1 ? b.c1 : b.c0);
    r.f = ((b.f >> a0) & 1) | (((b.f >> a1) & 1) << 1);
    return r;
}
__device__ __forceinline__ Comp shfl_up_comp(const Comp& v, int o) {
    Comp r;
    r.c0 = __shfl_up_sync(0xffffffffu, v.c0, o);
    r.c1 = __shfl_up_sync(0xffffffffu, v.c1, o);
    r.f = __shfl_up_sync(0xffffffffu, v.f, o);
    return r;
}
// ordered block scan (blockDim.x multiple of 32, <= 1024): exclusive prefix of `mine` and the block total
__device__ __forceinline__ void block_scan_comp(const Comp& mine, Comp& excl, Comp& total) {
    __shared__ Comp s_w[32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    Comp incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        Comp v = shfl_up_comp(incl, o);
        if (lane >= o) incl = compose(v, incl);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    Comp wpre = comp_identity();
    for (int i = 0; i < warp; i++) wpre = compose(wpre, s_w[i]);
    Comp prev = shfl_up_comp(incl, 1);
    excl = lane == 0 ? wpre : compose(wpre, prev);
    Comp t = comp_identity();
    for (int i = 0; i < nw; i++) t = compose(t, s_w[i]);
    total = t;
    __syncthreads();
}

template <bool FASTA>
__device__ __forceinline__ Comp line_comp(const McLineIdx& L, uint32_t i) {
    uint32_t s, e;
    bool marked;
    line_span(L, i, s, e, marked);
    if (FASTA) {
        if (marked) return comp_identity();   // comment line ('>' or ';')
        bool prev_comment = false;
        if (i > 0) {
            uint32_t ps, pe;
            line_span(L, i - 1, ps, pe, prev_comment);
        }
        return Comp{e - s, (i == 0 || prev_comment) ? 1u : 0u, 2};
    } else {
        bool blank = e == s;
        return Comp{0, 1, blank ? 0u : 1u};   // state 1 always consumes a data line; state 0: blank stays, else -> 1
    }
}

template <bool FASTA>
__global__ void __launch_bounds__(256) k_block_summaries(McLineIdx L, uint32_t n_lines, Comp* __restrict__ blk) {
    uint32_t first = blockIdx.x * LINES_PER_BLOCK + threadIdx.x * LINES_PER_THREAD;
    Comp mine = comp_identity();
    for (uint32_t j = 0; j < LINES_PER_THREAD; j++)
        if (first + j < n_lines) mine = compose(mine, line_comp<FASTA>(L, first + j));
    Comp excl, total;
    block_scan_comp(mine, excl, total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// single CTA: exclusive prefix of the block summaries; fills the header totals
template <bool FASTA>
__global__ void __launch_bounds__(1024) k_scan_blocks(const Comp* __restrict__ blk, Comp* __restrict__ blk_prefix, uint32_t n_blk,
                                                      McLineIdx L, uint32_t n, McParseHeader* h, int final_chunk) {
    const uint32_t* line_end = L.line_end;
    uint32_t per = (n_blk + 1023) / 1024;
    uint32_t lo = min(threadIdx.x * per, n_blk), hi = min(lo + per, n_blk);
    Comp mine = comp_identity();
    for (uint32_t i = lo; i < hi; i++) mine = compose(mine, blk[i]);
    Comp excl, total;
    block_scan_comp(mine, excl, total);
    Comp run = excl;
    for (uint32_t i = lo; i < hi; i++) {
        blk_prefix[i] = run;
        run = compose(run, blk[i]);
    }
    if (threadIdx.x == 0) {
        uint32_t n_lines = h->n_lines;
        // default: everything up to the end of the last complete line is consumed (k_emit lowers it when the
        // last record is incomplete)
        if (!final_chunk) h->consumed = n_lines ? line_end[n_lines - 1] + 1 : 0;
        if (FASTA) {
            uint32_t n_rec = total.c1;
            bool last_is_data = false;
            if (n_lines > 0) {
                uint32_t s, e;
                bool marked;
                line_span(L, n_lines - 1, s, e, marked);
                last_is_data = !marked;
            }
            h->data_lines = n_rec;
            h->total_seq_bytes = total.c0;
            // the last group may continue in the next chunk unless the input ends here
            h->n_records = (!final_chunk && last_is_data && n_rec > 0) ? n_rec - 1 : n_rec;
        } else {
            uint32_t D = total.c0, st = total.f & 1;   // started in state 0
            h->data_lines = D;
            h->final_state = st;
            h->n_records = D >> 1;
            if (final_chunk && ((D & 1) || st)) mc_report_error(h, MC_PE_FQ_EOF, n);
        }
    }
}

template <bool FASTA>
__global__ void __launch_bounds__(256) k_emit(McLineIdx L, uint32_t n_lines,
                                             const Comp* __restrict__ blk_prefix, McParseHeader* h, McRec* __restrict__ recs,
                                             uint32_t* __restrict__ line_dst, uint32_t* __restrict__ rec_off, int final_chunk) {
    uint32_t first = blockIdx.x * LINES_PER_BLOCK + threadIdx.x * LINES_PER_THREAD;
    Comp lc[LINES_PER_THREAD];
    Comp mine = comp_identity();
#pragma unroll
    for (uint32_t j = 0; j < LINES_PER_THREAD; j++) {
        lc[j] = (first + j < n_lines) ? line_comp<FASTA>(L, first + j) : comp_identity();
        mine = compose(mine, lc[j]);
    }
    Comp excl, total;
    block_scan_comp(mine, excl, total);
    Comp in = compose(blk_prefix[blockIdx.x], excl);
    const uint32_t D = h->data_lines, final_state = h->final_state, n_records = h->n_records;
    if (FASTA) {
        uint32_t dst = in.c0, rec = in.c1;
#pragma unroll
        for (uint32_t j = 0; j < LINES_PER_THREAD; j++) {
            uint32_t i = first + j;
            if (i >= n_lines) break;
            uint32_t s, e;
            bool comment;
            line_span(L, i, s, e, comment);
            if (comment) {
                line_dst[i] = 0xFFFFFFFFu;
                continue;
            }
            if (lc[j].c1) {   // first line of a record
                rec_off[rec] = dst;
                if (rec == n_records && n_records < D) h->consumed = s;   // the carried-over last group starts here
                rec++;
            }
            line_dst[i] = dst;
            dst += e - s;
        }
        if (first <= n_lines - 1 && n_lines - 1 < first + LINES_PER_THREAD) rec_off[D] = h->total_seq_bytes;
    } else {
        uint32_t st = in.f & 1, d = in.c0, longest = 0;
#pragma unroll
        for (uint32_t j = 0; j < LINES_PER_THREAD; j++) {
            uint32_t i = first + j;
            if (i >= n_lines) break;
            uint32_t s, e;
            bool marked;
            line_span(L, i, s, e, marked);
            if (st == 0) {
                if (e == s) continue;   // skipped blank line
                if (!marked) mc_report_error(h, MC_PE_FQ_HEADER, s);
                st = 1;
                // a header whose data line is not in this chunk: carry it over
                if (i == n_lines - 1 && !(D & 1)) h->consumed = s;
            } else {
                uint32_t r = d >> 1;
                if ((d & 1) == 0) {
                    recs[r].seq_off = s;
                    recs[r].seq_len = e - s;
                    longest = max(longest, e - s);
                    if ((D & 1) && d == D - 1) {   // sequence line whose quality line is not in this chunk: carry the record
                        uint32_t ps, pe;
                        bool pm;
                        line_span(L, i - 1, ps, pe, pm);
                        h->consumed = ps;
                    }
                } else {
                    recs[r].qual_off = s;
                    recs[r].qual_len = e - s;
                }
                d++;
                st = 0;
            }
        }
        longest = __reduce_max_sync(__activemask(), longest);
        if ((threadIdx.x & 31) == 0 && longest) atomicMax(&h->max_seq_len, longest);
    }
    (void)final_state;
    (void)final_chunk;
}

// FASTA: gather the sequence lines into one contiguous buffer (a warp per line)
__global__ void __launch_bounds__(256) k_fa_copy(const uint8_t* __restrict__ buf, McLineIdx L, uint32_t n_lines,
                                                const uint32_t* __restrict__ line_dst, uint8_t* __restrict__ seqbuf) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n_lines; i += n_warps) {
        uint32_t dst = line_dst[i];
        if (dst == 0xFFFFFFFFu) continue;
        uint32_t s, e;
        bool marked;
        line_span(L, i, s, e, marked);
        for (uint32_t p = s + lane; p < e; p += 32) seqbuf[dst + (p - s)] = buf[p];
    }
}
__global__ void __launch_bounds__(256) k_fa_recs(const uint32_t* __restrict__ rec_off, uint32_t n_records, McRec* __restrict__ recs, McParseHeader* h) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t l = 0;
    if (r < n_records) {
        uint32_t o = rec_off[r];
        l = rec_off[r + 1] - o;
        recs[r] = McRec{o, l, 0, l};
    }
    l = __reduce_max_sync(0xffffffffu, l);
    if ((threadIdx.x & 31) == 0 && l) atomicMax(&h->max_seq_len, l);
}

// --------------------------------------------------------------------------------------------
// host driver
// --------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void mc_parse_chunk(mc_ctx* ctx, const uint8_t* d_buf, uint32_t n, bool fasta, bool final_chunk, McParsed* out) {
    cudaStream_t st = ctx->stream;
    const uint32_t mis = (uint32_t)((uintptr_t)d_buf & 15);
    const uint8_t* abuf = d_buf - mis;
    const uint8_t pad = final_chunk ? 0 : '\n';
    const int fin = final_chunk ? 1 : 0;
    uint32_t n_tiles = (mis + n + TILE_BYTES - 1) / TILE_BYTES;
    if (n_tiles == 0) n_tiles = 1;
    // slot 1: header + tile counts + one terminator mask per 32 input bytes
    const uint32_t n_words = n_tiles * TILE_WORDS;
    const size_t off_masks = (256 + (size_t)n_tiles * 4 + 255) & ~(size_t)255;
    uint8_t* s1 = (uint8_t*)mc_scratch(ctx, 1, off_masks + (size_t)n_words * 4);
    McParseHeader* d_h = (McParseHeader*)s1;
    uint32_t* d_tiles = (uint32_t*)(s1 + 256);
    uint32_t* d_masks = (uint32_t*)(s1 + off_masks);
    k_count_terms<<<n_tiles, 256, 0, st>>>(abuf, mis, n, pad, d_tiles, d_masks);
    MC_LAUNCH(ctx);
    k_scan_tiles<<<1, 1024, 0, st>>>(d_tiles, n_tiles, d_buf, n, d_h, fin, fasta ? 1 : 0);
    MC_LAUNCH(ctx);
    MC_CUDA(cudaGetLastError());
    McParseHeader* hp = (McParseHeader*)ctx->h_pinned_small;
    MC_CUDA(cudaMemcpyAsync(hp, d_h, sizeof(McParseHeader), cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    uint32_t n_lines = hp->n_lines;

    uint32_t n_blk = (n_lines + LINES_PER_BLOCK - 1) / LINES_PER_BLOCK;
    if (n_blk == 0) n_blk = 1;
    size_t max_recs = fasta ? (size_t)n_lines + 1 : (size_t)n_lines / 2 + 1;
    // slot 0 layout
    size_t off_line_end = 0;
    size_t off_line_flags = off_line_end + ((size_t)n_lines + 1) * 4;
    size_t off_blk = align_up(off_line_flags + (size_t)n_lines + 1, 16);
    size_t off_blk_prefix = off_blk + (size_t)n_blk * sizeof(Comp);
    size_t off_recs = align_up(off_blk_prefix + (size_t)n_blk * sizeof(Comp), 16);
    size_t off_line_dst = off_recs + max_recs * sizeof(McRec);
    size_t off_rec_off = off_line_dst + (fasta ? ((size_t)n_lines + 1) * 4 : 0);
    size_t off_seqbuf = align_up(off_rec_off + (fasta ? (max_recs + 1) * 4 : 0), 256);
    size_t total = off_seqbuf + (fasta ? (size_t)n + 16 : 0);
    uint8_t* s0 = (uint8_t*)mc_scratch(ctx, 0, total);
    uint32_t* d_line_end = (uint32_t*)(s0 + off_line_end);
    uint8_t* d_line_flags = s0 + off_line_flags;
    const McLineIdx L{d_line_end, d_line_flags, hp->first_mark};
    Comp* d_blk = (Comp*)(s0 + off_blk);
    Comp* d_blk_prefix = (Comp*)(s0 + off_blk_prefix);
    McRec* d_recs = (McRec*)(s0 + off_recs);
    uint32_t* d_line_dst = (uint32_t*)(s0 + off_line_dst);
    uint32_t* d_rec_off = (uint32_t*)(s0 + off_rec_off);
    uint8_t* d_seqbuf = s0 + off_seqbuf;

    out->d_header = d_h;
    out->recs = d_recs;
    out->base = fasta ? d_seqbuf : d_buf;
    if (n_lines == 0) {
        out->h = *hp;
        if (!final_chunk) out->h.consumed = 0;   // no complete line yet: everything is carried over
        return;
    }
    if (fasta) k_write_terms<true><<<n_tiles, 256, 0, st>>>(abuf, mis, n, d_masks, d_tiles, d_line_end, d_line_flags, n_lines, fin);
    else k_write_terms<false><<<n_tiles, 256, 0, st>>>(abuf, mis, n, d_masks, d_tiles, d_line_end, d_line_flags, n_lines, fin);
    MC_LAUNCH(ctx);
    if (fasta) {
        k_block_summaries<true><<<n_blk, 256, 0, st>>>(L, n_lines, d_blk);
        k_scan_blocks<true><<<1, 1024, 0, st>>>(d_blk, d_blk_prefix, n_blk, L, n, d_h, final_chunk ? 1 : 0);
        k_emit<true><<<n_blk, 256, 0, st>>>(L, n_lines, d_blk_prefix, d_h, d_recs, d_line_dst, d_rec_off, final_chunk ? 1 : 0);
        int copy_blocks = (int)std::min<uint64_t>(((uint64_t)n_lines * 32 + 255) / 256, (uint64_t)ctx->sm_count * 16);
        k_fa_copy<<<copy_blocks, 256, 0, st>>>(d_buf, L, n_lines, d_line_dst, d_seqbuf);
        ctx->launches += 4;
    } else {
        k_block_summaries<false><<<n_blk, 256, 0, st>>>(L, n_lines, d_blk);
        k_scan_blocks<false><<<1, 1024, 0, st>>>(d_blk, d_blk_prefix, n_blk, L, n, d_h, final_chunk ? 1 : 0);
        k_emit<false><<<n_blk, 256, 0, st>>>(L, n_lines, d_blk_prefix, d_h, d_recs, nullptr, nullptr, final_chunk ? 1 : 0);
        ctx->launches += 3;
    }
    MC_CUDA(cudaGetLastError());
    MC_CUDA(cudaMemcpyAsync(hp, d_h, sizeof(McParseHeader), cudaMemcpyDeviceToHost, st));
    MC_CUDA(cudaStreamSynchronize(st));
    out->h = *hp;
    if (fasta && out->h.n_records > 0) {
        k_fa_recs<<<(out->h.n_records + 255) / 256, 256, 0, st>>>(d_rec_off, out->h.n_records, d_recs, d_h);
        MC_LAUNCH(ctx);
        MC_CUDA(cudaGetLastError());
        MC_CUDA(cudaMemcpyAsync(&hp->max_seq_len, &d_h->max_seq_len, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MC_CUDA(cudaStreamSynchronize(st));
        out->h.max_seq_len = hp->max_seq_len;
    }
}

void mc_throw_parse_error(unsigned long long packed, const uint8_t* d_buf, uint32_t n, mc_ctx* ctx) {
    uint32_t code = (uint32_t)(packed & 0xFF), pos = (uint32_t)(packed >> 8);
    char snippet[32] = {0};
    if (d_buf && pos < n) {
        uint32_t len = std::min<uint32_t>(23, n - pos);
        cudaMemcpyAsync(snippet, d_buf + pos, len, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        for (uint32_t i = 0; i < len; i++)
            if (snippet[i] == '\n' || snippet[i] == '\r') { snippet[i] = 0; break; }
    }
    switch (code) {
        case MC_PE_FQ_HEADER: {
            std::string s(snippet);
            if (s.size() > 20) s = s.substr(0, 20) + "...";
            mc_throw(MC_ERR_FORMAT, "Unknown structure of fastq file! Waiting \"@ID\" or \"+ID\" string, found \"%s\". Possibly file is corrupted/format mismatch. (byte %u)", s.c_str(), pos);
        }
        case MC_PE_FQ_EOF:
            mc_throw(MC_ERR_FORMAT, "Unexpected end of file. File is corrupted/Format mismatch.");
        case MC_PE_FQ_LEN:
            mc_throw(MC_ERR_FORMAT, "Bad DnaQ record: length of chars and quality is not the same. File is corrupted/Format mismatch. (byte %u)", pos);
        case MC_PE_NUC:
            mc_throw(MC_ERR_FORMAT, "Incorrect nucleotide char: \"%c\" (byte %u)", snippet[0], pos);
        case MC_PE_IUPAC:
            mc_throw(MC_ERR_FORMAT, "IUPAC ambiguity code \"%c\" (byte %u): the reference substitutes a random base (DnaTools.java:66-117); not reproducible, rejected", snippet[0], pos);
        case MC_PE_QUAL:
            mc_throw(MC_ERR_FORMAT, "Invalid quality code char: \"%c\" char code = %d (byte %u)", snippet[0], (int)(unsigned char)snippet[0], pos);
        case MC_PE_TOO_LONG:
            mc_throw(MC_ERR_FORMAT, "a single record is longer than the device chunk; raise mc_ctx_set_chunk_bytes");
        default:
            mc_throw(MC_ERR_FORMAT, "malformed input (code %u at byte %u)", code, pos);
    }
}
