// K1: byte-level record scan of a FASTA/FASTQ chunk on the device.
#pragma once
#include "mc_common.cuh"

// one sequence record of a chunk; offsets are relative to `base` in McParsed
struct McRec {
    uint32_t seq_off, seq_len, qual_off, qual_len;
};

enum McParseError {
    MC_PE_OK = 0,
    MC_PE_FQ_HEADER = 1,      // "Unknown structure of fastq file! Waiting "@ID" or "+ID" string" (FastqReader.java:99-103)
    MC_PE_FQ_EOF = 2,         // "Unexpected end of file. File is corrupted/Format mismatch." (:63,:107)
    MC_PE_FQ_LEN = 3,         // "Bad DnaQ record: length of chars and quality is not the same." (:67-70)
    MC_PE_NUC = 4,            // "Incorrect nucleotide char" (DnaTools.java:61-62)
    MC_PE_IUPAC = 5,          // IUPAC ambiguity code: random in the reference (DnaTools.java:66-117), rejected
    MC_PE_QUAL = 6,           // "Invalid quality code char" (Illumina.java:7-12 / Sanger.java:7-12)
    MC_PE_TOO_LONG = 7        // a single record does not fit in one device chunk
};

#define MC_NO_ERROR 0xFFFFFFFFFFFFFFFFULL

// device-resident result header (mirrored to the host)
struct McParseHeader {
    unsigned long long err_packed;   // (byte position << 8) | McParseError of the FIRST error in byte order; MC_NO_ERROR if none
    uint32_t n_lines;
    uint32_t n_records;       // complete records in this chunk
    uint32_t consumed;        // bytes consumed; [consumed, n) must be carried into the next chunk
    uint32_t data_lines;      // FASTQ: number of data lines seen; FASTA: number of records started
    uint32_t final_state;     // FASTQ reader FSM state after the last line
    uint32_t total_seq_bytes; // FASTA: bytes copied into the compacted sequence buffer
    uint32_t max_seq_len;     // longest sequence of the chunk's records (picks the extractor's lanes-per-record)
    uint32_t first_mark;      // the chunk's first byte opens a header / comment line ('@' '+' / '>' ';')
    uint32_t pad_;
};

struct McParsed {
    const uint8_t* base;      // chunk itself (FASTQ) or the compacted sequence buffer (FASTA)
    const McRec* recs;        // device array, n_records entries
    McParseHeader* d_header;  // device copy (K2 records its errors there too)
    McParseHeader h;          // host copy after parsing
};

// Parses d_buf[0..n) (starts at a record boundary).  `final_chunk`: end of input, nothing may be left over.
void mc_parse_chunk(mc_ctx* ctx, const uint8_t* d_buf, uint32_t n, bool fasta, bool final_chunk, McParsed* out);
// turns a device-reported error into the reference's exception text and throws McError(MC_ERR_FORMAT)
[[noreturn]] void mc_throw_parse_error(unsigned long long err_packed, const uint8_t* d_buf, uint32_t n, mc_ctx* ctx);

// device-side "first error wins" (smallest byte position)
__device__ __forceinline__ void mc_report_error(McParseHeader* h, uint32_t code, uint32_t pos) {
    atomicMin(&h->err_packed, ((unsigned long long)pos << 8) | code);
}
