// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of MetaCherchant's environment-finder hot path (reads -> k-mer
// count table -> BFS environment -> unitig graph -> graph.gfa / env.txt / seqs.fasta
// / tsvs) and of environment-finder-multi.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this; the product
// (metacherchant_b200/) never links, imports or executes it.
//
// Language note: the reference is Java and leans on java.util.HashMap<String,..>
// and String; the restatement is C++17 (g++), deliberately literal, one function
// per reference method, each citing the file:line it follows.
//   src/...   = /root/reference/src/...
//   ITMO!/... = inside /root/reference/lib/itmo-assembler-src.jar (ru/ifmo/genetics/...)
//
// PARITY PINNING: stage D (environment -> unitigs/GFA/TSV) and the Java HashMap
// iteration order are pinned against the reference's committed example outputs
// (Hi-C_pipline/example_work_dir/output/1/merged/*, fixtures derived by
// tests/golden/make_golden.py).  Stage C1 (BFS) is pinned on the same example at
// the level of SETS: from the example target's seeds, over a table holding the
// golden k-mers, the two passes visit exactly the 93 572 golden k-mers
// (tests/test_oracle_golden.py); the order of lines inside one HashMap bucket is
// not reproducible (the committed file comes from an older revision).  Stages A-B
// (parse, count) and the reads-classifier are PARITY UNPINNED by any reference
// fixture (the JVM, the compiled ITMO jar and the WGS reads are all absent); they
// are cross-checked against a second, independent pure-Python transcription
// (oracle/pyref.py).
#pragma once
#include <cstdint>
#include <cstddef>

extern "C" {

// ---- error text of the last failing call on this thread ---------------------------
const char* orc_last_error();

// ---- A0: format detection (ITMO!/io/ReadersUtils.java:27-54) ----------------------
// returns 0 = fasta, 1 = fastq, <0 = error ("Can't detect file format ...", or gz/bz2/binq unsupported)
int orc_detect_format(const char* path);

// ---- count table (B1) ---------------------------------------------------------------
typedef struct orc_table orc_table;
orc_table* orc_table_new(int k, int threads);   // threads = --available-processors (stripe count 2^(floor(log2 P)+4))
void orc_table_free(orc_table*);
// fmt: 0 fasta, 1 fastq (sniff quality offset on first 1000 records), 33 / 64 = fastq with forced offset
int orc_count_bytes(orc_table*, const uint8_t* buf, size_t len, int fmt);
int orc_count_file(orc_table*, const char* path);
// add (key,count) pairs with saturating add (Kmers2HMWorker, src/io/IOUtils.java:71-92)
int orc_table_add_pairs(orc_table*, const uint64_t* keys, const int16_t* counts, size_t n);
uint64_t orc_table_size(const orc_table*);
int16_t orc_table_get(const orc_table*, uint64_t canonical_key);        // -1 absent
void orc_table_get_many(const orc_table*, const uint64_t* keys, size_t n, int16_t* out);
// sorted by key; entries with count >= min_count; returns number written (<= cap); *total = matching entries
uint64_t orc_table_dump(const orc_table*, int min_count, uint64_t* keys, int16_t* counts, uint64_t cap, uint64_t* total);
// {entries, sum of counts, sum of digest(key) * count mod 2^64, largest count} over entries with count >= min_count
// (the digest of the library's mc_table_checksum)
void orc_table_checksum(const orc_table*, int min_count, uint64_t out[4]);
// reader statistics of everything counted so far
void orc_table_stats(const orc_table*, uint64_t* reads, uint64_t* runs, uint64_t* bases, uint64_t* kmers);

// ---- k-mer helpers (A2) -------------------------------------------------------------
uint64_t orc_revcomp(uint64_t kmer, int k);          // ITMO!/utils/KmerUtils.java:12-22
uint64_t orc_canonical(uint64_t kmer, int k);        // KmerUtils.java:58-60
uint64_t orc_encode(const char* s, int k);           // DnaTools.toLong, A0 G1 C2 T3, first base most significant
void orc_decode(uint64_t kmer, int k, char* out);    // KmerUtils.kmer2String (out gets k chars + NUL)
int32_t orc_java_string_hash(const char* s);         // java.lang.String.hashCode

// ---- C1: one BFS pass, literal (src/algo/OneSequenceCalculator.java:154-220) -------
// dir: -1 left, +1 right, 0 both.  max_kmers / max_radius: -1 = unset.
// out_kmers receives the oriented k-mers in DISCOVERY order (queue order with duplicate
// seeds removed), k chars each, no terminators; out_dist the BFS distance.  Returns the
// number of visited oriented k-mers, or -1 if there are no seeds ("fail"), -2 if cap too small.
int64_t orc_bfs(const orc_table*, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                int dir, char* out_kmers, int32_t* out_dist, uint64_t cap);

// ---- C0..C6: whole environment for every target of a FASTA (EnvironmentFinderMain.runImpl) ---
// Writes <out_dir>/<header>/{graph.txt,env.txt,seqs.fasta,graph.gfa,tsvs/nodes.tsv,tsvs/edges.tsv}.
// Returns number of targets processed (>=0) or <0 on error.
int orc_env_find_ex(const orc_table*, const char* seq_fasta, const char* hic_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                    int both_dirs, int chunk_len, int merge, int trim, const char* out_dir);
int orc_env_find(const orc_table*, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                 int both_dirs, int chunk_len, const char* out_dir);
// stage D only: from an env file (graph.txt lines "KMER COUNT", order kept) + target sequence string
int orc_env_from_graph_txt(const char* graph_txt, const char* target_seq, int k, int chunk_len, const char* out_dir);

// ---- M1: environment-finder-multi (src/tools/EnvironmentFinderMultiMain.java) ------
int orc_env_multi(const char* const* env_paths, int n, const char* seq_fasta, int gene_id, const char* out_dir);

// ---- f-1: kmer-counter dump (src/io/IOUtils.java:39-65), big-endian (int64 key,int16 count), count > threshold
int64_t orc_print_kmers(const orc_table*, int threshold, const char* out_file, const char* stat_file);
// loadKmers (src/io/IOUtils.java:94-126, src/io/KmersLoadWorker.java:9-32): add records with freq > threshold
int orc_load_kmers(orc_table*, const char* path, int freq_threshold);

// ---- CPU baseline leg (bench.py only): multi-threaded counting with the reference's structure:
// one serialized parser (ReadsDispatcher monitor) feeding P workers over lock-striped maps.
// Returns seconds of wall time, <0 on error.  `bases` gets the number of input bases.
double orc_bench_count(const uint8_t* buf, size_t len, int fmt, int k, int threads, uint64_t* bases, uint64_t* distinct);

// ---- f-4b: reads-classifier (src/algo/ReadsFinderInGraph.java:37-123, src/algo/PairFinder.java:30-55) -------------------------
// found[i] = 1 when read i of the FASTQ buffer is "found" in the graph; margins (optional) = distance of the decisive f64
// comparison from its boundary.  z = 1 (or 1.96 with --interval95), found_threshold = --found-threshold / 100.
long orc_classify_bytes(const orc_table* graph, const uint8_t* buf, size_t len, int fmt, double z, double found_threshold, int correction,
                        uint8_t* found, double* margins, size_t cap);
// the tool (src/tools/ReadsClassifier.java:141-224): reads2 may be NULL/"" (single-end); writes the six FASTQ files, input order
int orc_reads_classifier(const orc_table* graph, const char* reads1, const char* reads2, double z, double found_threshold, int correction,
                         const char* out_dir, long counts[4]);

// N of "Extending endings by N kmers" (src/algo/OneSequenceCalculator.java:265-295) per target; returns the number of targets
int orc_env_extend_counts(const orc_table* t, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                          long long* out, int cap);
void orc_java_format_6_2f(float v, char* out);   // String.format("%6.2f ", v); out: room for 48 characters

// CPU baseline of the environment stage (bench.py only): one OneSequenceCalculator per target on `threads` pool workers
// (src/tools/EnvironmentFinderMain.java:188,218-225).  Returns wall seconds, < 0 on error.
double orc_bench_env(const orc_table* t, const char* seq_fasta, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                     int chunk_len, int threads, const char* out_dir, int* done);

}  // extern "C"
