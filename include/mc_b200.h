/* libmcgpu — B200-native (sm_100a) drop-in for MetaCherchant's environment-finder hot path.
 *
 * C ABI: plain pointers and sizes, opaque handles, no torch/C++ types.  Every entry point names
 * the reference interface it replaces (file:line; src/ = /root/reference/src/, ITMO!/ = inside
 * lib/itmo-assembler-src.jar ru/ifmo/genetics/).  The Java side binds these through the JNI shim
 * in jni/mc_jni.cpp.
 *
 * Conventions
 *   - every function returns MC_OK (0) or a negative mc_status; mc_last_error() gives the message
 *     of the last failure on the calling thread (text follows the reference's exception messages).
 *   - k-mers are 2-bit packed, first base most significant, A=0 G=1 C=2 T=3
 *     (ITMO!/dna/DnaTools.java:31,123-129); table keys are min(fw, rc) (ITMO!/utils/KmerUtils.java:58-60).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with MC_ERR_CUDA.
 */
#ifndef MC_B200_H
#define MC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mc_status {
    MC_OK = 0,
    MC_ERR_CUDA = -1,        /* no device / CUDA runtime failure */
    MC_ERR_ARG = -2,         /* bad argument (k outside 1..31, null pointer, ...) */
    MC_ERR_IO = -3,          /* "Failed to read from file <path>" (src/io/IOUtils.java:292) */
    MC_ERR_FORMAT = -4,      /* malformed FASTA/FASTQ: the reference throws from its reader threads */
    MC_ERR_NOMEM = -5,       /* table does not fit in device memory */
    MC_ERR_UNSUPPORTED = -6, /* reference feature outside the GPU path (binq input, k > 31, ...) */
    MC_ERR_CAPACITY = -7     /* caller-provided output buffer too small */
} mc_status;

/* input formats: ITMO!/io/ReadersUtils.java:27-54 (by file suffix) and :63-77 (quality-offset sniffing) */
enum { MC_FMT_FASTA = 0, MC_FMT_FASTQ = 1 /* sniff offset on the first 1000 records */, MC_FMT_FASTQ33 = 33, MC_FMT_FASTQ64 = 64 };

typedef struct mc_ctx mc_ctx;       /* one per GPU: device, streams, staging buffers */
typedef struct mc_table mc_table;   /* the k-mer count map, resident in HBM */

const char* mc_last_error(void);
const char* mc_version(void);

/* ------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------- */
int mc_ctx_create(int device, mc_ctx** out);
void mc_ctx_destroy(mc_ctx* ctx);
/* chunk size (bytes of reads per device batch, default 256 MiB) — the analogue of READS_WORK_RANGE_SIZE (src/io/IOUtils.java:25) */
int mc_ctx_set_chunk_bytes(mc_ctx* ctx, size_t bytes);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
uint64_t mc_ctx_launches(const mc_ctx* ctx);
/* Device blocks freed by the library (tables, bins, temporaries) stay cached in the context and are reused by later
 * calls, so a steady sequence of count / filter / free calls makes no driver allocation after its first pass (the JVM
 * analogue: the map lives on the heap until tool exit, EnvironmentFinderMain.java:251-254).  mc_ctx_trim returns the
 * cache to the driver; mc_ctx_arena_stats reports the cached bytes and how many driver allocations were made so far. */
int mc_ctx_trim(mc_ctx* ctx);
int mc_ctx_arena_stats(mc_ctx* ctx, uint64_t* cached_bytes, uint64_t* driver_allocs);

/* ------------------------------------------------------------------------------------------------
 * Seam 1 — counting.  Replaces
 *   BigLong2ShortHashMap IOUtils.loadReads(File[] files, int k, int minSeqLen=0, int P, Logger)   src/io/IOUtils.java:217-248
 *   (readers ITMO!/io/readers/{FastqReader,FastaReader,FastaReaderFromXQSourceTrunc}.java, k-mers
 *    ITMO!/dna/kmers/ShortKmer.java:104-149, map ITMO!/structures/map/BigLong2ShortHashMap.java:40-77).
 * ---------------------------------------------------------------------------------------------- */
/* new BigLong2ShortHashMap(...) (src/io/IOUtils.java:220-221).  capacity_hint = expected distinct k-mers (0 = grow on demand).
 * shard_bits/shard_id: this table owns the keys whose hash has its top shard_bits bits == shard_id (multi-GPU, 0/0 = everything). */
int mc_table_create(mc_ctx* ctx, int k, uint64_t capacity_hint, int shard_bits, int shard_id, mc_table** out);
void mc_table_free(mc_table* t);
/* loadReads over files: format by suffix, FASTQ offset sniffed per file; counts into a fresh table */
int mc_count_files(mc_ctx* ctx, const char* const* paths, int n_paths, int k, mc_table** out);
/* one file into an existing table (the per-file body of IOUtils.run, src/io/IOUtils.java:284-314) */
int mc_count_file(mc_ctx* ctx, const char* path, mc_table* t);
/* streaming form: `buf` is HOST memory holding whole records (a file, or a record-aligned slice of one) */
int mc_count_bytes(mc_ctx* ctx, const void* buf, size_t len, int fmt, mc_table* t);
/* same, `dbuf` already resident in device memory (bench `value` leg; also the multi-GPU pipeline) */
int mc_count_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, mc_table* t);
/* BigLong2ShortHashMap.size() — "Hashtable size: N kmers" (src/tools/EnvironmentFinderMain.java:137) */
uint64_t mc_table_size(const mc_table* t);
/* reader statistics: sequences seen / runs after N- and quality-splitting / bases in runs / k-mer instances
 * ("N reads added", good/total statistics, src/io/IOUtils.java:239-245,313) */
int mc_table_stats(const mc_table* t, uint64_t* reads, uint64_t* runs, uint64_t* bases, uint64_t* kmers);
/* addAndBound(key, freq) for (key,count) pairs — Kmers2HMWorker (src/io/IOUtils.java:71-92) */
int mc_table_add_pairs(mc_table* t, const uint64_t* keys, const int16_t* counts, size_t n);

/* ------------------------------------------------------------------------------------------------
 * Seam 2 — lookups.  Replaces short BigLong2ShortHashMap.get(long) (-1 = absent,
 * ITMO!/structures/map/Long2ShortHashMap.java:159-175) and entryIterator() (src/io/IOUtils.java:47).
 * ---------------------------------------------------------------------------------------------- */
int mc_table_get(const mc_table* t, const uint64_t* keys, size_t n, int16_t* out);
/* entries with count >= min_count, sorted by key.  Pass cap = 0 to query *n only. */
int mc_table_dump(const mc_table* t, int min_count, uint64_t* keys, int16_t* counts, size_t cap, size_t* n);
/* The same two seams with DEVICE buffers (multi-GPU: every rank all-gathers the solid (key, count16) set of the others,
 * SURVEY.md 8e, so that BFS + unitigs run on a complete filtered table; the pairs never touch the host).  dump_device
 * writes the entries in table order (unsorted); cap = 0 queries *n only.  add_pairs_device skips keys another shard owns. */
int mc_table_dump_device(const mc_table* t, int min_count, uint64_t* d_keys, int16_t* d_counts, size_t cap, size_t* n);
int mc_table_add_pairs_device(mc_table* t, const uint64_t* d_keys, const int16_t* d_counts, size_t n);
/* coverage-threshold compaction (B2/K4): a new, densely sized table holding only count >= min_count */
int mc_table_filter(mc_ctx* ctx, const mc_table* t, int min_count, mc_table** solid);

/* ------------------------------------------------------------------------------------------------
 * Seam 3 — environments.  Replaces, batched over all targets,
 *   new OneSequenceCalculator(seq, k, minOcc, outPrefix, workPrefix, hasher=null, reads, logger,
 *                             bothDirs, chunkLength, termMode, trim=false).run()
 *   (src/tools/EnvironmentFinderMain.java:218-225, src/algo/OneSequenceCalculator.java:98-114).
 * max_kmers / max_radius: -1 = unset (at least one must be set, EnvironmentFinderMain.java:171-175).
 * Writes <out_dir>/<names[i]>/{graph.txt, env.txt, seqs.fasta, graph.gfa, tsvs/nodes.tsv, tsvs/edges.tsv}.
 * found[i] (optional) = 1 if target i had at least one seed k-mer, else 0 (the reference logs
 * "Could not find any k-mers of the target gene in the input, halting." and writes nothing).
 * ---------------------------------------------------------------------------------------------- */
int mc_env_find(mc_ctx* ctx, const mc_table* t, const char* const* seqs, const char* const* names, int n_targets, int k,
                int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs, int chunk_len, const char* out_dir,
                int* found);
/* The same with the tool's remaining switches:
 *   merge (--merge, EnvironmentFinderMain.java:226-233): ONE environment seeded by the k-mers of every target and then of every
 *     Hi-C sequence (--hicseq, OneSequenceCalculator.java:168-192), written to <out_dir>/merged/; gene nodes are decided against
 *     the targets only (:421-432); found[0] tells whether any seed was solid.  names may be NULL.
 *   trim (--trim, :208-216,241-262): after each pass keep only the k-mers from which, walking against the pass direction
 *     inside the visited set, a "last" k-mer (one with a solid neighbour that allowsAddition refused) is reached. */
int mc_env_find_ex(mc_ctx* ctx, const mc_table* t, const char* const* seqs, const char* const* names, int n_targets,
                   const char* const* hic_seqs, int n_hic, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int both_dirs,
                   int chunk_len, int merge, int trim, const char* out_dir, int* found);
/* extendEnvironment (OneSequenceCalculator.java:265-295) does not change an environment (`cont = kmer`, :276); what remains of it is
 * its log line "Extending endings by N kmers" (:291), N = environment k-mers with exactly one solid neighbour outside the environment.
 * out[i] = N of target i of the LAST mc_env_find / mc_env_find_ex call on this context (-1: no environment for it); *n = targets. */
int mc_env_extend_counts(mc_ctx* ctx, long long* out, int cap, int* n);
/* one BFS pass (runBfs, OneSequenceCalculator.java:154-220): oriented k-mers in discovery order.
 * dir: -1 left, +1 right, 0 both.  Returns in *n the number visited (0 = no seeds). */
int mc_bfs(mc_ctx* ctx, const mc_table* t, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius, int dir,
           uint64_t* kmers, int32_t* dist, int16_t* counts, size_t cap, size_t* n);
/* the environment of one target as (ASCII-canonical k-mer packed 2-bit, count), in graph.txt order */
int mc_env_kmers(mc_ctx* ctx, const mc_table* t, const char* seq, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                 int both_dirs, uint64_t* kmers, int32_t* counts, size_t cap, size_t* n);
/* stage D alone (createPicture, OneSequenceCalculator.java:326-339) from an environment file with "KMER COUNT" lines
 * (DeBruijnGraphUtils.loadGraph): writes the same six files under out_dir.  target_seq may be "" (no gene node). */
int mc_env_from_graph_txt(mc_ctx* ctx, const char* graph_txt, const char* target_seq, int k, int chunk_len, const char* out_dir);
/* FASTA target loader (src/io/RichFastaReader.java:38-76 + DnaQ(str,0).toString()): returns a handle to iterate */
typedef struct mc_targets mc_targets;
int mc_targets_load(const char* fasta_path, mc_targets** out);
int mc_targets_count(const mc_targets* t);
const char* mc_targets_seq(const mc_targets* t, int i);
const char* mc_targets_name(const mc_targets* t, int i);   /* header minus first char = output sub-directory */
void mc_targets_free(mc_targets* t);

/* ------------------------------------------------------------------------------------------------
 * Seam 4 — environment-finder-multi.  Replaces
 *   new MultiSequenceCalculator(seq, k, outPrefix, logger, graphs).run() + printGene + printProbability
 *   (src/tools/EnvironmentFinderMultiMain.java:81-168).
 * ---------------------------------------------------------------------------------------------- */
int mc_env_multi(mc_ctx* ctx, const char* const* env_paths, int n_envs, const char* seq_fasta, int gene_id, const char* out_dir);

/* ------------------------------------------------------------------------------------------------
 * "next" row f-1 — kmer-counter's on-disk table (src/io/IOUtils.java:39-65,94-126): big-endian
 * (int64 key, int16 count) records with count > threshold.
 * ---------------------------------------------------------------------------------------------- */
int mc_table_print_kmers(const mc_table* t, int threshold, const char* out_file, const char* stat_file, uint64_t* good);
int mc_table_load_kmers(mc_table* t, const char* path, int freq_threshold);

/* ------------------------------------------------------------------------------------------------
 * "next" row f-4b -- reads-classifier.  Replaces, for every read of a FASTQ buffer (host memory, whole records),
 *   ReadsFinderInGraph.findRead(dnaQ) / findReadWithCorrection(dnaQ)            src/algo/ReadsFinderInGraph.java:37-49,89-123
 * over `graph` = the k-mer count table the tool loads (ReadsClassifier.loadGraph, src/tools/ReadsClassifier.java:86-105):
 * found[i] = 1 when read i's coverage breadth is at least found_threshold and inside the z-sigma interval around
 * 1 - exp(-depth).  z = 1, or 1.96 with --interval95; found_threshold = --found-threshold / 100; correction = --correction.
 * Pass cap = 0 (found may be NULL) to count the reads only.  The pair logic and the six output files of the tool
 * (PairFinder.run, ReadsClassifier.runImpl :141-224) are host code in the CLI / Python layer.
 * ---------------------------------------------------------------------------------------------- */
int mc_classify_bytes(mc_ctx* ctx, const mc_table* graph, const void* buf, size_t len, int fmt, double z, double found_threshold,
                      int correction, uint8_t* found, size_t cap, size_t* n_reads);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU building blocks: one process per GPU; keys are exchanged by the host
 * layer (torch.distributed all_to_all over NCCL/NVLink) between these two calls.
 * ---------------------------------------------------------------------------------------------- */
/* parse + extract the canonical k-mers of a device-resident chunk, hashed and partitioned by owner
 * (top shard_bits bits of the hash): d_keys receives the hashed keys grouped by owner, counts[o] the
 * number for owner o.  d_keys must hold at least mc_max_kmers(len, fmt, k) entries. */
int mc_extract_partition_device(mc_ctx* ctx, const void* dbuf, size_t len, int fmt, int k, int shard_bits,
                                uint64_t* d_keys, uint64_t* counts /* host, 1<<shard_bits entries */, mc_table* stats_sink);
uint64_t mc_max_kmers(size_t len, int fmt, int k);
/* insert hashed keys (as produced above) that this table owns; device pointer */
int mc_table_insert_hashed_device(mc_table* t, const uint64_t* d_hkeys, size_t n);

/* Peer-memory exchange (one process per GPU of ONE node): the owner groups of a batch are written by the scatter kernel itself
 * into the owners' receive regions over NVLink (CUDA IPC mappings) -- no collective moves keys.  Protocol per batch b, slot = b & 1:
 *   every rank: mc_xchg_scatter_device(slot);   all ranks: barrier (host layer);   every rank: mc_xchg_insert(slot).
 * Setup: mc_xchg_create, exchange the 64-byte handles of mc_xchg_handle between the ranks, mc_xchg_open for every peer.
 * region_keys must be >= mc_max_kmers of the largest batch. */
typedef struct mc_xchg mc_xchg;
int mc_xchg_create(mc_ctx* ctx, int world, int rank, uint64_t region_keys, mc_xchg** out);
int mc_xchg_handle(mc_xchg* x, void* out64);
int mc_xchg_open(mc_xchg* x, int peer_rank, const void* handle64);
int mc_xchg_scatter_device(mc_xchg* x, const void* dbuf, size_t len, int fmt, int k, int slot, mc_table* stats_sink);
int mc_xchg_insert(mc_xchg* x, int slot, mc_table* t, uint64_t* n_received);
/* "Count locally, then merge": when a rank's own k-mers fit its GPU, every rank first counts ITS reads into a complete local
 * table (the single-GPU pipeline, repeats summed on the spot), and only the DISTINCT (key, count) pairs cross NVLink -- a
 * metagenome's k-mer instances are mostly repeats, so this moves several times fewer bytes than shipping every instance.
 *   every rank: mc_xchg_dump_owners(local, slot)   the table scan itself stores every entry into its owner's region (peer memory:
 *                                                  8-byte hashed keys, then 16-bit counts at the same positions = 10 bytes per
 *                                                  pair); region_keys of mc_xchg_create must be >= 1.25 * pairs + 64
 *   all ranks:  barrier (host layer)
 *   every rank: mc_xchg_merge(slot, shard)         adds the received pairs (saturating) into this rank's shard table
 * mc_table_copy_stats carries the reader statistics (reads, runs, bases, k-mers) of the local table over to the shard. */
int mc_xchg_dump_owners(mc_xchg* x, const mc_table* local, int slot, uint64_t* n_pairs);
/* The dump fused into the local table's LAST flush: `local` still holds pending binned updates (filled with the "defer_flush" tuning
 * on); they are applied group by group in hash order, and every hash range of the table is shipped to its owners on a second stream
 * as soon as the flush has finished it -- the NVLink transfer overlaps the aggregation of the following groups.  *fused = 0 when the
 * flush could not be followed range by range (spill lists, a small batch) and the whole table was dumped afterwards instead; -1 when a
 * receive region overflowed (nothing usable was shipped: create larger regions on every rank and call mc_xchg_dump_owners). */
int mc_xchg_flush_and_dump(mc_xchg* x, mc_table* local, int slot, uint64_t* n_pairs, int* fused);
int mc_xchg_merge(mc_xchg* x, int slot, mc_table* shard, uint64_t* n_pairs);
int mc_table_copy_stats(mc_table* dst, const mc_table* src);
void mc_xchg_destroy(mc_xchg* x);

/* ------------------------------------------------------------------------------------------------
 * Bench/test utility (NOT reference API): synthetic metagenome reads generated on the device
 * (include/mc_synth.h).  `params` is a mc_synth_params blob built by mc_synth_init.
 * ---------------------------------------------------------------------------------------------- */
int mc_synth_generate_device(mc_ctx* ctx, const void* params, size_t params_size, uint64_t first, uint64_t count, void* d_out);
/* plain device-memory helpers for hosts without their own allocator (the JNI shim, the CLI) */
int mc_device_alloc(mc_ctx* ctx, size_t bytes, void** d_ptr);
int mc_device_free(mc_ctx* ctx, void* d_ptr);
int mc_memcpy_d2h(mc_ctx* ctx, void* dst, const void* d_src, size_t bytes);
int mc_memcpy_h2d(mc_ctx* ctx, void* d_dst, const void* src, size_t bytes);
int mc_ctx_sync(mc_ctx* ctx);
/* CUDA-event timing on the context's stream (bench.py times kernels on the stream they are launched on) */
int mc_event_record(mc_ctx* ctx, int slot /*0..15*/);
int mc_event_elapsed_ms(mc_ctx* ctx, int slot_a, int slot_b, float* ms);
/* per-stage device times since the start of the last mc_count_* call, milliseconds:
 * [0]=parse [1]=extract [2]=grow/rehash [3]=bin scatter [4]=bin count [5]=coverage filter [6..7] reserved */
int mc_ctx_stage_ms(mc_ctx* ctx, float* ms3);          /* first three only (kept for older callers) */
int mc_ctx_stage_ms_all(mc_ctx* ctx, float* ms8);
/* host wall time inside the same phases since the last reset (launches and synchronisations included; [2] = reserving the bins,
 * [4] = the whole flush, [6] = resolving timers + final flag check): tells kernel time from host time when a step is slow */
int mc_ctx_host_ms(mc_ctx* ctx, double* ms8);
int mc_ctx_stage_reset(mc_ctx* ctx);                  /* mc_count_* reset the timers themselves; the multi-GPU building blocks do not */
/* tuning knobs of the binned update path (NOT reference API; tests and profiles use them):
 *   "part_bits" (-1 = automatic, else P = 2^value bins), "batch_keys" (capacity of one batch, 0 = automatic),
 *   "segment_keys" (keys applied per count launch, 0 = automatic), "count_tile", "count_variant" (k_bin_count shape),
 *   "extract_parts" (lanes per record in k_extract: 0 = by read length, 1, 4 or 32),
 *   "agg" (two-level bins + shared-memory aggregation in the flush: -1 = by batch size, 0 = off, 1 = on), "agg_unit" (keys per
 *   aggregation unit, 0 = automatic), "device_chunk_bytes" (mc_count_device: bytes per parse / extract pass over device-resident input, 0 = as large as the 32-bit record
 *   offsets allow; tests force small values), "defer_flush" (1: mc_count_* leave their last batch of bins pending for mc_xchg_flush_and_dump),
 *   "extract_tma" (1 = default: the extractor stages the byte span of a warp's records with one cp.async.bulk + mbarrier; 0: per-lane
 *   16-byte loads), "flush_overlap" (1: in the aggregated flush the sub-bin scatter of group g + 1 runs on a second stream beside the
 *   aggregation of group g; 0 = default: one after the other), "stage_timing" (0: no per-stage CUDA events), "scatter_tma" (1 = default: the scatter kernels hand every bin run to the copy engine with cp.async.bulk; 0: per-sector stores),
 *   "keep_stage_ms" (1: mc_count_* accumulate into the stage timers instead of resetting them --
 *   a step made of several count calls; reset with mc_ctx_stage_reset),
 *   "superkmers" (0 = default; 1: for k in [22,27] and [30,31] the extractor emits 16-byte super-k-mer records into minimizer bins that
 *   are de-duplicated and unpacked at flush time, mc_sk.cu: same table, measured slower at cfg2), "sk_bits" / "sk_cap" (log2 of the
 *   number of minimizer bins / records per bin; tests force spills with tiny bins), "agg_lane_loop" (0 = default; 1: the aggregation
 *   walks a lane's keys of a round in one probe loop: measured slower) */
int mc_ctx_set_tuning(mc_ctx* ctx, const char* name, long long value);
/* table updates are buffered in bins; every read of the table applies them first, this call does so explicitly */
int mc_table_flush_pending(mc_table* t);
/* "About `n` more k-mer instances will be added to `t`" (0 = unknown).  Without a capacity hint the table is sized from a
 * counted 1/64 sample of every batch of binned updates (mc_bins.cu); with this the estimate is scaled to the rest of the
 * input, so a stream of batches does not grow the table by repeated doubling.  mc_count_bytes sets it itself. */
int mc_table_expect_keys(mc_table* t, uint64_t n);
/* Order-independent digest of the table's entries with count >= min_count, computed on the device (tests at sizes where a
 * dump through the host is impractical): out[0] = entries, out[1] = sum of counts (each clamped to 32767, like get()),
 * out[2] = sum of digest(key) * count mod 2^64 with digest(x) = (x ^ (x >> 29)) * 0xBF58476D1CE4E5B9 + 1, out[3] = largest
 * count.  The sum of shard digests of a sharded table equals the digest of the whole table. */
int mc_table_checksum(const mc_table* t, int min_count, uint64_t out[4]);
/* Host only: the reader set-up of ReadersUtils -- format by file suffix (ITMO!/io/ReadersUtils.java:27-54; returns MC_FMT_FASTA or
 * MC_FMT_FASTQ, *gz = 1 for a ".gz" name and 2 for a ".bz2" name, or a negative mc_status with the reference's message) and the FASTQ quality offset
 * sniffed on the first 1000 records (:63-77; returns 33 or 64). */
int mc_detect_file_format(const char* path, int* gz);
int mc_sniff_quality_offset(const void* buf, size_t len);
/* Host only (no device needed; test surface of the environment stage's host logic): iteration order of a
 * java.util.HashMap<String,Integer> whose keys are the k-mers `codes` (2-bit, A0 G1 C2 T3, first base most significant), put in
 * this order -- the order of graph.txt / env.txt lines and of node ids (src/algo/OneSequenceCalculator.java:297-310,387-405).
 * `peak` = the largest size the map ever had (a java HashMap never shrinks; 0 = its final size).  order[i] = index into `codes`
 * of the i-th entry; *n_out = number of distinct keys. */
int mc_kmer_hashmap_order(const uint64_t* codes, size_t n, int k, size_t peak, uint32_t* order, size_t* n_out);
/* Host only (test surface): the text Java's String.format("%6.2f ", floatValue) produces -- the cells of Jacard_sym.txt /
 * Jacard_alt.txt (src/tools/EnvironmentFinderMultiMain.java:141-142,163-164): shortest round-trip digits of the widened double,
 * rounded HALF_UP.  `out` must have room for 48 characters. */
int mc_java_format_6_2f(float v, char* out);

#ifdef __cplusplus
}
#endif
#endif /* MC_B200_H */
