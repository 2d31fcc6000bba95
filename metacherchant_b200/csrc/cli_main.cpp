// `metacherchant` — command-line front end with the reference's tool flags over the C ABI of libmcgpu.
//
// Mirrors, for the hot path only:
//   Runner / ITMO Runner.run                     tool selection by --tool/-t, default environment-finder   src/Runner.java:8-22
//   ITMO Tool launch options                      -w/--work-dir, -p, -m, -c, --force, -v, -h               ITMO!/utils/tool/Tool.java:61-143
//   EnvironmentFinderMain parameters              src/tools/EnvironmentFinderMain.java:32-105
//   EnvironmentFinderMultiMain parameters         src/tools/EnvironmentFinderMultiMain.java:21-42
//   KmersCounter parameters ("next" row f-1)      src/tools/KmersCounter.java
//   ReadsClassifier parameters ("next" row f-4b)  src/tools/ReadsClassifier.java:40-84
// Option syntax follows commons-cli PosixParser as the ITMO Tool uses it: "--name value", "--name=value", short
// options, multi-valued options taking every following non-option token and re-tokenised on "[, ]"
// (Tool.java:853-855,888-895), booleans with an optional True/False argument (Parameter.java:51-58).
// Reference flags outside the GPU path (--forcehash, --hash, k > 31) fail loudly.
#include <sys/stat.h>

#include <algorithm>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/mc_b200.h"

namespace {

FILE* g_log = nullptr;
bool g_verbose = false;

void logline(const char* level, const char* fmt, va_list ap) {
    char msg[4096];
    vsnprintf(msg, sizeof msg, fmt, ap);
    time_t now = time(nullptr);
    struct tm tmv;
    localtime_r(&now, &tmv);
    char ts[32];
    strftime(ts, sizeof ts, "%H:%M:%S", &tmv);
    fprintf(stderr, "%s %s: %s\n", ts, level, msg);
    if (g_log) {
        fprintf(g_log, "%s %s: %s\n", ts, level, msg);
        fflush(g_log);
    }
}
void info(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    logline("INFO", fmt, ap);
    va_end(ap);
}
void debug(const char* fmt, ...) {
    if (!g_verbose) return;
    va_list ap;
    va_start(ap, fmt);
    logline("DEBUG", fmt, ap);
    va_end(ap);
}
[[noreturn]] void die(const char* fmt, ...) {   // ExecutionFailedException -> log + System.exit(1) (Tool.java:448-463)
    va_list ap;
    va_start(ap, fmt);
    logline("ERROR", fmt, ap);
    va_end(ap);
    if (g_log) fclose(g_log);
    exit(1);
}

void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i < path.size(); i++) {
        cur.push_back(path[i]);
        if (path[i] == '/' || i + 1 == path.size()) mkdir(cur.c_str(), 0777);
    }
}

enum Kind { INT, STR, BOOL, FILES };
struct Opt {
    const char* name;
    const char* short_name;
    Kind kind;
    bool mandatory;
    const char* help;
};

struct Args {
    std::map<std::string, std::vector<std::string>> values;
    bool has(const std::string& n) const { return values.count(n) != 0; }
    std::string str(const std::string& n, const std::string& dflt = "") const {
        auto it = values.find(n);
        return it == values.end() || it->second.empty() ? dflt : it->second[0];
    }
    bool boolean(const std::string& n, bool dflt) const {
        auto it = values.find(n);
        if (it == values.end()) return dflt;
        if (it->second.empty()) return true;
        std::string v = it->second[0];
        for (auto& c : v) c = (char)tolower((unsigned char)c);
        return v == "true";   // new Boolean(String): anything but "true" (any case) is false
    }
    bool integer(const std::string& n, long long& out) const {
        auto it = values.find(n);
        if (it == values.end() || it->second.empty()) return false;
        char* end = nullptr;
        out = strtoll(it->second[0].c_str(), &end, 10);
        if (*end || it->second[0].empty()) die("Can't convert value '%s' of parameter '%s' to type 'class java.lang.Integer'", it->second[0].c_str(), n.c_str());
        return true;
    }
};

const Opt LAUNCH_OPTS[] = {
    {"tool", "t", STR, false, "set certain tool to run"},
    {"memory", "m", STR, false, "memory to use (ignored: the table lives in HBM)"},
    {"available-processors", "p", INT, false, "available processors (host threads)"},
    {"work-dir", "w", STR, false, "working directory (default workDir)"},
    {"continue", "c", BOOL, false, "continue the previous run (accepted; the tool has a single stage)"},
    {"force", nullptr, BOOL, false, "force run with rewriting old results"},
    {"start", "s", STR, false, "first force run stage"},
    {"finish", "f", STR, false, "stop after running this stage"},
    {"verbose", "v", BOOL, false, "enable debug output"},
    {"help", "h", BOOL, false, "print short help message"},
    {"device", nullptr, INT, false, "[GPU path] CUDA device index (default 0)"},
};
const Opt ENV_FINDER_OPTS[] = {
    {"k", "k", INT, true, "k-mer size"},
    {"reads", "i", FILES, false, "FASTQ, FASTA reads"},
    {"seq", nullptr, STR, true, "FASTA file with sequences"},
    {"hicseq", nullptr, STR, false, "FASTA file with Hi-C sequences (seeds of the merged environment)"},
    {"output", "o", STR, true, "output directory"},
    {"maxkmers", nullptr, INT, false, "maximum number of k-mers in created subgraph"},
    {"maxradius", nullptr, INT, false, "maximum distance in k-mers from starting gene"},
    {"coverage", nullptr, INT, false, "minimum depth of k-mers to consider (default 1)"},
    {"bothdirs", nullptr, BOOL, false, "run graph search in both directions from starting sequence"},
    {"chunklength", nullptr, INT, false, "minimum node length for BLAST search (default 1)"},
    {"forcehash", nullptr, BOOL, false, "force k-mer hashing [not supported by the GPU path]"},
    {"hash", nullptr, STR, false, "hash function to use [not supported by the GPU path]"},
    {"trim", nullptr, BOOL, false, "trim all not maximal paths?"},
    {"merge", nullptr, BOOL, false, "single environment for multiple input sequences?"},
};
const Opt MULTI_OPTS[] = {
    {"env", "e", FILES, true, "environment files to build difference for"},
    {"seq", nullptr, STR, true, ".fasta file with nucleotide sequence[s]"},
    {"output", "o", STR, true, "output directory to write results to"},
    {"geneid", "g", INT, false, "gene id from .fasta file (default 1)"},
};
const Opt READS_CLASSIFIER_OPTS[] = {   // src/tools/ReadsClassifier.java:40-84
    {"k", "k", INT, true, "k-mer size"},
    {"input-files", "i", FILES, true, "file with paired input reads for De Bruijn graph OR file with k-mers in binary format"},
    {"read-files", "r", FILES, true, "files with paired reads to classify"},
    {"output-dir", "o", STR, false, "directory to output found reads (default <work-dir>/reads_classifier)"},
    {"correction", "corr", BOOL, false, "Do replacement of nucleotide in read with one low quality position"},
    {"hash", nullptr, STR, false, "hash function to use [k > 31 only: not supported by the GPU path]"},
    {"interval95", nullptr, BOOL, false, "Set the interval width to probability 0.95"},
    {"found-threshold", "found", INT, false, "Minimum coverage breadth for class `found` [0 - 100 %] (default 90)"},
};
const Opt KMER_COUNTER_OPTS[] = {
    {"k", "k", INT, true, "k-mer size"},
    {"reads", "i", FILES, true, "list of input files"},
    {"output-dir", "o", STR, false, "Output directory (default <work-dir>/kmers)"},
};

template <size_t N>
std::vector<Opt> as_vec(const Opt (&a)[N]) { return std::vector<Opt>(a, a + N); }

const Opt* find_opt(const std::vector<Opt>& opts, const std::string& tok, bool is_long) {
    for (const Opt& o : opts) {
        if (is_long && tok == o.name) return &o;
        if (!is_long && o.short_name && tok == o.short_name) return &o;
    }
    return nullptr;
}

void parse_args(const std::vector<std::string>& argv, const std::vector<Opt>& opts, Args& out) {
    size_t i = 0;
    while (i < argv.size()) {
        std::string tok = argv[i++];
        if (tok.size() < 2 || tok[0] != '-') die("Unrecognized option: %s", tok.c_str());
        bool is_long = tok[1] == '-';
        std::string name = tok.substr(is_long ? 2 : 1), inline_val;
        bool has_inline = false;
        size_t eq = name.find('=');
        if (eq != std::string::npos) {
            inline_val = name.substr(eq + 1);
            name = name.substr(0, eq);
            has_inline = true;
        }
        const Opt* o = find_opt(opts, name, is_long);
        if (!o && !is_long) o = find_opt(opts, name, true);   // commons-cli also resolves "-name"
        if (!o) die("Unrecognized option: %s", tok.c_str());
        std::vector<std::string>& vals = out.values[o->name];
        vals.clear();
        auto is_option = [&](const std::string& s) { return s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9'); };
        if (has_inline) vals.push_back(inline_val);
        if (o->kind == FILES) {
            while (i < argv.size() && !is_option(argv[i])) vals.push_back(argv[i++]);
            // Tool.parseArray: re-tokenise on "[, ]"
            std::vector<std::string> toks;
            for (const std::string& v : vals) {
                std::string cur;
                for (char c : v) {
                    if (c == '[' || c == ',' || c == ' ' || c == ']') {
                        if (!cur.empty()) toks.push_back(cur);
                        cur.clear();
                    } else cur.push_back(c);
                }
                if (!cur.empty()) toks.push_back(cur);
            }
            vals = toks;
        } else if (o->kind == BOOL) {
            if (!has_inline && i < argv.size() && !is_option(argv[i])) vals.push_back(argv[i++]);   // optional argument
        } else if (!has_inline) {
            if (i >= argv.size()) die("Missing argument for option: %s", o->name);
            vals.push_back(argv[i++]);
        }
    }
}

void check_mandatory(const std::vector<Opt>& opts, const Args& a) {   // Tool.checkArguments (Tool.java:725-736)
    for (const Opt& o : opts)
        if (o.mandatory && !a.has(o.name)) {
            std::string s = o.short_name ? std::string(" (-") + o.short_name + ")" : "";
            die("Mandatory argument --%s%s not set", o.name, s.c_str());
        }
}

void print_help(const char* tool, const std::vector<Opt>& opts) {
    printf("metacherchant (B200-native hot path) — tool %s\n\nUsage: metacherchant --tool %s [options]\n\n", tool, tool);
    for (const Opt& o : opts) printf("  %s%s%s--%-22s %s%s\n", o.short_name ? "-" : " ", o.short_name ? o.short_name : " ", o.short_name ? ", " : "  ", o.name, o.help,
                                    o.mandatory ? " (mandatory)" : "");
}

#define MC(call)                                                      \
    do {                                                              \
        int _rc = (call);                                             \
        if (_rc != MC_OK) die("%s", mc_last_error());                 \
    } while (0)

std::string base_name(const std::string& p) {
    size_t s = p.find_last_of('/');
    return s == std::string::npos ? p : p.substr(s + 1);
}

// IOUtils.loadReads with the reference's INFO lines (src/io/IOUtils.java:283-316)
mc_table* load_reads(mc_ctx* ctx, const std::vector<std::string>& files, int k) {
    mc_table* table = nullptr;
    MC(mc_table_create(ctx, k, 0, 0, 0, &table));
    for (const std::string& f : files) {
        info("Loading file %s...", base_name(f).c_str());
        uint64_t before = 0, after = 0;
        mc_table_stats(table, &before, nullptr, nullptr, nullptr);
        MC(mc_count_file(ctx, f.c_str(), table));   // one file at a time so the per-file "reads added" line can be printed
        mc_table_stats(table, &after, nullptr, nullptr, nullptr);
        char buf[64];
        snprintf(buf, sizeof buf, "%llu", (unsigned long long)(after - before));
        info("%s reads added", buf);
    }
    return table;
}

int run_environment_finder(const std::vector<std::string>& rest, const Args& launch) {
    std::vector<Opt> opts = as_vec(ENV_FINDER_OPTS);
    Args a;
    parse_args(rest, opts, a);
    check_mandatory(opts, a);
    for (const char* bad : {"forcehash", "hash"})
        if (a.has(bad) && !(std::string(bad) == "forcehash" && !a.boolean(bad, false)))
            die("--%s is not supported by the GPU path (the reference's lossy hashed k-mer mode is out of scope)", bad);
    const bool merge = a.boolean("merge", false), trim = a.boolean("trim", false);
    long long k = 0, max_kmers = -1, max_radius = -1, coverage = 1, chunk_len = 1, device = 0;
    a.integer("k", k);
    bool has_mk = a.integer("maxkmers", max_kmers), has_mr = a.integer("maxradius", max_radius);
    a.integer("coverage", coverage);
    a.integer("chunklength", chunk_len);
    launch.integer("device", device);
    if (k > 31) die("k > 31 selects the reference's lossy hashed mode (LargeKIOUtils), which is not supported by the GPU path");
    if (k < 1) die("k must be positive");
    // getTerminationMode (EnvironmentFinderMain.java:171-175)
    if (!has_mk && !has_mr) die("At least one of --maxkmers and --maxradius parameters should be set");
    if (!has_mk) max_kmers = -1;
    if (!has_mr) max_radius = -1;
    const bool both = a.boolean("bothdirs", false);
    std::vector<std::string> reads = a.values.count("reads") ? a.values.at("reads") : std::vector<std::string>();

    mc_ctx* ctx = nullptr;
    MC(mc_ctx_create((int)device, &ctx));
    mc_table* table = load_reads(ctx, reads, (int)k);
    info("Hashtable size: %llu kmers", (unsigned long long)mc_table_size(table));
    mc_targets* targets = nullptr;
    MC(mc_targets_load(a.str("seq").c_str(), &targets));
    const int n = mc_targets_count(targets);
    // --hicseq: its sequences seed the merged environment; loadInput also overwrites the output names with the Hi-C file's
    // header list (EnvironmentFinderMain.java:146-150), reproduced here
    mc_targets* hic = nullptr;
    std::vector<const char*> hic_seqs;
    if (a.has("hicseq")) {
        MC(mc_targets_load(a.str("hicseq").c_str(), &hic));
        for (int i = 0; i < mc_targets_count(hic); i++) hic_seqs.push_back(mc_targets_seq(hic, i));
    }
    std::vector<const char*> seqs, names;
    for (int i = 0; i < n; i++) {
        seqs.push_back(mc_targets_seq(targets, i));
        if (merge) continue;
        const char* name = mc_targets_name(hic ? hic : targets, i);
        if (!name) die("IndexOutOfBoundsException: sequence %d of %s has no header entry", i, a.str("seq").c_str());
        names.push_back(name);
        std::string s = seqs.back();
        if ((long long)s.size() >= 2 * k) s = s.substr(0, (size_t)k) + "..." + s.substr(s.size() - (size_t)k) + " (length=" + std::to_string(s.size()) + ")";
        info("Finding environment for sequence %s", s.c_str());   // OneSequenceCalculator.java:101, StringUtils.shortenLabel
    }
    if (merge) {
        info("hicSequences = %d", (int)hic_seqs.size());                                    // EnvironmentFinderMain.java:227
        info("Finding single environment for %d sequences", n);                               // OneSequenceCalculator.java:103
    }
    // B2/K4: coverage-threshold compaction; the BFS and the unitig stage probe the small solid table
    mc_table* solid = nullptr;
    MC(mc_table_filter(ctx, table, (int)coverage, &solid));
    debug("%llu k-mers with coverage >= %lld", (unsigned long long)mc_table_size(solid), coverage);
    std::vector<int> found((size_t)std::max(n, 1), 0);
    MC(mc_env_find_ex(ctx, solid, seqs.data(), merge ? nullptr : names.data(), n, hic_seqs.empty() ? nullptr : hic_seqs.data(), (int)hic_seqs.size(),
                      (int)k, (int)coverage, max_kmers, max_radius, both ? 1 : 0, (int)chunk_len, merge ? 1 : 0, trim ? 1 : 0,
                      a.str("output").c_str(), found.data()));
    {
        std::vector<long long> ext((size_t)std::max(n, 1), -1);
        int n_ext = 0;
        MC(mc_env_extend_counts(ctx, ext.data(), (int)ext.size(), &n_ext));
        for (int i = 0; i < (merge ? 1 : n); i++) {
            if (!found[(size_t)i]) info("Could not find any k-mers of the target gene in the input, halting.");
            else if (i < n_ext && ext[(size_t)i] >= 0) info("Extending endings by %lld kmers", ext[(size_t)i]);   // OneSequenceCalculator.java:291
        }
    }
    info("Finished processing all sequences!");
    mc_targets_free(targets);
    if (hic) mc_targets_free(hic);
    mc_table_free(solid);
    mc_table_free(table);
    mc_ctx_destroy(ctx);
    return 0;
}

int run_multi(const std::vector<std::string>& rest, const Args& launch) {
    std::vector<Opt> opts = as_vec(MULTI_OPTS);
    Args a;
    parse_args(rest, opts, a);
    check_mandatory(opts, a);
    long long gene_id = 1, device = 0;
    a.integer("geneid", gene_id);
    launch.integer("device", device);
    std::vector<const char*> envs;
    for (const std::string& e : a.values.at("env")) envs.push_back(e.c_str());
    if (envs.empty()) die("Zero environments given");
    if (envs.size() > 256) fprintf(stderr, "WARN: Found more than 256 environments. Grayscale graph may be not accurate.\n");
    mc_ctx* ctx = nullptr;
    MC(mc_ctx_create((int)device, &ctx));
    MC(mc_env_multi(ctx, envs.data(), (int)envs.size(), a.str("seq").c_str(), (int)gene_id, a.str("output").c_str()));
    info("Finished processing!");
    mc_ctx_destroy(ctx);
    return 0;
}

int run_kmer_counter(const std::vector<std::string>& rest, const Args& launch, const std::string& work_dir) {
    std::vector<Opt> opts = as_vec(KMER_COUNTER_OPTS);
    Args a;
    parse_args(rest, opts, a);
    check_mandatory(opts, a);
    long long k = 0, device = 0;
    a.integer("k", k);
    launch.integer("device", device);
    if (k > 31) die("k > 31 selects the reference's lossy hashed mode, which is not supported by the GPU path");
    std::vector<std::string> reads = a.values.at("reads");
    if (reads.empty()) die("Mandatory argument --reads (-i) not set");
    std::string out_dir = a.str("output-dir", work_dir + "/kmers");
    mkdirs(out_dir);
    mc_ctx* ctx = nullptr;
    MC(mc_ctx_create((int)device, &ctx));
    mc_table* table = load_reads(ctx, reads, (int)k);
    info("Hashtable size: %llu kmers", (unsigned long long)mc_table_size(table));
    // name of the first input without its extension (ReadersUtils.readDnaLazy(file).name(), KmersCounter.java:93)
    std::string name = base_name(reads[0]);
    for (const char* ext : {".fastq", ".fq", ".fasta", ".fa", ".fn", ".fna"}) {
        size_t m = strlen(ext);
        if (name.size() > m && name.compare(name.size() - m, m, ext) == 0) { name = name.substr(0, name.size() - m); break; }
    }
    std::string out_file = out_dir + "/" + name + ".kmers.bin", st_file = out_dir + "/" + name + ".stat.txt";
    uint64_t good = 0;
    MC(mc_table_print_kmers(table, 0, out_file.c_str(), st_file.c_str(), &good));
    uint64_t size = mc_table_size(table);
    info("%llu k-mers found, %llu (%.1f%%) of them is good (not erroneous)", (unsigned long long)size, (unsigned long long)good, size ? good * 100.0 / size : 0.0);
    info("k-mers printed to %s", out_file.c_str());
    mc_table_free(table);
    mc_ctx_destroy(ctx);
    return 0;
}

// ---- reads-classifier ("next" row f-4b) ---------------------------------------------------------------------------------
// One FASTQ record as FastqReader keeps it (ITMO!/io/readers/FastqReader.java:53-82): bases as nucleotides (N, n, . -> A) and
// the stored 6-bit phred; the tool writes its classes back through WritersUtils.writeDnaQsToFastqFile, i.e. with numbered
// headers, the stored nucleotides and Phred+64 qualities.
struct FqRead { std::string nucs, phred; };
std::vector<uint8_t> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("Failed to read from file %s", path.c_str());
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + got);
    fclose(f);
    return data;
}
std::vector<FqRead> parse_fastq(const std::vector<uint8_t>& d, int qoff) {
    std::vector<std::string> lines;   // BufferedReader.readLine
    size_t p = 0;
    while (p < d.size()) {
        size_t s = p;
        while (p < d.size() && d[p] != '\n' && d[p] != '\r') p++;
        lines.emplace_back((const char*)d.data() + s, p - s);
        if (p < d.size()) p += (d[p] == '\r' && p + 1 < d.size() && d[p + 1] == '\n') ? 2 : 1;
    }
    std::vector<FqRead> out;
    size_t i = 0;
    auto data_line = [&](std::string& v) -> bool {   // readNextDataLine (:84-110); structural errors were reported by the library already
        while (i < lines.size() && lines[i].empty()) i++;
        if (i >= lines.size()) return false;
        i++;
        if (i >= lines.size()) return false;
        v = lines[i++];
        return true;
    };
    std::string seq, qual;
    while (data_line(seq) && data_line(qual)) {
        FqRead r;
        r.nucs.resize(seq.size());
        r.phred.resize(seq.size());
        for (size_t j = 0; j < seq.size(); j++) {
            const char c = seq[j];
            const bool unknown = c == 'N' || c == 'n' || c == '.';
            char u = (char)toupper((unsigned char)c);
            r.nucs[j] = unknown ? 'A' : u;
            r.phred[j] = (char)(64 + (unknown ? 0 : (((unsigned char)qual[j] - qoff) & 63)));
        }
        out.push_back(std::move(r));
    }
    return out;
}
void write_fastq(const std::vector<const FqRead*>& reads, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) die("cannot write %s", path.c_str());
    long n = 0;
    for (const FqRead* r : reads) {
        if (r->nucs.empty()) { fclose(f); die("Empty DnaQ!"); }   // FastqDedicatedWriter.writeData
        fprintf(f, "@%ld\n%s\n+\n%s\n", ++n, r->nucs.c_str(), r->phred.c_str());
    }
    fclose(f);
}

int run_reads_classifier(const std::vector<std::string>& rest, const Args& launch, const std::string& work_dir) {
    std::vector<Opt> opts = as_vec(READS_CLASSIFIER_OPTS);
    Args a;
    parse_args(rest, opts, a);
    check_mandatory(opts, a);
    long long k = 0, device = 0, found_pct = 90;
    a.integer("k", k);
    a.integer("found-threshold", found_pct);
    launch.integer("device", device);
    if (k > 31) die("k > 31 selects the reference's lossy hashed mode (LargeKIOUtils), which is not supported by the GPU path");
    if (k < 1) die("k must be positive");
    const std::vector<std::string> inputs = a.values.at("input-files"), reads = a.values.at("read-files");
    if (inputs.empty()) die("Mandatory argument --input-files (-i) not set");
    if (reads.empty()) die("Mandatory argument --read-files (-r) not set");
    const std::string out_dir = a.str("output-dir", work_dir + "/reads_classifier");
    const bool correction = a.boolean("correction", false);
    const double z = a.boolean("interval95", false) ? 1.96 : 1.0, threshold = (double)found_pct / 100.0;
    mkdirs(out_dir);
    mc_ctx* ctx = nullptr;
    MC(mc_ctx_create((int)device, &ctx));
    // loadGraph (:86-105): k-mers in binary format, or reads
    mc_table* graph = nullptr;
    std::string first = base_name(inputs[0]);
    for (auto& c : first) c = (char)tolower((unsigned char)c);
    if (first.size() >= 9 && first.compare(first.size() - 9, 9, "kmers.bin") == 0) {
        MC(mc_table_create(ctx, (int)k, 0, 0, 0, &graph));
        for (const std::string& f : inputs) MC(mc_table_load_kmers(graph, f.c_str(), 0));
    } else {
        graph = load_reads(ctx, inputs, (int)k);
    }
    info("Hashtable size: %llu kmers", (unsigned long long)mc_table_size(graph));
    info("Loading reads...");
    std::vector<std::vector<FqRead>> src;
    std::vector<std::vector<uint8_t>> flags;
    for (size_t fi = 0; fi < reads.size() && fi < 2; fi++) {
        int gz = 0;
        const int fmt = mc_detect_file_format(reads[fi].c_str(), &gz);
        if (fmt < 0) die("%s", mc_last_error());
        if (fmt != MC_FMT_FASTQ || gz) die("reads-classifier on the GPU path reads plain FASTQ files (%s)", reads[fi].c_str());
        const std::vector<uint8_t> bytes = slurp(reads[fi]);
        const int qoff = mc_sniff_quality_offset(bytes.data(), bytes.size());
        size_t n = 0;
        MC(mc_classify_bytes(ctx, graph, bytes.data(), bytes.size(), qoff, z, threshold, correction ? 1 : 0, nullptr, 0, &n));
        flags.emplace_back(n);
        if (fi == 0) info(correction ? "Searching for corrected reads in graph..." : "Searching for reads in graph...");
        if (n) MC(mc_classify_bytes(ctx, graph, bytes.data(), bytes.size(), qoff, z, threshold, correction ? 1 : 0, flags.back().data(), n, &n));
        src.push_back(parse_fastq(bytes, qoff));
        if (src.back().size() != n) die("internal: the host reader sees %zu records in %s, the device %zu", src.back().size(), reads[fi].c_str(), n);
    }
    // PairFinder.run (src/algo/PairFinder.java:30-55) + the six output files (ReadsClassifier.java:204-217), reads in input order
    const bool paired = src.size() == 2;
    const size_t n = paired ? std::min(src[0].size(), src[1].size()) : src[0].size();
    const FqRead empty;
    std::vector<const FqRead*> found_1, found_2, not_found_1, not_found_2, found_s, not_found_s, tail_found, tail_not;
    long both_found = 0, first_found = 0, second_found = 0, both_not = 0;
    for (size_t i = 0; i < n; i++) {
        const FqRead& r1 = src[0][i];
        const FqRead& r2 = paired ? src[1][i] : empty;
        bool f1 = flags[0][i] != 0, f2 = paired ? flags[1][i] != 0 : false;
        if (r2.nucs.empty()) f2 = !f1;
        if (f1 && f2) { both_found++; found_1.push_back(&r1); found_2.push_back(&r2); }
        else if (f1) { first_found++; if (!r1.nucs.empty()) found_s.push_back(&r1); if (!r2.nucs.empty()) not_found_s.push_back(&r2); }
        else if (f2) { second_found++; if (!r2.nucs.empty()) tail_found.push_back(&r2); if (!r1.nucs.empty()) tail_not.push_back(&r1); }
        else { both_not++; not_found_1.push_back(&r1); not_found_2.push_back(&r2); }
    }
    found_s.insert(found_s.end(), tail_found.begin(), tail_found.end());
    not_found_s.insert(not_found_s.end(), tail_not.begin(), tail_not.end());
    const long total = 2 * (both_not + both_found + first_found + second_found), paired_n = 2 * (both_found + both_not);
    const long found_n = 2 * both_found + first_found + second_found, not_found_n = 2 * both_not + first_found + second_found;
    info("|\tTotal: %ld reads", total);
    info("|\tPaired: %ld reads", paired_n);
    info("|\tTotal quality: %.2f %%", 100.0 * (double)paired_n / (double)total);
    info("|\tFound: %ld reads", found_n);
    info("|\tPercent of found reads: %.2f %%", 100.0 * (double)found_n / (double)total);
    info("|\tQuality of found bin: %.2f %%", (double)both_found * 2 / (double)(both_found * 2 + first_found + second_found) * 100);
    info("|\tNot found: %ld reads", not_found_n);
    info("|\tPercent of not found reads: %.2f %%", 100.0 * (double)not_found_n / (double)total);
    info("|\tQuality of not found bin: %.2f %%", (double)both_not * 2 / (double)(both_not * 2 + first_found + second_found) * 100);
    info("Writing classified reads...");
    write_fastq(found_1, out_dir + "/found_1.fastq");
    write_fastq(found_2, out_dir + "/found_2.fastq");
    write_fastq(not_found_1, out_dir + "/not_found_1.fastq");
    write_fastq(not_found_2, out_dir + "/not_found_2.fastq");
    write_fastq(found_s, out_dir + "/found_s.fastq");
    write_fastq(not_found_s, out_dir + "/not_found_s.fastq");
    info("Reads have been written. Finishing...");
    mc_table_free(graph);
    mc_ctx_destroy(ctx);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    // split launch options from tool options (ITMO Runner.run strips --tool, Tool.mainImpl the rest)
    std::vector<Opt> launch_opts = as_vec(LAUNCH_OPTS);
    std::vector<std::string> launch_args, rest;
    for (int i = 1; i < argc;) {
        std::string tok = argv[i];
        bool is_long = tok.size() > 2 && tok[0] == '-' && tok[1] == '-';
        bool is_short = tok.size() >= 2 && tok[0] == '-' && tok[1] != '-';
        std::string name = is_long ? tok.substr(2) : is_short ? tok.substr(1) : tok;
        size_t eq = name.find('=');
        if (eq != std::string::npos) name = name.substr(0, eq);
        const Opt* o = (is_long || is_short) ? find_opt(launch_opts, name, is_long) : nullptr;
        // -k/-i/-o/-e/-g belong to the tools; launch short options are t m p w c s f v h
        if (o) {
            launch_args.push_back(tok);
            i++;
            bool takes = o->kind != BOOL ? eq == std::string::npos : false;
            if (o->kind == BOOL && eq == std::string::npos && i < argc) {
                std::string nx = argv[i];
                for (auto& c : nx) c = (char)tolower((unsigned char)c);
                if (nx == "true" || nx == "false") takes = true;
            }
            if (takes && i < argc) launch_args.push_back(argv[i++]);
        } else {
            rest.push_back(tok);
            i++;
        }
    }
    Args launch;
    parse_args(launch_args, launch_opts, launch);
    g_verbose = launch.boolean("verbose", false);
    std::string tool = launch.str("tool", "environment-finder");   // default tool (src/Runner.java:22)
    if (launch.boolean("help", false) || argc == 1) {
        if (tool == "environment-finder-multi") print_help(tool.c_str(), as_vec(MULTI_OPTS));
        else if (tool == "kmer-counter") print_help(tool.c_str(), as_vec(KMER_COUNTER_OPTS));
        else if (tool == "reads-classifier") print_help(tool.c_str(), as_vec(READS_CLASSIFIER_OPTS));
        else print_help("environment-finder", as_vec(ENV_FINDER_OPTS));
        printf("\nLaunch options:\n");
        print_help("(all tools)", launch_opts);
        return 0;
    }
    std::string work_dir = launch.str("work-dir", "workDir");
    mkdirs(work_dir);
    g_log = fopen((work_dir + "/log").c_str(), "a");   // Tool.updateFileLoggers (Tool.java:666-711)
    int rc;
    if (tool == "environment-finder") rc = run_environment_finder(rest, launch);
    else if (tool == "environment-finder-multi") rc = run_multi(rest, launch);
    else if (tool == "kmer-counter") rc = run_kmer_counter(rest, launch, work_dir);
    else if (tool == "reads-classifier") rc = run_reads_classifier(rest, launch, work_dir);
    else die("Unknown tool '%s': the GPU path provides environment-finder, environment-finder-multi, kmer-counter and reads-classifier", tool.c_str());
    if (g_log) fclose(g_log);
    return rc;
}
