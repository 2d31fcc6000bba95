/* Synthetic metagenome workload generator — BENCH/TEST UTILITY, not part of the
 * reference-facing API.  All-integer, counter-based (splitmix64 of seed^index), so the CPU twin
 * (oracle/mc_synth_cpu.cpp) and the CUDA twin (metacherchant_b200/csrc/mc_synth.cu) emit
 * byte-identical FASTQ/FASTA for the same parameters.
 *
 * Community: n_genomes genomes, genome g has length genome_unit*(g+1), i.i.d. uniform bases.
 * Targets ("AR genes"): n_targets sequences of target_len bases, each planted (overwriting) into 1..3
 * genomes; optional shared repeat of repeat_len bases planted right after every target copy.
 * Reads: read_len bases, uniform start, random strand, substitution errors (err_ppm per base),
 * 'N' at n_ppm per base, quality constant 'I' with '!' at bangq_ppm; record 0 carries one '5'
 * so the reference's sniffer (ReadersUtils.java:63-77) picks Phred+33.
 * FASTQ record: "@r" + 8 hex digits + "\n" + seq + "\n+\n" + qual + "\n"  (2*read_len + 15 bytes)
 * FASTA record: ">r" + 8 hex digits + "\n" + seq + "\n"                     (read_len + 12 bytes)
 */
#ifndef MC_SYNTH_H
#define MC_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD static inline
#endif

#define MC_SYNTH_MAX_GENOMES 16
#define MC_SYNTH_MAX_PLANTS 2048

typedef struct mc_synth_plant {
    uint64_t start;      /* global coordinate (genomes concatenated) of the first planted base */
    uint32_t len;        /* target_len + repeat_len */
    uint32_t target;     /* which target sequence */
} mc_synth_plant;

typedef struct mc_synth_params {
    uint64_t seed;
    uint32_t n_genomes;
    uint32_t read_len;
    uint64_t genome_unit;                          /* genome g has genome_unit*(g+1) bases */
    uint64_t genome_start[MC_SYNTH_MAX_GENOMES + 1]; /* cumulative global coordinates */
    uint32_t genome_cum_weight[MC_SYNTH_MAX_GENOMES]; /* cumulative 32-bit thresholds, abundance ~ 2^(-g/3) */
    uint32_t n_targets, target_len, repeat_len;
    uint32_t n_plants;                             /* sorted by start, non-overlapping */
    uint32_t err_ppm, n_ppm, bangq_ppm;            /* per-base probabilities in parts per million */
    uint32_t fasta;                                /* 0 = FASTQ, 1 = FASTA */
    mc_synth_plant plants[MC_SYNTH_MAX_PLANTS];
} mc_synth_params;

MC_HD uint64_t mc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* ASCII order here is only a choice of the generator: code 0..3 -> "ACGT" */
MC_HD char mc_synth_base_char(uint32_t code) { return "ACGT"[code & 3]; }

MC_HD uint32_t mc_synth_target_code(const mc_synth_params* p, uint32_t target, uint32_t j) {
    if (j < p->target_len) return (uint32_t)(mc_splitmix64(p->seed ^ 0x7461726765740000ULL ^ ((uint64_t)target << 20) ^ j) & 3);
    /* the shared repeat is the same sequence for every target */
    return (uint32_t)(mc_splitmix64(p->seed ^ 0x7265706561740000ULL ^ (uint64_t)(j - p->target_len)) & 3);
}

/* base code of the community at global coordinate g */
MC_HD uint32_t mc_synth_genome_code(const mc_synth_params* p, uint64_t g) {
    /* binary search: last plant with start <= g */
    uint32_t lo = 0, hi = p->n_plants;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (p->plants[mid].start <= g) lo = mid + 1; else hi = mid;
    }
    if (lo > 0) {
        const mc_synth_plant* pl = &p->plants[lo - 1];
        if (g - pl->start < pl->len) return mc_synth_target_code(p, pl->target, (uint32_t)(g - pl->start));
    }
    return (uint32_t)(mc_splitmix64(p->seed ^ 0x67656E6F6D650000ULL ^ (g * 0x9E3779B1ULL)) & 3);
}

MC_HD uint32_t mc_synth_record_bytes(const mc_synth_params* p) {
    return p->fasta ? p->read_len + 12u : 2u * p->read_len + 15u;
}

/* writes one record (mc_synth_record_bytes bytes) for read index i */
MC_HD void mc_synth_write_record(const mc_synth_params* p, uint64_t i, uint8_t* out) {
    const uint32_t L = p->read_len;
    uint64_t h = mc_splitmix64(p->seed ^ 0x7265616400000000ULL ^ (i * 4));
    uint32_t pick = (uint32_t)(h >> 32);
    uint32_t g = 0;
    while (g + 1 < p->n_genomes && pick >= p->genome_cum_weight[g]) g++;
    uint64_t glen = p->genome_start[g + 1] - p->genome_start[g];
    uint64_t h2 = mc_splitmix64(p->seed ^ 0x7265616400000000ULL ^ (i * 4 + 1));
    uint64_t start = p->genome_start[g] + h2 % (glen - L + 1);
    uint32_t strand = (uint32_t)(h & 1);
    uint32_t pos = 0;
    out[pos++] = p->fasta ? '>' : '@';
    out[pos++] = 'r';
    for (int d = 7; d >= 0; d--) out[pos++] = (uint8_t)"0123456789abcdef"[(i >> (4 * d)) & 15];
    out[pos++] = '\n';
    uint8_t* seq = out + pos;
    uint8_t* qual = out + pos + L + 3;
    for (uint32_t j = 0; j < L; j++) {
        uint32_t code = strand ? (3u - mc_synth_genome_code(p, start + (L - 1 - j))) : mc_synth_genome_code(p, start + j);
        uint64_t e = mc_splitmix64(p->seed ^ 0x6572726F72000000ULL ^ (i * 1024 + j));
        uint32_t r_err = (uint32_t)(e % 1000000u);
        uint32_t r_n = (uint32_t)((e >> 20) % 1000000u);
        uint32_t r_q = (uint32_t)((e >> 40) % 1000000u);
        if (r_err < p->err_ppm) code = (code + 1 + (uint32_t)((e >> 60) % 3)) & 3;
        char c = mc_synth_base_char(code);
        if (r_n < p->n_ppm) c = 'N';
        seq[j] = (uint8_t)c;
        if (!p->fasta) qual[j] = (uint8_t)((r_q < p->bangq_ppm) ? '!' : 'I');
    }
    pos += L;
    out[pos++] = '\n';
    if (!p->fasta) {
        out[pos++] = '+';
        out[pos++] = '\n';
        if (i == 0) qual[0] = '5';
        pos += L;
        out[pos++] = '\n';
    }
}

#ifndef __CUDACC__
#include <math.h>
#include <string.h>
/* Host-side parameter construction (shared by every user so CPU and GPU twins agree).
 * Returns 0, or -1 if the plants do not fit. */
static inline int mc_synth_init(mc_synth_params* p, uint64_t seed, uint32_t n_genomes, uint64_t genome_unit, uint32_t read_len,
                                uint32_t n_targets, uint32_t target_len, uint32_t repeat_len, uint32_t err_ppm, uint32_t n_ppm,
                                uint32_t bangq_ppm, uint32_t fasta) {
    memset(p, 0, sizeof *p);
    if (n_genomes < 1 || n_genomes > MC_SYNTH_MAX_GENOMES) return -1;
    p->seed = seed; p->n_genomes = n_genomes; p->genome_unit = genome_unit; p->read_len = read_len;
    p->n_targets = n_targets; p->target_len = target_len; p->repeat_len = repeat_len;
    p->err_ppm = err_ppm; p->n_ppm = n_ppm; p->bangq_ppm = bangq_ppm; p->fasta = fasta;
    double wsum = 0, w[MC_SYNTH_MAX_GENOMES];
    for (uint32_t g = 0; g < n_genomes; g++) { w[g] = pow(2.0, -(double)g / 3.0); wsum += w[g]; }
    double acc = 0;
    p->genome_start[0] = 0;
    for (uint32_t g = 0; g < n_genomes; g++) {
        acc += w[g] / wsum;
        double t = acc * 4294967296.0;
        p->genome_cum_weight[g] = (g + 1 == n_genomes || t >= 4294967295.0) ? 0xFFFFFFFFu : (uint32_t)t;
        p->genome_start[g + 1] = p->genome_start[g] + genome_unit * (g + 1);
    }
    /* plant every target into 1..3 genomes at slot-aligned positions; slots are plant_len*2 wide so plants never touch */
    uint32_t plant_len = target_len + repeat_len;
    uint64_t slot = (uint64_t)plant_len * 2;
    uint32_t np = 0;
    for (uint32_t t = 0; t < n_targets; t++) {
        uint32_t copies = 1 + (uint32_t)(mc_splitmix64(seed ^ 0x636F706965730000ULL ^ t) % 3);
        for (uint32_t c = 0; c < copies; c++) {
            uint64_t hh = mc_splitmix64(seed ^ 0x706C616E74000000ULL ^ ((uint64_t)t * 8 + c));
            for (uint32_t attempt = 0; attempt < 64; attempt++, hh = mc_splitmix64(hh)) {
                uint32_t g = (uint32_t)(hh % n_genomes);
                uint64_t glen = genome_unit * (g + 1);
                if (glen < slot * 2) continue;
                uint64_t nslots = glen / slot;
                uint64_t s = p->genome_start[g] + ((hh >> 16) % nslots) * slot + plant_len / 2;
                if (s + plant_len > p->genome_start[g + 1]) continue;
                int clash = 0;
                for (uint32_t q = 0; q < np; q++)
                    if (p->plants[q].start / slot == s / slot) { clash = 1; break; }
                if (clash) continue;
                if (np >= MC_SYNTH_MAX_PLANTS) return -1;
                p->plants[np].start = s; p->plants[np].len = plant_len; p->plants[np].target = t; np++;
                break;
            }
        }
    }
    /* insertion sort by start */
    for (uint32_t a = 1; a < np; a++) {
        mc_synth_plant v = p->plants[a];
        uint32_t b = a;
        while (b > 0 && p->plants[b - 1].start > v.start) { p->plants[b] = p->plants[b - 1]; b--; }
        p->plants[b] = v;
    }
    p->n_plants = np;
    return 0;
}
#endif /* !__CUDACC__ */

#endif /* MC_SYNTH_H */
