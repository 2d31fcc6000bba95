// ORACLE-SIDE (test infrastructure): CPU twin of the synthetic workload generator in include/mc_synth.h.
// Produces the exact bytes the CUDA generator (metacherchant_b200/csrc/mc_synth.cu) produces.
#include "../include/mc_synth.h"

#include <cstddef>
#include <thread>
#include <vector>

extern "C" {

size_t orc_synth_params_size() { return sizeof(mc_synth_params); }

int orc_synth_init(mc_synth_params* p, uint64_t seed, uint32_t n_genomes, uint64_t genome_unit, uint32_t read_len, uint32_t n_targets,
                   uint32_t target_len, uint32_t repeat_len, uint32_t err_ppm, uint32_t n_ppm, uint32_t bangq_ppm, uint32_t fasta) {
    return mc_synth_init(p, seed, n_genomes, genome_unit, read_len, n_targets, target_len, repeat_len, err_ppm, n_ppm, bangq_ppm, fasta);
}

uint32_t orc_synth_record_bytes(const mc_synth_params* p) { return mc_synth_record_bytes(p); }

// records [first, first+count) -> out (count * record_bytes bytes)
void orc_synth_generate(const mc_synth_params* p, uint64_t first, uint64_t count, uint8_t* out, int threads) {
    uint32_t rb = mc_synth_record_bytes(p);
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([=] {
            for (uint64_t i = (uint64_t)t; i < count; i += (uint64_t)threads) mc_synth_write_record(p, first + i, out + i * rb);
        });
    for (auto& x : th) x.join();
}

// target sequence t as ASCII (target_len chars, no terminator)
void orc_synth_target(const mc_synth_params* p, uint32_t t, char* out) {
    for (uint32_t j = 0; j < p->target_len; j++) out[j] = mc_synth_base_char(mc_synth_target_code(p, t, j));
}

uint32_t orc_synth_n_plants(const mc_synth_params* p) { return p->n_plants; }

}  // extern "C"
