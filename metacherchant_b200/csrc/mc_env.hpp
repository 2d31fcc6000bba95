// Host-side structures shared by the BFS (mc_bfs.cu), the unitig stage (mc_unitig.cu) and the
// environment driver / writers (mc_env.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mc_common.cuh"

struct McBfsRequest {
    std::vector<uint8_t> codes;   // seed sequence(s) as nucleotide codes; code 4 separates the sequences of a merged seed list
    int dir = 0;                  // -1 left, +1 right, 0 both
    bool trim = false;            // --trim: also report which k-mers are "lastKmers" (OneSequenceCalculator.java:208-216)
};
struct McBfsResult {
    std::vector<uint64_t> kmers;  // oriented, discovery order
    std::vector<int32_t> dist;
    std::vector<int16_t> counts;
    std::vector<uint8_t> last;    // with trim: 1 = some solid neighbour of this k-mer was refused by allowsAddition
};

std::vector<uint8_t> mc_codes_from_target(const std::string& seq);
void mc_run_bfs_passes(mc_ctx* ctx, const mc_table* t, int k, int min_cov, int64_t max_kmers, int64_t max_radius,
                       const std::vector<McBfsRequest>& reqs, std::vector<McBfsResult>& results);

// ---- unitig stage --------------------------------------------------------------------------------
// Input: the environment in `subgraph` iteration order: node 2i = env k-mer i in its printed (ASCII-canonical)
// orientation, node 2i+1 = its reverse complement (initializeStructures, OneSequenceCalculator.java:387-405).
struct McUnitigInput {
    int k = 0;
    std::vector<uint64_t> kmers;            // oriented code of node 2i
    std::vector<int32_t> counts;            // coverage of env k-mer i (summed over environments in multi mode)
    std::vector<uint64_t> gene_windows;     // sorted codes of every k-window of the target(s), forward strand
    std::vector<uint64_t> colors;           // multi mode: bitmask/hash of the set of environments holding k-mer i (else empty)
    // extendEnvironment's log line (OneSequenceCalculator.java:265-295): with a table, the stage also counts the environment k-mers
    // that have exactly ONE solid neighbour outside the environment ("Extending endings by N kmers")
    const mc_table* table = nullptr;
    int min_cov = 0;
    bool order_by_nucleotide = false;       // neighbour lists in A,G,C,T order of the appended base (MultiSequenceCalculator.java:90-99)
                                            // instead of ascending node id (OneSequenceCalculator.java:406-418)
};
struct McUnitigNodes {
    // per oriented node (2n entries)
    std::vector<uint8_t> deg;               // |neighbors| (palindromic neighbours count twice)
    std::vector<int32_t> nbr;               // 4 per node, ascending node id, -1 padded; a palindromic neighbour appears once (even id)
    std::vector<uint8_t> nbr_pal;           // bit s set: nbr[4*node+s] is palindromic, i.e. stands for ids (x, x+1)
    std::vector<uint8_t> gene;              // isGeneNode
    std::vector<int32_t> head;              // head of my merged path (myself when unmerged); -1 when the path is a cycle (deleted)
    std::vector<int32_t> rank;              // my position inside that path
    // Unitigs assembled on the device (doMerge / mergeNodes / mergeLabels, OneSequenceCalculator.java:434-462,312-324): one per
    // alive node = head of a chain, in node-id order.  `regular` = the environment has no palindromic k-mer, no start-to-own-
    // start link and no cyclic chain; only then are these arrays authoritative (the corner cases keep the literal host sweep,
    // whose outcome depends on the reference's sweep order).
    long long extend_count = -1;            // see McUnitigInput.table; -1 = not computed
    bool regular = false;
    std::vector<int32_t> u_node;            // alive node id (the chain's head)
    std::vector<int32_t> u_rc;              // its reverse-complement partner after all merges (tail of the chain ^ 1)
    std::vector<uint32_t> u_off, u_len;     // sequence = u_chars[u_off .. u_off + u_len)
    std::vector<int64_t> u_kc;              // sum of the chain's k-mer coverages + coverage of the last one * (k - 1) (GFAWriter.java:90-95)
    std::string u_chars;
};
void mc_unitig_nodes(mc_ctx* ctx, const McUnitigInput& in, McUnitigNodes& out);
