// seed_table.h -- minimizer seed-position table and the D-SOFT filter.
//
// Host-side producer of the hot path's input: every candidate it emits becomes
// one GACT extension (darwin.cpp:215-246).  The candidate stream must equal the
// reference's bit for bit, so the observable quirks are kept (each cited below):
//   * 2-bit packing, A/a C/c G/g T/t -> 0..3, everything else -> 0 (ntcoding.cpp:60-103)
//   * Thomas Wang hash masked to 4^k (ntcoding.cpp:74-85)
//   * minimizers report the LOOP position, not the arg-min position, and the
//     query side scans up to 16*ceil(len/16) - k - w (ntcoding.cpp:126-182)
//   * D-SOFT uses N+1 seeds (`num_seeds > N` test, seed_pos_table.cpp:128) and
//     remembers at most 25,000,000 non-zero bins (macro nz_bins, seed_pos_table.h:33)
#pragma once
#include <cstdint>
#include <vector>

namespace darwin {

class SeedTable {
public:
    // ref: concatenated, bin-padded reference string (darwin.cpp:530-543)
    SeedTable(const char *ref, uint32_t ref_len, int kmer_size, uint32_t seed_occurence_multiple,
              uint32_t bin_size, uint32_t window_size, int build_threads);
    ~SeedTable();
    SeedTable(const SeedTable &) = delete;
    SeedTable &operator=(const SeedTable &) = delete;

    uint32_t num_bins() const { return 1u + (ref_len_ >> log_bin_size_); }   // darwin.cpp:183

    // Per-thread scratch of DSOFT (darwin.cpp:193-199).
    struct Scratch {
        std::vector<uint64_t> bin_count_offset;   // num_bins entries, zero between calls
        std::vector<uint32_t> nz_bins;            // num_nz_bins entries
        std::vector<uint32_t> q2bit;
        std::vector<uint64_t> minimizers;
        Scratch(const SeedTable &t, int num_nz_bins);
    };

    // seed_pos_table.cpp:100-167.  candidates: (hit << 32) | offset, appended in
    // emission order; returns their number.
    int dsoft(const char *query, uint32_t query_len, int num_seeds, int threshold,
              int max_candidates, Scratch &s, std::vector<uint64_t> &candidates) const;

    uint32_t num_minimizers() const { return n_pos_; }
    // raw tables, for the device-side filter (gact_dsoft_create)
    const uint32_t *index_table() const { return index_table_; }
    uint64_t index_entries() const { return ((uint64_t)1 << (2 * k_)) + 1; }
    const uint32_t *pos_table() const { return pos_table_; }
    uint32_t kmer_max_occurence() const { return kmer_max_occurence_; }
    int kmer_size() const { return k_; }
    int window_size() const { return w_; }
    uint32_t bin_size() const { return bin_size_; }

private:
    uint32_t ref_len_, bin_size_, log_bin_size_, kmer_max_occurence_;
    int k_, w_;
    uint32_t *index_table_ = nullptr;     // 4^k + 1 entries: #minimizers with hash <= s
    uint32_t *pos_table_ = nullptr;
    uint32_t n_pos_ = 0;
};

// exposed for tests
void pack_two_bit(const char *seq, uint32_t len, uint32_t *out, uint32_t n_words);
uint32_t wang_hash32(uint32_t key, int k);

}  // namespace darwin
