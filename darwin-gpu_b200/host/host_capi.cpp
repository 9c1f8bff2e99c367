// host_capi.cpp -- C entry points over the host-side seeding code, so that the CPU test-suite
// can exercise it through ctypes (no GPU involved; the GACT path is not reachable from here).
#include <cstring>
#include <string>
#include <vector>

#include "darwin_config.h"
#include "fasta_io.h"
#include "seed_table.h"

using namespace darwin;

extern "C" {

void *dh_seed_table_new(const char *ref, uint32_t ref_len, int k, uint32_t occ_mult, uint32_t bin_size, uint32_t w, int threads)
{
    try { return new SeedTable(ref, ref_len, k, occ_mult, bin_size, w, threads); } catch (...) { return nullptr; }
}
void dh_seed_table_free(void *t) { delete (SeedTable *)t; }
uint32_t dh_seed_table_size(void *t) { return ((SeedTable *)t)->num_minimizers(); }

// returns the number of candidates; writes up to cap of them
int dh_dsoft(void *t, const char *q, uint32_t qlen, int num_seeds, int threshold, int max_cand, int num_nz_bins,
             uint64_t *out, int cap)
{
    SeedTable *tab = (SeedTable *)t;
    SeedTable::Scratch s(*tab, num_nz_bins);
    std::vector<uint64_t> c;
    int n = tab->dsoft(q, qlen, num_seeds, threshold, max_cand, s, c);
    for (int i = 0; i < n && i < cap; i++) out[i] = c[i];
    return n;
}

// raw tables of a SeedTable (to hand to gact_dsoft_create)
void dh_seed_table_arrays(void *t, const uint32_t **index, uint64_t *index_entries, const uint32_t **pos, uint64_t *n_pos,
                          uint32_t *max_occ)
{
    SeedTable *tab = (SeedTable *)t;
    *index = tab->index_table(); *index_entries = tab->index_entries();
    *pos = tab->pos_table(); *n_pos = tab->num_minimizers(); *max_occ = tab->kmer_max_occurence();
}

uint32_t dh_hash32(uint32_t key, int k) { return wang_hash32(key, k); }

int dh_read_fasta(const char *path, char *names_out, int names_cap, long long *lens_out, int max_seqs)
{
    FastaSet fs;
    std::string err;
    if (!read_fasta(path, &fs, &err)) return -1;
    std::string joined;
    for (size_t i = 0; i < fs.names.size(); i++) { joined += fs.names[i]; joined += '\n'; }
    strncpy(names_out, joined.c_str(), names_cap - 1);
    names_out[names_cap - 1] = 0;
    for (size_t i = 0; i < fs.seqs.size() && (int)i < max_seqs; i++) lens_out[i] = (long long)fs.seqs[i].size();
    return (int)fs.seqs.size();
}

int dh_params(const char *path, int *out16)
{
    try {
        Params p = Params::from_file(path);
        int v[16] = {p.match, p.mismatch, p.gap_open, p.gap_extend, p.seed_size, (int)p.bin_size, (int)p.window_size,
                     p.threshold, p.num_seeds, p.seed_occurence_multiple, p.max_candidates, p.num_nz_bins,
                     p.first_tile_size, p.first_tile_score_threshold, p.tile_size, p.tile_overlap};
        memcpy(out16, v, sizeof(v));
        return 0;
    } catch (...) { return -1; }
}

}
