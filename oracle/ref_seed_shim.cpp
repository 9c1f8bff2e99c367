// oracle/ref_seed_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points around the reference's own SeedPosTable / DSOFT
// (/root/reference/seed_pos_table.cpp:46-167, ntcoding.cpp), compiled in place
// by oracle/Makefile into oracle/_ref/libseed_ref.so.  Used to pin the host-side
// seeding code of darwin-gpu_b200/host/seed_table.cpp.
#include <cstdint>
#include <cstring>
#include <vector>
#define private public          // test-only view of the reference's tables; the class layout is unchanged
#include "seed_pos_table.h"
#undef private

extern "C" void *ref_seed_table_new(const char *ref, uint32_t ref_len, int k, uint32_t occ_mult, uint32_t bin_size, uint32_t w)
{
    return new SeedPosTable((char *)ref, ref_len, k, occ_mult, bin_size, w);
}

extern "C" int ref_dsoft(void *t, const char *q, uint32_t qlen, uint32_t ref_len, uint32_t bin_size,
                         int num_seeds, int threshold, int max_cand, int num_nz_bins, uint64_t *out, int cap)
{
    SeedPosTable *sa = (SeedPosTable *)t;
    // per-thread scratch exactly as AlignReads sets it up (darwin.cpp:183-199)
    uint32_t log_bin = (uint32_t)log2(bin_size);
    int num_bins = 1 + (ref_len >> log_bin);
    std::vector<uint64_t> bins(num_bins, 0), cand(max_cand);
    std::vector<uint32_t> nz(num_nz_bins);
    int n = sa->DSOFT((char *)q, qlen, num_seeds, threshold, cand.data(), bins.data(), nz.data(), max_cand);
    for (int i = 0; i < n && i < cap; i++) out[i] = cand[i];
    return n;
}

// index_table_ (4^k + 1 entries) and pos_table_ (index_table_[4^k] entries) as the reference built them
extern "C" void ref_seed_table_arrays(void *t, const uint32_t **index, uint64_t *index_entries, const uint32_t **pos,
                                      uint64_t *n_pos)
{
    SeedPosTable *sa = (SeedPosTable *)t;
    *index = sa->index_table_;
    *index_entries = sa->index_table_size_;
    *pos = sa->pos_table_;
    *n_pos = sa->index_table_[sa->index_table_size_ - 1];
}
