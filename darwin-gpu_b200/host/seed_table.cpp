#include "seed_table.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace darwin {

static inline uint32_t nt_code(char c)
{
    switch (c) {
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 2;
        case 't': case 'T': return 3;
        default: return 0;                       // A, a and everything else
    }
}

void pack_two_bit(const char *seq, uint32_t len, uint32_t *out, uint32_t n_words)
{
    memset(out, 0, (size_t)n_words * sizeof(uint32_t));
    for (uint32_t i = 0; i < len; i++) out[i >> 4] |= nt_code(seq[i]) << (2 * (i & 15));
}

uint32_t wang_hash32(uint32_t key, int k)
{
    const uint32_t m = (1u << (2 * k)) - 1;
    key = (~key + (key << 21)) & m;
    key = key ^ (key >> 24);
    key = ((key + (key << 3)) + (key << 8)) & m;
    key = key ^ (key >> 14);
    key = ((key + (key << 2)) + (key << 4)) & m;
    key = key ^ (key >> 28);
    key = (key + (key << 31)) & m;
    return key;
}

static inline uint32_t seed_at(const uint32_t *w, uint32_t pos, uint32_t mask)
{
    const uint32_t idx = pos >> 4, sh = 2 * (pos & 15);
    const uint64_t cat = ((uint64_t)w[idx + 1] << 32) | w[idx];
    return (uint32_t)(cat >> sh) & mask;
}

// Shared minimizer scan (ntcoding.cpp:126-182).  emit(m, p) is called with the window
// minimum and the loop position p whenever the reference would append an entry.
// `last_m`/`last_p` start at 0 exactly like the reference's locals.
template <class Emit>
static void scan_minimizers(const uint32_t *two_bit, uint32_t s_len, int k, int w, Emit emit)
{
    const uint32_t mask = (1u << (2 * k)) - 1;
    const uint32_t end = 16u * s_len - (uint32_t)k - (uint32_t)w;     // unsigned, as in the reference
    if (16u * s_len < (uint32_t)(k + w)) return;                      // the reference would run off the array here
    uint32_t win[32];
    for (int p = 0; p < w; p++) win[p] = 0;
    for (int p = 0; p < w - 1; p++) win[p] = wang_hash32(seed_at(two_bit, (uint32_t)p, mask), k);
    uint64_t last_m = 0;
    uint32_t last_p = 0;
    for (uint32_t p = (uint32_t)w - 1; p < end; p++) {
        win[p % (uint32_t)w] = wang_hash32(seed_at(two_bit, p, mask), k);
        uint32_t m = 0xffffffffu;
        for (int x = 0; x < w; x++) m = std::min(m, win[x]);
        if (m != last_m || p - last_p >= (uint32_t)w) {
            emit(m, p);
            last_m = m;
            last_p = p;
        }
    }
}

SeedTable::SeedTable(const char *ref, uint32_t ref_len, int kmer_size, uint32_t seed_occurence_multiple,
                     uint32_t bin_size, uint32_t window_size, int build_threads)
    : ref_len_(ref_len), bin_size_(bin_size), k_(kmer_size), w_((int)window_size)
{
    if (!(kmer_size <= 15 && kmer_size > 3 && kmer_size > (int)window_size) || window_size == 0 || window_size > 32)
        throw std::runtime_error("seed_size/window_size out of range (3 < k <= 15, w < k)");   // seed_pos_table.cpp:48-50
    log_bin_size_ = (uint32_t)log2((double)bin_size);
    kmer_max_occurence_ = seed_occurence_multiple * (1 + (ref_len >> (2 * kmer_size)));

    const uint32_t n_words = 1 + ref_len / 16;
    std::vector<uint32_t> two_bit(n_words + 1);
    pack_two_bit(ref, ref_len, two_bit.data(), n_words + 1);

    // The scan carries (last_m, last_p) state, so it is sequential; it is a small part of the
    // build next to the sort and the 4^k-entry index fill.
    std::vector<uint64_t> mins;
    mins.reserve((size_t)ref_len / 2 + 16);
    scan_minimizers(two_bit.data(), n_words, k_, w_, [&](uint32_t m, uint32_t p) { mins.push_back(((uint64_t)m << 32) | p); });

    // sort by (hash, position); keys are unique, any correct sort gives the reference's order
    const int T = std::max(1, build_threads);
    if (T > 1 && mins.size() > (1u << 16)) {
        const size_t n = mins.size();
        std::vector<size_t> cut(T + 1);
        for (int t = 0; t <= T; t++) cut[t] = n * (size_t)t / T;
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t] { std::sort(mins.begin() + cut[t], mins.begin() + cut[t + 1]); });
        for (auto &x : th) x.join();
        for (int width = 1; width < T; width *= 2) {
            th.clear();
            for (int t = 0; t + width < T; t += 2 * width) {
                const size_t a = cut[t], b = cut[t + width], c = cut[std::min(T, t + 2 * width)];
                th.emplace_back([&, a, b, c] { std::inplace_merge(mins.begin() + a, mins.begin() + b, mins.begin() + c); });
            }
            for (auto &x : th) x.join();
        }
    } else {
        std::sort(mins.begin(), mins.end());
    }

    n_pos_ = (uint32_t)mins.size();
    const uint32_t index_size = (1u << (2 * kmer_size)) + 1;
    index_table_ = (uint32_t *)malloc((size_t)index_size * sizeof(uint32_t));
    pos_table_ = (uint32_t *)malloc(std::max<size_t>(1, n_pos_) * sizeof(uint32_t));
    if (!index_table_ || !pos_table_) throw std::bad_alloc();
    // index_table_[s] = number of minimizers with hash <= s  (seed_pos_table.cpp:79-93)
    uint32_t cur = 0;
    for (uint32_t i = 0; i < n_pos_; i++) {
        const uint32_t seed = (uint32_t)(mins[i] >> 32);
        pos_table_[i] = (uint32_t)mins[i];
        while (cur < seed) index_table_[cur++] = i;
    }
    while (cur < index_size) index_table_[cur++] = n_pos_;
}

SeedTable::~SeedTable()
{
    free(index_table_);
    free(pos_table_);
}

SeedTable::Scratch::Scratch(const SeedTable &t, int num_nz_bins)
    : bin_count_offset(t.num_bins(), 0), nz_bins((size_t)std::max(num_nz_bins, 1)) {}

int SeedTable::dsoft(const char *query, uint32_t query_len, int N, int threshold, int max_candidates,
                     Scratch &s, std::vector<uint64_t> &candidates) const
{
    const size_t first = candidates.size();
    const uint32_t q_words = (query_len + 15) / 16;          // seed_pos_table.cpp:109
    const uint32_t alloc_words = 1 + query_len / 16;          // what SeqToTwoBit allocates (ntcoding.cpp:88)
    s.q2bit.resize((size_t)std::max(q_words, alloc_words) + 2);
    pack_two_bit(query, query_len, s.q2bit.data(), (uint32_t)s.q2bit.size());
    s.minimizers.clear();
    scan_minimizers(s.q2bit.data(), q_words, k_, w_, [&](uint32_t m, uint32_t p) { s.minimizers.push_back(((uint64_t)p << 32) | m); });

    uint64_t *bins = s.bin_count_offset.data();
    uint32_t *nz = s.nz_bins.data();
    const uint64_t nz_cap = std::min<uint64_t>(25000000ull, s.nz_bins.size());   // macro nz_bins; bounded by the array we own
    uint64_t n_nz = 0;
    int seeds = 0, n_cand = 0;
    const uint32_t thr = (uint32_t)threshold, kk = (uint32_t)k_;

    for (size_t x = 0; x < s.minimizers.size(); x++) {
        const uint32_t offset = (uint32_t)(s.minimizers[x] >> 32);
        const uint32_t index = (uint32_t)s.minimizers[x];
        const uint32_t b = index > 0 ? index_table_[index - 1] : 0, e = index_table_[index];
        if (e - b > kmer_max_occurence_) continue;
        if (seeds > N) break;                                  // N+1 seeds in total (seed_pos_table.cpp:128-131)
        seeds++;
        for (uint32_t j = b; j < e; j++) {
            const uint32_t hit = pos_table_[j];
            if (hit < offset) continue;
            const uint32_t bin = (hit - offset) / bin_size_;
            const uint32_t count = (uint32_t)(bins[bin] >> 32), last = (uint32_t)bins[bin];
            if (count >= thr) continue;
            const uint32_t nc = ((offset - last > kk) || count == 0) ? count + kk : count + (offset - last);
            bins[bin] = ((uint64_t)nc << 32) + offset;
            if (nc >= thr) {
                if (n_cand >= max_candidates) break;           // leaves only the inner loop, like the reference
                candidates.push_back(((uint64_t)hit << 32) + offset);
                n_cand++;
            }
            if (count == 0 && n_nz < nz_cap) nz[n_nz++] = bin;
        }
    }
    for (uint64_t x = 0; x < n_nz; x++) bins[nz[x]] = 0;
    (void)first;
    return n_cand;
}

}  // namespace darwin
