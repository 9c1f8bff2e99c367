// dsoft.cuh -- D-SOFT candidate filter on the GPU ("next" row 1).
//
// Device restatement of the reference's query side of seeding:
//   QTwoBitToMinimizers()   ntcoding.cpp:155-182   (window minimum of Wang-hashed k-mers,
//                                                   loop-position reporting, re-emission every w)
//   SeedPosTable::DSOFT()   seed_pos_table.cpp:100-167 (bucket lookup, N+1 seeds, per-bin
//                                                   (count, last_offset) update, threshold crossing)
// The seed-position table itself (index_table_ / pos_table_) is built on the host
// (host/seed_table.cpp, pinned against the reference) and uploaded once.
//
// One warp per query strand.  Positions are hashed 32 at a time; the window minimum, the
// "minimizer changed / w positions since the last emission" rule and the seed budget are
// evaluated with shuffles, ballots and prefix counts; hits of one seed are fetched with one
// coalesced load.  The per-bin state lives in a per-warp open-addressing table in global
// memory (L2-resident), touched slots are remembered and cleared after the query -- the
// reference's bin_count_offset_array + nz_bins_array (darwin.cpp:183-199) without the
// num_bins-sized dense array per thread.  Only as many positions are scanned as the seed budget
// needs (the reference computes every minimizer of the read first).
#pragma once
#include "gact_common.cuh"

namespace gact {

struct DsoftParams {
    const uint32_t *index_table;     // 4^k + 1 entries
    const uint32_t *pos_table;
    int k, w;
    uint32_t bin_size;
    uint32_t max_occ;                // kmer_max_occurence_
    int num_seeds;                   // N
    int threshold;
    int max_candidates;
    uint32_t table_cap;              // slots per warp (power of two)
    SeqSetDev sets[GACT_MAX_SETS];
};

struct DsoftQuery { long long start; int len; int set; };

__device__ __forceinline__ uint32_t dsoft_hash32(uint32_t key, uint32_t m)
{
    key = (~key + (key << 21)) & m;
    key = key ^ (key >> 24);
    key = ((key + (key << 3)) + (key << 8)) & m;
    key = key ^ (key >> 14);
    key = ((key + (key << 2)) + (key << 4)) & m;
    key = key ^ (key >> 28);
    key = (key + (key << 31)) & m;
    return key;
}

// 2-bit code of base `idx` of a set (ntcoding.cpp:60-72: A/a 0, C/c 1, G/g 2, T/t 3, other 0)
__device__ __forceinline__ uint32_t dsoft_code(const SeqSetDev &s, long long idx)
{
    if (s.packed) return (__ldg(s.packed + (idx >> 4)) >> (2 * (int)(idx & 15))) & 3u;
    const int c = __ldg(s.bytes + idx) | 0x20;           // fold case
    return c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 0u;
}

// k-mer at query position p (GetSeedAtPos, ntcoding.cpp:115-124); bases past the read are 0
__device__ __forceinline__ uint32_t dsoft_kmer(const SeqSetDev &s, long long start, int len, int p, int k)
{
    uint32_t v = 0;
    if (s.packed) {
        const long long b = start + p;
        const long long wi = b >> 4;
        const int sh = 2 * (int)(b & 15);
        const uint64_t cat = ((uint64_t)__ldg(s.packed + wi + 1) << 32) | __ldg(s.packed + wi);
        v = (uint32_t)(cat >> sh);
        const int valid = min(max(len - p, 0), k);
        v &= (valid >= 16) ? 0xffffffffu : ((1u << (2 * valid)) - 1u);
        v &= (1u << (2 * k)) - 1u;
    } else {
        for (int x = 0; x < k; x++)
            if (p + x < len) v |= dsoft_code(s, start + p + x) << (2 * x);
    }
    return v;
}

// out record: query index, sequence number inside the query, hit (reference position), offset (query position)
struct DsoftCand { int query; int seq; uint32_t hit; uint32_t offset; };

__global__ void __launch_bounds__(128)
dsoft_kernel(const __grid_constant__ DsoftParams P, const DsoftQuery *__restrict__ queries, int n_queries,
             uint32_t *tab_keys, unsigned long long *tab_vals, uint32_t *touched,
             DsoftCand *out, unsigned long long out_cap, unsigned long long *out_count, int *counter)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t *keys = tab_keys + (size_t)gwarp * P.table_cap;          // 0 = empty, else bin + 1
    unsigned long long *vals = tab_vals + (size_t)gwarp * P.table_cap;   // (count << 32) | last_offset
    uint32_t *tch = touched + (size_t)gwarp * P.table_cap;
    const uint32_t capm = P.table_cap - 1;
    const int k = P.k, w = P.w;
    const uint32_t hmask = (1u << (2 * k)) - 1u;

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(counter, 1);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_queries) break;
        const DsoftQuery q = queries[qi];
        const SeqSetDev &set = P.sets[q.set];
        const uint32_t s_len = ((uint32_t)q.len + 15u) / 16u;
        if (16u * s_len < (uint32_t)(k + w)) continue;              // the reference would run off its arrays here
        const int end = (int)(16u * s_len - (uint32_t)k - (uint32_t)w);   // positions w-1 .. end-1

        int seeds = 0, n_cand = 0, n_touched = 0;
        bool done = false;
        // carries between chunks of 32 positions
        uint32_t carry_h = 0;                // lane L holds hash of position (chunk_base - 32 + L) from the previous chunk
        uint32_t carry_m = 0;                // m of the last position of the previous chunk (virtual 0 before the first)
        int carry_s = 0;                     // run start at the end of the previous chunk (virtual run starts at 0)

        // pre-fill: hashes of positions 0 .. w-2 belong to the first window; treat chunk_base so that
        // position w-1 is lane 0's: chunk covers positions [base, base+32), first base = w-1, and the
        // "previous chunk" supplies positions base-32 .. base-1 (only base-(w-1) .. base-1 are read)
        {
            const int p = (w - 1) - 32 + lane;
            carry_h = (p >= 0) ? dsoft_hash32(dsoft_kmer(set, q.start, q.len, p, k), hmask) : 0u;
        }
        for (int base = w - 1; base < end && !done; base += 32) {
            const int p = base + lane;
            const bool inr = p < end;
            const uint32_t h = inr ? dsoft_hash32(dsoft_kmer(set, q.start, q.len, p, k), hmask) : 0xffffffffu;
            // window minimum over h[p-w+1 .. p]
            uint32_t m = h;
            for (int x = 1; x < w; x++) {
                uint32_t o = __shfl_up_sync(FULL, h, x);
                const uint32_t oc = __shfl_sync(FULL, carry_h, (32 - x + lane) & 31);   // previous chunk's lane 32-x+lane
                if (lane < x) o = oc;
                m = min(m, o);
            }
            // run structure: change_p = (m_p != m_{p-1})
            uint32_t mprev = __shfl_up_sync(FULL, m, 1);
            if (lane == 0) mprev = carry_m;
            const bool change = inr && (m != mprev);
            int s = change ? p : -1;                      // prefix max of run starts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, s, o);
                if (lane >= o) s = max(s, t);
            }
            if (s < 0) s = carry_s;
            const bool emit = inr && (change || ((p - s) % w == 0));
            // carries for the next chunk (from the last in-range lane; chunks are full except the last)
            carry_h = h;
            carry_m = __shfl_sync(FULL, m, 31);
            carry_s = __shfl_sync(FULL, s, 31);

            // bucket bounds of emitted minimizers
            uint32_t b0 = 0, b1 = 0;
            if (emit) {
                b0 = m > 0 ? __ldg(P.index_table + m - 1) : 0u;
                b1 = __ldg(P.index_table + m);
            }
            const bool ok = emit && (b1 - b0 <= P.max_occ);
            // most buckets hold one hit: fetch every bucket's first hit now, all lanes at once, so that the
            // seed-after-seed loop below does not wait for one more dependent load per seed
            const uint32_t hit0 = (ok && b1 > b0) ? __ldg(P.pos_table + b0) : 0u;
            const unsigned okmask = __ballot_sync(FULL, ok);
            const int rank = __popc(okmask & ((1u << lane) - 1u));
            // seed budget: the j-th usable minimizer (0-based, over the whole query) is used iff j <= N;
            // the first one beyond that ends the query (seed_pos_table.cpp:128-131)
            const bool used = ok && (seeds + rank <= P.num_seeds);
            unsigned usedmask = __ballot_sync(FULL, used);
            if (seeds + __popc(okmask) > P.num_seeds) done = true;
            seeds += __popc(okmask);

            // hits of the used minimizers, in query order
            while (usedmask) {
                const int src = __ffs(usedmask) - 1;
                usedmask &= usedmask - 1;
                const uint32_t hb = __shfl_sync(FULL, b0, src), he = __shfl_sync(FULL, b1, src);
                const uint32_t offset = (uint32_t)(base + src);
                bool stop_bucket = false;
                const uint32_t first_hit = __shfl_sync(FULL, hit0, src);
                for (uint32_t j0 = hb; j0 < he && !stop_bucket; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    uint32_t hit = 0u;
                    if (he - hb == 1) hit = first_hit;                    // (only lane 0's copy is read below)
                    else if (j < he) hit = __ldg(P.pos_table + j);
                    const int cnt = min(32u, he - j0);
                    for (int x = 0; x < cnt; x++) {
                        const uint32_t ht = __shfl_sync(FULL, hit, x);
                        if (ht < offset) continue;
                        // lane 0 owns the per-bin table (sequential semantics); the verdict is broadcast
                        int verdict = 0;                         // bit 0: new slot, bit 1: candidate, bit 2: stop bucket
                        if (lane == 0) {
                            const uint32_t bin = (ht - offset) / P.bin_size;
                            uint32_t slot = (bin * 2654435761u) & capm;
                            uint32_t key;
                            for (;;) {
                                key = keys[slot];
                                if (key == bin + 1u || key == 0u) break;
                                slot = (slot + 1) & capm;
                            }
                            const unsigned long long val = (key == 0u) ? 0ull : vals[slot];
                            const uint32_t count = (uint32_t)(val >> 32), last = (uint32_t)val;
                            if (count < (uint32_t)P.threshold) {
                                const uint32_t nc = ((offset - last > (uint32_t)k) || count == 0u) ? count + (uint32_t)k
                                                                                                   : count + (offset - last);
                                if (key == 0u) { keys[slot] = bin + 1u; tch[n_touched] = slot; verdict |= 1; }
                                vals[slot] = ((unsigned long long)nc << 32) + offset;
                                if (nc >= (uint32_t)P.threshold) {
                                    if (n_cand >= P.max_candidates) verdict |= 4;          // inner break of the reference
                                    else {
                                        const unsigned long long o = atomicAdd(out_count, 1ull);
                                        if (o < out_cap) out[o] = DsoftCand{qi, n_cand, ht, offset};
                                        verdict |= 2;
                                    }
                                }
                            }
                        }
                        verdict = __shfl_sync(FULL, verdict, 0);
                        n_touched += verdict & 1;
                        n_cand += (verdict >> 1) & 1;
                        if (verdict & 4) { stop_bucket = true; break; }
                    }
                }
            }
        }
        // clear the touched slots (the reference's nz_bins reset, seed_pos_table.cpp:160-163)
        __syncwarp();
        for (int x = lane; x < n_touched; x += 32) { const uint32_t sl = tch[x]; keys[sl] = 0u; }
        __syncwarp();
    }
}

}  // namespace gact
