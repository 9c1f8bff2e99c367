// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry point around the reference's own, unmodified AlignWithBT()
// (/root/reference/align.cpp:60-233).  This file is ours; the reference
// sources are compiled where they lie (see oracle/Makefile) and only the
// resulting library lands in oracle/_ref/.  It is used to validate the C
// restatement in gact_oracle.c and, in bench.py, as the "reference" CPU arm.
#include <cassert>
#include <cstdint>
#include <queue>
#include <string>
#include <vector>
#include "align.h"

extern "C" int ref_align_with_bt(const char *ref_seq, int ref_len,
                                 const char *query_seq, int query_len,
                                 int match_score, int mismatch_score,
                                 int gap_open, int gap_extend,
                                 int reverse, int first, int early_terminate,
                                 int32_t *out, int out_cap)
{
    // same argument convention as every call site in gact.cpp:87-94,149-156:
    // query_pos = query_len, ref_pos = ref_len
    std::queue<int> q = AlignWithBT((char *)ref_seq, ref_len, (char *)query_seq, query_len,
                                    match_score, mismatch_score, gap_open, gap_extend,
                                    query_len, ref_len, reverse != 0, first != 0, early_terminate);
    int n = 0;
    while (!q.empty()) {
        if (n < out_cap) out[n] = q.front();
        q.pop();
        n++;
    }
    return n;
}

// Batch form for the CPU-baseline arm: every tile through AlignWithBT, OpenMP
// over tiles (the reference itself parallelises over reads with std::thread,
// darwin.cpp:619-629; tiles of different reads are equally independent).
struct ref_tile_desc { int64_t ref_off, query_off; int32_t ref_len, query_len, reverse, first; };

extern "C" long long ref_align_batch(const char *ref_buf, const char *query_buf,
                                     const ref_tile_desc *descs, int n_tiles,
                                     int match_score, int mismatch_score,
                                     int gap_open, int gap_extend,
                                     int early_terminate, int n_threads, int32_t *scores)
{
    long long cells = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : cells)
    for (int t = 0; t < n_tiles; t++) {
        const ref_tile_desc &d = descs[t];
        std::queue<int> q = AlignWithBT((char *)ref_buf + d.ref_off, d.ref_len,
                                        (char *)query_buf + d.query_off, d.query_len,
                                        match_score, mismatch_score, gap_open, gap_extend,
                                        d.query_len, d.ref_len, d.reverse != 0, d.first != 0,
                                        early_terminate);
        if (scores) scores[t] = q.front();
        cells += (long long)d.ref_len * d.query_len;
    }
    return cells;
}
