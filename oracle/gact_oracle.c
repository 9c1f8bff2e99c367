/*
 * oracle/gact_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's GACT tile aligner and
 * per-candidate extension loop.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file; the
 * product path (darwin-gpu_b200/csrc + host) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (a) the known-answer vectors produced from the unmodified reference
 *       AlignWithBT() (committed in tests/golden/), and
 *   (b) oracle/_ref/libalign_ref.so -- the reference's own align.cpp compiled
 *       in place from /root/reference (recipe: oracle/Makefile) -- on random
 *       tiles, whenever that library has been built.
 *
 * What is restated (reference file:line):
 *   oracle_align_tile      <- AlignWithBT()        align.cpp:60-233
 *   oracle_gact_extend     <- GACT()               gact.cpp:48-228
 * The restatement keeps the reference's exact recurrence (clamped M matrix,
 * -2^30 borders, >= tie rules, last-max-wins, early-terminate check before
 * the push) but stores the 4-bit direction codes in a flat byte array instead
 * of a 2050x2050 vector<vector<int>> (align.cpp:85), which is where the
 * reference spends most of its time.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_NEG_INF (-(1 << 30))          /* align.h:18  INF (1 << 30)      */
enum { ST_Z = 0, ST_D = 1, ST_I = 2, ST_M = 3 };   /* align.h:22-23           */

typedef struct {
    int32_t score;      /* first: max_score ; else H[ref_len][query_len]       */
    int32_t max_i;      /* traceback start row (ref index, 1-based)            */
    int32_t max_j;      /* traceback start col (query index, 1-based)          */
    int32_t n_states;   /* number of traceback states written                  */
    int32_t i_steps;    /* reference bases consumed by the traceback           */
    int32_t j_steps;    /* query bases consumed by the traceback               */
} oracle_tile_result;

/* scratch big enough for one tile; caller may pass NULL (malloc per call). */
typedef struct {
    int32_t *rows;      /* 8 * (max_len + 1) ints                              */
    uint8_t *dir;       /* (max_len + 1)^2 bytes                               */
    int      max_len;
} oracle_scratch;

oracle_scratch *oracle_scratch_new(int max_len)
{
    oracle_scratch *s = (oracle_scratch *)malloc(sizeof(*s));
    if (!s) return NULL;
    s->max_len = max_len;
    s->rows = (int32_t *)malloc(sizeof(int32_t) * 8 * (size_t)(max_len + 1));
    s->dir = (uint8_t *)malloc((size_t)(max_len + 1) * (size_t)(max_len + 1));
    if (!s->rows || !s->dir) { free(s->rows); free(s->dir); free(s); return NULL; }
    return s;
}

void oracle_scratch_free(oracle_scratch *s)
{
    if (!s) return;
    free(s->rows); free(s->dir); free(s);
}

/*
 * One tile.  Follows align.cpp:60-233 statement by statement.
 *   ref_seq/query_seq : raw bytes, compared with == (align.cpp:134)
 *   reverse           : CPU-build sense (align.cpp:130-131): 0 = natural order,
 *                       1 = both sequences read back to front
 *   first             : start the traceback at the last maximum (align.cpp:190)
 *   states            : out, up to 2*early_terminate entries, values 1..3
 * query_pos/ref_pos of the reference signature are always query_len/ref_len at
 * every call site (gact.cpp:93,155), so they are not parameters here.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int oracle_align_tile(const char *ref_seq, int ref_len,
                      const char *query_seq, int query_len,
                      int match_score, int mismatch_score,
                      int gap_open, int gap_extend,
                      int reverse, int first, int early_terminate,
                      oracle_scratch *scratch,
                      oracle_tile_result *res, uint8_t *states, int states_cap)
{
    if (ref_len < 0 || query_len < 0 || !res) return -1;
    int own = 0;
    int need = ref_len > query_len ? ref_len : query_len;
    if (!scratch || scratch->max_len < need) {
        scratch = oracle_scratch_new(need);
        if (!scratch) return -1;
        own = 1;
    }
    const int W = query_len + 1;                 /* dir row pitch            */
    int32_t *h_wr = scratch->rows + 0 * (scratch->max_len + 1);
    int32_t *m_wr = scratch->rows + 1 * (scratch->max_len + 1);
    int32_t *i_wr = scratch->rows + 2 * (scratch->max_len + 1);
    int32_t *d_wr = scratch->rows + 3 * (scratch->max_len + 1);
    int32_t *h_rd = scratch->rows + 4 * (scratch->max_len + 1);
    int32_t *m_rd = scratch->rows + 5 * (scratch->max_len + 1);
    int32_t *i_rd = scratch->rows + 6 * (scratch->max_len + 1);
    int32_t *d_rd = scratch->rows + 7 * (scratch->max_len + 1);
    uint8_t *dir = scratch->dir;

    /* align.cpp:87-97 */
    for (int j = 0; j <= query_len; j++) {
        h_rd[j] = 0; m_rd[j] = 0; i_rd[j] = ORACLE_NEG_INF; d_rd[j] = ORACLE_NEG_INF;
        h_wr[j] = 0; m_wr[j] = 0; i_wr[j] = ORACLE_NEG_INF; d_wr[j] = ORACLE_NEG_INF;
    }
    /* align.cpp:101-107 */
    for (int i = 0; i <= ref_len; i++) dir[(size_t)i * W] = ST_Z;
    for (int j = 0; j <= query_len; j++) dir[j] = ST_Z;

    int max_score = 0, pos_score = 0, max_i = 0, max_j = 0;

    for (int i = 1; i <= ref_len; i++) {
        /* align.cpp:115-120: previous row <- current row (index 0 keeps its
         * border value because the reference copies from k = 1) */
        for (int k = 1; k <= query_len; k++) {
            m_rd[k] = m_wr[k]; h_rd[k] = h_wr[k];
            i_rd[k] = i_wr[k]; d_rd[k] = d_wr[k];
        }
        /* align.cpp:130 */
        const char ref_nt = reverse ? ref_seq[ref_len - i] : ref_seq[i - 1];
        for (int j = 1; j <= query_len; j++) {
            /* align.cpp:131 */
            const char query_nt = reverse ? query_seq[query_len - j] : query_seq[j - 1];
            const int match = (query_nt == ref_nt) ? match_score : mismatch_score;   /* :134 */

            /* align.cpp:138-147 */
            int m;
            if (m_rd[j - 1] > i_rd[j - 1] && m_rd[j - 1] > d_rd[j - 1]) m = m_rd[j - 1] + match;
            else if (i_rd[j - 1] > d_rd[j - 1])                         m = i_rd[j - 1] + match;
            else                                                        m = d_rd[j - 1] + match;
            if (m < 0) m = 0;
            m_wr[j] = m;

            /* align.cpp:149-156 */
            const int ins_open   = m_rd[j] + gap_open;
            const int ins_extend = i_rd[j] + gap_extend;
            const int del_open   = m_wr[j - 1] + gap_open;
            const int del_extend = d_wr[j - 1] + gap_extend;
            const int iv = (ins_open > ins_extend) ? ins_open : ins_extend;
            const int dv = (del_open > del_extend) ? del_open : del_extend;
            i_wr[j] = iv;
            d_wr[j] = dv;

            /* align.cpp:158-160 */
            const int max1 = m > iv ? m : iv;
            const int max2 = dv > 0 ? dv : 0;
            const int h = max1 > max2 ? max1 : max2;
            h_wr[j] = h;

            /* align.cpp:162-171 */
            int d = (m >= iv) ? ((m >= dv) ? ST_M : ST_D) : ((iv >= dv) ? ST_I : ST_D);
            if (m <= 0 && iv <= 0 && dv <= 0) d = ST_Z;
            d += (ins_open >= ins_extend) ? 8 : 0;      /* 2 << INSERT_OP */
            d += (del_open >= del_extend) ? 4 : 0;      /* 2 << DELETE_OP */
            dir[(size_t)i * W + j] = (uint8_t)d;

            /* align.cpp:173-177: >= means the LAST maximum wins */
            if (h >= max_score) { max_score = h; max_i = i; max_j = j; }
            /* align.cpp:179-181 with ref_pos = ref_len, query_pos = query_len */
            if (i == ref_len && j == query_len) pos_score = h;
        }
    }

    /* align.cpp:185-199 */
    int i_curr = ref_len, j_curr = query_len;
    int i_steps = 0, j_steps = 0, n = 0;
    if (first) { i_curr = max_i; j_curr = max_j; res->score = max_score; }
    else       { res->score = pos_score; }
    res->max_i = i_curr;
    res->max_j = j_curr;

    /* align.cpp:201-230 */
    int state = dir[(size_t)i_curr * W + j_curr] % 4;
    int rc = 0;
    while (state != ST_Z) {
        if (i_steps >= early_terminate || j_steps >= early_terminate) break;
        /* The reference would index dir[-1] if a gap state reached row or
         * column 0 and flipped to M; that needs a positive gap score and
         * cannot happen with gap_open, gap_extend <= 0 (the state at a border
         * cell is always Z then).  Stop instead of reading out of bounds. */
        if (i_curr <= 0 || j_curr <= 0) break;
        if (n >= states_cap) { rc = -1; break; }
        states[n++] = (uint8_t)state;
        if (state == ST_M) {
            state = dir[(size_t)(i_curr - 1) * W + (j_curr - 1)] % 4;
            i_curr--; j_curr--; i_steps++; j_steps++;
        } else if (state == ST_I) {
            state = (dir[(size_t)i_curr * W + j_curr] & 8) ? ST_M : ST_I;
            i_curr--; i_steps++;
        } else { /* ST_D */
            state = (dir[(size_t)i_curr * W + j_curr] & 4) ? ST_M : ST_D;
            j_curr--; j_steps++;
        }
    }
    res->n_states = n;
    res->i_steps = i_steps;
    res->j_steps = j_steps;
    if (own) oracle_scratch_free(scratch);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* Batch helper (OpenMP when compiled with -fopenmp): used for the parity     */
/* checks at scale and as the "port" CPU baseline of bench.py.                */
typedef struct {
    int64_t ref_off;    /* offset of the tile's first base in ref buffer      */
    int64_t query_off;
    int32_t ref_len;
    int32_t query_len;
    int32_t reverse;    /* CPU-build sense                                    */
    int32_t first;
} oracle_tile_desc;

int oracle_align_batch(const char *ref_buf, const char *query_buf,
                       const oracle_tile_desc *descs, int n_tiles,
                       int match_score, int mismatch_score, int gap_open, int gap_extend,
                       int early_terminate, int max_len, int n_threads,
                       oracle_tile_result *results, uint8_t *states, int states_pitch)
{
    int bad = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
    {
        oracle_scratch *s = oracle_scratch_new(max_len);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int t = 0; t < n_tiles; t++) {
            const oracle_tile_desc *d = &descs[t];
            if (!s || oracle_align_tile(ref_buf + d->ref_off, d->ref_len,
                                        query_buf + d->query_off, d->query_len,
                                        match_score, mismatch_score, gap_open, gap_extend,
                                        d->reverse, d->first, early_terminate, s,
                                        &results[t], states + (size_t)t * states_pitch,
                                        states_pitch) != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                bad = 1;
            }
        }
        oracle_scratch_free(s);
    }
    (void)n_threads;
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------- */
/* Per-candidate extension, gact.cpp:48-228.                                  */
typedef struct {
    int32_t ab, ae, bb, be;   /* ref begin/end, query begin/end (gact.cpp:219-222) */
    int32_t score;            /* recomputed total score (gact.cpp:197-210)         */
    int32_t first_tile_score;
    int32_t n_tiles;          /* tiles aligned for this candidate                  */
    int64_t n_cells;          /* sum of ref_len*query_len over those tiles         */
    int32_t n_columns;        /* length of the aligned strings                     */
} oracle_gact_result;

/* optional tile log: every tile the extension aligned, in call order */
typedef struct {
    int32_t ref_start, query_start, ref_len, query_len, reverse, first;
} oracle_tile_log;

typedef struct { char *p; size_t n, cap; } sbuf;
static int sbuf_push(sbuf *b, char c)
{
    if (b->n == b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 1024;
        char *np = (char *)realloc(b->p, nc);
        if (!np) return -1;
        b->p = np; b->cap = nc;
    }
    b->p[b->n++] = c;
    return 0;
}

int oracle_gact_extend(const char *ref_str, int ref_length,
                       const char *query_str, int query_length,
                       int tile_size, int tile_overlap,
                       int ref_pos, int query_pos, int first_tile_score_threshold,
                       int match_score, int mismatch_score, int gap_open, int gap_extend,
                       oracle_gact_result *out,
                       oracle_tile_log *tile_log, int tile_log_cap)
{
    const int et = tile_size - tile_overlap;
    oracle_scratch *sc = oracle_scratch_new(tile_size);
    uint8_t *states = (uint8_t *)malloc((size_t)2 * tile_size + 8);
    /* left part is produced anchor-outwards (the reference prepends,
     * gact.cpp:116-128); keep it reversed and flip at the end */
    sbuf lr = {0, 0, 0}, lq = {0, 0, 0}, rr = {0, 0, 0}, rq = {0, 0, 0};
    if (!sc || !states) { oracle_scratch_free(sc); free(states); return -1; }
    int rc = 0;
    int n_tiles = 0; int64_t n_cells = 0;

    int rev_ref_pos = ref_pos, rev_query_pos = query_pos;       /* gact.cpp:72-73 */
    int i = 0, j = 0;
    int first_tile_score = 0;
    int first_tile = 1;
    oracle_tile_result tr;

    /* gact.cpp:82-134 : extension towards position 0 */
    while (ref_pos > 0 && query_pos > 0 && ((i > 0 && j > 0) || first_tile)) {
        int rtl = ref_pos > tile_size ? tile_size : ref_pos;
        int qtl = query_pos > tile_size ? tile_size : query_pos;
        if (tile_log && n_tiles < tile_log_cap) {
            oracle_tile_log *l = &tile_log[n_tiles];
            l->ref_start = ref_pos - rtl; l->query_start = query_pos - qtl;
            l->ref_len = rtl; l->query_len = qtl; l->reverse = 0; l->first = first_tile;
        }
        if (oracle_align_tile(ref_str + ref_pos - rtl, rtl, query_str + query_pos - qtl, qtl,
                              match_score, mismatch_score, gap_open, gap_extend,
                              0, first_tile, et, sc, &tr, states, 2 * tile_size + 8) != 0) { rc = -1; goto done; }
        n_tiles++; n_cells += (int64_t)rtl * qtl;
        i = 0; j = 0;
        if (first_tile) {
            ref_pos = ref_pos - rtl + tr.max_i;
            query_pos = query_pos - qtl + tr.max_j;
            rev_ref_pos = ref_pos; rev_query_pos = query_pos;
            first_tile_score = tr.score;
            if (tr.score < first_tile_score_threshold) break;
        }
        for (int s = 0; s < tr.n_states; s++) {
            first_tile = 0;
            int st = states[s];
            if (st == ST_M) {
                rc |= sbuf_push(&lr, ref_str[ref_pos - j - 1]);
                rc |= sbuf_push(&lq, query_str[query_pos - i - 1]);
                i++; j++;
            } else if (st == ST_I) {
                rc |= sbuf_push(&lr, ref_str[ref_pos - j - 1]);
                rc |= sbuf_push(&lq, '-');
                j++;
            } else if (st == ST_D) {
                rc |= sbuf_push(&lr, '-');
                rc |= sbuf_push(&lq, query_str[query_pos - i - 1]);
                i++;
            }
        }
        ref_pos -= j; query_pos -= i;
        if (first_tile && tr.n_states == 0 && tr.score >= first_tile_score_threshold) {
            /* the reference would spin forever here (only reachable with a
             * non-positive threshold); stop instead */
            break;
        }
    }

    int abpos = ref_pos, bbpos = query_pos;                     /* gact.cpp:136-141 */
    ref_pos = rev_ref_pos; query_pos = rev_query_pos;
    i = tile_size; j = tile_size;

    /* gact.cpp:144-195 : extension towards the sequence ends */
    while (ref_pos < ref_length && query_pos < query_length && ((i > 0 && j > 0) || first_tile)) {
        int rtl = (ref_pos + tile_size < ref_length) ? tile_size : ref_length - ref_pos;
        int qtl = (query_pos + tile_size < query_length) ? tile_size : query_length - query_pos;
        if (tile_log && n_tiles < tile_log_cap) {
            oracle_tile_log *l = &tile_log[n_tiles];
            l->ref_start = ref_pos; l->query_start = query_pos;
            l->ref_len = rtl; l->query_len = qtl; l->reverse = 1; l->first = first_tile;
        }
        if (oracle_align_tile(ref_str + ref_pos, rtl, query_str + query_pos, qtl,
                              match_score, mismatch_score, gap_open, gap_extend,
                              1, first_tile, et, sc, &tr, states, 2 * tile_size + 8) != 0) { rc = -1; goto done; }
        n_tiles++; n_cells += (int64_t)rtl * qtl;
        i = 0; j = 0;
        if (first_tile) {
            ref_pos = ref_pos + rtl - tr.max_i;
            query_pos = query_pos + qtl - tr.max_j;
            first_tile_score = tr.score;
            if (tr.score < first_tile_score_threshold) break;
        }
        for (int s = 0; s < tr.n_states; s++) {
            first_tile = 0;
            int st = states[s];
            if (st == ST_M) {
                rc |= sbuf_push(&rr, ref_str[ref_pos + j]);
                rc |= sbuf_push(&rq, query_str[query_pos + i]);
                i++; j++;
            } else if (st == ST_I) {
                rc |= sbuf_push(&rr, ref_str[ref_pos + j]);
                rc |= sbuf_push(&rq, '-');
                j++;
            } else if (st == ST_D) {
                rc |= sbuf_push(&rr, '-');
                rc |= sbuf_push(&rq, query_str[query_pos + i]);
                i++;
            }
        }
        ref_pos += j; query_pos += i;
        if (first_tile && tr.n_states == 0 && tr.score >= first_tile_score_threshold) break;
    }

    /* gact.cpp:197-210 : total score over the concatenated columns */
    {
        int total = 0; int open = 1;
        size_t ncol = lr.n + rr.n;
        for (size_t c = 0; c < ncol; c++) {
            char rn, qn;
            if (c < lr.n) { rn = lr.p[lr.n - 1 - c]; qn = lq.p[lq.n - 1 - c]; }
            else          { rn = rr.p[c - lr.n];     qn = rq.p[c - lr.n]; }
            if (rn == '-' || qn == '-') { total += open ? gap_open : gap_extend; open = 0; }
            else { total += (qn == rn) ? match_score : mismatch_score; open = 1; }
        }
        out->ab = abpos; out->ae = ref_pos; out->bb = bbpos; out->be = query_pos;
        out->score = total; out->first_tile_score = first_tile_score;
        out->n_tiles = n_tiles; out->n_cells = n_cells; out->n_columns = (int32_t)ncol;
    }
done:
    free(lr.p); free(lq.p); free(rr.p); free(rq.p);
    free(states); oracle_scratch_free(sc);
    return rc ? -1 : 0;
}
