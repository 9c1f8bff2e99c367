// gact_kernels_it.cuh -- "inter-task" mapping of the packed s16x2 GACT tile kernel: one LANE per pair of tiles.
//
// The wavefront kernels (gact_kernels_s16h.cuh) spread one tile over 16 or 32 lanes.  Their lanes sit on different
// anti-diagonals, so whenever ANY lane is inside the traceback window the whole warp has to run the tagged
// (pointer-producing) loop body: 63 % of a full 320 x 320 tile's steps, although a traceback only ever visits a narrow
// band around the tile's main diagonal.  Here a lane owns two whole tiles (one in the low, one in the high half-word of
// every register) and a warp 64 tiles of IDENTICAL shape, so every lane is at the same cell (i, j) of its own tiles at
// the same time: band membership is warp-uniform, the tagged body runs only for the cells of a band of half-width W
// around the diagonal inside the traceback window (14 % of the row-steps at W = 32), there is no wavefront skew, no
// shuffle and no per-step edge bookkeeping.  Price: a lane sweeps its tiles strip by strip (CS columns in registers),
// and the strip's right edge (H and D of every row) goes through a per-warp scratch array in global memory to the next
// strip -- 16 bytes per row and lane, coalesced, about 1 byte per cell pair.  The reference rows' score tables cost no
// memory traffic: the tile's reference is re-packed into DP order in shared memory (2 bits per row, 16 rows per word, the
// same alignment for every lane) and the 4-byte table of a row is one PRMT in "backward 4 extract" mode on the
// constant (match, mismatch, mismatch, mismatch), selected by the two low bits of the shifted word.
//
// Eligible tiles (the engine routes them, check_descs): ref_len == query_len == tile_size, not a first tile, query
// window and reference set free of exceptions (2-bit codes then say everything about a base), tile_size a multiple of CS.  A traceback that leaves the band is
// detected (the lane knows the band geometry) and the tile is appended to an escape list, which the wavefront kernel
// then redoes with the full window; results are bit-exact either way (same arithmetic, same tie rules).
//
// Arithmetic, tags and traceback rules are those of gact_kernels_s16.cuh / align.cpp:60-233; only the mapping differs.
#pragma once
#include "gact_kernels_s16h.cuh"

namespace gact {

constexpr int IT_CS = 16;             // columns of a strip (per tile); 4 code words per row
constexpr int IT_PF = 8;              // rows ahead an edge entry is prefetched into L1
constexpr int IT_MAX_STRIPS = 64;     // tile_size <= 1024

struct ITGeom {
    int T, S, W;                      // tile size, strips, band half-width
    int i0, j0, s0;                   // first window row / column, first window strip
    int band_lo[IT_MAX_STRIPS];       // first / last tagged row of a strip (lo > hi: none)
    int band_hi[IT_MAX_STRIPS];
    int win_off[IT_MAX_STRIPS + 1];   // offset of a strip's code words (in units of IT_CS / 4 words per row), per lane
};

inline ITGeom it_geometry(int T, int et, int W)
{
    ITGeom g;
    g.T = T; g.S = T / IT_CS; g.W = W;
    g.i0 = (T - et > 1) ? T - et : 1;
    g.j0 = g.i0;
    g.s0 = (g.j0 - 1) / IT_CS;
    int off = 0;
    for (int s = 0; s < IT_MAX_STRIPS; s++) {
        g.band_lo[s] = 1; g.band_hi[s] = 0;
        g.win_off[s] = off;
        if (s < g.S && s >= g.s0) {
            const int lo = s * IT_CS + 1 - W, hi = s * IT_CS + IT_CS + W;
            g.band_lo[s] = lo > g.i0 ? lo : g.i0;
            g.band_hi[s] = hi < T ? hi : T;
            if (g.band_hi[s] >= g.band_lo[s]) off += g.band_hi[s] - g.band_lo[s] + 1;
        }
    }
    g.win_off[IT_MAX_STRIPS] = off;
    return g;
}
// per resident warp, in bytes
inline size_t it_edge_bytes(int T) { return (size_t)(T + 1 + IT_PF) * 32 * sizeof(uint2); }
// shared memory per warp: both tiles' reference rows (and, QS, query columns), 2 bits per base
inline size_t it_smem_per_warp(int T, bool qs) { return (size_t)(qs ? 4 : 2) * ((T + 15) / 16) * 32 * sizeof(uint32_t); }
inline size_t it_win_bytes(const ITGeom &g) { return (size_t)g.win_off[IT_MAX_STRIPS] * (IT_CS / 4) * 32 * sizeof(uint32_t); }

__device__ __forceinline__ void it_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 2-bit code of base `pos` of a packed set
__device__ __forceinline__ uint32_t it_code(const uint32_t *__restrict__ packed, long long pos)
{
    return (__ldg(packed + (pos >> 4)) >> (2 * (int)(pos & 15))) & 3u;
}

// w >> 2 as a multiply-high: the FMA pipe has room, the ALU pipe (shifts) is the one this kernel is bound by
__device__ __forceinline__ uint32_t it_shr2(uint32_t w)
{
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 0x40000000;" : "=r"(r) : "r"(w));
    return r;
}

// 4-byte score table of a reference row: `base` = (match, mismatch, mismatch, mismatch) * 16 rotated so that byte [code]
// is the match score; only the two low bits of `w` (the row's 2-bit code) are looked at
__device__ __forceinline__ uint32_t it_row_table(uint32_t base, uint32_t w)
{
    uint32_t r;
    asm("prmt.b32.b4e %0, %1, %1, %2;" : "=r"(r) : "r"(base), "r"(w));
    return r;
}

// One tile's view of the code words its lane wrote: a tagged row of a strip is one 16-byte record per lane
// ([strip][row][lane][4 words]: a warp's row is 512 contiguous bytes, a lane's record sits in one 32-byte sector).
struct ITWin {
    const uint32_t *w;        // this lane's record of the first tagged row
    int half;                 // 0: low half-word, 1: high
    // word index of cell (i, j)'s code word, or -1 if the cell lies outside the band
    __device__ __forceinline__ int index(const ITGeom &g, int i, int j) const
    {
        const int s = (j - 1) / IT_CS, c = (j - 1) - s * IT_CS;
        if (s < g.s0 || i < g.band_lo[s] || i > g.band_hi[s]) return -1;
        return (g.win_off[s] + (i - g.band_lo[s])) * (IT_CS / 4) * 32 + (c >> 2);
    }
    __device__ __forceinline__ int load(const ITGeom &g, int i, int j) const
    {
        const int x = index(g, i, j);
        if (x < 0) return -1;
        return (int)((w[x] >> (16 * half + 4 * (3 - ((j - 1) & 3)))) & 15u);
    }
    // pull the record of row i of the strip column j lies in towards L1 (no-op outside the band)
    __device__ __forceinline__ void prefetch(const ITGeom &g, int i, int j) const
    {
        if (i < 1 || j < 1) return;
        const int x = index(g, i, j);
        if (x >= 0) it_prefetch(w + x);
    }
};

// 2-bit code of DP row / column k (1-based) from the lane's re-packed words in shared memory (16 per word, 32 words apart)
__device__ __forceinline__ uint32_t it_smem_code(const uint32_t *words, int k)
{
    return (words[(size_t)((k - 1) >> 4) * 32] >> (2 * ((k - 1) & 15))) & 3u;
}

// Traceback of one tile by one lane (align.cpp:185-230): the reference's state machine, with the score of the cell the
// cursor stands on tracked instead of a stored ZERO code (H <= 0 <=> stop).  Returns false if the path left the band.
// The walk is a chain of dependent loads of the lane's own code words (DRAM: the scratch of all resident warps is far
// larger than L2), so the record IT_TB_PF rows up the diagonal is prefetched at every step (both strips when the
// prediction is near a strip boundary); the bases come from shared memory (reference rows always, query columns if QS).
constexpr int IT_TB_PF = 8;
template <bool QS>
__device__ __forceinline__ bool it_traceback(const ITGeom &g, const KParams &P, const ITWin &win,
                                             const uint32_t *rows, const uint32_t *cols,
                                             const uint32_t *__restrict__ qpk, long long qoff, int reverse,
                                             int corner, uint32_t *__restrict__ states_out, gact_tile_result *res)
{
    const int T = g.T, et = P.et;
    int i = T, j = T, v = corner, ri = et, rj = et, cnt = 0;
    if (v > 0) {
#pragma unroll
        for (int k = 1; k < IT_TB_PF; k++) win.prefetch(g, T - k, T - k);
    }
    int code = (v > 0) ? win.load(g, i, j) : 0;
    if (code < 0) return false;
    int state = code >> 2;
    uint32_t accw = 0;                       // 16 states per output word
    while (state != 0 && ri > 0 && rj > 0) {
        accw |= (uint32_t)state << (2 * (cnt & 15));
        cnt++;
        if ((cnt & 15) == 0) { states_out[(cnt >> 4) - 1] = accw; accw = 0; }
        {
            const int pi = i - IT_TB_PF, pj = j - IT_TB_PF;
            win.prefetch(g, pi, pj + 3);
            if (((pj + 2) ^ (pj - 4)) & ~(IT_CS - 1)) win.prefetch(g, pi, pj - 3);
        }
        if (state == 3) {
            // score of this cell's substitution (no exceptions on this path: equal codes <=> equal bytes)
            const uint32_t rc = it_smem_code(rows, i);
            const uint32_t qc = QS ? it_smem_code(cols, j) : it_code(qpk, qoff + (reverse ? (T - j) : (j - 1)));
            i--; j--; ri--; rj--;
            // the next cell's code is asked for before the score is known (same address either way)
            int nxt = 0;
            const bool inside = ri > 0 && rj > 0 && i >= g.i0 && j >= g.j0;
            if (inside) nxt = win.load(g, i, j);
            v -= (rc == qc) ? P.match : P.mismatch;
            if (ri <= 0 || rj <= 0) break;                      // early terminate: the next state is never looked at
            if (inside && v > 0) {
                if (nxt < 0) return false;
                code = nxt;
                state = code >> 2;
            } else {
                state = 0;
            }
        } else {
            const bool open = (state == 2) ? (code & 2) : (code & 1);
            v -= open ? P.gap_open : P.gap_extend;
            if (state == 2) { i--; ri--; } else { j--; rj--; }
            if (i <= 0 || j <= 0) { state = 0; }
            else if (open) { state = 3; }                       // M at the new cell, whatever its own code (align.cpp:218-229)
            else if (ri > 0 && rj > 0) {                        // the gap goes on: its flag sits in the new cell's code
                code = win.load(g, i, j);
                if (code < 0) return false;
            }
        }
    }
    if (cnt & 15) states_out[cnt >> 4] = accw;
    res->score = corner;
    res->max_i = T; res->max_j = T;
    res->n_states = cnt;
    res->i_steps = et - ri;
    res->j_steps = et - rj;
    return true;
}

// counters: [0] next batch of 64 tiles, [1] number of escaped tiles
template <bool QS>
__global__ void __launch_bounds__(128, 4)
gact_tile_it_kernel(const __grid_constant__ KParams P, const __grid_constant__ ITGeom G,
                    const gact_tile_desc *__restrict__ descs, const int *__restrict__ order, int n_batches,
                    gact_tile_result *__restrict__ results, uint32_t *__restrict__ states, int pitch_words,
                    int *counters, int *__restrict__ escaped,
                    uint2 *edge_scratch, uint32_t *win_scratch, size_t win_words_per_warp)
{
    extern __shared__ __align__(16) uint32_t it_smem[];
    constexpr int CS = IT_CS, NW = CS / 4;
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int T = G.T, S = G.S;
    uint2 *edge = edge_scratch + (size_t)gwarp * (T + 1 + IT_PF) * 32 + lane;   // edge[i * 32]: (H pair, D pair) of row i
    const int RW = (T + 15) / 16;                                               // words of reference rows per tile
    uint32_t *rowsA = it_smem + (size_t)(threadIdx.x >> 5) * (QS ? 4 : 2) * RW * 32 + lane;  // rowsA[k * 32]: rows 16k+1 .. 16k+16 of tile A
    uint32_t *rowsB = rowsA + (size_t)RW * 32;
    uint32_t *colsA = rowsB + (size_t)RW * 32, *colsB = colsA + (size_t)RW * 32;   // QS only: the query columns, same packing
    uint32_t *win = win_scratch + (size_t)gwarp * win_words_per_warp + lane * NW;   // this lane's 16-byte record of a tagged row
    // constants of the biased x16 domain (as SegCtx)
    const int B = P.s16_bias;
    const uint32_t Bp = pk16(B), ge16 = pk16(P.gap_extend * 16);
    const int KO = (P.gap_open * 16) * 65537, KI = (P.gap_open * 16 - 5) * 65537, KD = (P.gap_open * 16 - 10) * 65537;
    const int ONE = P.one;
    // score table of a reference row with code c: byte [c] = match * 16, the others mismatch * 16 = this constant
    // rotated by c bytes (prmt.b4e with the code in the two low bits of the selector)
    const uint32_t lut_base = ((uint32_t)((P.mismatch * 16) & 0xff) * 0x01010100u) | (uint32_t)((P.match * 16) & 0xff);
    // left border of the tile: H[i][0] = 0; D[i][0] is chosen so that D[i][1] = gap_open with the open flag set, which is
    // what the reference's -inf border gives (align.cpp:87-97)
    const uint32_t borderG = Bp, borderD = pk16(B + P.gap_open * 16), borderD_tag = pk16(B + P.gap_open * 16 + 5);
    // EDGE_D: the D value handed to the next strip belongs to that strip's first column, and so does its flag
    // (del_open >= del_extend): the edge always carries D TAGGED (....0101 open / ....0100 extend), also from untagged
    // rows, because the next strip may be inside its band where this strip is not.  In an untagged row the strip's last
    // column therefore takes its D maximum over the tagged candidates (clean values are multiples of 16, so the tags
    // decide ties only: open wins) -- same instruction count as the clean one.  H travels clean.
    const uint32_t ge16t = ge16 + pk16(4);
    const int KO5 = (P.gap_open * 16 + 5) * 65537;

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(&counters[0], 1);
        b = __shfl_sync(FULL, b, 0);
        if (b >= n_batches) break;
        const int tA = order[(size_t)b * 64 + lane], tB = order[(size_t)b * 64 + 32 + lane];
        const gact_tile_desc dA = descs[tA], dB = descs[tB];
        const SeqSetDev &rsA = P.sets[dA.ref_set], &qsA = P.sets[dA.query_set], &rsB = P.sets[dB.ref_set], &qsB = P.sets[dB.query_set];

        // ---- both tiles' reference rows in DP order, 2 bits per row, into shared memory ----
        for (int k = 0; k < RW; k++) {
            uint32_t wa = 0, wb = 0;
            for (int r = 0; r < 16; r++) {
                const int i = 16 * k + 1 + r;
                if (i <= T) {
                    wa |= it_code(rsA.packed, dA.ref_off + (dA.reverse ? (T - i) : (i - 1))) << (2 * r);
                    wb |= it_code(rsB.packed, dB.ref_off + (dB.reverse ? (T - i) : (i - 1))) << (2 * r);
                }
            }
            rowsA[(size_t)k * 32] = wa;
            rowsB[(size_t)k * 32] = wb;
            if (QS) {
                uint32_t ca = 0, cb = 0;
                for (int r = 0; r < 16; r++) {
                    const int j = 16 * k + 1 + r;
                    if (j <= T) {
                        ca |= it_code(qsA.packed, dA.query_off + (dA.reverse ? (T - j) : (j - 1))) << (2 * r);
                        cb |= it_code(qsB.packed, dB.query_off + (dB.reverse ? (T - j) : (j - 1))) << (2 * r);
                    }
                }
                colsA[(size_t)k * 32] = ca;
                colsB[(size_t)k * 32] = cb;
            }
        }
        __syncwarp();

        uint32_t cornerG = Bp;
        for (int s = 0; s < S; s++) {
            // ---- PRMT selectors of the strip's columns: byte [code] of the row table and its sign, per half ----
            uint32_t q[CS];
            static_assert(CS == 16, "a strip is one word of re-packed query columns");
            const uint32_t qwA = QS ? colsA[(size_t)s * 32] : 0u, qwB = QS ? colsB[(size_t)s * 32] : 0u;
#pragma unroll
            for (int c = 0; c < CS; c++) {
                const int j = s * CS + c + 1;
                const uint32_t l2 = QS ? ((qwA >> (2 * c)) & 3u) : it_code(qsA.packed, dA.query_off + (dA.reverse ? (T - j) : (j - 1)));
                const uint32_t h2 = QS ? ((qwB >> (2 * c)) & 3u) : it_code(qsB.packed, dB.query_off + (dB.reverse ? (T - j) : (j - 1)));
                q[c] = l2 | ((8u | l2) << 4) | ((4u | h2) << 8) | ((12u | h2) << 12);
            }
            uint32_t Gup[CS], IoUp[CS], IcUp[CS];
#pragma unroll
            for (int c = 0; c < CS; c++) {
                Gup[c] = Bp;                                            // H[0][j] = 0
                IoUp[c] = pk16(B + P.gap_open * 16);                    // M[0][j] + gap_open
                IcUp[c] = pk16(S16_NEG);                                // I[0][j] = -inf
            }
            uint32_t diag = Bp;                                         // H[0][j0 - 1]
            const bool first_strip = (s == 0), last_strip = (s == S - 1);
            const int blo = G.band_lo[s], bhi = G.band_hi[s];
            const bool has_band = bhi >= blo;
            const int u1_end = has_band ? blo - 1 : T;                  // rows 1 .. u1_end untagged

            uint32_t wA = 0, wB = 0;                                    // reference rows of the current group of 16, shifted as they are used
            int i = 1;
            // the edge entry of row i + 1 is fetched at the top of row i (rows i + 1 .. are still the previous strip's:
            // this strip overwrites entry i at the end of row i), the L1 prefetch runs IT_PF rows ahead of that
            uint2 enext = make_uint2(0u, 0u);
            if (!first_strip) {
                enext = edge[(size_t)32];
#pragma unroll
                for (int k = 2; k < IT_PF; k++) it_prefetch(edge + (size_t)k * 32);
            }
            // ---------------- untagged rows above the band ----------------
            for (; i <= u1_end; i++) {
                uint32_t inG = borderG, inD = borderD;
                if (!first_strip) { inG = enext.x; inD = enext.y & 0xfff0fff0u; enext = edge[(size_t)(i + 1) * 32]; it_prefetch(edge + (size_t)(i + IT_PF) * 32); }
                if (((i - 1) & 15) == 0) { wA = rowsA[(size_t)((i - 1) >> 4) * 32]; wB = rowsB[(size_t)((i - 1) >> 4) * 32]; }
                const uint2 rw = make_uint2(it_row_table(lut_base, wA), it_row_table(lut_base, wB));
                wA = it_shr2(wA); wB = it_shr2(wB);
                uint32_t hd = diag, dv = inD;
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const uint32_t sc = subst_score<true>(q[c], rw.x, rw.y, 0u, 0u);
                    const uint32_t mc = __viaddmax_s16x2(hd, sc, Bp);
                    hd = Gup[c];
                    const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
                    Gup[c] = __vimax3_s16x2(mc, iv, dv);
                    IoUp[c] = (uint32_t)((int)mc * ONE + KO);
                    IcUp[c] = iv;
                    // the last column's D only travels to the next strip: computed with its tag (EDGE_D above)
                    if (c < CS - 1) dv = __viaddmax_s16x2(dv, ge16, IoUp[c]);
                    else dv = __viaddmax_s16x2(dv, ge16t, (uint32_t)((int)mc * ONE + KO5));
                }
                diag = inG;
                if (!last_strip) edge[(size_t)i * 32] = make_uint2(Gup[CS - 1], dv);
            }
            if (has_band) {
                // ---------------- into the tagged domain ----------------
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    IoUp[c] = __vadd2(IoUp[c], pk16(10));
                    IcUp[c] = __vadd2(IcUp[c], pk16(8));
                }
                uint32_t *wp = win + (size_t)G.win_off[s] * NW * 32;
                for (; i <= bhi; i++) {
                    uint32_t inG = borderG, inD = borderD_tag;
                    if (!first_strip) { inG = enext.x; inD = enext.y; enext = edge[(size_t)(i + 1) * 32]; it_prefetch(edge + (size_t)(i + IT_PF) * 32); }
                    if (((i - 1) & 15) == 0) { wA = rowsA[(size_t)((i - 1) >> 4) * 32]; wB = rowsB[(size_t)((i - 1) >> 4) * 32]; }
                    const uint2 rw = make_uint2(it_row_table(lut_base, wA), it_row_table(lut_base, wB));
                    wA = it_shr2(wA); wB = it_shr2(wB);
                    uint32_t hd = diag, dv = inD;                             // tagged, with the flag of this strip's first column
                    uint32_t acc[NW];
#pragma unroll
                    for (int c = 0; c < CS; c++) {
                        const uint32_t sc = subst_score<true>(q[c], rw.x, rw.y, 0u, 0u);
                        const uint32_t mt = __viaddmax_s16x2(hd, sc, Bp) | 0x000f000fu;
                        hd = Gup[c];
                        const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
                        const uint32_t g = __vimax3_s16x2(mt, iv, dv);
                        const uint32_t code = (g & 0x000c000cu) | ((iv | dv) & 0x00030003u);
                        if ((c & 3) == 0) acc[c >> 2] = code; else acc[c >> 2] = acc[c >> 2] * 16u + code;
                        Gup[c] = g;
                        IoUp[c] = (uint32_t)((int)mt * ONE + KI);
                        IcUp[c] = iv & 0xfffdfffdu;
                        dv = __viaddmax_s16x2(dv & 0xfffefffeu, ge16, (uint32_t)((int)mt * ONE + KD));
                    }
                    diag = inG;
                    if (!last_strip) edge[(size_t)i * 32] = make_uint2(Gup[CS - 1] & 0xfff0fff0u, dv);
                    static_assert(NW == 4, "one 16-byte record per lane and tagged row");
                    *reinterpret_cast<uint4 *>(wp) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                    wp += NW * 32;
                }
                // ---------------- back to the untagged domain ----------------
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    Gup[c] &= 0xfff0fff0u;
                    IoUp[c] &= 0xfff0fff0u;
                    IcUp[c] &= 0xfff0fff0u;
                }
                for (; i <= T; i++) {
                    uint32_t inG = borderG, inD = borderD;
                    if (!first_strip) { inG = enext.x; inD = enext.y & 0xfff0fff0u; enext = edge[(size_t)(i + 1) * 32]; it_prefetch(edge + (size_t)(i + IT_PF) * 32); }
                    if (((i - 1) & 15) == 0) { wA = rowsA[(size_t)((i - 1) >> 4) * 32]; wB = rowsB[(size_t)((i - 1) >> 4) * 32]; }
                    const uint2 rw = make_uint2(it_row_table(lut_base, wA), it_row_table(lut_base, wB));
                    wA = it_shr2(wA); wB = it_shr2(wB);
                    uint32_t hd = diag, dv = inD;
#pragma unroll
                    for (int c = 0; c < CS; c++) {
                        const uint32_t sc = subst_score<true>(q[c], rw.x, rw.y, 0u, 0u);
                        const uint32_t mc = __viaddmax_s16x2(hd, sc, Bp);
                        hd = Gup[c];
                        const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
                        Gup[c] = __vimax3_s16x2(mc, iv, dv);
                        IoUp[c] = (uint32_t)((int)mc * ONE + KO);
                        IcUp[c] = iv;
                        // the last column's D only travels to the next strip: computed with its tag (EDGE_D above)
                        if (c < CS - 1) dv = __viaddmax_s16x2(dv, ge16, IoUp[c]);
                        else dv = __viaddmax_s16x2(dv, ge16t, (uint32_t)((int)mc * ONE + KO5));
                    }
                    diag = inG;
                    if (!last_strip) edge[(size_t)i * 32] = make_uint2(Gup[CS - 1], dv);
                }
            }
            cornerG = Gup[CS - 1];
        }
        // the code words of this warp must be visible to its own lanes' loads below (same thread wrote them: program order)
        // ---- tracebacks: each lane walks its two tiles ----
        const int cA = (((int)(short)(cornerG & 0xffffu)) - B) >> 4, cB = (((int)(short)(cornerG >> 16)) - B) >> 4;
        {
            gact_tile_result r;
            ITWin w; w.w = win; w.half = 0;
            if (it_traceback<QS>(G, P, w, rowsA, colsA, qsA.packed, dA.query_off, dA.reverse, cA,
                             states + (size_t)tA * pitch_words, &r))
                results[tA] = r;
            else
                escaped[atomicAdd(&counters[1], 1)] = tA;
        }
        {
            gact_tile_result r;
            ITWin w; w.w = win; w.half = 1;
            if (it_traceback<QS>(G, P, w, rowsB, colsB, qsB.packed, dB.query_off, dB.reverse, cB,
                             states + (size_t)tB * pitch_words, &r))
                results[tB] = r;
            else
                escaped[atomicAdd(&counters[1], 1)] = tB;
        }
        __syncwarp();
    }
}

}  // namespace gact
