// gact_kernels_s16h.cuh -- packed s16x2 DPX GACT tile kernel, LANES lanes per tile.
//
// Same arithmetic as gact_kernels_s16.cuh (tagged x16 domain, see there), different mapping:
// a tile is owned by a SEGMENT of LANES = 16 lanes, so a warp aligns 32 / LANES = 2 tiles at a
// time and each lane owns two strips of CS = tile_size / (2 * LANES) columns (10 at T = 320).
// Compared with one tile per warp this halves the wavefront skew (n + 31 instead of n + 63 steps)
// and spreads the per-step overhead (edge shuffle, predicates, row fetch) over twice as many
// cells.  The direction window lives in per-segment global scratch (L2-resident).
#pragma once
#include <type_traits>
#include "gact_kernels_s16.cuh"

// build-time tuning knobs of the wavefront loops (defaults = what measured best on B200, profiles/r2_latency_tuning.txt)
#ifndef GACT_TAG_UNROLL
#define GACT_TAG_UNROLL 4            // unroll factor of the tagged (window-row) loop of the latency kernels (one warp alone on
#endif                               // a sub-partition: 50.6 k -> 44.1 k cycles per tile); the throughput kernels use 2 (+1 %)
#ifndef GACT_RR_PREFETCH
#define GACT_RR_PREFETCH 1           // fetch the next step's reference word one step ahead
#endif

namespace gact {

template <int CS, int LANES>
struct TagUnroll { static constexpr int value = (LANES == 32 && CS <= 5) ? GACT_TAG_UNROLL : 2; };

#ifdef GACT_CHECK
// Bounds-checked build (libgact_b200_check.so, tools/sanitize_run.py): every access to the direction window and to the
// per-segment shared-memory arrays is range-checked; violations are counted per site.  Stands in for compute-sanitizer,
// which is closed on the GPU pool.
__device__ unsigned long long g_check_fail[8];
#define GACT_CHK(site, cond) do { if (!(cond)) atomicAdd(&g_check_fail[site], 1ull); } while (0)
#else
#define GACT_CHK(site, cond)
#endif

template <int CS>
struct DirWinH {
    static constexpr int NW = CS / 4;                 // 32-bit words per lane-step (4 columns x 2 strips each)
    static constexpr int R = CS % 4;                  // leftover columns, kept in one 16-bit field (R <= 2)
    static_assert(R <= 2, "unsupported strip width");
    // leftover columns of a lane-step (both strips): one byte for R = 1 (4 bits per strip), two for R = 2
    struct HT1 { typedef uint8_t type; };
    struct HT2 { typedef uint16_t type; };
    typedef typename std::conditional<R == 1, HT1, HT2>::type::type HT;
    int i0, lane0, nl;
    uint32_t *w;
    HT *h;
#ifdef GACT_CHECK
    int rows;                  // rows allocated (win_rows + 1)
#endif
    __device__ __forceinline__ void init(void *base, int n, int m, const KParams &P)
    {
#ifdef GACT_CHECK
        rows = P.win_rows + 1;
#endif
        i0 = max(n - P.et, 1);
        const int j0 = max(m - P.et, 1);
        lane0 = ((j0 - 1) / CS) >> 1;
        nl = P.win_lanes;
        w = reinterpret_cast<uint32_t *>(base);
        h = reinterpret_cast<HT *>(w + (size_t)(P.win_rows + 1) * nl * NW);
    }
    // the leftover field of a lane-step from the accumulator word (low strip's codes in bits 7:0, high strip's in 23:16)
    static __device__ __forceinline__ HT pack_rest(uint32_t acc)
    {
        if (R == 1) return (HT)((acc & 0xfu) | ((acc >> 12) & 0xf0u));
        return (HT)((acc & 0xffu) | ((acc >> 8) & 0xff00u));
    }
    // 4-bit code of cell (i, j): bits 3:2 = M/I/D tag, bit 1 = ins flag, bit 0 = del flag
    __device__ __forceinline__ int load(int i, int j) const
    {
        const int s = (j - 1) / CS, c = (j - 1) - s * CS;
        const int lane = s >> 1, half = s & 1;
        const int e = (i + half - i0) * nl + (lane - lane0);
        GACT_CHK(0, i + half - i0 >= 0 && i + half - i0 < rows && lane - lane0 >= 0 && lane - lane0 < nl);
        if (c < NW * 4) return (w[e * NW + (c >> 2)] >> (16 * half + 4 * (3 - (c & 3)))) & 15;
        if (R == 1) return (h[e] >> (4 * half)) & 15;
        return (h[e] >> (8 * half + 4 * (R - 1 - (c - NW * 4)))) & 15;
    }
    static __host__ __device__ size_t bytes(int win_rows, int win_lanes)
    {
        size_t s = (size_t)(win_rows + 1) * win_lanes * (NW * 4 + R);
        return (s + 15) & ~(size_t)15;
    }
};

// ---------------------------------------------------------------------------
// Per-segment context and the building blocks shared by the tile kernel, the first-tile kernel and
// the chain kernel.  Every function below is executed by the WHOLE warp (its shuffles and ballots use
// the full mask); each 16-lane segment works on its own tile, segments without work pass n = m = 0.
// Shared-memory bytes of one segment's sequence arrays.  rr[] carries PAD = 2 * LANES sentinel words on the left (rows
// <= 0: lanes that have not reached row 1 yet run pseudo rows against the sentinel, which reproduce the border values,
// so the wavefront loop needs no per-lane "am I active" branch) and PAD words of slack on the right (rows past the
// end of a tile: their results are never used, the words only have to be readable).
// Column capacity of a mapping is TS = CS * 2 * LANES; its ROW capacity TR is the same except for the narrow mappings
// (16 lanes, strips of 4 / 5 columns), which exist for the tiles of a batch whose query window is at most half a tile
// wide: they take reference windows of the full tile size (TR = 2 TS) at half the work per wavefront step.
__host__ __device__ constexpr int s16h_row_cap(int CS, int LANES)
{
    return (LANES == 16 && CS <= 5) ? 4 * CS * LANES : 2 * CS * LANES;
}
__host__ __device__ constexpr size_t s16h_seq_bytes(int CS, int LANES)
{
    const int TS = CS * 2 * LANES, TR = s16h_row_cap(CS, LANES), PAD = 2 * LANES;
    return (size_t)((((PAD + TR + 2 + PAD) * 4 + (TS + 2 + TR + 2) * 2) + 15) & ~15) +
           (size_t)((((TR / 16 + 2) + (TS / 16 + 2) + (TR / 32 + 2)) * 4 + 15) & ~15);
}

template <int CS, int LANES>
struct SegCtx {
    static constexpr int TS = CS * 2 * LANES;              // columns
    static constexpr int TR = s16h_row_cap(CS, LANES);     // rows
    static constexpr int PAD = 2 * LANES;
    // layout of a segment's carve-out (init): rr | qs, rb | wr, wq, we -- it must fit what the host reserves
    static constexpr size_t OFF_WORDS = (size_t)((((PAD + TR + 2 + PAD) * 4 + (TS + 2 + TR + 2) * 2) + 15) & ~15);
    static constexpr size_t OFF_END = OFF_WORDS + (size_t)((TR / 16 + 2) + (TS / 16 + 2) + (TR / 32 + 2)) * 4;
    static_assert(OFF_END <= s16h_seq_bytes(CS, LANES), "segment arrays exceed s16h_seq_bytes");
    static_assert((PAD + TR + 2 + PAD) * 4 % 2 == 0 && OFF_WORDS % 16 == 0, "alignment of the 16- and 32-bit arrays");
    int lane, seg, sl, segbase;
    uint32_t *rr;            // rr[i]: substitution table of R[i] (LUT) or enc(R[i]) | enc(R[i-1]) << 16
    uint16_t *qs, *rb;       // enc(Q[j]), enc(R[i])
    uint32_t *wr, *wq;       // the tile's 2-bit packed words as loaded from HBM (reference <= TR/16 + 2, query <= TS/16 + 2)
    uint32_t *we;            // exception bitmap words of the reference window (<= TR/32 + 2), sets with exceptions only
    void *dirbase;
    // constants of the biased x16 domain
    int B, KO, KI, KD, ONE, et, match, mismatch, gap_open, gap_extend;
    uint32_t Bp, ma16, mi16, ge16, borderD_tag, borderD_raw, lut_mis, lut_delta;

    // seq_stride: bytes between consecutive segments' carve-outs (>= s16h_seq_bytes; the latency chain kernel keeps
    // the direction window behind the sequence arrays).  gscratch: global scratch of the direction windows or nullptr.
    template <bool SMEMWIN = false>
    __device__ __forceinline__ void init(const KParams &P, uint8_t *smem, int warp, size_t seq_stride, uint8_t *gscratch,
                                         size_t dir_bytes, int global_warp, bool lut)
    {
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        seg = lane / LANES; sl = lane % LANES; segbase = seg * LANES;
        constexpr int TPW = 32 / LANES;
        uint8_t *my = smem + ((size_t)warp * TPW + seg) * seq_stride;
        rr = reinterpret_cast<uint32_t *>(my) + PAD;
        qs = reinterpret_cast<uint16_t *>(my + (PAD + TR + 2 + PAD) * 4);
        rb = qs + (TS + 2);
        wr = reinterpret_cast<uint32_t *>(my + OFF_WORDS);
        wq = wr + (TR / 16 + 2);
        we = wq + (TS / 16 + 2);
        // direction window: per-segment global scratch, or (SMEMWIN) the shared memory behind this segment's sequence arrays
        if (SMEMWIN) dirbase = (void *)(my + s16h_seq_bytes(CS, LANES));
        else dirbase = gscratch ? (void *)(gscratch + ((size_t)global_warp * TPW + seg) * dir_bytes) : nullptr;
        B = P.s16_bias; Bp = pk16(B);
        match = P.match; mismatch = P.mismatch; gap_open = P.gap_open; gap_extend = P.gap_extend; et = P.et;
        ma16 = pk16(P.match * 16); mi16 = pk16(P.mismatch * 16); ge16 = pk16(P.gap_extend * 16);
        KO = (P.gap_open * 16) * 65537;
        KI = (P.gap_open * 16 - 5) * 65537;
        KD = (P.gap_open * 16 - 10) * 65537;
        ONE = P.one;
        borderD_tag = ((uint32_t)(B + P.gap_open * 16 + 5) << 16) | (uint32_t)B;
        borderD_raw = ((uint32_t)(B + P.gap_open * 16) << 16) | (uint32_t)B;
        lut_mis = (uint32_t)((P.mismatch * 16) & 0xff) * 0x01010101u;
        lut_delta = (uint32_t)(((P.match ^ P.mismatch) * 16) & 0xff);
        // sentinel rows <= 0 (written once: staging only writes rows >= 1) and defined slack on the right
        const uint32_t sent = lut ? lut_mis : (SENT_R | (SENT_R << 16));
        for (int x = sl; x <= PAD; x += LANES) rr[-x] = sent;
        for (int x = 1 + sl; x < TR + 2 + PAD; x += LANES) rr[x] = sent;
        __syncwarp();
    }
};

// Base x (1-based, DP order) of a tile whose packed words sit in shared memory: word 0 holds base
// (off & ~15) of the set.
__device__ __forceinline__ int smem_base(const uint32_t *w, int off_in_word0, int len, int reverse, int x)
{
    const int pos = off_in_word0 + (reverse ? (len - x) : (x - 1));
    const int code = (w[pos >> 4] >> (2 * (pos & 15))) & 3;
    return (0x54474341u >> (8 * code)) & 0xff;            // "ACGT"
}

// Stage the tile's bases (DP order) into the segment's shared-memory arrays.  2-bit packed sets are
// fetched with coalesced 32-bit loads (one word = 16 bases per lane) into shared memory and expanded
// from there; 8-bit sets are read per base.
template <int CS, int LANES, bool LUT>
__device__ __forceinline__ void seg_stage(const SegCtx<CS, LANES> &cx, const SeqSetDev &rset, const SeqSetDev &qset,
                                          long long ref_off, int ref_len, long long query_off, int query_len,
                                          int reverse, int n, int m)
{
    __syncwarp();
    const bool work = (n > 0 && m > 0);
    const bool rpk = rset.packed && !rset.bytes, qpk = qset.packed && !qset.bytes;    // raw bytes win where a set keeps them
    if (work && rpk) {
        const long long w0 = ref_off >> 4;
        const int nw = (int)(((ref_off + ref_len - 1) >> 4) - w0) + 1;
        for (int x = cx.sl; x < nw; x += LANES) cx.wr[x] = __ldg(rset.packed + w0 + x);
    }
    if (work && qpk) {
        const long long w0 = query_off >> 4;
        const int nw = (int)(((query_off + query_len - 1) >> 4) - w0) + 1;
        for (int x = cx.sl; x < nw; x += LANES) cx.wq[x] = __ldg(qset.packed + w0 + x);
    }
    __syncwarp();
    if (work) {
        const int ro = (int)(ref_off & 15), qo = (int)(query_off & 15);
        for (int x = 1 + cx.sl; x <= n + 1; x += LANES) {
            const bool in = (x <= n);
            const int base = !in ? 0 : rpk ? smem_base(cx.wr, ro, ref_len, reverse, x)
                                           : tile_base(rset, ref_off, ref_len, reverse, x);
            cx.rb[x] = (uint16_t)(in ? enc_base(base) : SENT_R);
            if (LUT) {
                const int code = (base == 'A') ? 0 : (base == 'C') ? 1 : (base == 'G') ? 2 : (base == 'T') ? 3 : 4;
                uint32_t w = cx.lut_mis;
                if (in && code < 4) w ^= cx.lut_delta << (8 * code);
                cx.rr[x] = w;
            }
        }
        for (int x = cx.sl; x <= m; x += LANES) {
            const int base = x < 1 ? 0 : qpk ? smem_base(cx.wq, qo, query_len, reverse, x)
                                             : tile_base(qset, query_off, query_len, reverse, x);
            cx.qs[x] = (x >= 1) ? (uint16_t)enc_base(base) : (uint16_t)SENT_Q;
        }
    }
    __syncwarp();
    if (!LUT && work) {
        for (int x = 1 + cx.sl; x <= n + 1; x += LANES)
            cx.rr[x] = (uint32_t)cx.rb[x] | ((uint32_t)(x >= 2 ? cx.rb[x - 1] : (uint16_t)SENT_R) << 16);
    }
    __syncwarp();
}

// Staging from the 2-bit packed words (every kernel that scores with the one-PRMT table): both word loads are in flight
// together, reference and query are expanded in one loop without the byte-set branches, and rb[] / qs[] hold what
// seg_load_q and the traceback's match test expect (the ASCII encoding of seg_stage), computed from the 2-bit code by
// one shift.  The query window must be free of exceptions (the host routes other tiles to the raw-byte kernels); a
// reference base that is an exception (rset.exc) becomes a sentinel row: it mismatches every query base, exactly what
// raw byte equality gives against an ACGT-only query (align.cpp:134).
template <int CS, int LANES>
__device__ __forceinline__ void seg_stage_packed(const SegCtx<CS, LANES> &cx, const SeqSetDev &rset, const SeqSetDev &qset,
                                                 long long ref_off, int ref_len, long long query_off, int query_len,
                                                 int reverse, int n, int m)
{
    __syncwarp();
    const bool work = (n > 0 && m > 0);
    if (work) {
        const long long rw0 = ref_off >> 4, qw0 = query_off >> 4;
        const int rnw = (int)(((ref_off + ref_len - 1) >> 4) - rw0) + 1;
        const int qnw = (int)(((query_off + query_len - 1) >> 4) - qw0) + 1;
        for (int x = cx.sl; x < max(rnw, qnw); x += LANES) {
            uint32_t a = 0, b = 0;
            if (x < rnw) a = __ldg(rset.packed + rw0 + x);
            if (x < qnw) b = __ldg(qset.packed + qw0 + x);
            if (x < rnw) cx.wr[x] = a;
            if (x < qnw) cx.wq[x] = b;
        }
        if (rset.exc) {
            const long long e0 = ref_off >> 5;
            const int enw = (int)(((ref_off + ref_len - 1) >> 5) - e0) + 1;
            for (int x = cx.sl; x < enw; x += LANES) cx.we[x] = __ldg(rset.exc + e0 + x);
        }
    }
    __syncwarp();
    if (work) {
        const int ro = (int)(ref_off & 15), qo = (int)(query_off & 15);
        // DP index x (1-based) -> position in the staged words: natural order, or back to front for reverse tiles
        const int rbase = reverse ? ro + ref_len : ro - 1, qbase = reverse ? qo + query_len : qo - 1, step = reverse ? -1 : 1;
        const int eo = (int)(ref_off & 31) - ro;             // bit position of a base in we[] = its position in wr[] + eo
        const bool has_exc = rset.exc != nullptr;
        const int top = max(n + 1, m);
        for (int x = cx.sl; x <= top; x += LANES) {
            if (x >= 1 && x <= n + 1) {
                const int pos = rbase + step * x;
                bool in = (x <= n);
                if (has_exc && in && ((cx.we[(pos + eo) >> 5] >> ((pos + eo) & 31)) & 1u)) in = false;
                const uint32_t code = in ? (cx.wr[pos >> 4] >> (2 * (pos & 15))) & 3u : 0u;
                cx.rb[x] = (uint16_t)(in ? enc_base((0x54474341u >> (8 * code)) & 0xffu) : SENT_R);
                cx.rr[x] = in ? (cx.lut_mis ^ (cx.lut_delta << (8 * code))) : cx.lut_mis;
            }
            if (x <= m) {
                const int pos = qbase + step * x;
                const uint32_t code = (x >= 1) ? (cx.wq[pos >> 4] >> (2 * (pos & 15))) & 3u : 0u;
                cx.qs[x] = (uint16_t)((x >= 1) ? enc_base((0x54474341u >> (8 * code)) & 0xffu) : SENT_Q);
            }
        }
    }
    __syncwarp();
}

// per-lane query registers: enc pairs (general) or PRMT selectors (LUT)
template <int CS, int LANES, bool LUT>
__device__ __forceinline__ void seg_load_q(const SegCtx<CS, LANES> &cx, int m, uint32_t (&q)[CS])
{
#pragma unroll
    for (int c = 0; c < CS; c++) {
        const int jl = (2 * cx.sl) * CS + c + 1, jh = (2 * cx.sl + 1) * CS + c + 1;
        const uint32_t el = jl <= m ? (uint32_t)cx.qs[jl] : SENT_Q + c, eh = jh <= m ? (uint32_t)cx.qs[jh] : SENT_Q + c;
        if (LUT) {
            const uint32_t tl = (el >> 1) & 3u, th = (eh >> 1) & 3u;       // A0 C1 G3 T2 -> 0 1 2 3 below
            const uint32_t l2 = (jl <= m) ? (tl ^ (tl >> 1)) : 0u, h2 = (jh <= m) ? (th ^ (th >> 1)) : 0u;
            q[c] = l2 | ((8u | l2) << 4) | ((4u | h2) << 8) | ((12u | h2) << 12);
        } else {
            q[c] = el | (eh << 16);
        }
    }
}

// DP of one tile per segment: fills the direction window, returns the corner score H[n][m].
// The wavefront loops carry no per-lane activity test: a lane that has not reached row 1 yet computes pseudo rows
// <= 0 against the sentinel words left of rr[1] (every score a mismatch: H stays 0, I and D stay at gap_open, which
// gives row 1 the same values and the same open-flags as the reference's -inf borders, align.cpp:87-97), and a lane
// past its last row computes rows whose results nobody reads (its stores are predicated off, the corner is taken at
// the corner cell's own step).
template <int CS, int LANES, bool LUT>
__device__ __forceinline__ int seg_dp(const SegCtx<CS, LANES> &cx, const uint32_t (&q)[CS], int n, int m,
                                      const DirWinH<CS> &dw)
{
    constexpr int NW = DirWinH<CS>::NW;
    constexpr int R = DirWinH<CS>::R;
    const int sl = cx.sl, B = cx.B;
    const uint32_t Bp = cx.Bp, ge16 = cx.ge16;
    const int laststrip = (m > 0) ? (m - 1) / CS : -1;
    const int lastlane = laststrip >> 1;                       // segment-local
    const int c_lane = max(lastlane, 0), c_half = laststrip & 1, c_col = (m > 0) ? (m - 1) - laststrip * CS : 0;
    const int kc = n + 2 * c_lane + c_half;                    // step in which the corner cell (n, m) is computed
    const bool work = (n > 0 && m > 0);
    const int steps_seg = work ? kc : 0;                       // no real cell is left after the corner's step
    // the first lane that keeps direction codes (lane0) reaches window row i0 at step i0 + 2*lane0;
    // until then every lane can stay in the cheaper untagged loop
    int steps = steps_seg, k1 = work ? min(dw.i0 - 1 + 2 * dw.lane0, steps_seg) : 0x3fffffff;
    int kc_min = work ? kc : 0x3fffffff;                       // the corner step of the segment that finishes first
#pragma unroll
    for (int o = LANES; o < 32; o <<= 1) {
        steps = max(steps, __shfl_xor_sync(FULL, steps, o));
        k1 = min(k1, __shfl_xor_sync(FULL, k1, o));
        kc_min = min(kc_min, __shfl_xor_sync(FULL, kc_min, o));
    }
    k1 = min(k1, steps);
    kc_min = min(kc_min, steps);
    // codes are kept from window row i0 on, by the lanes that own window columns, while the lane has real rows
    const bool keeps = work && sl >= dw.lane0 && sl <= lastlane;
    const int kstore = keeps ? dw.i0 + 2 * sl : 0x3fffffff;
    const int kend = keeps ? n + 2 * sl + 1 : -1;            // last step with a real row in either half
    const uint32_t *rrp = cx.rr - 2 * sl;               // rrp[k] = rr[k - 2*sl]

    // ---------------- phase 1: rows above every segment's window, score only ----------------
    uint32_t Gup[CS], IoUp[CS], IcUp[CS];
#pragma unroll
    for (int c = 0; c < CS; c++) {
        Gup[c] = Bp;
        IoUp[c] = pk16(B + cx.gap_open * 16, S16_NEG);
        IcUp[c] = pk16(S16_NEG, S16_NEG);
    }
    uint32_t eG = Bp, eD = pk16(S16_NEG), diag = Bp;
    uint32_t rprev = LUT ? rrp[0] : 0u;                 // reference word of the previous step = this step's high half
    uint32_t rnext = rrp[1];                            // this step's word, fetched one step ahead (GACT_RR_PREFETCH)
    int k = 1;
    for (; k <= k1; k++) {
        const uint32_t pack = __byte_perm(eG, eD, 0x7632);
        uint32_t recv = __shfl_up_sync(FULL, pack, 1, LANES);
        if (sl == 0) recv = cx.borderD_raw;
        const uint32_t inG = __byte_perm(recv, eG, 0x5410);
        const uint32_t inD = __byte_perm(recv, eD, 0x5432);
        GACT_CHK(2, k - 2 * sl >= -cx.PAD && k + 1 - 2 * sl < cx.TR + 2 + cx.PAD);
        const uint32_t rlo = GACT_RR_PREFETCH ? rnext : rrp[k], rhi = rprev;
        if (GACT_RR_PREFETCH) rnext = rrp[k + 1];
        uint32_t hd = diag, dv = inD;
#pragma unroll
        for (int c = 0; c < CS; c++) {
            const uint32_t sc = subst_score<LUT>(q[c], rlo, rhi, cx.ma16, cx.mi16);
            const uint32_t mc = __viaddmax_s16x2(hd, sc, Bp);
            hd = Gup[c];
            const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
            Gup[c] = __vimax3_s16x2(mc, iv, dv);
            const uint32_t mo = (uint32_t)((int)mc * cx.ONE + cx.KO);
            IoUp[c] = mo;
            IcUp[c] = iv;
            dv = __viaddmax_s16x2(dv, ge16, mo);
        }
        eG = Gup[CS - 1];
        eD = dv;
        diag = inG;
        if (LUT) rprev = rlo;
    }
    // ---------------- switch to the tagged domain ----------------
#pragma unroll
    for (int c = 0; c < CS; c++) {
        IoUp[c] = __vadd2(IoUp[c], pk16(10));
        IcUp[c] = __vadd2(IcUp[c], pk16(8));
    }
    eD = __vadd2(eD, pk16(4));

    // ---------------- phase 2: window rows, tagged values + direction codes ----------------
    uint32_t *wptr = dw.w + ((k - 2 * sl - dw.i0) * dw.nl + (sl - dw.lane0)) * NW;
    typename DirWinH<CS>::HT *hptr = dw.h + ((k - 2 * sl - dw.i0) * dw.nl + (sl - dw.lane0));
    int corner16 = B;
    // the loop stops at the corner step of each segment (at most two different ones per warp) so that the corner
    // value is picked out of the registers outside the loop
    for (int stop = kc_min;; stop = steps) {
#pragma unroll (TagUnroll<CS, LANES>::value)
    for (; k <= stop; k++) {
        const uint32_t pack = __byte_perm(eG, eD, 0x7632);
        uint32_t recv = __shfl_up_sync(FULL, pack, 1, LANES);
        if (sl == 0) recv = cx.borderD_tag;
        const uint32_t inG = __byte_perm(recv, eG, 0x5410);
        const uint32_t inD = __byte_perm(recv, eD, 0x5432);
        GACT_CHK(2, k - 2 * sl >= -cx.PAD && k + 1 - 2 * sl < cx.TR + 2 + cx.PAD);
        const uint32_t rlo = GACT_RR_PREFETCH ? rnext : rrp[k], rhi = rprev;
        if (GACT_RR_PREFETCH) rnext = rrp[k + 1];
        uint32_t hd = diag, dv = inD;
        uint32_t acc[NW + 1];
#pragma unroll
        for (int c = 0; c < CS; c++) {
            const uint32_t sc = subst_score<LUT>(q[c], rlo, rhi, cx.ma16, cx.mi16);
            const uint32_t mt = __viaddmax_s16x2(hd, sc, Bp) | 0x000f000fu;
            hd = Gup[c];
            const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
            const uint32_t g = __vimax3_s16x2(mt, iv, dv);
            const uint32_t code = (g & 0x000c000cu) | ((iv | dv) & 0x00030003u);
            if ((c & 3) == 0) acc[c >> 2] = code; else acc[c >> 2] = acc[c >> 2] * 16u + code;
            Gup[c] = g;
            IoUp[c] = (uint32_t)((int)mt * cx.ONE + cx.KI);
            IcUp[c] = iv & 0xfffdfffdu;
            dv = __viaddmax_s16x2(dv & 0xfffefffeu, ge16, (uint32_t)((int)mt * cx.ONE + cx.KD));
        }
        eG = Gup[CS - 1];
        eD = dv;
        diag = inG;
        if (LUT) rprev = rlo;
        if (k >= kstore && k <= kend) {
            GACT_CHK(1, k - 2 * sl - dw.i0 >= 0 && k - 2 * sl - dw.i0 < dw.rows && sl - dw.lane0 >= 0 && sl - dw.lane0 < dw.nl &&
                            wptr == dw.w + ((k - 2 * sl - dw.i0) * dw.nl + (sl - dw.lane0)) * NW);
#pragma unroll
            for (int x = 0; x < NW; x++) wptr[x] = acc[x];
            if (R) *hptr = DirWinH<CS>::pack_rest(acc[NW]);
        }
        wptr += dw.nl * NW;
        hptr += dw.nl;
    }
        if (work && kc == stop) {
            uint32_t gsel = 0;
#pragma unroll
            for (int c = 0; c < CS; c++) if (c == c_col) gsel = Gup[c];
            corner16 = (int)(short)(c_half ? (gsel >> 16) : (gsel & 0xffffu));
        }
        if (stop >= steps) break;
    }
    int corner = (__shfl_sync(FULL, corner16, cx.segbase + c_lane) - B) >> 4;
    if (!work) corner = 0;
    __threadfence_block();
    __syncwarp();
    return corner;
}

// Result of one traceback (align.cpp:185-230), all lanes of the segment hold the same values.
struct SegTrace {
    int cnt;             // number of states
    int is, js;          // reference / query bases consumed
    int col_score;       // sum of the per-column scores of gact.cpp:197-210 over these states
    int first_gap;       // was the first state a gap column (valid if cnt > 0)
    int last_gap;        // prev_gap after the last state
};

// Traceback by the segment.  EMIT: write the states (one byte each) into stbuf.  prev_gap: whether the
// column scored just before this tile's first state was a gap column (for the open/extend accounting).
// In state M the lanes inspect the next LANES cells down the diagonal at once; ballot + popc give the
// scores along the run, ffs its length.
template <int CS, int LANES, bool EMIT>
__device__ __forceinline__ SegTrace seg_traceback(const SegCtx<CS, LANES> &cx, const DirWinH<CS> &dw, int n, int m,
                                                  int corner, uint8_t *stbuf, int prev_gap)
{
    constexpr unsigned SEGMASK = (LANES == 32) ? 0xffffffffu : ((1u << LANES) - 1u);
    const int sl = cx.sl, et = cx.et;
    const int ma = cx.match, mi = cx.mismatch, go = cx.gap_open, ge = cx.gap_extend;
    const bool work = (n > 0 && m > 0);
    const int i0 = dw.i0, j0 = max(m - et, 1);
    int i = n, j = m, cnt = 0, v = corner, ri = et, rj = et;
    int col = 0, first_gap = 0, pg = prev_gap;
    // `code` is the direction code of the cell the cursor stands on whenever the cursor is in a gap state: an M run
    // ends on a cell whose code one of the lanes has just loaded, so only a gap that goes on needs another load
    int code = (work && v > 0) ? dw.load(i, j) : 0;
    int state = code >> 2;
    bool act = (state != 0);
    while (LANES == 32 ? act : __any_sync(FULL, act)) {
        // one gap column (if the segment is in I or D), then the M run that follows it
        if (act && state != 3) {
            const bool open = (state == 2) ? (code & 2) : (code & 1);
            if (EMIT && sl == 0) stbuf[cnt] = (uint8_t)state;
            if (cnt == 0) first_gap = 1;
            cnt++;
            v -= open ? go : ge;
            col += pg ? ge : go;
            pg = 1;
            if (state == 2) { i--; ri--; } else { j--; rj--; }
            state = open ? 3 : state;
            if (i <= 0 || j <= 0) state = 0;
            act = (state != 0 && ri > 0 && rj > 0);          // the early-terminate test precedes every push (align.cpp:205-207)
            if (act && state != 3) code = dw.load(i, j);     // the gap goes on
        }
        const bool inM = act && state == 3;
        const int it = i - sl, jt = j - sl;
        const bool inb = inM && it >= i0 && jt >= j0;
        const int code_t = inb ? dw.load(it, jt) : 0;
        GACT_CHK(3, !inb || (it >= 1 && it <= cx.TR + 1 && jt >= 1 && jt <= cx.TS + 1));
        const bool match_t = inb && (cx.rb[it] == cx.qs[jt]);
        const unsigned mm = (__ballot_sync(FULL, match_t) >> cx.segbase) & SEGMASK;
        unsigned run;
        // Walking back over t cells lowers the score by at most t * match (mismatch <= 0 only raises it), so while
        // v > LANES * match no cell of this step can be a zero crossing: the run length then does not depend on the scores
        // along the diagonal and the two ballots are independent.  Only a warp-wide segment may branch on it (ballots
        // need the whole warp).
        if (LANES == 32 && v > LANES * ma) {
            const bool isM_t = (sl == 0) || (inb && (code_t >> 2) == 3);
            run = __ballot_sync(FULL, isM_t);
        } else {
            const int below = __popc(mm & ((1u << sl) - 1u));
            const int v_t = v - (below * ma + (sl - below) * mi);
            const bool isM_t = (sl == 0) || (inb && v_t > 0 && (code_t >> 2) == 3);
            run = (__ballot_sync(FULL, isM_t) >> cx.segbase) & SEGMASK;
        }
        int L = __ffs(~run) - 1;
        if (L < 0 || L > LANES - 1) L = LANES - 1;
        L = min(L, min(ri, rj));
        const int codeL = __shfl_sync(FULL, code_t, cx.segbase + L);
        if (inM) {
            GACT_CHK(4, !EMIT || cnt + L <= 2 * et);
            if (EMIT && sl < L) stbuf[cnt + sl] = 3;
            const int bl = __popc(mm & ((1u << L) - 1u));
            const int d = bl * ma + (L - bl) * mi;
            cnt += L; ri -= L; rj -= L;
            v -= d; col += d;
            if (L > 0) pg = 0;
            i -= L; j -= L;
            code = codeL;
            state = (i >= i0 && j >= j0 && v > 0) ? (codeL >> 2) : 0;
        }
        act = (state != 0 && ri > 0 && rj > 0);
    }
    __syncwarp();
    SegTrace t;
    t.cnt = cnt; t.is = et - ri; t.js = et - rj; t.col_score = col; t.first_gap = first_gap; t.last_gap = pg;
    return t;
}

// First-tile pass (align.cpp:173-177,190): score only + position of the LAST maximum in (i outer, j inner)
// order.  Per column and half-word the running maximum and the row where it was last reached are kept
// (VIMNMX.S16x2 with predicate outputs = the reference's `>=` update); columns are compared by (H, i, j).
template <int CS, int LANES, bool LUT>
__device__ __forceinline__ void seg_first_pass(const SegCtx<CS, LANES> &cx, const uint32_t (&q)[CS], int n, int m,
                                               int *max_i, int *max_j)
{
    const int sl = cx.sl, B = cx.B;
    const uint32_t Bp = cx.Bp, ge16 = cx.ge16;
    const bool work = (n > 0 && m > 0);
    const int laststrip = (m > 0) ? (m - 1) / CS : -1;
    const int lastlane = laststrip >> 1;
    int steps = work ? n + 1 + 2 * lastlane : 0;
#pragma unroll
    for (int o = LANES; o < 32; o <<= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, o));
    const uint32_t *rrp = cx.rr - 2 * sl;

    uint32_t Gup[CS], IoUp[CS], IcUp[CS], best[CS], brow[CS];
#pragma unroll
    for (int c = 0; c < CS; c++) {
        Gup[c] = Bp;
        IoUp[c] = pk16(B + cx.gap_open * 16, S16_NEG);
        IcUp[c] = pk16(S16_NEG, S16_NEG);
        best[c] = 0;                               // below every H (H >= B > 0): the first valid cell always updates
        brow[c] = 0;
    }
    uint32_t eG = Bp, eD = pk16(S16_NEG), diag = Bp;
    uint32_t rprev = LUT ? rrp[0] : 0u;
    // no per-lane activity branch (see seg_dp): rows outside 1..n are masked out of the maximum search by `keep`
    for (int k = 1; k <= steps; k++) {
        const uint32_t pack = __byte_perm(eG, eD, 0x7632);
        uint32_t recv = __shfl_up_sync(FULL, pack, 1, LANES);
        if (sl == 0) recv = cx.borderD_raw;
        const uint32_t inG = __byte_perm(recv, eG, 0x5410);
        const uint32_t inD = __byte_perm(recv, eD, 0x5432);
        const int ilo = k - 2 * sl;
        const uint32_t rlo = rrp[k], rhi = rprev;
        // low half: row ilo, high half: row ilo - 1; only rows 1..n are recorded
        const uint32_t keep = (((unsigned)(ilo - 1) < (unsigned)n) ? 0x0000ffffu : 0u) | (((unsigned)(ilo - 2) < (unsigned)n) ? 0xffff0000u : 0u);
        const uint32_t ipair = pk16(ilo, ilo - 1);
        uint32_t hd = diag, dv = inD;
#pragma unroll
        for (int c = 0; c < CS; c++) {
            const uint32_t sc = subst_score<LUT>(q[c], rlo, rhi, cx.ma16, cx.mi16);
            const uint32_t mc = __viaddmax_s16x2(hd, sc, Bp);
            hd = Gup[c];
            const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
            const uint32_t h = __vimax3_s16x2(mc, iv, dv);
            Gup[c] = h;
            const uint32_t mo = (uint32_t)((int)mc * cx.ONE + cx.KO);
            IoUp[c] = mo;
            IcUp[c] = iv;
            dv = __viaddmax_s16x2(dv, ge16, mo);
            bool ph, pl;
            best[c] = __vibmax_s16x2(h & keep, best[c], &ph, &pl);      // pred = (h >= best): last maximum wins
            if (pl) brow[c] = __byte_perm(brow[c], ipair, 0x3254);
            if (ph) brow[c] = __byte_perm(brow[c], ipair, 0x7610);
        }
        eG = Gup[CS - 1];
        eD = dv;
        diag = inG;
        if (LUT) rprev = rlo;
    }
    // best cell of this lane by (H, i, j); i, j <= 1024 take 11 bits each
    long long key = -1;
#pragma unroll
    for (int c = 0; c < CS; c++) {
        const int jl = (2 * sl) * CS + c + 1, jh = (2 * sl + 1) * CS + c + 1;
        const int hl = (int)(best[c] & 0xffffu), hh = (int)(best[c] >> 16);
        const int il = (int)(brow[c] & 0xffffu), ih = (int)(brow[c] >> 16);
        if (work && jl <= m && il >= 1) key = max(key, ((long long)((hl - B) >> 4) << 22) | ((long long)il << 11) | jl);
        if (work && jh <= m && ih >= 1) key = max(key, ((long long)((hh - B) >> 4) << 22) | ((long long)ih << 11) | jh);
    }
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
    if (key < 0) { *max_i = 0; *max_j = 0; }
    else { *max_j = (int)(key & 2047); *max_i = (int)((key >> 11) & 2047); }
}

// ---------------------------------------------------------------------------
// tile kernel: one tile per segment, results + packed states to global memory
template <int CS, int LANES, bool LUT>
__global__ void __launch_bounds__(128, (CS <= 10 ? 4 : 2))
gact_tile_s16h_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                      const int *__restrict__ order, int n_tiles, const EffLen *__restrict__ eff,
                      gact_tile_result *__restrict__ results, uint32_t *__restrict__ states,
                      int pitch_words, int *counter, size_t seq_bytes, uint8_t *gscratch, size_t dir_bytes,
                      const int *__restrict__ n_ptr)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int TPW = 32 / LANES;
    if (n_ptr) n_tiles = *n_ptr;              // list length produced on the device (tiles the inter-task kernel handed back)
    SegCtx<CS, LANES> cx;
    cx.init(P, smem, threadIdx.x >> 5, seq_bytes, gscratch, dir_bytes, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), LUT);
    const int lane = cx.lane, seg = cx.seg, sl = cx.sl;

    for (;;) {
        int t0 = 0;
        if (lane == 0) t0 = atomicAdd(counter, TPW);
        t0 = __shfl_sync(FULL, t0, 0);
        if (t0 >= n_tiles) break;
        const bool valid = (t0 + seg) < n_tiles;
        const int t = valid ? (order ? order[t0 + seg] : t0 + seg) : 0;

        gact_tile_desc d;
        d.ref_off = 0; d.query_off = 0; d.ref_len = 0; d.query_len = 0; d.ref_set = 0; d.query_set = 0; d.reverse = 0; d.first = 0;
        if (valid) d = descs[t];
        int n = d.ref_len, m = d.query_len;
        if (valid && d.first) { n = eff[t].n; m = eff[t].m; }

        if constexpr (LUT) seg_stage_packed<CS, LANES>(cx, P.sets[d.ref_set], P.sets[d.query_set], d.ref_off, d.ref_len, d.query_off,
                                                       d.query_len, d.reverse, n, m);
        else seg_stage<CS, LANES, false>(cx, P.sets[d.ref_set], P.sets[d.query_set], d.ref_off, d.ref_len, d.query_off,
                                         d.query_len, d.reverse, n, m);
        uint32_t q[CS];
        seg_load_q<CS, LANES, LUT>(cx, m, q);
        DirWinH<CS> dw;
        dw.init(cx.dirbase, n, m, P);
        const int corner = seg_dp<CS, LANES, LUT>(cx, q, n, m, dw);
        uint8_t *stbuf = reinterpret_cast<uint8_t *>(cx.rr + 1);  // rows >= 1 of rr[] are dead after the DP (rr[<= 0] must stay sentinel)
        const SegTrace tr = seg_traceback<CS, LANES, true>(cx, dw, n, m, corner, stbuf, 0);
        if (valid) {
            uint32_t *out = states + (size_t)t * pitch_words;
            for (int w = sl; w * 16 < tr.cnt; w += LANES) {
                uint32_t a = 0;
#pragma unroll
                for (int x = 0; x < 16; x++) {
                    const int idx = w * 16 + x;
                    const uint32_t st = (idx < tr.cnt) ? stbuf[idx] : 0u;
                    a |= st << (2 * x);
                }
                out[w] = a;
            }
            if (sl == 0) {
                gact_tile_result r;
                r.score = corner;
                r.max_i = d.first ? n : d.ref_len;
                r.max_j = d.first ? m : d.query_len;
                r.n_states = tr.cnt;
                r.i_steps = tr.is;
                r.j_steps = tr.js;
                results[t] = r;
            }
        }
        __syncwarp();
    }
}

// first-tile kernel: (max_i, max_j) of every first tile
template <int CS, int LANES, bool LUT>
__global__ void __launch_bounds__(128, (CS <= 10 ? 4 : 2))
gact_first_s16h_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                       const int *__restrict__ first_list, int n_first, EffLen *__restrict__ eff, int *counter,
                       size_t seq_bytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int TPW = 32 / LANES;
    SegCtx<CS, LANES> cx;
    cx.init(P, smem, threadIdx.x >> 5, seq_bytes, nullptr, 0, 0, LUT);
    const int lane = cx.lane, seg = cx.seg, sl = cx.sl;

    for (;;) {
        int t0 = 0;
        if (lane == 0) t0 = atomicAdd(counter, TPW);
        t0 = __shfl_sync(FULL, t0, 0);
        if (t0 >= n_first) break;
        const bool valid = (t0 + seg) < n_first;
        const int t = valid ? first_list[t0 + seg] : 0;
        gact_tile_desc d;
        d.ref_off = 0; d.query_off = 0; d.ref_len = 0; d.query_len = 0; d.ref_set = 0; d.query_set = 0; d.reverse = 0; d.first = 0;
        if (valid) d = descs[t];
        const int n = d.ref_len, m = d.query_len;
        if constexpr (LUT) seg_stage_packed<CS, LANES>(cx, P.sets[d.ref_set], P.sets[d.query_set], d.ref_off, d.ref_len, d.query_off,
                                                       d.query_len, d.reverse, n, m);
        else seg_stage<CS, LANES, false>(cx, P.sets[d.ref_set], P.sets[d.query_set], d.ref_off, d.ref_len, d.query_off,
                                         d.query_len, d.reverse, n, m);
        uint32_t q[CS];
        seg_load_q<CS, LANES, LUT>(cx, m, q);
        int mi = 0, mj = 0;
        seg_first_pass<CS, LANES, LUT>(cx, q, n, m, &mi, &mj);
        if (valid && sl == 0) { EffLen e; e.n = mi; e.m = mj; eff[t] = e; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// chain kernel ("next" row 2): the whole GACT() of gact.cpp:48-228 per candidate on
// the device -- left extension, right extension from the first tile's maximum, first-tile threshold,
// total score -- so a candidate costs no host round trip per tile.  A segment walks one candidate's
// tile chain; when it finishes it takes the next candidate from the queue.  The per-column score of
// gact.cpp:197-210 is accumulated inside the traceback (an M run contributes its matches/mismatches,
// a gap column gap_open or gap_extend depending on the previous column), so no state string is kept.
// Only for ACGT-only (2-bit packed) sets: there a '-' column test (gact.cpp:201) can never fire.
struct ChainCall {
    long long ref_start, query_start;     // offsets of the two sequences in their sets
    int ref_len, query_len;               // full sequence lengths
    int ref_pos, query_pos;               // D-SOFT anchor (darwin.cpp:216-224)
    int query_set, pad;
};
struct ChainResult {
    int ab, ae, bb, be, score, first_tile_score, n_tiles, pad;
    long long n_cells;
};

#ifdef GACT_PROF
// per-phase clock accounting of the chain kernel (tools/chain_latency.py, libgact_b200_prof.so only)
__device__ unsigned long long g_chain_prof[8];
#define PROF_DECL long long prof_t0 = clock64(); unsigned long long prof_acc[7] = {0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(slot) do { const long long prof_t1 = clock64(); prof_acc[slot] += (unsigned long long)(prof_t1 - prof_t0); prof_t0 = prof_t1; } while (0)
#define PROF_COUNT(slot, v) prof_acc[slot] += (unsigned long long)(v)
#define PROF_FLUSH(cond) do { if (cond) { for (int x = 0; x < 7; x++) if (prof_acc[x]) atomicAdd(&g_chain_prof[x], prof_acc[x]); } } while (0)
#else
#define PROF_DECL
#define PROF_MARK(slot)
#define PROF_COUNT(slot, v)
#define PROF_FLUSH(cond)
#endif

// Claim bookkeeping of one chain-kernel launch (cleared before every launch).  Calls arrive sorted by
// expected serial length, longest first.  The first claim of every segment is DEALT (a fixed rank per
// segment, so that the longest chains start at once and spread over the machine); later claims take
// the next rank from a queue.  A dealt rank that no segment came for (CTA placement is the hardware's
// choice) is picked up by the scan at the end; `taken` makes every rank run exactly once either way.
//   deal 0: queue only.
//   deal 1: rank = segment-in-CTA * gridDim + blockIdx: consecutive ranks go to different CTAs.
//   deal 2: rank = sm + n_sm * (warp + warps_per_cta * cta_slot_on_that_sm): the 4 * n_sm longest chains each get an
//           SM sub-partition (one warp scheduler and its ALU pipe) of their own -- a lone warp runs a tile in about a third of
//           the time it takes with three neighbours (profiles/r2_chain_latency_anatomy_before.txt).  The SM is
//           identified by %smid, mapped to a dense index in arrival order.
struct ChainAux {
    int q_next;               // queue: next undealt rank - q0
    int scan_next;            // final scan over the dealt ranks [0, q0)
    int n_vsm;                // deal 2: dense SM indices handed out so far
    int pad;
    int vmap[1024];           // deal 2: %smid -> dense index + 1 (0: none yet, -1: being assigned)
    int sm_slots[1024];       // deal 2: CTAs that have arrived per dense SM index
};

__device__ __forceinline__ int chain_claim(ChainAux *aux, int *taken, int n_calls, int q0)
{
    for (;;) {
        const int r = q0 + atomicAdd(&aux->q_next, 1);
        if (r >= n_calls) break;
        return r;                                  // ranks >= q0 are never dealt
    }
    for (;;) {
        const int r = atomicAdd(&aux->scan_next, 1);
        if (r >= q0) return n_calls;
        if (atomicExch(&taken[r], 1) == 0) return r;
    }
}

// SMEMWIN: the direction window lives in shared memory behind the segment's sequence arrays (latency
// mapping, tile_size <= 320: the traceback's dependent loads take ~30 cycles instead of an L2 round trip).
template <int CS, int LANES, bool SMEMWIN>
__global__ void __launch_bounds__(128, (SMEMWIN ? 2 : (CS <= 10 ? 3 : 2)))
gact_chain_s16h_kernel(const __grid_constant__ KParams P, const ChainCall *__restrict__ calls, int n_calls,
                       ChainResult *__restrict__ results, int thr, ChainAux *aux, int *taken, size_t seq_stride,
                       uint8_t *gscratch, size_t dir_bytes, int deal, int q0, int n_sm)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr unsigned SEGBITS = (LANES == 32) ? 0xffffffffu : ((1u << LANES) - 1u);
    SegCtx<CS, LANES> cx;
    cx.template init<SMEMWIN>(P, smem, threadIdx.x >> 5, seq_stride, gscratch, dir_bytes,
                              blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), true);
    const int sl = cx.sl;
    const unsigned segmask = SEGBITS << cx.segbase;
    const int T = P.tile_size;
    const SeqSetDev &rset = P.sets[GACT_SET_REF];

    // candidate state, replicated in every lane of the segment (the locals of GACT(), gact.cpp:51-80)
    int call = -1;
    ChainCall c;
    c.ref_start = 0; c.query_start = 0; c.ref_len = 0; c.query_len = 0; c.ref_pos = 0; c.query_pos = 0; c.query_set = 0; c.pad = 0;
    int rp = 0, qp = 0, rrp = 0, rqp = 0, ab = 0, bb = 0, score = 0, fts = 0, n_tiles = 0;
    long long n_cells = 0;
    int phase = 2, first_tile = 0, prev_gap = 0, anchor_gap = 0, left_any = 0, adv = 1;
    bool alive = true;                       // false once the queue is empty for this segment
    constexpr int SEGS = 32 / LANES;
    const int lseg = (threadIdx.x >> 5) * SEGS + cx.segbase / LANES;
    // the rank dealt to this segment (-1: none)
    int dealt = -1;
    if (deal == 1) dealt = lseg * (int)gridDim.x + (int)blockIdx.x;
    if (deal == 2) {
        __shared__ int s_rank0;
        if (threadIdx.x == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            sm &= 1023u;
            int v = atomicCAS(&aux->vmap[sm], 0, -1);
            if (v == 0) { v = atomicAdd(&aux->n_vsm, 1) + 1; atomicExch(&aux->vmap[sm], v); }
            else while (v < 0) v = atomicAdd(&aux->vmap[sm], 0);          // the assigning thread is already running
            v = (v - 1) & 1023;
            const int slot = atomicAdd(&aux->sm_slots[v], 1);
            s_rank0 = v + n_sm * (int)(blockDim.x >> 5) * SEGS * slot;
        }
        __syncthreads();
        dealt = s_rank0 + n_sm * lseg;
    }
    bool first_claim = deal != 0;
    PROF_DECL;

    for (;;) {
        // ---- next tile of this segment's candidate (loop heads of gact.cpp:82 and :144) ----
        int t_rl = 0, t_ql = 0, reverse = 0;
        long long roff = 0, qoff = 0;
        bool have = false;
        while (alive && !have) {
            if (phase == 2) {
                if (call >= 0 && sl == 0) {
                    ChainResult r;
                    r.ab = ab; r.bb = bb; r.ae = rp; r.be = qp; r.score = score; r.first_tile_score = fts;
                    r.n_tiles = n_tiles; r.pad = 0; r.n_cells = n_cells;
                    results[call] = r;
                }
                int nxt = n_calls;
                if (sl == 0) {
                    if (first_claim && dealt >= 0 && dealt < q0 && atomicExch(&taken[dealt], 1) == 0) nxt = dealt;
                    else nxt = chain_claim(aux, taken, n_calls, q0);
                }
                first_claim = false;
                nxt = __shfl_sync(segmask, nxt, cx.segbase);
                if (nxt >= n_calls) { alive = false; call = -1; break; }
                call = nxt;
                c = calls[call];
                rp = rrp = c.ref_pos; qp = rqp = c.query_pos;
                ab = bb = 0; score = 0; fts = 0; n_tiles = 0; n_cells = 0;
                phase = 0; first_tile = 1; prev_gap = 0; anchor_gap = 0; left_any = 0; adv = 1;
            }
            if (phase == 0) {
                if (rp > 0 && qp > 0 && (adv || first_tile)) {
                    t_rl = rp > T ? T : rp;
                    t_ql = qp > T ? T : qp;
                    roff = c.ref_start + rp - t_rl;
                    qoff = c.query_start + qp - t_ql;
                    reverse = 0;
                    have = true;
                } else {
                    ab = rp; bb = qp; rp = rrp; qp = rqp;          // gact.cpp:136-141
                    phase = 1;
                    prev_gap = left_any ? anchor_gap : 0;
                    adv = 1;
                }
            }
            if (phase == 1 && !have) {
                if (rp < c.ref_len && qp < c.query_len && (adv || first_tile)) {
                    t_rl = (rp + T < c.ref_len) ? T : c.ref_len - rp;
                    t_ql = (qp + T < c.query_len) ? T : c.query_len - qp;
                    roff = c.ref_start + rp;
                    qoff = c.query_start + qp;
                    reverse = 1;
                    have = true;
                } else {
                    phase = 2;
                }
            }
        }
        if (!__any_sync(FULL, have)) break;
        PROF_MARK(0);

        // ---- the tile: stage, (first pass), DP, traceback ----
        int n = have ? t_rl : 0, m = have ? t_ql : 0;
        const SeqSetDev &qset = P.sets[c.query_set];
        seg_stage_packed<CS, LANES>(cx, rset, qset, roff, t_rl, qoff, t_ql, reverse, n, m);
        uint32_t q[CS];
        seg_load_q<CS, LANES, true>(cx, m, q);
        PROF_MARK(1);
        const bool is_first = have && first_tile;
        int mi = n, mj = m;
        if (__any_sync(FULL, is_first)) {
            int fi = 0, fj = 0;
            seg_first_pass<CS, LANES, true>(cx, q, is_first ? n : 0, is_first ? m : 0, &fi, &fj);
            if (is_first) { mi = fi; mj = fj; n = fi; m = fj; }           // sub-tile ending at the last maximum
        }
        PROF_MARK(2);
        DirWinH<CS> dw;
        dw.init(cx.dirbase, n, m, P);
        const int tile_score = seg_dp<CS, LANES, true>(cx, q, n, m, dw);
        PROF_MARK(3);
        const SegTrace tr = seg_traceback<CS, LANES, false>(cx, dw, n, m, tile_score, nullptr, prev_gap);
        PROF_MARK(4);
        PROF_COUNT(6, 1);

        // ---- consume (bodies of the loops at gact.cpp:95-133 and :158-194) ----
        if (have) {
            const bool left = (phase == 0);
            n_tiles++;
            n_cells += (long long)t_rl * t_ql;
            bool skip = false;
            if (first_tile) {
                if (left) { rp = rp - t_rl + mi; qp = qp - t_ql + mj; rrp = rp; rqp = qp; }
                else      { rp = rp + t_rl - mi; qp = qp + t_ql - mj; }
                fts = tile_score;
                if (tile_score < thr) {                                    // gact.cpp:107-109 / :168-170
                    if (left) { ab = rp; bb = qp; rp = rrp; qp = rqp; phase = 1; prev_gap = 0; adv = 1; }
                    else      { phase = 2; }
                    skip = true;
                }
            }
            if (!skip) {
                if (tr.cnt > 0) {
                    first_tile = 0;
                    score += tr.col_score;
                    prev_gap = tr.last_gap;
                    if (left && !left_any) { left_any = 1; anchor_gap = tr.first_gap; }
                }
                if (left) { rp -= tr.is; qp -= tr.js; } else { rp += tr.is; qp += tr.js; }
                if (first_tile && tr.cnt == 0) { first_tile = 0; adv = 0; }
                else adv = (tr.is > 0 && tr.js > 0) ? 1 : 0;
            }
        }
        __syncwarp();
        PROF_MARK(5);
    }
    PROF_FLUSH(cx.lane == 0);
}

// ---------------------------------------------------------------------------
// host side
struct S16HPlan {
    bool ok = false;
    int CS = 0, lanes = 16, win_rows = 0, win_lanes = 0, warps_per_cta = 4, ctas = 0, bias = 0;
    bool lut_ok = false;
    bool smem_window = false;            // latency chain plan: direction window in shared memory
    size_t seq_bytes = 0, smem = 0, dir_bytes = 0;
    uint8_t *d_scratch = nullptr;        // GACT_MAX_INFLIGHT regions of scratch_bytes: kernels of consecutive batches may overlap
    size_t scratch_bytes = 0;
    int tpw() const { return 32 / lanes; }
    int cols() const { return CS * 2 * lanes; }                      // widest query window of a tile
    int slots() const { return ctas * warps_per_cta * tpw(); }       // tiles / chains resident at once
};

typedef void (*s16h_fn)(const KParams, const gact_tile_desc *, const int *, int, const EffLen *, gact_tile_result *,
                        uint32_t *, int, int *, size_t, uint8_t *, size_t, const int *);
typedef void (*s16h_first_fn)(const KParams, const gact_tile_desc *, const int *, int, EffLen *, int *, size_t);
typedef void (*s16h_chain_fn)(const KParams, const ChainCall *, int, ChainResult *, int, ChainAux *, int *, size_t, uint8_t *,
                              size_t, int, int, int);

// (strip width, lanes per tile): T <= 256: (8,16), <= 320: (10,16), <= 512: (8,32), <= 1024: (16,32)
#define S16H_DISPATCH(CSV, LANESV, EXPR_CS_LANES)                         \
    do {                                                                  \
        if ((CSV) == 8 && (LANESV) == 16) { EXPR_CS_LANES(8, 16); }       \
        else if ((CSV) == 10 && (LANESV) == 16) { EXPR_CS_LANES(10, 16); } \
        else if ((CSV) == 8 && (LANESV) == 32) { EXPR_CS_LANES(8, 32); }  \
        else { EXPR_CS_LANES(16, 32); }                                   \
    } while (0)

inline s16h_fn s16h_pick(int CS, int lanes, bool lut)
{
    // narrow mappings: score-table tile kernel only
    if (lanes == 16 && CS == 5) return lut ? gact_tile_s16h_kernel<5, 16, true> : nullptr;
    if (lanes == 16 && CS == 4) return lut ? gact_tile_s16h_kernel<4, 16, true> : nullptr;
    s16h_fn f = nullptr;
#define S16H_X(C, L) f = lut ? gact_tile_s16h_kernel<C, L, true> : gact_tile_s16h_kernel<C, L, false>
    S16H_DISPATCH(CS, lanes, S16H_X);
#undef S16H_X
    return f;
}
inline s16h_first_fn s16h_pick_first(int CS, int lanes, bool lut)
{
    s16h_first_fn f = nullptr;
#define S16H_X(C, L) f = lut ? gact_first_s16h_kernel<C, L, true> : gact_first_s16h_kernel<C, L, false>
    S16H_DISPATCH(CS, lanes, S16H_X);
#undef S16H_X
    return f;
}
inline s16h_chain_fn s16h_pick_chain(int CS, int lanes, bool smem_window)
{
    // latency variants of the chain kernel: one tile per warp also for tile_size <= 320, window in shared memory
    if (smem_window) return CS == 4 ? gact_chain_s16h_kernel<4, 32, true> : gact_chain_s16h_kernel<5, 32, true>;
    s16h_chain_fn f = nullptr;
#define S16H_X(C, L) f = gact_chain_s16h_kernel<C, L, false>
    S16H_DISPATCH(CS, lanes, S16H_X);
#undef S16H_X
    return f;
}
inline size_t s16h_dir_bytes(int CS, int rows, int lanes)
{
    switch (CS) {
        case 4: return DirWinH<4>::bytes(rows, lanes);
        case 5: return DirWinH<5>::bytes(rows, lanes);
        case 8: return DirWinH<8>::bytes(rows, lanes);
        case 10: return DirWinH<10>::bytes(rows, lanes);
        default: return DirWinH<16>::bytes(rows, lanes);
    }
}

inline void s16h_free_plan(S16HPlan *pl)
{
    if (pl->d_scratch) cudaFree(pl->d_scratch);
    pl->d_scratch = nullptr;
}

// Two tiles per warp for tile_size <= 320, one tile per warp up to 1024.
// latency = true: plan for the chain kernel only, tile_size <= 320: one tile per warp (32 lanes, strips of 4 / 5
// columns) with the direction window in shared memory and two CTAs (8 warps) per SM at most: used when a shard has
// few candidates, so that the longest read's serial tile chain finishes sooner.
// narrow = true: plan for the tile kernel only, tile_size <= 320: the two-tiles-per-warp mapping with strips of half
// the width (4 / 5 columns) for tiles whose query window is at most narrow_cols() wide.
inline int s16h_make_plan(const gact_params &p, int num_sms, int warps_per_sm, S16HPlan *pl, bool latency = false,
                          bool narrow = false)
{
    s16h_free_plan(pl);
    *pl = S16HPlan();
    const int T = p.tile_size, et = p.tile_size - p.tile_overlap;
    const int bias = 16 * (-p.gap_open + 2);
    pl->bias = bias;
    pl->lut_ok = (p.match * 16 <= 127 && p.mismatch * 16 >= -128);
    const long hi = (long)T * (p.match > 0 ? p.match : 0) * 16 + 16 + bias;
    if (hi > 30000 || p.mismatch > 0 || p.match < 0 || p.gap_open < -500 || p.gap_extend < -500 || p.mismatch < -1000)
        return 0;
    int CS, lanes;
    if (T <= 256) { CS = 8; lanes = 16; } else if (T <= 320) { CS = 10; lanes = 16; }
    else if (T <= 512) { CS = 8; lanes = 32; } else { CS = 16; lanes = 32; }
    if (latency) {
        if (T > 320 || !pl->lut_ok) return 0;        // larger tiles already run one per warp; their window does not fit
        CS = (T <= 256) ? 4 : 5; lanes = 32;
        pl->smem_window = true;
    }
    if (narrow) {
        if (T > 320 || T < 64 || !pl->lut_ok) return 0;
        CS = (T <= 256) ? 4 : 5; lanes = 16;
    }
    pl->CS = CS; pl->lanes = lanes;
    pl->win_rows = (et + 1 < T) ? et + 1 : T;
    int wl = et / (2 * CS) + 2;
    pl->win_lanes = wl > lanes ? lanes : wl;
    pl->seq_bytes = s16h_seq_bytes(CS, lanes);
    pl->dir_bytes = s16h_dir_bytes(CS, pl->win_rows, pl->win_lanes);
    int wps = warps_per_sm > 0 ? warps_per_sm : 16;
    const int wps_max = latency ? 8 : (CS <= 10) ? 16 : 8;
    if (wps > wps_max) wps = wps_max;
    pl->warps_per_cta = 4;
    const int c = (wps + 3) / 4;
    pl->ctas = c * num_sms;
    if (latency) {
        pl->seq_bytes += pl->dir_bytes;               // stride between the warps' carve-outs: sequences, then the window
        pl->smem = (size_t)pl->warps_per_cta * pl->seq_bytes;
        if (pl->smem > 113 * 1024) return 0;          // two CTAs per SM must fit
    } else {
        pl->smem = (size_t)pl->warps_per_cta * pl->tpw() * pl->seq_bytes;
        pl->scratch_bytes = (size_t)pl->ctas * pl->warps_per_cta * pl->tpw() * pl->dir_bytes;
        if (cudaMalloc(&pl->d_scratch, GACT_MAX_INFLIGHT * pl->scratch_bytes) != cudaSuccess) { cudaGetLastError(); pl->d_scratch = nullptr; return 0; }
    }
    if (narrow) {
        if (cudaFuncSetAttribute((const void *)s16h_pick(CS, lanes, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pl->smem) != cudaSuccess)
            return -1;
        pl->ok = true;
        return 0;
    }
    for (int lut = 0; lut < 2 && !latency; lut++)
        if (cudaFuncSetAttribute((const void *)s16h_pick(CS, lanes, lut != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pl->smem) != cudaSuccess ||
            cudaFuncSetAttribute((const void *)s16h_pick_first(CS, lanes, lut != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pl->smem) != cudaSuccess)
            return -1;
    if (cudaFuncSetAttribute((const void *)s16h_pick_chain(CS, lanes, pl->smem_window), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)pl->smem) != cudaSuccess)
        return -1;
    pl->ok = true;
    return 0;
}

inline int s16h_grid(const S16HPlan &pl, int n_items)
{
    const int per_cta = pl.warps_per_cta * pl.tpw();
    const int need = (n_items + per_cta - 1) / per_cta;
    return need < pl.ctas ? need : pl.ctas;
}

// One chain-kernel launch.  aux / taken: claim bookkeeping of this launch (cleared here); grid_ctas: CTAs to launch
// (<= pl.ctas); deal: see ChainAux; scratch_region: which third of the plan's direction-window scratch to use.
inline void s16h_launch_chain(const S16HPlan &pl, KParams kp, const ChainCall *calls, int n_calls, ChainResult *results,
                              int thr, ChainAux *aux, int *taken, cudaStream_t st, int deal, int grid_ctas, int num_sms,
                              int scratch_region = 0)
{
    kp.win_rows = pl.win_rows;
    kp.win_lanes = pl.win_lanes;
    kp.s16_bias = pl.bias;
    kp.one = 1;
    const int per_cta = pl.warps_per_cta * pl.tpw();
    int grid = grid_ctas < 1 ? 1 : grid_ctas;
    if (grid > pl.ctas) grid = pl.ctas;
    const int need = (n_calls + per_cta - 1) / per_cta;
    if (need < grid) grid = need;
    const long long dealt = deal ? (long long)grid * per_cta : 0;
    const int q0 = (int)(dealt < n_calls ? dealt : n_calls);
    cudaMemsetAsync(aux, 0, sizeof(ChainAux), st);
    if (q0 > 0) cudaMemsetAsync(taken, 0, (size_t)q0 * sizeof(int), st);
    s16h_pick_chain(pl.CS, pl.lanes, pl.smem_window)<<<grid, pl.warps_per_cta * 32, pl.smem, st>>>(
        kp, calls, n_calls, results, thr, aux, taken, pl.seq_bytes,
        pl.d_scratch ? pl.d_scratch + (size_t)scratch_region * pl.scratch_bytes : nullptr, pl.dir_bytes, deal, q0,
        grid < num_sms ? grid : num_sms);
}

// lut: score with the one-PRMT table from the packed words (tiles whose query window has no exception) or compare raw bytes
inline void s16h_launch_first(const S16HPlan &pl, KParams kp, const gact_tile_desc *descs, const int *first_list,
                              int n_first, EffLen *eff, int *counter, cudaStream_t st, bool lut)
{
    kp.s16_bias = pl.bias;
    kp.one = 1;
    s16h_pick_first(pl.CS, pl.lanes, lut)<<<s16h_grid(pl, n_first), pl.warps_per_cta * 32, pl.smem, st>>>(
        kp, descs, first_list, n_first, eff, counter, pl.seq_bytes);
}

inline void s16h_launch(const S16HPlan &pl, KParams kp, const gact_tile_desc *descs, const int *order, int n,
                        const EffLen *eff, gact_tile_result *results, uint32_t *states, int pitch_words, int *counter,
                        cudaStream_t st, int scratch_region, bool lut, const int *n_ptr = nullptr)
{
    kp.win_rows = pl.win_rows;
    kp.win_lanes = pl.win_lanes;
    kp.s16_bias = pl.bias;
    kp.one = 1;
    // n_ptr: the list length is on the device (n is then only an upper bound used to size the grid)
    s16h_pick(pl.CS, pl.lanes, lut)<<<s16h_grid(pl, n), pl.warps_per_cta * 32, pl.smem, st>>>(
        kp, descs, order, n, eff, results, states, pitch_words, counter, pl.seq_bytes,
        pl.d_scratch + (size_t)scratch_region * pl.scratch_bytes, pl.dir_bytes, n_ptr);
}

}  // namespace gact
