// gact_kernels_s16.cuh -- packed s16x2 DPX GACT tile kernel (one warp per tile).
//
// Same contract as the int32 kernel (AlignWithBT, align.cpp:60-233), half the
// instructions per cell: every 32-bit register holds two DP cells.
//
// Mapping.  The tile's query columns are cut into 64 strips of CS columns;
// lane L owns strips 2L (low half-word) and 2L+1 (high half-word).  In step k
// the low half works on reference row k-2L and the high half on row k-2L-1, so
// the high strip's left neighbour (the lane's own low strip, same row) was
// finished one step earlier and the low strip's left neighbour (lane L-1's
// high strip) likewise; one SHFL per step moves the strip edge.
//
// Arithmetic.  Scores are scaled by 16 and the low nibble of every value is a
// TAG, so that the max instructions themselves produce the traceback pointer
// of align.cpp:162-171:
//      M operand  ....1100     I: open ....1010 / extend ....1000
//                              D: open ....0101 / extend ....0100
//   * I = VIADDMNMX.S16x2(Iclean_up, ge, Mup+go-2)  -> bit 1 = (ins_open >= ins_extend)
//   * D = VIADDMNMX.S16x2(Dclean_lf, ge, Mlf+go-7)  -> bit 0 = (del_open >= del_extend)
//   * G = VIMNMX3.S16x2(M, I, D)                    -> bits 3:2 = 3/2/1 = M/I/D with the
//     reference's tie order M >= I >= D (equal scores are decided by the tag)
//   * M = VIADDMNMX.S16x2.RELU(Gdiag, s, 0), s from one HSET2 + LOP3 (raw byte equality)
// The ZERO state (align.cpp:166-168, H <= 0) is not stored: the traceback
// tracks the score of the cell it stands on and stops when it reaches 0.
// Border rows/columns are not special-cased: the high half runs a pseudo row 0
// against a sentinel base that reproduces the border values exactly.
#pragma once
#include <cuda_fp16.h>
#include "gact_common.cuh"

namespace gact {

__device__ __forceinline__ uint32_t pk16(int x) { return ((uint32_t)x & 0xffffu) | ((uint32_t)x << 16); }
__device__ __forceinline__ uint32_t pk16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
// base byte -> a normal fp16 bit pattern (2.0 + b/512): HSET2 compares them exactly
__device__ __forceinline__ uint32_t enc_base(int b) { return 0x4000u | (uint32_t)b; }
static constexpr uint32_t SENT_R = 0x4100u;      // reference sentinel (row 0 / past the end)
static constexpr uint32_t SENT_Q = 0x4200u;      // query sentinel (columns past the end)
static constexpr int S16_NEG = -16384;           // "minus infinity" in the x16 tagged domain

template <int CS>
struct DirWin16 {
    static constexpr int NW = CS / 4;                 // 32-bit words per lane-step (4 columns x 2 strips each)
    static constexpr bool HAS_B = (CS % 4) != 0;      // CS = 5: one extra byte (1 column x 2 strips)
    int i0, lane0, nl;
    uint32_t *w;
    uint8_t *b;
    __device__ __forceinline__ void init(void *base, int n, int m, const KParams &P)
    {
        i0 = max(n - P.et, 1);
        const int j0 = max(m - P.et, 1);
        lane0 = ((j0 - 1) / CS) >> 1;
        nl = P.win_lanes;
        w = reinterpret_cast<uint32_t *>(base);
        b = reinterpret_cast<uint8_t *>(w + (size_t)(P.win_rows + 1) * nl * NW);
    }
    // 4-bit code of cell (i, j): bits 3:2 = M/I/D tag, bit 1 = ins flag, bit 0 = del flag
    __device__ __forceinline__ int load(int i, int j) const
    {
        const int s = (j - 1) / CS, c = (j - 1) - s * CS;
        const int lane = s >> 1, half = s & 1;
        const int e = (i + half - i0) * nl + (lane - lane0);
        if (c < NW * 4) return (w[e * NW + (c >> 2)] >> (16 * half + 4 * (3 - (c & 3)))) & 15;
        return (b[e] >> (4 * half)) & 15;
    }
    static __host__ __device__ size_t bytes(int win_rows, int win_lanes)
    {
        size_t s = (size_t)(win_rows + 1) * win_lanes * (NW * 4 + (HAS_B ? 1 : 0));
        return (s + 15) & ~(size_t)15;
    }
};

// traceback, align.cpp:185-230, one lane.  v = score of the current state's cell.
template <int CS>
__device__ __forceinline__ void traceback_tile16(const DirWin16<CS> &dw, const uint32_t *rr, const uint16_t *qs,
                                                 int n, int m, int score, const KParams &P,
                                                 uint32_t *states, gact_tile_result *res,
                                                 int out_max_i, int out_max_j)
{
    int i = n, j = m, is = 0, js = 0, cnt = 0, v = score;
    const int et = P.et;
    uint32_t acc = 0;
    int state = (i > 0 && j > 0 && v > 0) ? (dw.load(i, j) >> 2) : 0;
    while (state != 0) {
        if (is >= et || js >= et) break;
        if (i <= 0 || j <= 0) break;
        acc |= (uint32_t)state << (2 * (cnt & 15));
        if ((cnt & 15) == 15) { states[cnt >> 4] = acc; acc = 0; }
        cnt++;
        if (state == 3) {
            const int s = ((rr[i] & 0xffffu) == qs[j]) ? P.match : P.mismatch;
            v -= s;                                   // H[i-1][j-1] = M[i][j] - s   (M > 0 on the path)
            i--; j--; is++; js++;
            state = (i > 0 && j > 0 && v > 0) ? (dw.load(i, j) >> 2) : 0;
        } else if (state == 2) {
            const bool open = dw.load(i, j) & 2;
            v -= open ? P.gap_open : P.gap_extend;
            state = open ? 3 : 2;
            i--; is++;
        } else {
            const bool open = dw.load(i, j) & 1;
            v -= open ? P.gap_open : P.gap_extend;
            state = open ? 3 : 1;
            j--; js++;
        }
    }
    if (cnt & 15) states[cnt >> 4] = acc;
    res->score = score;
    res->max_i = out_max_i;
    res->max_j = out_max_j;
    res->n_states = cnt;
    res->i_steps = is;
    res->j_steps = js;
}

template <int CS>
__global__ void __launch_bounds__(256)
gact_tile_s16_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                     int n_tiles, const EffLen *__restrict__ eff,
                     gact_tile_result *__restrict__ results, uint32_t *__restrict__ states,
                     int pitch_words, int *counter, size_t per_warp_bytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int TS = CS * 64;
    constexpr int NW = DirWin16<CS>::NW;

    // per-warp carve-out: rr[TS+2] words | qs[TS+2] halves | direction window
    uint8_t *my = smem + (size_t)warp * per_warp_bytes;
    uint32_t *rr = reinterpret_cast<uint32_t *>(my);                 // rr[i] = enc(R[i]) | enc(R[i-1]) << 16
    uint16_t *qs = reinterpret_cast<uint16_t *>(my + (TS + 2) * 4);  // qs[j] = enc(Q[j])
    void *dirbase = my + (TS + 2) * 4 + (((TS + 2) * 2 + 15) & ~15);

    const uint32_t ma16 = pk16(P.match * 16), mi16 = pk16(P.mismatch * 16);
    const uint32_t ge16 = pk16(P.gap_extend * 16);
    const uint32_t goI = pk16(P.gap_open * 16 - 2);      // M tag 1100 -> I-open tag 1010
    const uint32_t goD = pk16(P.gap_open * 16 - 7);      // M tag 1100 -> D-open tag 0101
    const uint32_t CLEAN = 0xfff0fff0u, TAGM = 0x000c000cu;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_tiles) break;

        const gact_tile_desc d = descs[t];
        int n = d.ref_len, m = d.query_len;
        if (d.first) { n = eff[t].n; m = eff[t].m; }
        const SeqSetDev &rset = P.sets[d.ref_set];
        const SeqSetDev &qset = P.sets[d.query_set];

        __syncwarp();
        for (int x = lane; x <= n + 1; x += 32) {
            const uint32_t cur = (x >= 1 && x <= n) ? enc_base(tile_base(rset, d.ref_off, d.ref_len, d.reverse, x)) : SENT_R;
            const uint32_t prv = (x >= 2 && x <= n + 1) ? enc_base(tile_base(rset, d.ref_off, d.ref_len, d.reverse, x - 1)) : SENT_R;
            rr[x] = cur | (prv << 16);
        }
        for (int x = lane; x <= m; x += 32)
            qs[x] = (x >= 1) ? (uint16_t)enc_base(tile_base(qset, d.query_off, d.query_len, d.reverse, x)) : (uint16_t)SENT_Q;
        __syncwarp();

        uint32_t q[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) {
            const int jl = (2 * lane) * CS + c + 1, jh = (2 * lane + 1) * CS + c + 1;
            q[c] = (jl <= m ? (uint32_t)qs[jl] : SENT_Q + c) | ((jh <= m ? (uint32_t)qs[jh] : SENT_Q + c) << 16);
        }

        DirWin16<CS> dw;
        dw.init(dirbase, n, m, P);
        const int laststrip = (m > 0) ? (m - 1) / CS : -1;
        const int lastlane = laststrip >> 1;
        // where the corner H[n][m] will appear
        const int c_lane = lastlane, c_half = laststrip & 1, c_col = (m > 0) ? (m - 1) - laststrip * CS : 0;

        // state before the lane's first step: low half = border row 0 already applied,
        // high half = "row -1" (its first step is the pseudo row 0)
        uint32_t Gup[CS], IoUp[CS], IcUp[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) {
            Gup[c] = 0;
            IoUp[c] = pk16(P.gap_open * 16 + 10, S16_NEG);      // (0|1100) + go16 - 2
            IcUp[c] = pk16(S16_NEG + 8, S16_NEG + 8);
        }
        uint32_t eG = 0;                                   // G of my last column (row just finished)
        uint32_t eD = pk16(S16_NEG + 4);                   // D value for the column right of my strip
        uint32_t diag = 0;                                 // G[i-1][first column - 1] for both halves
        const uint32_t borderD = pk16(P.gap_open * 16 + 5);  // D[i][1] = 0 + go, open flag set
        int corner16 = 0;

        const int steps = (n > 0 && m > 0) ? n + 1 + 2 * lastlane : 0;
        for (int k = 1; k <= steps; k++) {
            const int ilo = k - 2 * lane;                  // low half: row ilo, high half: row ilo-1
            // edge of the strip to the left: low half <- lane-1's high strip, high half <- my low strip
            const uint32_t pack = __byte_perm(eG, eD, 0x7632);          // (eG.hi, eD.hi)
            uint32_t recv = __shfl_up_sync(FULL, pack, 1);
            if (lane == 0) recv = (borderD << 16);                       // G border 0, D border
            const uint32_t inG = __byte_perm(recv, eG, 0x5410);          // lo: recv.lo16 (G), hi: my eG.lo
            const uint32_t inD = __byte_perm(recv, eD, 0x5432);          // lo: recv.hi16 (D), hi: my eD.lo
            if (ilo >= 1 && ilo <= n + 1 && lane <= lastlane) {
                const uint32_t rp = rr[ilo];
                const __half2 rh = *reinterpret_cast<const __half2 *>(&rp);
                uint32_t hd = diag, dv = inD;
                uint32_t acc[NW + 1];
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const __half2 qh = *reinterpret_cast<const __half2 *>(&q[c]);
                    const uint32_t eq = __heq2_mask(qh, rh);
                    const uint32_t s = (eq & ma16) | (~eq & mi16);
                    const uint32_t mraw = __viaddmax_s16x2_relu(hd, s, 0);
                    const uint32_t mt = (mraw & CLEAN) | TAGM;
                    hd = Gup[c];
                    const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
                    const uint32_t g = __vimax3_s16x2(mt, iv, dv);
                    const uint32_t code = (g & 0x000c000cu) | ((iv | dv) & 0x00030003u);
                    if ((c & 3) == 0) acc[c >> 2] = code; else acc[c >> 2] = acc[c >> 2] * 16u + code;
                    Gup[c] = g;
                    IoUp[c] = __vadd2(mt, goI);
                    IcUp[c] = iv & 0xfffdfffdu;
                    // D of the next column (or of the neighbouring strip's first column)
                    dv = __viaddmax_s16x2(dv & 0xfffefffeu, ge16, __vadd2(mt, goD));
                }
                eG = Gup[CS - 1];
                eD = dv;
                diag = inG;
                if (lane == c_lane) {
                    const int irow = ilo - c_half;
                    if (irow == n) {
                        uint32_t gsel = 0;
#pragma unroll
                        for (int c = 0; c < CS; c++) if (c == c_col) gsel = Gup[c];
                        corner16 = (int)(short)(c_half ? (gsel >> 16) : (gsel & 0xffffu));
                    }
                }
                if (ilo >= dw.i0 && lane >= dw.lane0) {
                    const int e = (ilo - dw.i0) * dw.nl + (lane - dw.lane0);
#pragma unroll
                    for (int x = 0; x < NW; x++) dw.w[e * NW + x] = acc[x];
                    if (DirWin16<CS>::HAS_B) dw.b[e] = (uint8_t)((acc[NW] & 15u) | ((acc[NW] >> 12) & 0xf0u));
                }
            }
        }
        int corner = __shfl_sync(FULL, corner16, max(c_lane, 0)) >> 4;
        if (n == 0 || m == 0) corner = 0;
        __syncwarp();
        if (lane == 0) {
            traceback_tile16<CS>(dw, rr, qs, n, m, corner, P, states + (size_t)t * pitch_words, &results[t],
                                 d.first ? n : d.ref_len, d.first ? m : d.query_len);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host-side planning / launch
template <int CS>
inline size_t s16_warp_bytes(int win_rows, int win_lanes)
{
    constexpr int TS = CS * 64;
    return (size_t)(TS + 2) * 4 + (((TS + 2) * 2 + 15) & ~15) + DirWin16<CS>::bytes(win_rows, win_lanes);
}

typedef void (*s16_fn)(const KParams, const gact_tile_desc *, int, const EffLen *, gact_tile_result *,
                       uint32_t *, int, int *, size_t);
inline s16_fn s16_pick(int CS)
{
    switch (CS) {
        case 4: return gact_tile_s16_kernel<4>;
        case 5: return gact_tile_s16_kernel<5>;
        case 8: return gact_tile_s16_kernel<8>;
        default: return nullptr;
    }
}

// win_rows/win_lanes for the s16 kernel are derived here (they differ from the int32 kernel's).
struct S16Plan {
    bool ok = false;
    int CS = 0, win_rows = 0, win_lanes = 0, warps_per_cta = 0, ctas = 0;
    size_t per_warp_bytes = 0, smem = 0;
};

inline int s16_make_plan(const gact_params &p, int num_sms, S16Plan *pl)
{
    *pl = S16Plan();
    const int T = p.tile_size, et = p.tile_size - p.tile_overlap;
    // value range of the x16 tagged domain and the pseudo-row trick
    const long hi = (long)T * (p.match > 0 ? p.match : 0) * 16 + 16;
    if (hi > 30000 || p.mismatch > 0 || p.match < 0 || p.gap_open < -500 || p.gap_extend < -500 || p.mismatch < -1000)
        return 0;
    int CS;
    if (T <= 256) CS = 4; else if (T <= 320) CS = 5; else if (T <= 512) CS = 8; else return 0;
    pl->CS = CS;
    pl->win_rows = (et + 1 < T) ? et + 1 : T;
    int wl = et / (2 * CS) + 2;
    pl->win_lanes = wl > 32 ? 32 : wl;
    size_t pw = 0;
    switch (CS) {
        case 4: pw = s16_warp_bytes<4>(pl->win_rows, pl->win_lanes); break;
        case 5: pw = s16_warp_bytes<5>(pl->win_rows, pl->win_lanes); break;
        default: pw = s16_warp_bytes<8>(pl->win_rows, pl->win_lanes); break;
    }
    pw = (pw + 15) & ~(size_t)15;
    pl->per_warp_bytes = pw;
    const size_t SM = 228 * 1024, CTA_MAX = 227 * 1024;
    int best_w = 0, best_c = 0, best_total = 0;
    for (int w = 1; w <= 8; w++) {
        const size_t cta = (size_t)w * pw;
        if (cta > CTA_MAX) break;
        int c = (int)(SM / (cta + 1024));
        if (c > 16) c = 16;
        if (c * w > 48) c = 48 / w;
        if (c < 1) continue;
        if (c * w > best_total || (c * w == best_total && w > best_w)) { best_total = c * w; best_w = w; best_c = c; }
    }
    if (best_total < 4) return 0;              // window too large for shared memory: int32/L2-scratch kernel instead
    pl->warps_per_cta = best_w;
    pl->ctas = best_c * num_sms;
    pl->smem = (size_t)best_w * pw;
    if (cudaFuncSetAttribute((const void *)s16_pick(CS), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem) != cudaSuccess)
        return -1;
    pl->ok = true;
    return 0;
}

inline void s16_launch(const S16Plan &pl, KParams kp, const gact_tile_desc *descs, int n, const EffLen *eff,
                       gact_tile_result *results, uint32_t *states, int pitch_words, int *counter, cudaStream_t st)
{
    kp.win_rows = pl.win_rows;
    kp.win_lanes = pl.win_lanes;
    int ctas = pl.ctas;
    const int need = (n + pl.warps_per_cta - 1) / pl.warps_per_cta;
    if (need < ctas) ctas = need;
    s16_pick(pl.CS)<<<ctas, pl.warps_per_cta * 32, pl.smem, st>>>(kp, descs, n, eff, results, states, pitch_words,
                                                                  counter, pl.per_warp_bytes);
}

}  // namespace gact
