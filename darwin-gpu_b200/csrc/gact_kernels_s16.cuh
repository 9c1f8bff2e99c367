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

// traceback, align.cpp:185-230, by the whole warp.  All lanes hold the same cursor
// (i, j, state, v = score of the current state's cell, remaining step budgets).
//   * state M: the 32 lanes look at the 32 cells down the diagonal at once -- lane t loads the
//     code of cell (i-t, j-t) and the match bit of that cell; the scores along the diagonal follow
//     from a ballot + popc, so the length of the M run (the common case: ~85 % of all states) is
//     one ffs away.  Up to 31 states are emitted per iteration.
//   * states I / D: one step per iteration (gap runs are short).
// States are first written one byte each into stbuf (shared memory), then packed 16 per word.
template <int CS>
__device__ __forceinline__ void traceback_tile16(const DirWin16<CS> &dw, const uint16_t *rb, const uint16_t *qs,
                                                 uint8_t *stbuf, int lane,
                                                 int n, int m, int score, const KParams &P,
                                                 uint32_t *states, gact_tile_result *res,
                                                 int out_max_i, int out_max_j)
{
    const int et = P.et, ma = P.match, mi = P.mismatch, go = P.gap_open, ge = P.gap_extend;
    const int i0 = dw.i0, j0 = max(m - et, 1);
    int i = n, j = m, cnt = 0, v = score;
    int ri = et, rj = et;                       // remaining step budget per dimension
    int state = (i > 0 && j > 0 && v > 0) ? (dw.load(i, j) >> 2) : 0;
    while (state != 0 && ri > 0 && rj > 0) {
        if (state == 3) {
            const int it = i - lane, jt = j - lane;
            const bool inb = (it >= i0 && jt >= j0);
            const int code_t = inb ? dw.load(it, jt) : 0;
            const bool match_t = inb && (rb[it] == qs[jt]);
            const unsigned mm = __ballot_sync(FULL, match_t);
            const int below = __popc(mm & ((1u << lane) - 1u));
            const int v_t = v - (below * ma + (lane - below) * mi);        // H of cell t, if cells 0..t-1 are all M
            const bool isM_t = (lane == 0) || (inb && v_t > 0 && (code_t >> 2) == 3);
            const unsigned run = __ballot_sync(FULL, isM_t);
            int L = __ffs(~run) - 1;                                        // leading M cells
            if (L < 0 || L > 31) L = 31;                                    // cell L must be covered by lane L
            L = min(L, min(ri, rj));
            if (lane < L) stbuf[cnt + lane] = 3;
            cnt += L; ri -= L; rj -= L;
            const int bl = __popc(mm & ((1u << L) - 1u));
            v -= bl * ma + (L - bl) * mi;
            i -= L; j -= L;
            const int codeL = __shfl_sync(FULL, code_t, L);
            state = (i >= i0 && j >= j0 && v > 0) ? (codeL >> 2) : 0;
        } else {
            const int code = dw.load(i, j);
            const bool open = (state == 2) ? (code & 2) : (code & 1);
            if (lane == 0) stbuf[cnt] = (uint8_t)state;
            cnt++;
            v -= open ? go : ge;
            if (state == 2) { i--; ri--; } else { j--; rj--; }
            state = open ? 3 : state;
            if (i <= 0 || j <= 0) state = 0;                                // unreachable for gap scores <= 0
        }
    }
    __syncwarp();
    for (int w = lane; w * 16 < cnt; w += 32) {
        uint32_t acc = 0;
#pragma unroll
        for (int x = 0; x < 16; x++) {
            const int idx = w * 16 + x;
            const uint32_t st = (idx < cnt) ? stbuf[idx] : 0u;
            acc |= st << (2 * x);
        }
        states[w] = acc;
    }
    if (lane == 0) {
        res->score = score;
        res->max_i = out_max_i;
        res->max_j = out_max_j;
        res->n_states = cnt;
        res->i_steps = et - ri;
        res->j_steps = et - rj;
    }
}

// substitution score of one column pair.
//   LUT mode (both sets 2-bit packed, |score*16| < 128): one PRMT -- the row registers hold a
//     4-byte table (score*16 per query code) for the low and the high strip's reference base,
//     the per-column selector picks byte [code] and its sign extension for each half;
//   general mode: HSET2 equality mask on fp16-encoded raw bytes + one LOP3 select.
template <bool LUT>
__device__ __forceinline__ uint32_t subst_score(uint32_t qc, uint32_t rlo, uint32_t rhi, uint32_t ma16, uint32_t mi16)
{
    if (LUT) {
        // prmt in its default mode: selector bit 3 of a nibble replicates the selected byte's sign
        // (__byte_perm would mask that bit away)
        uint32_t r;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(rlo), "r"(rhi), "r"(qc));
        return r;
    }
    const __half2 qh = *reinterpret_cast<const __half2 *>(&qc);
    const __half2 rh = *reinterpret_cast<const __half2 *>(&rlo);
    const uint32_t eq = __heq2_mask(qh, rh);
    return (eq & ma16) | (~eq & mi16);
}

// DIRG = false: the direction window lives in the warp's shared-memory carve-out;
// DIRG = true : it lives in a per-warp scratch area in global memory that the warp rewrites for
//               every tile (it stays L2-resident); shared memory then only holds the sequences,
//               so occupancy is no longer bounded by the window size (needed for tile_size 1024).
template <int CS, bool LUT, bool DIRG>
__global__ void __launch_bounds__(256)
gact_tile_s16_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                     int n_tiles, const EffLen *__restrict__ eff,
                     gact_tile_result *__restrict__ results, uint32_t *__restrict__ states,
                     int pitch_words, int *counter, size_t per_warp_bytes, uint8_t *gscratch, size_t dir_bytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = threadIdx.x >> 5;
    constexpr int TS = CS * 64;
    constexpr int NW = DirWin16<CS>::NW;

    // per-warp carve-out: rr[TS+2] words | qs[TS+2] halves | rb[TS+2] halves | direction window
    //   general mode: rr[i] = enc(R[i]) | enc(R[i-1]) << 16;  LUT mode: rr[i] = score table of R[i]
    //   (the traceback needs raw equality: it uses rb[] / qs[])
    uint8_t *my = smem + (size_t)warp * per_warp_bytes;
    uint32_t *rr = reinterpret_cast<uint32_t *>(my);
    uint16_t *qs = reinterpret_cast<uint16_t *>(my + (TS + 2) * 4);               // qs[j] = enc(Q[j])
    uint16_t *rb = reinterpret_cast<uint16_t *>(my + (TS + 2) * 6);               // rb[i] = enc(R[i])
    void *dirbase = DIRG ? (void *)(gscratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * dir_bytes)
                         : (void *)(my + (((TS + 2) * 8 + 15) & ~15));

    // biased x16 domain: stored = 16*score + tag + B; B keeps every half-word that takes part in a
    // plain 32-bit add non-negative, so those adds can run as IMAD on the FMA pipe
    const int B = P.s16_bias;
    const uint32_t Bp = pk16(B);
    const uint32_t ma16 = pk16(P.match * 16), mi16 = pk16(P.mismatch * 16);
    const uint32_t ge16 = pk16(P.gap_extend * 16);
    const int KO = (P.gap_open * 16) * 65537;            // + go            (phase 1, untagged)
    const int KI = (P.gap_open * 16 - 5) * 65537;        // M tag 1111 -> I-open tag 1010
    const int KD = (P.gap_open * 16 - 10) * 65537;       // M tag 1111 -> D-open tag 0101
    const int ONE = P.one;                               // opaque 1: keeps the adds on IMAD
    const uint32_t borderD_tag = ((uint32_t)(B + P.gap_open * 16 + 5) << 16) | (uint32_t)B;   // lane 0: G = 0 | D[i][1] = go, open
    const uint32_t borderD_raw = ((uint32_t)(B + P.gap_open * 16) << 16) | (uint32_t)B;
    const uint32_t lut_mis = (uint32_t)((P.mismatch * 16) & 0xff) * 0x01010101u;

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
            const bool in = (x >= 1 && x <= n);
            const int base = in ? tile_base(rset, d.ref_off, d.ref_len, d.reverse, x) : 0;
            rb[x] = (uint16_t)(in ? enc_base(base) : SENT_R);
            if (LUT) {
                // byte [code] = match*16 where code == base's 2-bit code, mismatch*16 elsewhere
                const int code = (base == 'A') ? 0 : (base == 'C') ? 1 : (base == 'G') ? 2 : (base == 'T') ? 3 : 4;
                uint32_t w = lut_mis;
                if (in && code < 4) w ^= (uint32_t)(((P.match ^ P.mismatch) * 16) & 0xff) << (8 * code);
                rr[x] = w;
            }
        }
        if (!LUT) {
            __syncwarp();
            for (int x = lane; x <= n + 1; x += 32)
                rr[x] = (uint32_t)rb[x] | ((uint32_t)(x >= 1 ? rb[x - 1] : (uint16_t)SENT_R) << 16);
        }
        for (int x = lane; x <= m; x += 32)
            qs[x] = (x >= 1) ? (uint16_t)enc_base(tile_base(qset, d.query_off, d.query_len, d.reverse, x)) : (uint16_t)SENT_Q;
        __syncwarp();

        uint32_t q[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) {
            const int jl = (2 * lane) * CS + c + 1, jh = (2 * lane + 1) * CS + c + 1;
            const uint32_t el = jl <= m ? (uint32_t)qs[jl] : SENT_Q + c, eh = jh <= m ? (uint32_t)qs[jh] : SENT_Q + c;
            if (LUT) {
                // ASCII A=0x41 C=0x43 G=0x47 T=0x54: (b >> 1) & 3 = 0,1,3,2 -> 2-bit code 0,1,2,3
                const uint32_t tl = (el >> 1) & 3u, th = (eh >> 1) & 3u;
                const uint32_t l2 = (jl <= m) ? (tl ^ (tl >> 1)) : 0u, h2 = (jh <= m) ? (th ^ (th >> 1)) : 0u;
                q[c] = l2 | ((8u | l2) << 4) | ((4u | h2) << 8) | ((12u | h2) << 12);
            } else {
                q[c] = el | (eh << 16);
            }
        }

        DirWin16<CS> dw;
        dw.init(dirbase, n, m, P);
        const int laststrip = (m > 0) ? (m - 1) / CS : -1;
        const int lastlane = laststrip >> 1;
        // the corner H[n][m] appears in lane c_lane, half c_half, column c_col, at step kc
        const int c_lane = max(lastlane, 0), c_half = laststrip & 1, c_col = (m > 0) ? (m - 1) - laststrip * CS : 0;
        const int kc = n + 2 * c_lane + c_half;
        const int steps = (n > 0 && m > 0) ? n + 1 + 2 * lastlane : 0;
        // lane-private step windows: active for k in [kfirst, kfirst + n], stores from kstore on
        const int kfirst = (lane <= lastlane) ? 2 * lane + 1 : 0x3fffffff;
        const int kstore = (lane >= dw.lane0) ? dw.i0 + 2 * lane : 0x3fffffff;
        const uint32_t *rrp = rr - 2 * lane;             // rrp[k] = rr[k - 2*lane]

        // ---------------- phase 1: rows above the traceback window, score only ----------------
        // untagged biased values; low half = after border row 0, high half = "row -1"
        uint32_t Gup[CS], IoUp[CS], IcUp[CS];
#pragma unroll
        for (int c = 0; c < CS; c++) {
            Gup[c] = Bp;
            IoUp[c] = pk16(B + P.gap_open * 16, S16_NEG);      // M[0][j] + go
            IcUp[c] = pk16(S16_NEG, S16_NEG);
        }
        uint32_t eG = Bp, eD = pk16(S16_NEG), diag = Bp;
        // phase 1 covers steps 1..k1: the first lane that keeps direction codes (lane0) reaches
        // window row i0 at step i0 + 2*lane0
        const int k1 = min(dw.i0 - 1 + 2 * dw.lane0, steps);
        int k = 1;
        for (; k <= k1; k++) {
            const uint32_t pack = __byte_perm(eG, eD, 0x7632);
            uint32_t recv = __shfl_up_sync(FULL, pack, 1);
            if (lane == 0) recv = borderD_raw;
            const uint32_t inG = __byte_perm(recv, eG, 0x5410);
            const uint32_t inD = __byte_perm(recv, eD, 0x5432);
            if ((unsigned)(k - kfirst) <= (unsigned)n) {
                const uint32_t rlo = rrp[k], rhi = LUT ? rrp[k - 1] : 0u;
                uint32_t hd = diag, dv = inD;
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const uint32_t sc = subst_score<LUT>(q[c], rlo, rhi, ma16, mi16);
                    const uint32_t mc = __viaddmax_s16x2(hd, sc, Bp);
                    hd = Gup[c];
                    const uint32_t iv = __viaddmax_s16x2(IcUp[c], ge16, IoUp[c]);
                    Gup[c] = __vimax3_s16x2(mc, iv, dv);
                    const uint32_t mo = (uint32_t)((int)mc * ONE + KO);
                    IoUp[c] = mo;
                    IcUp[c] = iv;
                    dv = __viaddmax_s16x2(dv, ge16, mo);
                }
                eG = Gup[CS - 1];
                eD = dv;
                diag = inG;
            }
        }
        // ---------------- switch to the tagged domain ----------------
#pragma unroll
        for (int c = 0; c < CS; c++) {
            IoUp[c] = __vadd2(IoUp[c], pk16(10));              // (M|1111) + go - 5
            IcUp[c] = __vadd2(IcUp[c], pk16(8));               // I tag 1000
        }
        eD = __vadd2(eD, pk16(4));                             // D tag 0100 (flag irrelevant above the window)

        // ---------------- phase 2: window rows, tagged values + direction codes ----------------
        uint32_t *wptr = dw.w + ((k - 2 * lane - dw.i0) * dw.nl + (lane - dw.lane0)) * NW;
        uint8_t *bptr = dw.b + ((k - 2 * lane - dw.i0) * dw.nl + (lane - dw.lane0));
        int corner16 = B;
        for (; k <= steps; k++) {
            const uint32_t pack = __byte_perm(eG, eD, 0x7632);           // (eG.hi, eD.hi)
            uint32_t recv = __shfl_up_sync(FULL, pack, 1);
            if (lane == 0) recv = borderD_tag;
            const uint32_t inG = __byte_perm(recv, eG, 0x5410);          // lo: left strip's G, hi: my low strip's G
            const uint32_t inD = __byte_perm(recv, eD, 0x5432);          // same for D
            if ((unsigned)(k - kfirst) <= (unsigned)n) {
                const uint32_t rlo = rrp[k], rhi = LUT ? rrp[k - 1] : 0u;
                uint32_t hd = diag, dv = inD;
                uint32_t acc[NW + 1];
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const uint32_t sc = subst_score<LUT>(q[c], rlo, rhi, ma16, mi16);
                    const uint32_t mt = __viaddmax_s16x2(hd, sc, Bp) | 0x000f000fu;     // M, tag 1111
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
                eG = Gup[CS - 1];
                eD = dv;
                diag = inG;
                if (k >= kstore) {
#pragma unroll
                    for (int x = 0; x < NW; x++) wptr[x] = acc[x];
                    if (DirWin16<CS>::HAS_B) *bptr = (uint8_t)((acc[NW] & 15u) | ((acc[NW] >> 12) & 0xf0u));
                }
            }
            wptr += dw.nl * NW;
            bptr += dw.nl;
            if (k == kc) {                                   // warp-uniform: the corner row just finished
                uint32_t gsel = 0;
#pragma unroll
                for (int c = 0; c < CS; c++) if (c == c_col) gsel = Gup[c];
                corner16 = (int)(short)(c_half ? (gsel >> 16) : (gsel & 0xffffu));
            }
        }
        int corner = (__shfl_sync(FULL, corner16, c_lane) - B) >> 4;
        if (n == 0 || m == 0) corner = 0;
        if (DIRG) __threadfence_block();
        __syncwarp();
        // rr[] (substitution tables) is dead now: reuse it as the per-state byte buffer (2*et <= 4*(TS+2))
        traceback_tile16<CS>(dw, rb, qs, reinterpret_cast<uint8_t *>(rr), lane, n, m, corner, P,
                             states + (size_t)t * pitch_words, &results[t],
                             d.first ? n : d.ref_len, d.first ? m : d.query_len);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host-side planning / launch
template <int CS>
inline size_t s16_seq_bytes()
{
    constexpr int TS = CS * 64;
    return (size_t)(((TS + 2) * 8 + 15) & ~15);
}
inline size_t s16_seq_bytes(int CS)
{
    switch (CS) { case 4: return s16_seq_bytes<4>(); case 5: return s16_seq_bytes<5>(); case 8: return s16_seq_bytes<8>(); default: return s16_seq_bytes<16>(); }
}
inline size_t s16_dir_bytes(int CS, int rows, int lanes)
{
    switch (CS) {
        case 4: return DirWin16<4>::bytes(rows, lanes);
        case 5: return DirWin16<5>::bytes(rows, lanes);
        case 8: return DirWin16<8>::bytes(rows, lanes);
        default: return DirWin16<16>::bytes(rows, lanes);
    }
}

typedef void (*s16_fn)(const KParams, const gact_tile_desc *, int, const EffLen *, gact_tile_result *,
                       uint32_t *, int, int *, size_t, uint8_t *, size_t);
template <int CS>
inline s16_fn s16_pick_cs(bool lut, bool dirg)
{
    if (dirg) return lut ? gact_tile_s16_kernel<CS, true, true> : gact_tile_s16_kernel<CS, false, true>;
    return lut ? gact_tile_s16_kernel<CS, true, false> : gact_tile_s16_kernel<CS, false, false>;
}
inline s16_fn s16_pick(int CS, bool lut, bool dirg)
{
    switch (CS) {
        case 4: return s16_pick_cs<4>(lut, dirg);
        case 5: return s16_pick_cs<5>(lut, dirg);
        case 8: return s16_pick_cs<8>(lut, dirg);
        case 16: return s16_pick_cs<16>(lut, dirg);
        default: return nullptr;
    }
}

// win_rows/win_lanes for the s16 kernel are derived here (they differ from the int32 kernel's).
struct S16Plan {
    bool ok = false;
    int CS = 0, win_rows = 0, win_lanes = 0, warps_per_cta = 0, ctas = 0, bias = 0;
    bool lut_ok = false;      // scores fit the one-PRMT substitution table
    bool dirg = false;        // direction window in global (L2-resident) scratch instead of shared memory
    size_t per_warp_bytes = 0, smem = 0, dir_bytes = 0;
    uint8_t *d_scratch = nullptr;
};

inline void s16_free_plan(S16Plan *pl)
{
    if (pl->d_scratch) cudaFree(pl->d_scratch);
    pl->d_scratch = nullptr;
}

// mode: 0 = default: global (L2-resident) scratch window -- measured 18 % faster than the
//           shared-memory window at tile_size 320 because occupancy is no longer bounded by the
//           22 KB window (24 instead of 9 warps per SM; profiles/r1_window_sweep.txt);
//       1 = force shared memory (falls back to global when fewer than 4 warps per SM would fit);
//       2 = force global scratch.  warps_per_sm: 0 = default (24).
inline int s16_make_plan(const gact_params &p, int num_sms, int mode, int warps_per_sm, S16Plan *pl)
{
    s16_free_plan(pl);
    *pl = S16Plan();
    const int T = p.tile_size, et = p.tile_size - p.tile_overlap;
    // value range of the x16 tagged domain and the pseudo-row trick
    const int bias = 16 * (-p.gap_open + 2);
    pl->bias = bias;
    pl->lut_ok = (p.match * 16 <= 127 && p.mismatch * 16 >= -128);
    const long hi = (long)T * (p.match > 0 ? p.match : 0) * 16 + 16 + bias;
    if (hi > 30000 || p.mismatch > 0 || p.match < 0 || p.gap_open < -500 || p.gap_extend < -500 || p.mismatch < -1000)
        return 0;
    int CS;
    if (T <= 256) CS = 4; else if (T <= 320) CS = 5; else if (T <= 512) CS = 8; else CS = 16;
    pl->CS = CS;
    pl->win_rows = (et + 1 < T) ? et + 1 : T;
    int wl = et / (2 * CS) + 2;
    pl->win_lanes = wl > 32 ? 32 : wl;
    const size_t seqb = s16_seq_bytes(CS);
    pl->dir_bytes = s16_dir_bytes(CS, pl->win_rows, pl->win_lanes);
    const size_t SM = 228 * 1024, CTA_MAX = 227 * 1024;

    auto plan_smem = [&]() -> bool {
        const size_t pw = (seqb + pl->dir_bytes + 15) & ~(size_t)15;
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
        if (best_total < 4) return false;
        pl->dirg = false;
        pl->per_warp_bytes = pw;
        pl->warps_per_cta = best_w;
        pl->ctas = best_c * num_sms;
        pl->smem = (size_t)best_w * pw;
        return true;
    };
    auto plan_global = [&]() -> bool {
        int wps = warps_per_sm > 0 ? warps_per_sm : 24;
        if (wps > 32) wps = 32;
        const int w = 4;                                   // warps per CTA
        const int c = (wps + w - 1) / w;
        pl->dirg = true;
        pl->per_warp_bytes = seqb;
        pl->warps_per_cta = w;
        pl->ctas = c * num_sms;
        pl->smem = (size_t)w * seqb;
        const size_t total = (size_t)pl->ctas * w * pl->dir_bytes;
        if (cudaMalloc(&pl->d_scratch, total) != cudaSuccess) { cudaGetLastError(); pl->d_scratch = nullptr; return false; }
        return true;
    };
    bool ok = false;
    if (mode == 1) ok = plan_smem() || plan_global();
    else ok = plan_global() || plan_smem();
    if (!ok) return 0;
    for (int lut = 0; lut < 2; lut++)
        if (cudaFuncSetAttribute((const void *)s16_pick(CS, lut != 0, pl->dirg), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pl->smem) != cudaSuccess)
            return -1;
    pl->ok = true;
    return 0;
}

inline void s16_launch(const S16Plan &pl, KParams kp, const gact_tile_desc *descs, int n, const EffLen *eff,
                       gact_tile_result *results, uint32_t *states, int pitch_words, int *counter, cudaStream_t st)
{
    kp.win_rows = pl.win_rows;
    kp.win_lanes = pl.win_lanes;
    kp.s16_bias = pl.bias;
    kp.one = 1;
    // one-PRMT substitution table only when every set in use is 2-bit packed (ACGT only)
    bool lut = pl.lut_ok;
    for (int i = 0; i < GACT_MAX_SETS; i++) if (kp.sets[i].bytes) lut = false;
    int ctas = pl.ctas;
    const int need = (n + pl.warps_per_cta - 1) / pl.warps_per_cta;
    if (need < ctas) ctas = need;
    s16_pick(pl.CS, lut, pl.dirg)<<<ctas, pl.warps_per_cta * 32, pl.smem, st>>>(kp, descs, n, eff, results, states,
                                                                              pitch_words, counter, pl.per_warp_bytes,
                                                                              pl.d_scratch, pl.dir_bytes);
}

}  // namespace gact
