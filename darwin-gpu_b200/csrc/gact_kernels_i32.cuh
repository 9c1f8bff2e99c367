// gact_kernels_i32.cuh -- int32 DPX GACT tile kernels (one warp per tile).
//
// What is computed: exactly AlignWithBT() of the reference (align.cpp:60-233),
// see include/gact_b200.h.  How: lane L of a warp owns the C query columns
// [L*C+1, L*C+C] and sweeps the reference rows; lanes run one row apart
// (anti-diagonal wavefront), the right-edge H / M / D of a strip travel to the
// next lane by warp shuffle.  Per cell: VIADDMNMX.RELU for M, two add+max pairs
// for I and D, VIMNMX3 for H, then the 4-bit direction code of align.cpp:162-171
// which is nibble-packed and written to the tile's window in shared memory
// (or an L2-resident scratch area for the largest tile sizes).  Only the
// window the traceback can reach -- the last ET+1 rows and columns, ET =
// tile_size - tile_overlap (align.cpp:205) -- is kept.  First tiles
// (align.cpp:190) run a score-only pass that finds the last maximum, then the
// same kernel on the sub-tile ending at that cell.
#pragma once
#include "gact_common.cuh"

namespace gact {

// ---------------------------------------------------------------------------
// geometry of the direction-code window of one tile
template <int C>
struct DirWin {
    static constexpr int WPL = C / 8;               // full 32-bit words per lane-row
    static constexpr bool HAS_B = (C % 8) != 0;     // plus one byte (C = 10: 2 nibbles)
    int i0, lane0, nl;
    uint32_t *w;
    uint8_t *b;
    __device__ __forceinline__ void init(void *base, int n, int m, const KParams &P)
    {
        i0 = max(n - P.et, 1);
        const int j0 = max(m - P.et, 1);
        lane0 = (j0 - 1) / C;
        nl = P.win_lanes;
        w = reinterpret_cast<uint32_t *>(base);
        b = reinterpret_cast<uint8_t *>(w + (size_t)P.win_rows * nl * WPL);
    }
    // direction nibble of cell (i, j), i >= i0, j >= j0
    __device__ __forceinline__ int load(int i, int j) const
    {
        const int lane = (j - 1) / C, c = (j - 1) - lane * C;
        const int e = (i - i0) * nl + (lane - lane0);
        if (c < WPL * 8) return (w[e * WPL + (c >> 3)] >> (4 * (c & 7))) & 15;
        return (b[e] >> (4 * (c - WPL * 8))) & 15;
    }
    static __host__ __device__ size_t bytes(int win_rows, int win_lanes)
    {
        size_t s = (size_t)win_rows * win_lanes * (WPL * 4 + (HAS_B ? 1 : 0));
        return (s + 15) & ~(size_t)15;
    }
};

// ---------------------------------------------------------------------------
// traceback, align.cpp:185-230, executed by one lane.  `states` receives 2-bit
// codes, 16 per word.
template <int C>
__device__ __forceinline__ void traceback_tile(const DirWin<C> &dw, int n, int m, int score, int et,
                                               uint32_t *states, gact_tile_result *res,
                                               int out_max_i, int out_max_j)
{
    int i = n, j = m, is = 0, js = 0, cnt = 0;
    uint32_t acc = 0;
    int state = (i > 0 && j > 0) ? (dw.load(i, j) & 3) : 0;
    while (state != 0) {
        if (is >= et || js >= et) break;
        if (i <= 0 || j <= 0) break;            // unreachable for gap scores <= 0
        acc |= (uint32_t)state << (2 * (cnt & 15));
        if ((cnt & 15) == 15) { states[cnt >> 4] = acc; acc = 0; }
        cnt++;
        if (state == 3) {
            i--; j--; is++; js++;
            state = (i > 0 && j > 0) ? (dw.load(i, j) & 3) : 0;
        } else if (state == 2) {
            state = (dw.load(i, j) & 8) ? 3 : 2;
            i--; is++;
        } else {
            state = (dw.load(i, j) & 4) ? 3 : 1;
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

// ---------------------------------------------------------------------------
// main kernel: DP + direction window + traceback.  One warp per tile, tiles
// handed out through an atomic counter (persistent CTAs).
template <int C, bool DIR_GLOBAL>
__global__ void __launch_bounds__(256)
gact_tile_i32_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                     int n_tiles, const EffLen *__restrict__ eff,
                     gact_tile_result *__restrict__ results, uint32_t *__restrict__ states,
                     int pitch_words, int *counter, uint8_t *gscratch, size_t per_warp_bytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int TS = C * 32;

    // per-warp carve-out: [reference bases, one byte each][direction window]
    uint8_t *my = DIR_GLOBAL ? smem + (size_t)warp * TS : smem + (size_t)warp * per_warp_bytes;
    uint8_t *rs = my;                                           // rs[i-1] = R[i]
    void *dirbase = DIR_GLOBAL
        ? (void *)(gscratch + ((size_t)blockIdx.x * warps_per_cta + warp) * per_warp_bytes)
        : (void *)(my + TS);

    const int ma = P.match, mi = P.mismatch, go = P.gap_open, ge = P.gap_extend;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_tiles) break;

        const gact_tile_desc d = descs[t];
        int n = d.ref_len, m = d.query_len;
        if (d.first) { n = eff[t].n; m = eff[t].m; }            // sub-tile ending at the last maximum
        const SeqSetDev &rset = P.sets[d.ref_set];
        const SeqSetDev &qset = P.sets[d.query_set];

        __syncwarp();
        for (int x = lane; x < n; x += 32) rs[x] = (uint8_t)tile_base(rset, d.ref_off, d.ref_len, d.reverse, x + 1);
        int q[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = lane * C + c + 1;
            q[c] = (j <= m) ? tile_base(qset, d.query_off, d.query_len, d.reverse, j) : 256 + c;
        }
        __syncwarp();

        DirWin<C> dw;
        dw.init(dirbase, n, m, P);
        const int lastlane = (m > 0) ? (m - 1) / C : -1;

        int Hup[C], Mup[C], Iup[C];
#pragma unroll
        for (int c = 0; c < C; c++) { Hup[c] = 0; Mup[c] = 0; Iup[c] = NEG_BORDER; }
        int eH = 0, eM = 0, eD = NEG_BORDER;      // right edge of my strip, current row
        int diag = 0;                             // H[i-1][first column - 1]

        const int steps = (n > 0 && m > 0) ? n + lastlane : 0;
        for (int k = 1; k <= steps; k++) {
            const int i = k - lane;
            int rH = __shfl_up_sync(FULL, eH, 1);
            int rM = __shfl_up_sync(FULL, eM, 1);
            int rD = __shfl_up_sync(FULL, eD, 1);
            if (lane == 0) { rH = 0; rM = 0; rD = NEG_BORDER; }
            if (i >= 1 && i <= n && lane <= lastlane) {
                const int r = rs[i - 1];
                int hd = diag, ml = rM, dl = rD;
                uint32_t code[C];
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int s = (q[c] == r) ? ma : mi;
                    const int mc = __viaddmax_s32_relu(hd, s, 0);            // align.cpp:138-147
                    hd = Hup[c];
                    const int io = Mup[c] + go, ie = Iup[c] + ge;              // align.cpp:149-150
                    const int dopen = ml + go, dext = dl + ge;                 // align.cpp:151-152
                    const int iv = max(io, ie), dv = max(dopen, dext);
                    const int h = __vimax3_s32(mc, iv, dv);                    // align.cpp:158-160 (mc >= 0)
                    int st = (mc >= iv) ? ((mc >= dv) ? 3 : 1) : ((iv >= dv) ? 2 : 1);   // :162-164
                    if (h <= 0) st = 0;                                        // :166-168
                    code[c] = st | ((io >= ie) ? 8 : 0) | ((dopen >= dext) ? 4 : 0);      // :170-171
                    Hup[c] = h; Mup[c] = mc; Iup[c] = iv;
                    ml = mc; dl = dv;
                }
                eH = Hup[C - 1]; eM = ml; eD = dl;
                diag = rH;
                if (i >= dw.i0 && lane >= dw.lane0) {
                    const int e = (i - dw.i0) * dw.nl + (lane - dw.lane0);
#pragma unroll
                    for (int x = 0; x < DirWin<C>::WPL; x++) {
                        uint32_t wv = 0;
#pragma unroll
                        for (int c = 0; c < 8; c++) wv |= code[x * 8 + c] << (4 * c);
                        dw.w[e * DirWin<C>::WPL + x] = wv;
                    }
                    if (DirWin<C>::HAS_B) {
                        uint32_t bv = 0;
#pragma unroll
                        for (int c = DirWin<C>::WPL * 8; c < C; c++) bv |= code[c] << (4 * (c - DirWin<C>::WPL * 8));
                        dw.b[e] = (uint8_t)bv;
                    }
                }
            }
        }
        // corner score H[n][m] lives in lane `lastlane`, column (m-1)%C
        int corner = 0;
        {
            const int cm = (m > 0) ? (m - 1) - lastlane * C : 0;
#pragma unroll
            for (int c = 0; c < C; c++) if (c == cm) corner = Hup[c];
            corner = __shfl_sync(FULL, corner, max(lastlane, 0));
            if (n == 0 || m == 0) corner = 0;
        }
        __syncwarp();
        if (DIR_GLOBAL) __threadfence_block();
        if (lane == 0) {
            traceback_tile<C>(dw, n, m, corner, P.et, states + (size_t)t * pitch_words, &results[t],
                              d.first ? n : d.ref_len, d.first ? m : d.query_len);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// first-tile pass: score only, finds the last maximum in (i outer, j inner)
// order (align.cpp:173-177).  Per column the running key H*2048 + i keeps the
// largest H and, among equals, the largest row; columns are then compared by
// (key, j).
template <int C>
__global__ void __launch_bounds__(256)
gact_first_i32_kernel(const __grid_constant__ KParams P, const gact_tile_desc *__restrict__ descs,
                      const int *__restrict__ first_list, int n_first, EffLen *__restrict__ eff, int *counter)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int TS = C * 32;
    uint8_t *rs = smem + (size_t)warp * TS;
    const int ma = P.match, mi = P.mismatch, go = P.gap_open, ge = P.gap_extend;

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= n_first) break;
        const int t = first_list[idx];
        const gact_tile_desc d = descs[t];
        const int n = d.ref_len, m = d.query_len;
        const SeqSetDev &rset = P.sets[d.ref_set];
        const SeqSetDev &qset = P.sets[d.query_set];

        __syncwarp();
        for (int x = lane; x < n; x += 32) rs[x] = (uint8_t)tile_base(rset, d.ref_off, d.ref_len, d.reverse, x + 1);
        int q[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = lane * C + c + 1;
            q[c] = (j <= m) ? tile_base(qset, d.query_off, d.query_len, d.reverse, j) : 256 + c;
        }
        __syncwarp();
        const int lastlane = (m > 0) ? (m - 1) / C : -1;

        int Hup[C], Mup[C], Iup[C], key[C];
#pragma unroll
        for (int c = 0; c < C; c++) { Hup[c] = 0; Mup[c] = 0; Iup[c] = NEG_BORDER; key[c] = -1; }
        int eH = 0, eM = 0, eD = NEG_BORDER, diag = 0;
        const int steps = (n > 0 && m > 0) ? n + lastlane : 0;
        for (int k = 1; k <= steps; k++) {
            const int i = k - lane;
            int rH = __shfl_up_sync(FULL, eH, 1);
            int rM = __shfl_up_sync(FULL, eM, 1);
            int rD = __shfl_up_sync(FULL, eD, 1);
            if (lane == 0) { rH = 0; rM = 0; rD = NEG_BORDER; }
            if (i >= 1 && i <= n && lane <= lastlane) {
                const int r = rs[i - 1];
                int hd = diag, ml = rM, dl = rD;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int s = (q[c] == r) ? ma : mi;
                    const int mc = __viaddmax_s32_relu(hd, s, 0);
                    hd = Hup[c];
                    const int iv = __viaddmax_s32(Iup[c], ge, Mup[c] + go);
                    const int dv = __viaddmax_s32(dl, ge, ml + go);
                    const int h = __vimax3_s32(mc, iv, dv);
                    key[c] = max(key[c], h * 2048 + i);
                    Hup[c] = h; Mup[c] = mc; Iup[c] = iv;
                    ml = mc; dl = dv;
                }
                eH = Hup[C - 1]; eM = ml; eD = dl;
                diag = rH;
            }
        }
        long long best = -1;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = lane * C + c + 1;
            if (j <= m && key[c] >= 0) best = max(best, ((long long)key[c] << 11) | j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
        if (lane == 0) {
            EffLen e;
            if (best < 0) { e.n = 0; e.m = 0; }
            else { e.m = (int)(best & 2047); e.n = (int)((best >> 11) & 2047); }
            eff[t] = e;
        }
    }
}

}  // namespace gact
