// Host-side check of the inter-task kernel's band geometry (csrc/gact_kernels_it.cuh: it_geometry, it_win_bytes,
// it_edge_bytes): every window cell within W of the main diagonal lies in a tagged row of its strip, the per-strip
// record offsets are the running sum of the band heights, and the byte counts the engine allocates match.
#include <cstdio>
#include <cstdlib>
#include "gact_kernels_it.cuh"

using namespace gact;

int main()
{
    int bad = 0, cases = 0;
    const int Ts[] = {64, 128, 256, 320, 512, 1024};
    for (int T : Ts) {
        for (int et : {T, T - T / 8, (T * 5) / 8, T / 2, 10}) {
            if (et < 1) continue;
            for (int W : {4, 5, 32, (et + 7) / 8 > 32 ? (et + 7) / 8 : 32, T}) {
                const ITGeom g = it_geometry(T, et, W);
                cases++;
                if (g.S != T / IT_CS || g.i0 != (T - et > 1 ? T - et : 1) || g.j0 != g.i0 || g.s0 != (g.j0 - 1) / IT_CS) bad++;
                int off = 0;
                for (int s = 0; s < IT_MAX_STRIPS; s++) {
                    if (g.win_off[s] != off) bad++;
                    const bool has = g.band_hi[s] >= g.band_lo[s];
                    if (has && (s >= g.S || s < g.s0)) bad++;            // bands only in strips that reach the window
                    if (has) {
                        if (g.band_lo[s] < g.i0 || g.band_hi[s] > T) bad++;
                        off += g.band_hi[s] - g.band_lo[s] + 1;
                    }
                }
                if (g.win_off[IT_MAX_STRIPS] != off) bad++;
                if (it_win_bytes(g) != (size_t)off * (IT_CS / 4) * 32 * sizeof(uint32_t)) bad++;
                // every window cell within W of the diagonal is covered; cells left of the window's first strip never are
                for (int i = g.i0; i <= T; i++)
                    for (int j = 1; j <= T; j++) {
                        const int s = (j - 1) / IT_CS;
                        const bool in_band = s >= g.s0 && i >= g.band_lo[s] && i <= g.band_hi[s];
                        const int d = i > j ? i - j : j - i;
                        if (j >= g.j0 && d <= W && !in_band) bad++;
                        if (s < g.s0 && in_band) bad++;
                    }
            }
        }
        if (it_edge_bytes(T) != (size_t)(T + 1 + IT_PF) * 32 * 8) bad++;
        if (it_smem_per_warp(T, true) != 2 * it_smem_per_warp(T, false)) bad++;
    }
    if (bad) printf("FAIL %d violations in %d geometries\n", bad, cases);
    else printf("OK %d geometries\n", cases);
    return bad ? 1 : 0;
}
