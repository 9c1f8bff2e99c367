// gact_kernels_s16.cuh -- packed s16x2 DPX GACT tile kernel (placeholder: not yet enabled).
#pragma once
#include "gact_common.cuh"

namespace gact {

inline int s16_plan(const gact_params &, int, const KParams &, bool *ok, int *C, size_t *per_warp,
                    int *warps_per_cta, int *ctas, size_t *smem)
{
    *ok = false; *C = 0; *per_warp = 0; *warps_per_cta = 0; *ctas = 0; *smem = 0;
    return 0;
}
inline void s16_launch(int, const KParams &, const gact_tile_desc *, int, const EffLen *, gact_tile_result *,
                       uint32_t *, int, int *, size_t, int, int, size_t, cudaStream_t) {}

}  // namespace gact
