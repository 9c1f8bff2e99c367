// gact_common.cuh -- device-side types shared by the GACT tile kernels.
//
// Data layout in HBM
//   * sequence sets: one concatenated buffer per set, 2-bit packed (16 bases per
//     32-bit word, base b at bits [2b, 2b+1], A/a=0 C/c=1 G/g=2 T/t=3, anything
//     else 0: the coding of ntcoding.cpp:60-72, which is what D-SOFT hashes).  A set
//     that holds any byte other than upper-case ACGT also keeps an EXCEPTION bitmap
//     (1 bit per base) and its raw bytes: the reference compares raw bytes
//     (align.cpp:134: 'N'=='N' matches, 'a'!='A'), so the one-PRMT score table is
//     exact only where the query window has no exception (an exception in the
//     reference window then mismatches every query base); tiles whose query window
//     holds one run on the raw-byte kernels;
//   * tile descriptors: gact_tile_desc (32 B, include/gact_b200.h);
//   * results: gact_tile_result (24 B) + 2-bit packed traceback states.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/gact_b200.h"

namespace gact {

struct SeqSetDev {
    const uint32_t *packed;   // 2-bit words (nullptr for an empty set)
    const uint8_t *bytes;     // raw bytes, only for sets with exceptions (else nullptr)
    const uint32_t *exc;      // exception bitmap, bit b of word w = base 32 w + b is not one of "ACGT" (else nullptr)
    long long len;
};

struct KParams {
    int match, mismatch, gap_open, gap_extend;
    int et;                  // early_terminate = tile_size - tile_overlap
    int tile_size;
    int win_rows;            // rows of direction codes kept per tile (<= et + 1)
    int win_lanes;           // lanes (column strips) of direction codes kept
    int s16_bias;            // packed kernel: bias of the x16 domain
    int one;                 // 1, opaque to the compiler (IMAD-as-add on the FMA pipe)
    SeqSetDev sets[GACT_MAX_SETS];
};

// (max_i, max_j) found by the first-tile pass; consumed by the main pass.
struct EffLen { int n, m; };

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int NEG_BORDER = -(1 << 30);     // align.h:18

__device__ __forceinline__ int fetch_base(const SeqSetDev &s, long long idx)
{
    if (s.bytes) return __ldg(s.bytes + idx);       // raw bytes where the set has exceptions
    if (s.packed) {
        const uint32_t w = __ldg(s.packed + (idx >> 4));
        const int code = (w >> (2 * (int)(idx & 15))) & 3;
        // 'A' 'C' 'G' 'T' packed into one constant, one byte each
        return (0x54474341u >> (8 * code)) & 0xff;
    }
    return __ldg(s.bytes + idx);
}

// Base j (1-based, DP order) of a tile: natural order for reverse = 0,
// back to front for reverse = 1 (align.cpp:130-131, CPU-build sense).
__device__ __forceinline__ int tile_base(const SeqSetDev &s, long long off, int len, int reverse, int j)
{
    return fetch_base(s, reverse ? off + (len - j) : off + (j - 1));
}

}  // namespace gact
