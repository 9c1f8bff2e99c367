// gact_kernels_s16.cuh -- arithmetic helpers of the packed s16x2 DPX GACT kernels.
//
// The kernels themselves live in gact_kernels_s16h.cuh (LANES lanes per tile: two tiles per warp for
// tile_size <= 320, one per warp up to 1024).  Same contract as the int32 kernel (AlignWithBT,
// align.cpp:60-233), half the instructions per cell: every 32-bit register holds two DP cells.
//
// Arithmetic.  Scores are scaled by 16 and the low nibble of every value is a TAG, so that the max
// instructions themselves produce the traceback pointer of align.cpp:162-171:
//      M operand  ....1111     I: open ....1010 / extend ....1000
//                              D: open ....0101 / extend ....0100
//   * I = VIADDMNMX.S16x2(Iclean_up, ge, Mup+go-5)  -> bit 1 = (ins_open >= ins_extend)
//   * D = VIADDMNMX.S16x2(Dclean_lf, ge, Mlf+go-10) -> bit 0 = (del_open >= del_extend)
//   * G = VIMNMX3.S16x2(M, I, D)                    -> bits 3:2 = 3/2/1 = M/I/D with the
//     reference's tie order M >= I >= D (equal scores are decided by the tag)
//   * M = VIADDMNMX.S16x2(Gdiag, s, bias) | 0xF, s from one PRMT (2-bit sets) or HSET2 + LOP3 (raw bytes)
// The ZERO state (align.cpp:166-168, H <= 0) is not stored: the traceback tracks the score of the
// cell it stands on and stops when it reaches 0.  Border rows/columns are not special-cased: the
// high half runs a pseudo row 0 against a sentinel base that reproduces the border values exactly.
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

}  // namespace gact
