// epilogue.cuh — accumulator (TMEM) -> bias (+ skip) -> ReLU -> bf16 -> next layer's B operand.
//
// The accumulator is D^T[Cout lane][slot column]; the next layer wants the K-major record
// [Cin/8][slot][8 channels] (16 bytes per slot and channel chunk).  tcgen05.ld.16x256b hands each
// warp the mma C-fragment of a 16-lane x 8-column block, and stmatrix.trans writes four 8x8
// fragments as 8 slots x 8 channels records: 256 elements per store instruction, no shuffles, no
// bank conflicts (8 consecutive 16-byte records per matrix).  The skip connection comes back the
// same way with ldmatrix.trans from the very records about to be overwritten.
#pragma once
#include <stdint.h>

#include "umma.cuh"

namespace nsb {

// slot n = 100*pos + 10*row + col; column 9 and row 9 are the permanent zero padding
__host__ __device__ constexpr bool is_real_slot(int n) { return (n % 100) < 90 && ((n % 100) % 10) < 9; }

// Per-thread masks for the packed bf16 pairs of 8-column group G (G = 0 .. 4*NCG-1): the pair
// holds slots col0 + 8G + 2(lane%4) + {0,1}; padding slots are forced to zero with one AND.
template <int NCG>
struct EpilogueMask {
    uint32_t m[NCG * 4];
    template <int COL0>
    __device__ __forceinline__ void init_at(int lane) {
#pragma unroll
        for (int G = 0; G < NCG * 4; ++G) {
            uint32_t v = 0;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {  // compile-time slots; select on lane % 4
                const int n = COL0 + 8 * G + 2 * q4;
                const uint32_t c = (is_real_slot(n) ? 0x0000FFFFu : 0u) | (is_real_slot(n + 1) ? 0xFFFF0000u : 0u);
                v = (lane & 3) == q4 ? c : v;
            }
            m[G] = v;
        }
    }
    __device__ __forceinline__ void init(int col0, int lane) {  // col0 in {0, 96}
        if (col0 == 0) init_at<0>(lane);
        else init_at<96>(lane);
    }
};

// One warp handles 16 of its 32 TMEM lanes (lane block lb = 0 / 1 -> 16 output channels = chunks
// chunk0 + 2*lb, +1) for the 32*NCG columns starting at col0.  `taddr` = TMEM address of (lane
// quadrant base, column col0 of this accumulator); `buf` = shared address of the output buffer's
// slot 0 of chunk 0 (guard already added); `pitch` = bytes between channel chunks.
// bias[lb*2 + h] is the bias of channel 32q + 16*lb + 8*h + lane/4.  Per element: one FADD
// (+ unpack and FADD for the skip), half a cvt.rn.relu.bf16x2 and half an AND.
// LBS = chunk distance between the two lane blocks (2: channels in accumulator-row order; 8: the
// TS kernel's row permutation, where lane block lb holds K block lb of the output channels).
template <int NCG, bool kResidual, int LBS = 2>
__device__ __forceinline__ void epilogue_half(int lb, uint32_t taddr, uint32_t buf, uint32_t pitch, int chunk0,
                                              int col0, const float (&bias)[4], const EpilogueMask<NCG>& mask,
                                              int lane) {
    const int mk = lane >> 3, mi = lane & 7;
    uint32_t v[NCG][16];
#pragma unroll
    for (int cg = 0; cg < NCG; ++cg) tmem_ld_16x256b_x4(taddr + ((uint32_t)(lb * 16) << 16) + cg * 32, v[cg]);
    const uint32_t row_addr =
        buf + (uint32_t)(chunk0 + lb * LBS + (mk & 1)) * pitch + (uint32_t)(col0 + 8 * (mk >> 1) + mi) * 16u;
    uint32_t xr[NCG][2][4];
    if (kResidual) {
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                ldmatrix_x4_trans(row_addr + (uint32_t)(cg * 32 + h * 16) * 16u, xr[cg][h][0], xr[cg][h][1],
                                  xr[cg][h][2], xr[cg][h][3]);
    }
    tmem_ld_wait();
    const float b0 = lb ? bias[2] : bias[0], b1 = lb ? bias[3] : bias[1];
    const uint64_t bb0 = pack_f32x2(b0, b0), bb1 = pack_f32x2(b1, b1);
#pragma unroll
    for (int cg = 0; cg < NCG; ++cg)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t p[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {  // m: bit0 = row half (lane/4 [+8]), bit1 = column group
                const int g = 2 * h + (m >> 1), rh = m & 1;
                // the accumulator pair (two adjacent slots of one channel) goes through packed adds
                uint64_t x = add_f32x2(pack_u32x2(v[cg][4 * g + 2 * rh + 0], v[cg][4 * g + 2 * rh + 1]), rh ? bb1 : bb0);
                if (kResidual) x = add_f32x2(x, pack_u32x2(xr[cg][h][m] << 16, xr[cg][h][m] & 0xFFFF0000u));
                p[m] = pack_relu_bf16x2(x) & mask.m[cg * 4 + g];
            }
            stmatrix_x4_trans(row_addr + (uint32_t)(cg * 32 + h * 16) * 16u, p[0], p[1], p[2], p[3]);
        }
}

// Both lane blocks: the warp's 32 TMEM lanes (= 32 output channels, 4 chunks starting at chunk0).
template <int NCG, bool kResidual>
__device__ __forceinline__ void epilogue_warp(uint32_t taddr, uint32_t buf, uint32_t pitch, int chunk0, int col0,
                                              const float (&bias)[4], const EpilogueMask<NCG>& mask, int lane) {
#pragma unroll
    for (int lb = 0; lb < 2; ++lb)
        epilogue_half<NCG, kResidual>(lb, taddr, buf, pitch, chunk0, col0, bias, mask, lane);
}

}  // namespace nsb
