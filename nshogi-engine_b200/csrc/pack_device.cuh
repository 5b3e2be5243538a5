// pack_device.cuh — stage 1 of feature extraction for one position by one warp: nsb_position
// (108 B) -> the 86 feature bitboards of reference src/evaluate/preset.h:20-66 (semantics: SURVEY.md
// App. A.2, builder-defined because libnshogi is absent).  Shared by pack_positions_kernel
// (stages.cu: bitboards to HBM, the reference's FeatureStackComptime contract) and by the trunk
// kernels' prologue (trunk_common.cuh: bitboards straight into shared memory).
//
// The record stays in registers (lane l holds word l; bytes are fetched with shuffles), so the only
// scratch is 28 x 16 B for the occupancy words of the 28 board planes: lane s holds the piece code
// of squares s, s+32, s+64 and MATCH.ANY hands every lane the occupancy word of its own piece code
// in one instruction - three match instructions per position.
#ifndef NSB_PACK_DEVICE_CUH
#define NSB_PACK_DEVICE_CUH

#include <stdint.h>

#include "nsb_internal.h"

namespace nsb {

// P1-6 L1-4 N1-4 S1-4 G1-4 B1-2 R1-2: stand channel k (0..25) = "at least `need` of `piece` in hand"
__device__ __forceinline__ int stand_piece_of(int k, int* need) {
    if (k < 6) { *need = k + 1; return 0; }
    if (k < 10) { *need = k - 5; return 1; }
    if (k < 14) { *need = k - 9; return 2; }
    if (k < 18) { *need = k - 13; return 3; }
    if (k < 22) { *need = k - 17; return 4; }
    if (k < 24) { *need = k - 21; return 5; }
    *need = k - 23;
    return 6;
}

// All 32 lanes of a warp call this together.  `occ` = 28 uint4 of shared scratch owned by the warp;
// it may alias the first 28 output slots when emit(c, f) writes slot c (lane c % 32 is the only
// reader and the only writer of slot c after the internal __syncwarp).  emit is called once for
// every channel c in 0..85 with the finished 16-byte bitboard.
template <typename Emit>
__device__ __forceinline__ void pack_position_warp(const nsb_position* __restrict__ pos, int lane, uint4* occ, Emit emit) {
    constexpr unsigned kFull = 0xffffffffu;
    const uint32_t w = lane < 27 ? __ldg(reinterpret_cast<const uint32_t*>(pos) + lane) : 0u;  // 108 B = 27 words
    auto byte_at = [&](int off) { return (__shfl_sync(kFull, w, off >> 2) >> ((off & 3) * 8)) & 0xFFu; };
    if (lane < 28) occ[lane] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const int me = (int)(byte_at(81) & 1u), op = me ^ 1;
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        const int sq = blk * 32 + lane;
        const uint32_t raw = byte_at(sq < 81 ? sq : 0);
        const int code = sq < 81 ? (int)raw : 0;                  // 0 = empty, else 1 + type + 14 * colour
        const uint32_t same = __match_any_sync(kFull, code);
        if (code >= 1 && code <= 28 && (int)(__ffs(same) - 1) == lane) {
            const int colour = (code - 1) / 14, pt = (code - 1) % 14;
            uint32_t* o = reinterpret_cast<uint32_t*>(occ + (colour == me ? 0 : 14) + pt);
            o[blk] = same;
        }
    }
    __syncwarp();
    const uint32_t plies = __shfl_sync(kFull, w, 24);             // ply | max_ply << 16
    const uint32_t bdv = __shfl_sync(kFull, w, 25), wdv = __shfl_sync(kFull, w, 26);
    const uint64_t rot = (uint64_t)me << 24;
    const uint64_t one = (uint64_t)0x3F800000u << 32;
    const uint64_t all_lo = (1ull << 63) - 1ull, all_hi = 0x3FFFFull;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int c = lane + 32 * it;
        // stand channels: shuffle the hand count in (every lane takes part; others read offset 82)
        int need = 0, hand_off = 82;
        if (c >= 28 && c < 80) {
            const int side = (c - 28) / 26, k = (c - 28) % 26;
            const int piece = stand_piece_of(k, &need);
            hand_off = 82 + 7 * (side == 0 ? me : op) + piece;
        }
        const int in_hand = (int)byte_at(hand_off);
        if (c >= NSB_FEATURE_CHANNELS) continue;
        uint64_t lo = 0, hi = 0, val = one;
        if (c < 28) {   // squares 0..62 -> lo bits 0..62, squares 63..80 -> hi bits 0..17
            const uint4 o = occ[c];
            lo = (uint64_t)o.x | ((uint64_t)(o.y & 0x7FFFFFFFu) << 32);
            hi = (uint64_t)(o.y >> 31) | ((uint64_t)(o.z & 0x1FFFFu) << 1);
        } else if (c < 80) {
            const int on = in_hand >= need;
            lo = on ? all_lo : 0;
            hi = on ? all_hi : 0;
        } else if (c < 82) {
            const int on = (c - 80) == me;
            lo = on ? all_lo : 0;
            hi = on ? all_hi : 0;
        } else {
            lo = all_lo;
            hi = all_hi;
            const uint32_t ply = plies & 0xFFFFu, max_ply = plies >> 16;
            const float maxply = (float)(max_ply ? max_ply : 1u);
            float v;
            if (c == 82) v = (float)ply / maxply;
            else if (c == 83) v = 1.0f / maxply;
            else if (c == 84) v = __uint_as_float(me == 0 ? bdv : wdv);
            else v = __uint_as_float(me == 0 ? wdv : bdv);
            val = (uint64_t)__float_as_uint(v) << 32;
        }
        hi |= rot | val;
        emit(c, make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)));
    }
}

}  // namespace nsb

#endif
