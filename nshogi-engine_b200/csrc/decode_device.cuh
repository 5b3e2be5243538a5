// decode_device.cuh — warp-per-position policy decode shared by the standalone decode kernel
// (logits in HBM) and the fused trunk epilogue (logits still in shared memory).
// Semantics: reference src/mcts/feedworker.cc:100-136 (gather at ml::getMoveIndex slots,
// 1-move shortcut, NaN fallback, softmax_ T=1) and src/selfplay/frame.cc:96-114 (raw logits).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/nsb.h"

namespace nsb {

constexpr int kDecodePerLane = (NSB_MAX_LEGAL_MOVES + 31) / 32;  // 19 moves per lane at most

// reference src/math/math.h:23-39: NaN test on the bit pattern (survives fast-math).
__device__ __forceinline__ bool isnan_bits(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu) != 0u;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Rank order of a decoded row: order[r] = index (within the row) of the move with the r-th largest value, ties
// by lower index first - the permutation that the reference's Node::sort() (std::sort of the edges by decreasing
// probability, src/mcts/node.h:163-168, run on a feed thread for every leaf, feedworker.cc:129) applies, made
// deterministic.  Lane l holds v[k] = value of move l + 32 k.  Rank counting with warp shuffles: m^2 / 32
// comparisons per lane, ~1.6 k cycles for the typical 80 moves; nothing leaves registers.
__device__ __forceinline__ void warp_rank_row(const float (&v)[kDecodePerLane], int m, int lane,
                                              uint16_t* __restrict__ order) {
    int rank[kDecodePerLane];
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) rank[k] = 0;
#pragma unroll
    for (int kp = 0; kp < kDecodePerLane; ++kp) {
        if (32 * kp >= m) break;
        const float mine = v[kp];
        const int lim = m - 32 * kp < 32 ? m - 32 * kp : 32;
        for (int src = 0; src < lim; ++src) {
            const float x = __shfl_sync(0xffffffffu, mine, src);
            const int j = 32 * kp + src;
#pragma unroll
            for (int k = 0; k < kDecodePerLane; ++k) {
                if (32 * k >= m) break;
                rank[k] += (x > v[k] || (x == v[k] && j < lane + 32 * k)) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int i = lane + 32 * k;
        if (i < m) order[rank[k]] = (uint16_t)i;
    }
}

// One warp decodes one position.  `logits` may point to shared or global memory.  `order` (optional):
// the row's rank order (warp_rank_row) of the values written to `out`; identity for a row with NaNs.
// Returns the row's NaN flag (uniform across the warp).
__device__ __forceinline__ bool warp_decode_row(const float* logits, const uint16_t* __restrict__ idx,
                                                int m, int mode, float win, float draw,
                                                float* __restrict__ out, int lane,
                                                uint16_t* __restrict__ order = nullptr) {
    if (m > NSB_MAX_LEGAL_MOVES) m = NSB_MAX_LEGAL_MOVES;
    float v[kDecodePerLane];
    bool bad = isnan_bits(win) || isnan_bits(draw);
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        v[k] = 0.f;
        if (j < m) {  // gather: feedworker.cc:119-125, frame.cc:101-106
            v[k] = logits[idx[j]];
            bad |= isnan_bits(v[k]);
            mx = fmaxf(mx, v[k]);
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (mode == NSB_DECODE_LOGITS) {  // self-play caches raw logits (frame.cc:110-114)
#pragma unroll
        for (int k = 0; k < kDecodePerLane; ++k) {
            const int j = lane + 32 * k;
            if (j < m) out[j] = v[k];
        }
        if (order != nullptr) {
            if (bad) {
                for (int j = lane; j < m; j += 32) order[j] = (uint16_t)j;
            } else {
                warp_rank_row(v, m, lane, order);
            }
        }
        return bad;
    }
    if (m <= 0) return bad;
    if (m == 1) {  // feedworker.cc:101-103
        if (lane == 0) {
            out[0] = 1.0f;
            if (order != nullptr) order[0] = 0;
        }
        return bad;
    }
    if (bad) {  // NaN fallback: every legal logit := 1 before the softmax (feedworker.cc:111-118)
        mx = 1.0f;
#pragma unroll
        for (int k = 0; k < kDecodePerLane; ++k) v[k] = 1.0f;
    } else {
        mx = warp_max(mx);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        if (j < m) {
            v[k] = expf(v[k] - mx);  // softmax_(x, n, 1.0f): feedworker.cc:127
            sum += v[k];
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        if (j < m) out[j] = v[k] * inv;
    }
    if (order != nullptr) {
        if (bad) {  // uniform probabilities: keep the generation order
            for (int j = lane; j < m; j += 32) order[j] = (uint16_t)j;
        } else {    // exp is monotonic and `inv` positive: ranking the unnormalised values ranks the probabilities,
                    // except where two different exponentials round to one probability - rank what was written
#pragma unroll
            for (int k = 0; k < kDecodePerLane; ++k) v[k] *= inv;
            warp_rank_row(v, m, lane, order);
        }
    }
    return bad;
}

}  // namespace nsb
