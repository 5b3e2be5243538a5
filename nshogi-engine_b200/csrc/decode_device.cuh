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
// deterministic.  Rank counting over a row staged in shared memory: warp `part` of the `nparts` warps that share
// the row ranks the 32-move slices part, part + nparts, ...; each lane compares its move with all m values
// (broadcast shared-memory reads), ~4 m cycles per slice.
__device__ __forceinline__ void rank_row_coop(const float* vals, int m, int part, int nparts, int lane,
                                              uint16_t* __restrict__ order) {
    if (m > NSB_MAX_LEGAL_MOVES) m = NSB_MAX_LEGAL_MOVES;
    for (int k = part; 32 * k < m; k += nparts) {
        const int i = lane + 32 * k;
        const float mine = i < m ? vals[i] : 0.f;
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const float x = vals[j];
            rank += (x > mine || (x == mine && j < i)) ? 1 : 0;
        }
        if (i < m) order[rank] = (uint16_t)i;
    }
}

// Copies a row's rank order from its shared-memory staging to `dst` (global or mapped host memory) with one
// 2-byte store per lane and consecutive addresses across the warp, slices shared like in rank_row_coop (the
// scattered stores of the ranking itself would each be a PCIe transaction when dst is host memory).
__device__ __forceinline__ void rank_row_copy_out(const uint16_t* staged, int m, int part, int nparts, int lane,
                                                  uint16_t* __restrict__ dst) {
    if (m > NSB_MAX_LEGAL_MOVES) m = NSB_MAX_LEGAL_MOVES;
    for (int i = lane + 32 * part; i < m; i += 32 * nparts) dst[i] = staged[i];
}

// One warp decodes one position.  `logits` may point to shared or global memory.  `stage` (kStage only,
// shared memory, may alias the row's own logits): receives a copy of the m values written to `out` for
// rank_row_coop - all equal for a row with NaNs, so that its rank order is the identity.
// Returns the row's NaN flag (uniform across the warp).
template <bool kStage = false>
__device__ __forceinline__ bool warp_decode_row(const float* logits, const uint16_t* __restrict__ idx,
                                                int m, int mode, float win, float draw,
                                                float* __restrict__ out, int lane, float* stage = nullptr) {
    if (m > NSB_MAX_LEGAL_MOVES) m = NSB_MAX_LEGAL_MOVES;
    float v[kDecodePerLane];
    bool bad = isnan_bits(win) || isnan_bits(draw);
    float mx = -CUDART_INF_F;
    // all index loads first, back to back: with direct I/O they cross PCIe, and one round trip per 32-move
    // slice (a load, then the gather that depends on it, then the next load) costs microseconds
    uint32_t id[kDecodePerLane];
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        id[k] = j < m ? (uint32_t)__ldg(idx + j) : 0u;
    }
    // (keeps the compiler from sinking each load down to its gather: it did, one register for all of them)
    asm volatile("" : "+r"(id[0]), "+r"(id[1]), "+r"(id[2]), "+r"(id[3]), "+r"(id[4]), "+r"(id[5]), "+r"(id[6]), "+r"(id[7]),
                      "+r"(id[8]), "+r"(id[9]), "+r"(id[10]), "+r"(id[11]), "+r"(id[12]), "+r"(id[13]), "+r"(id[14]),
                      "+r"(id[15]), "+r"(id[16]), "+r"(id[17]), "+r"(id[18]));
    static_assert(kDecodePerLane == 19, "operand list above");
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        v[k] = 0.f;
        if (j < m) {  // gather: feedworker.cc:119-125, frame.cc:101-106
            v[k] = logits[id[k]];
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
        if (kStage) {
            __syncwarp();  // every lane has gathered: the logits may be overwritten
#pragma unroll
            for (int k = 0; k < kDecodePerLane; ++k) {
                const int j = lane + 32 * k;
                if (j < m) stage[j] = bad ? 0.f : v[k];
            }
        }
        return bad;
    }
    if (m <= 0) return bad;
    if (m == 1) {  // feedworker.cc:101-103
        if (lane == 0) {
            out[0] = 1.0f;
            if (kStage) stage[0] = 1.0f;
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
    if (kStage) {  // (a NaN row is uniform here: all ties, identity order)
#pragma unroll
        for (int k = 0; k < kDecodePerLane; ++k) {
            const int j = lane + 32 * k;
            if (j < m) stage[j] = v[k] * inv;
        }
    }
    return bad;
}

}  // namespace nsb
