// decode_device.cuh — warp-per-position policy decode shared by the standalone decode kernel
// (logits in HBM) and the fused trunk epilogue (logits still in shared memory).
// Semantics: reference src/mcts/feedworker.cc:56-136 (gather at ml::getMoveIndex slots, 1-move shortcut,
// softmax_ T=1; with NaNFallbackEnabled: :58-85 win / draw NaN = NaNFound only, :106-118 a NaN logit = uniform
// row) and src/selfplay/frame.cc:93-118 (raw logits to the cache, then the softmax unless Gumbel root).
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

// What the decode hands to the evaluation cache: called once per row, by every lane, with the lane's values
// (lane l holds row elements l + 32 k) - raw logits in the LOGITS / BOTH modes (frame.cc:110-114), probabilities
// in PROBS mode (feedworker.cc:134-135) - and only for rows whose NaN flag is clear.
struct NoCacheStore {
    __device__ __forceinline__ void operator()(const float (&)[kDecodePerLane]) const {}
};

// One warp decodes one position.  `logits` may point to shared or global memory.
//   mode      : NSB_DECODE_PROBS / LOGITS / BOTH, optionally | NSB_DECODE_NAN_FALLBACK (include/nsb.h)
//   row_flags : NSB_ROW_* bits of this position (BOTH mode)
//   logits_out: BOTH mode, optional: the raw gathered logits next to the probabilities in `out`
//   stage     : kStage only (shared memory, may alias the row's own logits): a copy of the m values written to
//               `out` for rank_row_coop - all equal when they contain a NaN, so that the rank order is the identity
// Returns the row's NaN flag = the reference's NaNFound (uniform across the warp); always false without
// NSB_DECODE_NAN_FALLBACK, as in the reference's default build of feedResult (src/context.h:103).
template <bool kStage = false, typename Store = NoCacheStore>
__device__ __forceinline__ bool warp_decode_row(const float* logits, const uint16_t* __restrict__ idx, int m, int mode,
                                                int row_flags, float win, float draw, float* __restrict__ out,
                                                float* __restrict__ logits_out, int lane, float* stage = nullptr,
                                                Store store = Store()) {
    if (m > NSB_MAX_LEGAL_MOVES) m = NSB_MAX_LEGAL_MOVES;
    if (m < 0) m = 0;
    const bool fallback = (mode & NSB_DECODE_NAN_FALLBACK) != 0;
    const int kind = mode & NSB_DECODE_MODE_MASK;
    // feedworker.cc:58-85: a NaN win / draw rate is NaNFound, but leaves the policy row alone
    const bool value_nan = isnan_bits(win) || isnan_bits(draw);
    float v[kDecodePerLane];
    if (kind == NSB_DECODE_PROBS && m == 1) {  // feedworker.cc:101-103: no gather, no NaN test of the logit
#pragma unroll
        for (int k = 0; k < kDecodePerLane; ++k) v[k] = 1.0f;
        if (lane == 0) {
            out[0] = 1.0f;
            if (kStage) stage[0] = 1.0f;
        }
        const bool nan_found = fallback && value_nan;
        if (!nan_found) store(v);
        return nan_found;
    }
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
    bool logit_nan = false;
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        v[k] = 0.f;
        if (j < m) {  // gather: feedworker.cc:106-110,119-125, frame.cc:96-107
            // ml::getMoveIndex never leaves [0, 2187); a slot beyond it (a caller's bug) reads the last one
            v[k] = logits[min(id[k], (uint32_t)(NSB_POLICY_SIZE - 1))];
            logit_nan |= isnan_bits(v[k]);
            mx = fmaxf(mx, v[k]);
        }
    }
    logit_nan = __any_sync(0xffffffffu, logit_nan);
    const bool nan_found = fallback && (logit_nan || value_nan);
    if (kind != NSB_DECODE_PROBS) {  // self-play: the raw logits are what the cache keeps (frame.cc:110-114)
        float* raw = kind == NSB_DECODE_LOGITS ? out : logits_out;
        if (raw != nullptr) {
#pragma unroll
            for (int k = 0; k < kDecodePerLane; ++k) {
                const int j = lane + 32 * k;
                if (j < m) raw[j] = v[k];
            }
        }
        if (!nan_found) store(v);
        if (kind == NSB_DECODE_LOGITS || (row_flags & NSB_ROW_SKIP_SOFTMAX)) {  // frame.cc:116-118 (Gumbel root)
            if (kind != NSB_DECODE_LOGITS) {
#pragma unroll
                for (int k = 0; k < kDecodePerLane; ++k) {
                    const int j = lane + 32 * k;
                    if (j < m) out[j] = v[k];
                }
            }
            if (kStage) {
                __syncwarp();  // every lane has gathered: the logits may be overwritten
#pragma unroll
                for (int k = 0; k < kDecodePerLane; ++k) {
                    const int j = lane + 32 * k;
                    if (j < m) stage[j] = logit_nan ? 0.f : v[k];
                }
            }
            return nan_found;
        }
    }
    if (kind == NSB_DECODE_PROBS && fallback && logit_nan) {
        // NaN fallback: every legal logit := 1 before the softmax (feedworker.cc:111-118)
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
            v[k] = expf(v[k] - mx);  // softmax_(x, n, 1.0f): feedworker.cc:127, frame.cc:117
            sum += v[k];
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    bool out_nan = false;
#pragma unroll
    for (int k = 0; k < kDecodePerLane; ++k) {
        const int j = lane + 32 * k;
        v[k] *= inv;
        if (j < m) {
            out[j] = v[k];
            out_nan |= isnan_bits(v[k]);
        }
    }
    if (kind == NSB_DECODE_PROBS && !nan_found) store(v);  // feedworker.cc:134-135
    if (kStage) {  // (a row of the NaN fallback is uniform here: all ties, identity order)
        out_nan = __any_sync(0xffffffffu, out_nan);
#pragma unroll
        for (int k = 0; k < kDecodePerLane; ++k) {
            const int j = lane + 32 * k;
            if (j < m) stage[j] = out_nan ? 0.f : v[k];
        }
    }
    return nan_found;
}

}  // namespace nsb
