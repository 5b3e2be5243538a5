// stages.cu — the HBM-bound kernels of the leaf-evaluation path:
//   pack    : position record -> 86 feature bitboards   (stage 1; libnshogi FeatureStackComptime,
//             called at reference src/selfplay/evaluationworker.cc:87-92)
//   extract : feature bitboards -> fp32 planes           (stage 2; reference src/cuda/extractbit.cu)
//   decode  : dense logits -> per-legal-move softmax     (reference src/mcts/feedworker.cc:100-136,
//             src/selfplay/frame.cc:93-118)
// All three are integer / byte movers: coalesced, vectorised, grid sized to the data.
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode_device.cuh"
#include "nsb_internal.h"

namespace nsb {

// ---------------------------------------------------------------------------------------------
// extract, channels-first.  The NCHW output is a flat [planes][81] array (plane = b*C + c), so
// a block owns kPlanesPerBlock consecutive planes: 64 * 81 * 4 B = 20736 B, a multiple of 16, so
// every block's output window is 16-byte aligned and is written with float4 stores.  The 64
// feature bitboards (1 KB) are staged in shared memory once; each output element is the
// reference's integer expression (extractbit.cu:20-37): bit(sq) * value-bits, never converted.
// ---------------------------------------------------------------------------------------------
constexpr int kPlanesPerBlock = 64;
constexpr int kExtractThreads = 256;

__device__ __forceinline__ uint32_t expand_bit(uint64_t lo, uint64_t hi, int t) {
    const int rotate = (int)((hi >> 24) & 1ull);          // extractbit.cu:20
    const uint32_t value = (uint32_t)(hi >> 32);          // :21
    const int sq = rotate ? 80 - t : t;                   // :26
    const int use_hi = sq >= 63;                          // :30-34
    const uint64_t word = use_hi ? hi : lo;
    const int sh = sq - 63 * use_hi;
    return ((uint32_t)(word >> sh) & 1u) * value;         // :36-37
}

__global__ void __launch_bounds__(kExtractThreads)
extract_nchw_kernel(const uint4* __restrict__ fb, long long planes_total, uint32_t* __restrict__ out) {
    __shared__ uint4 s_fb[kPlanesPerBlock];
    const long long plane0 = (long long)blockIdx.x * kPlanesPerBlock;
    const int nplanes = (int)min((long long)kPlanesPerBlock, planes_total - plane0);
    if (threadIdx.x < nplanes) s_fb[threadIdx.x] = __ldg(&fb[plane0 + threadIdx.x]);
    __syncthreads();
    const int nelem = nplanes * 81;
    uint32_t* dst = out + plane0 * 81;
    const int nvec = nelem >> 2;
    for (int v = threadIdx.x; v < nvec; v += kExtractThreads) {
        uint32_t r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = v * 4 + e;
            const int pl = idx / 81, t = idx - pl * 81;
            const uint4 f = s_fb[pl];
            r[e] = expand_bit(((uint64_t)f.y << 32) | f.x, ((uint64_t)f.w << 32) | f.z, t);
        }
        reinterpret_cast<uint4*>(dst)[v] = make_uint4(r[0], r[1], r[2], r[3]);
    }
    for (int idx = (nvec << 2) + threadIdx.x; idx < nelem; idx += kExtractThreads) {  // ragged tail
        const int pl = idx / 81, t = idx - pl * 81;
        const uint4 f = s_fb[pl];
        dst[idx] = expand_bit(((uint64_t)f.y << 32) | f.x, ((uint64_t)f.w << 32) | f.z, t);
    }
}

// extract, channels-last (reference extractbit.cu:41-68; dead under the reference's current
// config, globalconfig.h:20, but part of the extractBits<> API and of test_extractbit.cc).
// One block per position; the position's C bitboards are staged in shared memory and the
// [81][C] output window is written with coalesced 4-byte stores.
__global__ void __launch_bounds__(256)
extract_nhwc_kernel(const uint4* __restrict__ fb, int channels, uint32_t* __restrict__ out) {
    extern __shared__ uint4 s_fbx[];
    const long long b = blockIdx.x;
    for (int c = threadIdx.x; c < channels; c += blockDim.x) s_fbx[c] = __ldg(&fb[b * channels + c]);
    __syncthreads();
    uint32_t* dst = out + b * 81 * channels;
    const int nelem = 81 * channels;
    for (int idx = threadIdx.x; idx < nelem; idx += blockDim.x) {
        const int t = idx / channels, c = idx - t * channels;
        const uint4 f = s_fbx[c];
        dst[idx] = expand_bit(((uint64_t)f.y << 32) | f.x, ((uint64_t)f.w << 32) | f.z, t);
    }
}

int launch_extract(const nsb_feature_bitboard* d_fb, size_t n, int channels, int channels_first,
                   float* d_planes, cudaStream_t s) {
    if (n == 0) return 0;
    if (((uintptr_t)d_fb & 15) || ((uintptr_t)d_planes & 15)) {
        set_error("extract: pointers must be 16-byte aligned");
        return NSB_ERR_INVALID;
    }
    if (channels_first) {
        const long long planes = (long long)n * channels;
        const unsigned grid = (unsigned)((planes + kPlanesPerBlock - 1) / kPlanesPerBlock);
        extract_nchw_kernel<<<grid, kExtractThreads, 0, s>>>(reinterpret_cast<const uint4*>(d_fb),
                                                             planes, reinterpret_cast<uint32_t*>(d_planes));
    } else {
        extract_nhwc_kernel<<<(unsigned)n, 256, (size_t)channels * 16, s>>>(
            reinterpret_cast<const uint4*>(d_fb), channels, reinterpret_cast<uint32_t*>(d_planes));
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------
// pack: one warp per position.  Lanes scan the 81-byte board out of shared memory; lane c builds
// channel c, c+32, c+64 (channel order: reference src/evaluate/preset.h:20-66, semantics
// SURVEY.md App. A.2 — builder-defined, libnshogi absent).  Output: 86 x 16 B, coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int kPackWarps = 4;

__device__ __forceinline__ int stand_piece_of(int k /*0..25*/, int* need) {
    // P1-6 L1-4 N1-4 S1-4 G1-4 B1-2 R1-2
    if (k < 6) { *need = k + 1; return 0; }
    if (k < 10) { *need = k - 5; return 1; }
    if (k < 14) { *need = k - 9; return 2; }
    if (k < 18) { *need = k - 13; return 3; }
    if (k < 22) { *need = k - 17; return 4; }
    if (k < 24) { *need = k - 21; return 5; }
    *need = k - 23;
    return 6;
}

__global__ void __launch_bounds__(kPackWarps * 32)
pack_positions_kernel(const nsb_position* __restrict__ pos, int n, uint4* __restrict__ fb) {
    __shared__ __align__(16) unsigned char s_pos[kPackWarps][112];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kPackWarps + warp;
    if (b >= n) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pos + b);  // 108 B = 27 words
    if (lane < 27) reinterpret_cast<uint32_t*>(s_pos[warp])[lane] = __ldg(src + lane);
    __syncwarp();
    const nsb_position* p = reinterpret_cast<const nsb_position*>(s_pos[warp]);
    const int me = p->side & 1, op = me ^ 1;
    const uint64_t rot = (uint64_t)me << 24;
    const uint64_t one = (uint64_t)0x3F800000u << 32;
    const uint64_t all_lo = (1ull << 63) - 1ull, all_hi = 0x3FFFFull;
    for (int c = lane; c < NSB_FEATURE_CHANNELS; c += 32) {
        uint64_t lo = 0, hi = 0, val = one;
        if (c < 28) {
            const int colour = c < 14 ? me : op, pt = c < 14 ? c : c - 14;
            const int code = 1 + pt + 14 * colour;
            for (int s = 0; s < 63; ++s) lo |= (uint64_t)(p->board[s] == code) << s;
            for (int s = 63; s < 81; ++s) hi |= (uint64_t)(p->board[s] == code) << (s - 63);
        } else if (c < 80) {
            const int side = (c - 28) / 26, k = (c - 28) % 26;
            int need;
            const int piece = stand_piece_of(k, &need);
            const int on = p->hands[side == 0 ? me : op][piece] >= need;
            lo = on ? all_lo : 0;
            hi = on ? all_hi : 0;
        } else if (c < 82) {
            const int on = (c - 80) == me;
            lo = on ? all_lo : 0;
            hi = on ? all_hi : 0;
        } else {
            lo = all_lo;
            hi = all_hi;
            const float maxply = (float)(p->max_ply ? p->max_ply : 1);
            float v;
            if (c == 82) v = (float)p->ply / maxply;
            else if (c == 83) v = 1.0f / maxply;
            else if (c == 84) v = me == 0 ? p->black_draw_value : p->white_draw_value;
            else v = me == 0 ? p->white_draw_value : p->black_draw_value;
            val = (uint64_t)__float_as_uint(v) << 32;
        }
        hi |= rot | val;
        fb[(long long)b * NSB_FEATURE_CHANNELS + c] =
            make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
}

int launch_pack_positions(const nsb_position* d_pos, size_t n, nsb_feature_bitboard* d_fb,
                          cudaStream_t s) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + kPackWarps - 1) / kPackWarps);
    pack_positions_kernel<<<grid, kPackWarps * 32, 0, s>>>(d_pos, (int)n,
                                                           reinterpret_cast<uint4*>(d_fb));
    return 1;
}

// ---------------------------------------------------------------------------------------------
// decode from dense logits in HBM: one warp per position (decode_device.cuh holds the shared
// warp routine, also used by the fused trunk epilogue on logits that never left shared memory).
// ---------------------------------------------------------------------------------------------
constexpr int kDecodeWarps = 4;

__global__ void __launch_bounds__(kDecodeWarps * 32)
decode_kernel(const float* __restrict__ policy, const float* __restrict__ win,
              const float* __restrict__ draw, int n, const uint32_t* __restrict__ off,
              const uint16_t* __restrict__ idx, int mode, float* __restrict__ out,
              uint8_t* __restrict__ flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kDecodeWarps + warp;
    if (i >= n) return;
    const uint32_t b = off[i], e = off[i + 1];
    const bool bad = warp_decode_row(policy + (size_t)i * kPolicySize, idx + b, (int)(e - b), mode,
                                     win[i], draw[i], out + b, lane);
    if (flag && lane == 0) flag[i] = bad ? 1 : 0;
}

int launch_decode(const float* d_policy, const float* d_win, const float* d_draw, size_t n,
                  const uint32_t* d_off, const uint16_t* d_idx, int mode, float* d_out,
                  uint8_t* d_flag, cudaStream_t s) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + kDecodeWarps - 1) / kDecodeWarps);
    decode_kernel<<<grid, kDecodeWarps * 32, 0, s>>>(d_policy, d_win, d_draw, (int)n, d_off, d_idx,
                                                     mode, d_out, d_flag);
    return 1;
}

}  // namespace nsb
