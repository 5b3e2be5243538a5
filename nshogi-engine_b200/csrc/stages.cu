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
#include "pack_device.cuh"

namespace nsb {

// ---------------------------------------------------------------------------------------------
// extract (stage 2).  Both layouts first turn every feature bitboard into {w0, w1, w2, value}:
// bit t of the 81-bit string w2:w1:w0 is the plane's occupancy at OUTPUT position t, i.e. the
// reference's rotation (extractbit.cu:20,26: sq = rotate ? 80 - t : t) and its lo/hi split
// (:30-34) are resolved once per plane instead of once per element; an output element is then
// value-bits AND -(bit), the reference's integer expression (:36-37), never converted.
// ---------------------------------------------------------------------------------------------
constexpr int kPlanesPerBlock = 64;
constexpr int kExtractThreads = 256;

__device__ __forceinline__ uint4 plane_string(uint4 f) {
    uint32_t s0 = f.x;                                   // squares 0..31
    uint32_t s1 = (f.y & 0x7FFFFFFFu) | (f.z << 31);     // squares 32..63 (63 = hi bit 0)
    uint32_t s2 = (f.z >> 1) & 0x1FFFFu;                 // squares 64..80
    if ((f.z >> 24) & 1u) {                              // out[t] = in[80 - t] == (96-bit reversal) >> 15
        const uint32_t r0 = __brev(s2), r1 = __brev(s1), r2 = __brev(s0);
        s0 = __funnelshift_r(r0, r1, 15);
        s1 = __funnelshift_r(r1, r2, 15);
        s2 = r2 >> 15;
    }
    return make_uint4(s0, s1, s2, f.w);
}

// the 32 bits of the plane string starting at output position t (t < 96)
__device__ __forceinline__ uint32_t string_bits_from(uint4 p, int t) {
    const int wi = t >> 5;
    const uint32_t lo = wi == 0 ? p.x : (wi == 1 ? p.y : p.z);
    const uint32_t hi = wi == 0 ? p.y : (wi == 1 ? p.z : 0u);
    return __funnelshift_r(lo, hi, t & 31);
}

// Channels-first.  The NCHW output is a flat [planes][81] array (plane = b*C + c), so a block owns
// kPlanesPerBlock consecutive planes: 64 * 81 * 4 B = 20736 B, a multiple of 16, so every block's
// output window is 16-byte aligned and is written with 16-byte stores (4 consecutive positions of
// one plane, or the seam between two planes).
__global__ void __launch_bounds__(kExtractThreads)
extract_nchw_kernel(const uint4* __restrict__ fb, long long planes_total, uint32_t* __restrict__ out) {
    __shared__ uint4 s_pl[kPlanesPerBlock + 1];
    const long long plane0 = (long long)blockIdx.x * kPlanesPerBlock;
    const int nplanes = (int)min((long long)kPlanesPerBlock, planes_total - plane0);
    if (threadIdx.x < nplanes) s_pl[threadIdx.x] = plane_string(__ldg(&fb[plane0 + threadIdx.x]));
    if (threadIdx.x == kPlanesPerBlock) s_pl[kPlanesPerBlock] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int nelem = nplanes * 81;
    uint32_t* dst = out + plane0 * 81;
    const int nvec = nelem >> 2;
    for (int v = threadIdx.x; v < nvec; v += kExtractThreads) {
        const int idx = v * 4;
        const int pl = idx / 81, t = idx - pl * 81;
        const uint4 p = s_pl[pl];
        uint32_t bits = string_bits_from(p, t);      // positions t .. t+3 of this plane (bits past 80 are zero)
        uint32_t v0 = p.w, v1 = p.w, v2 = p.w, v3 = p.w;
        if (t > 77) {                                  // seam: the last 81 - t elements are this plane's
            const uint4 pn = s_pl[pl + 1];
            const int k = 81 - t;                      // 1..3 elements left in this plane
            bits |= pn.x << k;
            v1 = k > 1 ? p.w : pn.w;
            v2 = k > 2 ? p.w : pn.w;
            v3 = pn.w;
        }
        reinterpret_cast<uint4*>(dst)[v] = make_uint4(v0 & (0u - (bits & 1u)), v1 & (0u - ((bits >> 1) & 1u)),
                                                      v2 & (0u - ((bits >> 2) & 1u)), v3 & (0u - ((bits >> 3) & 1u)));
    }
    for (int idx = (nvec << 2) + threadIdx.x; idx < nelem; idx += kExtractThreads) {  // ragged tail
        const int pl = idx / 81, t = idx - pl * 81;
        const uint4 p = s_pl[pl];
        dst[idx] = p.w & (0u - (string_bits_from(p, t) & 1u));
    }
}

// Channels-last (reference extractbit.cu:41-68; dead under the reference's current config,
// globalconfig.h:20, but part of the extractBits<> API and of test_extractbit.cc).  One block per
// position; the [81][C] output window is written with coalesced 8-byte stores when C is even
// (the window is then 8-byte aligned for every position), 4-byte stores otherwise.
__global__ void __launch_bounds__(256)
extract_nhwc_kernel(const uint4* __restrict__ fb, int channels, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t s_x[];   // [3][C] plane-string words (word-major), then [C] values
    uint32_t* s_w = s_x;
    uint32_t* s_v = s_x + 3 * channels;
    const long long b = blockIdx.x;
    for (int c = threadIdx.x; c < channels; c += blockDim.x) {
        const uint4 p = plane_string(__ldg(&fb[b * channels + c]));
        s_w[c] = p.x;
        s_w[channels + c] = p.y;
        s_w[2 * channels + c] = p.z;
        s_v[c] = p.w;
    }
    __syncthreads();
    uint32_t* dst = out + b * 81 * channels;
    const int nelem = 81 * channels;
    if ((channels & 1) == 0) {
        const int npair = nelem >> 1, stride = (int)blockDim.x * 2;
        const int dt = stride / channels, dc = stride - dt * channels;
        int idx = 2 * (int)threadIdx.x;
        int t = idx / channels, c = idx - t * channels;
        for (int v = threadIdx.x; v < npair; v += blockDim.x) {
            const uint2 w = *reinterpret_cast<const uint2*>(&s_w[(t >> 5) * channels + c]);
            const uint2 val = *reinterpret_cast<const uint2*>(&s_v[c]);
            const int sh = t & 31;
            reinterpret_cast<uint2*>(dst)[v] =
                make_uint2(val.x & (0u - ((w.x >> sh) & 1u)), val.y & (0u - ((w.y >> sh) & 1u)));
            c += dc;
            t += dt;
            if (c >= channels) { c -= channels; ++t; }
        }
    } else {
        for (int idx = threadIdx.x; idx < nelem; idx += blockDim.x) {
            const int t = idx / channels, c = idx - t * channels;
            dst[idx] = s_v[c] & (0u - ((s_w[(t >> 5) * channels + c] >> (t & 31)) & 1u));
        }
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
        extract_nhwc_kernel<<<(unsigned)n, 256, (size_t)channels * 16 + 16, s>>>(
            reinterpret_cast<const uint4*>(d_fb), channels, reinterpret_cast<uint32_t*>(d_planes));
    }
    return 1;
}

// ---------------------------------------------------------------------------------------------
// pack: one warp per position (pack_device.cuh holds the warp routine, which the trunk kernels also
// run in their prologue when they are handed packed positions).  Output: 86 x 16 B, coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int kPackWarps = 4;

__global__ void __launch_bounds__(kPackWarps * 32)
pack_positions_kernel(const nsb_position* __restrict__ pos, int n, uint4* __restrict__ fb) {
    __shared__ uint4 s_occ[kPackWarps][28];   // occupancy words of the 28 board planes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kPackWarps + warp;
    if (b >= n) return;
    uint4* dst = fb + (long long)b * NSB_FEATURE_CHANNELS;
    pack_position_warp(pos + b, lane, s_occ[warp], [&](int c, uint4 f) { dst[c] = f; });
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
              uint8_t* __restrict__ flag, const uint8_t* __restrict__ row_flags, float* __restrict__ logits_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kDecodeWarps + warp;
    if (i >= n) return;
    const uint32_t b = off[i], e = off[i + 1];
    const bool bad = warp_decode_row(policy + (size_t)i * kPolicySize, idx + b, (int)(e - b), mode,
                                     row_flags ? (int)row_flags[i] : 0, win[i], draw[i], out + b,
                                     logits_out ? logits_out + b : nullptr, lane);
    if (flag && lane == 0) flag[i] = bad ? 1 : 0;
}

int launch_decode(const float* d_policy, const float* d_win, const float* d_draw, size_t n,
                  const uint32_t* d_off, const uint16_t* d_idx, int mode, float* d_out,
                  uint8_t* d_flag, cudaStream_t s, const uint8_t* d_row_flags, float* d_logits_out) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + kDecodeWarps - 1) / kDecodeWarps);
    decode_kernel<<<grid, kDecodeWarps * 32, 0, s>>>(d_policy, d_win, d_draw, (int)n, d_off, d_idx,
                                                     mode, d_out, d_flag, d_row_flags, d_logits_out);
    return 1;
}

}  // namespace nsb
