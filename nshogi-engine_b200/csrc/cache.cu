// cache.cu — kernels of the device-resident evaluation cache (cache_device.cuh): batch probe with
// miss compaction in front of the trunk launch, standalone batch store, clear.  The store that
// follows an evaluation is fused into the trunk kernels' tail (trunk_common.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "cache_device.cuh"
#include "decode_device.cuh"
#include "nsb_internal.h"

namespace nsb {

namespace {
constexpr int kCacheWarps = 4;

// One warp per position: a hit writes the position's outputs (legal-move row, win, draw), a miss
// appends the position to the list the trunk launch will evaluate.  A hit of a NSB_DECODE_BOTH request holds
// the raw logits self-play stores (frame.cc:110-114) and takes the rest of Frame::setEvaluation from there:
// logits_out (optional), then the softmax unless the row is a Gumbel root (frame.cc:116-118).
__global__ void __launch_bounds__(kCacheWarps * 32)
cache_probe_kernel(const DeviceCache c, const uint64_t* __restrict__ hashes, int n, const uint32_t* __restrict__ off,
                   float* __restrict__ legal, float* __restrict__ win, float* __restrict__ draw,
                   uint8_t* __restrict__ hit, uint8_t* __restrict__ nan_flag, int* __restrict__ miss_idx,
                   int* __restrict__ miss_count, uint16_t* __restrict__ order, int mode,
                   const uint8_t* __restrict__ row_flags, float* __restrict__ logits_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kCacheWarps + warp;
    if (b >= n) return;
    const uint32_t mb = off[b], me = off[b + 1];
    const int m = (int)(me - mb);
    float w = 0.f, d = 0.f;
    __shared__ float s_row[kCacheWarps][kCacheRowPerLane * 32];
    float v[kCacheRowPerLane];
    const bool found = cache_load_warp(c, hashes[b], m, v, &w, &d, lane);
    if (found) {
        bool has_nan = false;
        if ((mode & NSB_DECODE_MODE_MASK) == NSB_DECODE_BOTH) {
            if (logits_out != nullptr) {
#pragma unroll
                for (int k = 0; k < kCacheRowPerLane; ++k)
                    if (lane + 32 * k < m) logits_out[mb + lane + 32 * k] = v[k];
            }
            if (!(row_flags && (row_flags[b] & NSB_ROW_SKIP_SOFTMAX))) {  // the same arithmetic as warp_decode_row
                float mx = -CUDART_INF_F;
#pragma unroll
                for (int k = 0; k < kCacheRowPerLane; ++k)
                    if (lane + 32 * k < m) mx = fmaxf(mx, v[k]);
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < kCacheRowPerLane; ++k)
                    if (lane + 32 * k < m) {
                        v[k] = expf(v[k] - mx);
                        sum += v[k];
                    }
                sum = warp_sum(sum);
                const float inv = 1.0f / sum;
#pragma unroll
                for (int k = 0; k < kCacheRowPerLane; ++k) v[k] *= inv;
            }
        }
#pragma unroll
        for (int k = 0; k < kCacheRowPerLane; ++k)
            if (lane + 32 * k < m) {
                legal[mb + lane + 32 * k] = v[k];
                has_nan |= isnan_bits(v[k]);
            }
        if (order != nullptr) {  // a hit is ranked like an evaluated row (a row with NaNs: identity)
            has_nan = __any_sync(0xffffffffu, has_nan);
#pragma unroll
            for (int k = 0; k < kCacheRowPerLane; ++k) s_row[warp][lane + 32 * k] = has_nan ? 0.f : v[k];
            __syncwarp();
            rank_row_coop(s_row[warp], m, 0, 1, lane, order + mb);
        }
    }
    if (lane == 0) {
        hit[b] = found ? 1 : 0;
        if (found) {
            win[b] = w;
            draw[b] = d;
            if (nan_flag) nan_flag[b] = 0;  // rows with NaNFound are never stored (feedworker.cc:134)
        } else {
            miss_idx[atomicAdd(miss_count, 1)] = b;
        }
    }
}

__global__ void __launch_bounds__(kCacheWarps * 32)
cache_store_kernel(const DeviceCache c, const uint64_t* __restrict__ hashes, int n, const uint32_t* __restrict__ off,
                   const float* __restrict__ legal, const float* __restrict__ win, const float* __restrict__ draw,
                   const uint8_t* __restrict__ skip, uint8_t* __restrict__ stored) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kCacheWarps + warp;
    if (b >= n) return;
    bool ok = false;
    if (!(skip && skip[b])) {
        const uint32_t mb = off[b], me = off[b + 1];
        ok = cache_store_warp(c, hashes[b], (int)(me - mb), legal + mb, win[b], draw[b], lane);
    }
    if (stored && lane == 0) stored[b] = ok ? 1 : 0;
}

__global__ void cache_clear_kernel(const DeviceCache c) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < c.num_bundles;
         i += (unsigned long long)gridDim.x * blockDim.x)
        c.meta[i] = kCacheInitMeta;
}
}  // namespace

int launch_cache_probe(const DeviceCache& c, const uint64_t* d_hashes, size_t n, const uint32_t* d_off, float* d_legal,
                       float* d_win, float* d_draw, uint8_t* d_hit, uint8_t* d_nan_flag, int* d_miss_idx, int* d_miss_count,
                       cudaStream_t s, uint16_t* d_order, int mode, const uint8_t* d_row_flags, float* d_logits_out) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + kCacheWarps - 1) / kCacheWarps);
    cache_probe_kernel<<<grid, kCacheWarps * 32, 0, s>>>(c, d_hashes, (int)n, d_off, d_legal, d_win, d_draw, d_hit,
                                                         d_nan_flag, d_miss_idx, d_miss_count, d_order, mode, d_row_flags,
                                                         d_logits_out);
    return 1;
}

int launch_cache_store(const DeviceCache& c, const uint64_t* d_hashes, size_t n, const uint32_t* d_off,
                       const float* d_legal, const float* d_win, const float* d_draw, const uint8_t* d_skip,
                       uint8_t* d_stored, cudaStream_t s) {
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + kCacheWarps - 1) / kCacheWarps);
    cache_store_kernel<<<grid, kCacheWarps * 32, 0, s>>>(c, d_hashes, (int)n, d_off, d_legal, d_win, d_draw, d_skip,
                                                         d_stored);
    return 1;
}

int launch_cache_clear(const DeviceCache& c, cudaStream_t s) {
    if (c.num_bundles == 0) return 0;
    cache_clear_kernel<<<592, 256, 0, s>>>(c);
    return 1;
}

}  // namespace nsb
