// umma_probe.cu — microbenchmark + correctness probe for operand layouts of tcgen05.mma
// (diagnostics, not on the product path): measures the issue-to-retire rate of M128 x N x K16
// MMAs for (a) the SWIZZLE_NONE 16-byte-row layout with a row-shifted B start and (b) the
// SWIZZLE_128B layout (128-byte rows) with a row-shifted B start, and checks (b) numerically.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "nsb_internal.h"
#include "umma.cuh"

namespace nsb {
namespace {

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;            // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows x 128 B
    d |= 1ull << 46;
    d |= 2ull << 61;                   // SWIZZLE_128B
    return d;
}

// layout 0: [k/8][row][8] (pitch rows), layout 1: [k/64][row][64] with 16-byte chunks XOR (row & 7)
__device__ __forceinline__ uint32_t elem_off(int layout, int r, int k, int pitch_rows) {
    if (layout == 0) return (uint32_t)(((k >> 3) * pitch_rows + r) * 16 + (k & 7) * 2);
    const int kc = k >> 6, c = (k >> 3) & 7;
    return (uint32_t)(kc * pitch_rows * 128 + r * 128 + ((c ^ (r & 7)) << 4) + (k & 7) * 2);
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, int N, int K, int rows,
                  int shift, int layout, int iters, float* __restrict__ d, unsigned long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_rows = 128, b_rows = (rows + 7) & ~7;
    const size_t a_bytes = (size_t)a_rows * K * 2, b_bytes = (size_t)b_rows * K * 2;
    uint8_t* sa = smem;
    uint8_t* sb = sa + ((a_bytes + 1023) & ~(size_t)1023);
    uint8_t* tail = sb + ((b_bytes + 1023) & ~(size_t)1023);
    const uint32_t bar = smem_u32(tail);
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(tail + 8);
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x)
        *reinterpret_cast<uint16_t*>(sa + elem_off(layout, i / K, i % K, a_rows)) = a[i];
    for (int i = threadIdx.x; i < rows * K; i += blockDim.x)
        *reinterpret_cast<uint16_t*>(sb + elem_off(layout, i / K, i % K, b_rows)) = b[i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 16, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *holder;
    if (warp == 0) {   // warp-uniform loop, elected lane issues (same idiom as the trunk)
        const uint32_t idesc = make_idesc_bf16_f32(128, N);
        uint32_t parity = 0;
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
        for (int it = 0; it < iters + 1; ++it) {
            const unsigned long long t0 = clock64();
            const int reps = it == 0 ? 1 : 16;
            for (int rep = 0; rep < reps; ++rep) {
                if (elect_one()) {
                    if (layout == 0) {
                        // the trunk kernels' issue loop: descriptor low words advance by one integer add per
                        // K step (building both descriptors from scratch costs more than a small-N MMA takes)
                        const uint32_t a_lo = smem_desc_lo(a0, a_rows * 16);
                        const uint32_t b_lo = smem_desc_lo(b0 + (uint32_t)(shift * 16), b_rows * 16);
                        for (int k4 = 0; k4 < K / 64; ++k4) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int k16 = k4 * 4 + kk;
                                umma_bf16(tmem_base, smem_desc_from(a_lo + (uint32_t)(k16 * 2 * a_rows), 128),
                                          smem_desc_from(b_lo + (uint32_t)(k16 * 2 * b_rows), 128), idesc,
                                          (it == 0 && rep == 0 && k16 == 0) ? 0u : 1u);
                            }
                            if (iters & 1) umma_commit(bar + 16);
                        }
                    } else
                    for (int k16 = 0; k16 < K / 16; ++k16) {
                        uint64_t ad, bd;
                        if (layout == 0) {
                            ad = make_smem_desc(a0 + (uint32_t)(k16 * 2 * a_rows * 16), a_rows * 16, 128);
                            bd = make_smem_desc(b0 + (uint32_t)((k16 * 2 * b_rows + shift) * 16), b_rows * 16, 128);
                        } else {
                            const int kc = k16 >> 2, kk = k16 & 3;
                            ad = make_desc_sw128(a0 + (uint32_t)(kc * a_rows * 128 + kk * 32));
                            bd = make_desc_sw128(b0 + (uint32_t)(kc * b_rows * 128 + shift * 128 + kk * 32));
                        }
                        umma_bf16(tmem_base, ad, bd, idesc, (it == 0 && rep == 0 && k16 == 0) ? 0u : 1u);
                        // odd iteration counts: a tcgen05.commit after every 4th MMA, like the trunk's ring
                        if ((iters & 1) && (k16 & 3) == 3) umma_commit(bar + 16);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar);
            __syncwarp();
            mbar_wait(bar, parity);
            parity ^= 1u;
            const unsigned long long t1 = clock64();
            if (threadIdx.x == 0) {
                if (it == 1) cycles[0] = 0;
                if (it >= 1) cycles[0] += t1 - t0;
            }
        }
    }
    // only the first iteration's result is checked: re-run it cleanly after timing
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16_f32(128, N);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t ad, bd;
            if (layout == 0) {
                ad = make_smem_desc(smem_u32(sa) + (uint32_t)(k16 * 2 * a_rows * 16), a_rows * 16, 128);
                bd = make_smem_desc(smem_u32(sb) + (uint32_t)((k16 * 2 * b_rows + shift) * 16), b_rows * 16, 128);
            } else {
                const int kc = k16 >> 2, kk = k16 & 3;
                ad = make_desc_sw128(smem_u32(sa) + (uint32_t)(kc * a_rows * 128 + kk * 32));
                bd = make_desc_sw128(smem_u32(sb) + (uint32_t)(kc * b_rows * 128 + shift * 128 + kk * 32));
            }
            umma_bf16(tmem_base, ad, bd, idesc, k16 != 0);
        }
        umma_commit(bar);
        mbar_wait(bar, (uint32_t)((iters + 1) & 1));
    }
    __syncthreads();
    tc_fence_after();
    for (int j = 0; j < N / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) d[(size_t)threadIdx.x * N + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// The same probe with the A operand in TENSOR MEMORY (layout 4): every thread stores its row's K
// elements into TMEM columns [256, 256 + K/2) with tcgen05.st; the MMAs then read A from TMEM and
// only B from shared memory.  `smem_noise` > 0 makes warps 1-3 stream 16-byte stores through a
// scratch area of shared memory during the timed loop (the traffic a weight ring would add).
__global__ void __launch_bounds__(128, 1)
umma_ts_probe_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, int N, int K, int rows,
                     int shift, int smem_noise, int iters, float* __restrict__ d, unsigned long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_rows = (rows + 7) & ~7;
    const size_t b_bytes = (size_t)b_rows * K * 2;
    uint8_t* sb = smem;
    uint8_t* noise = sb + ((b_bytes + 1023) & ~(size_t)1023);   // 32 KB scratch
    uint8_t* tail = noise + 32768;
    const uint32_t bar = smem_u32(tail);
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(tail + 8);
    volatile int* stop = reinterpret_cast<volatile int*>(tail + 16);
    for (int i = threadIdx.x; i < rows * K; i += blockDim.x)
        *reinterpret_cast<uint16_t*>(sb + elem_off(0, i / K, i % K, b_rows)) = b[i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        *stop = 0;
    }
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *holder;
    const uint32_t a_tmem = tmem_base + 256;
    {   // my row of A -> TMEM, 8 columns per K = 16 step
        const uint32_t lane_addr = a_tmem + ((uint32_t)(warp * 32) << 16);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = (uint32_t)a[(size_t)threadIdx.x * K + k16 * 16 + 2 * j] |
                       ((uint32_t)a[(size_t)threadIdx.x * K + k16 * 16 + 2 * j + 1] << 16);
            tmem_st_32x32b_x8(lane_addr + k16 * 8, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16_f32(128, N);
    const uint32_t b0 = smem_u32(sb);
    auto issue_all = [&](bool fresh) {
        for (int k16 = 0; k16 < K / 16; ++k16) {
            const uint64_t bd = make_smem_desc(b0 + (uint32_t)((k16 * 2 * b_rows + shift) * 16), b_rows * 16, 128);
            umma_bf16_ts(tmem_base, a_tmem + k16 * 8, bd, idesc, (fresh && k16 == 0) ? 0u : 1u);
        }
    };
    // smem_noise == 2: the trunk's ring protocol around every `ring_mmas` MMAs (odd iters: 2, even: 4): a
    // wait on an already complete mbarrier, tcgen05.fence, elect, the MMAs, a commit, __syncwarp
    const int ring_mmas = (iters & 1) ? 2 : 4;
    if (threadIdx.x == 0) {
        mbar_init(bar + 32, 1);
        mbar_init(bar + 40, 1);
        fence_mbar_init();
        mbar_arrive(bar + 40);  // phase 0 of this one is complete for good
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t parity = 0;
        for (int it = 0; it < iters + 1; ++it) {
            const unsigned long long t0 = clock64();
            const int reps = it == 0 ? 1 : 16;
            for (int rep = 0; rep < reps; ++rep) {
                if (smem_noise == 2) {
                    for (int k0 = 0; k0 < K / 16; k0 += ring_mmas) {
                        mbar_wait(bar + 40, 0);
                        tc_fence_after();
                        if (elect_one()) {
                            for (int k16 = k0; k16 < k0 + ring_mmas; ++k16) {
                                const uint64_t bd =
                                    make_smem_desc(b0 + (uint32_t)((k16 * 2 * b_rows + shift) * 16), b_rows * 16, 128);
                                umma_bf16_ts(tmem_base, a_tmem + k16 * 8, bd, idesc, (it == 0 && rep == 0 && k16 == 0) ? 0u : 1u);
                            }
                            umma_commit(bar + 32);
                        }
                        __syncwarp();
                    }
                    continue;
                }
                if (elect_one()) issue_all(it == 0 && rep == 0);
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar);
            __syncwarp();
            mbar_wait(bar, parity);
            parity ^= 1u;
            const unsigned long long t1 = clock64();
            if (threadIdx.x == 0) {
                if (it == 1) cycles[0] = 0;
                if (it >= 1) cycles[0] += t1 - t0;
            }
        }
        if (elect_one()) {
            issue_all(true);
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, parity);
        if (threadIdx.x == 0) *stop = 1;
    } else if (smem_noise == 1) {
        uint4* dst = reinterpret_cast<uint4*>(noise);
        unsigned k = threadIdx.x;
        while (*stop == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                dst[k & 2047] = make_uint4(k, k, k, k);
                k += 96;
            }
        }
    }
    __syncthreads();
    tc_fence_after();
    for (int j = 0; j < N / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) d[(size_t)threadIdx.x * N + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// The same probe on a CTA pair: one cta_group::2 MMA of M = 256 x N x K16; CTA r holds A rows
// [128r, 128r + 128) and N/2 rows of B (b_rows rows stored per CTA, start shifted by `shift`).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_pair_probe_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, int N, int K, int rows,
                       int shift, int layout, int iters, float* __restrict__ d, unsigned long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t rank = cluster_ctarank();
    const int a_rows = 128, b_rows = (rows + 7) & ~7;
    const size_t a_bytes = (size_t)a_rows * K * 2, b_bytes = (size_t)b_rows * K * 2;
    uint8_t* sa = smem;
    uint8_t* sb = sa + ((a_bytes + 1023) & ~(size_t)1023);
    uint8_t* tail = sb + ((b_bytes + 1023) & ~(size_t)1023);
    const uint32_t bar = smem_u32(tail);
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(tail + 8);
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x)
        *reinterpret_cast<uint16_t*>(sa + elem_off(layout, i / K, i % K, a_rows)) = a[(size_t)rank * 128 * K + i];
    for (int i = threadIdx.x; i < rows * K; i += blockDim.x)
        *reinterpret_cast<uint16_t*>(sb + elem_off(layout, i / K, i % K, b_rows)) = b[(size_t)rank * rows * K + i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(holder)), 256);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *holder;
    const uint32_t idesc = make_idesc_bf16_f32(256, N);
    const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
    auto issue_all = [&](bool fresh) {
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t ad, bd;
            if (layout == 0) {
                ad = make_smem_desc(a0 + (uint32_t)(k16 * 2 * a_rows * 16), a_rows * 16, 128);
                bd = make_smem_desc(b0 + (uint32_t)((k16 * 2 * b_rows + shift) * 16), b_rows * 16, 128);
            } else {
                const int kc = k16 >> 2, kk = k16 & 3;
                ad = make_desc_sw128(a0 + (uint32_t)(kc * a_rows * 128 + kk * 32));
                bd = make_desc_sw128(b0 + (uint32_t)(kc * b_rows * 128 + shift * 128 + kk * 32));
            }
            umma_bf16_pair(tmem_base, ad, bd, idesc, (fresh && k16 == 0) ? 0u : 1u);
        }
    };
    uint32_t parity = 0;
    if (warp == 0) {
        for (int it = 0; it < iters + 1; ++it) {
            const unsigned long long t0 = clock64();
            if (rank == 0) {
                const int reps = it == 0 ? 1 : 16;
                for (int rep = 0; rep < reps; ++rep) {
                    if (elect_one()) issue_all(it == 0 && rep == 0);
                    __syncwarp();
                }
                if (elect_one()) umma_commit_pair(bar, 3);
                __syncwarp();
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
            const unsigned long long t1 = clock64();
            if (threadIdx.x == 0 && rank == 0) {
                if (it == 1) cycles[0] = 0;
                if (it >= 1) cycles[0] += t1 - t0;
            }
        }
        // only a clean single pass is checked
        if (rank == 0) {
            if (elect_one()) {
                issue_all(true);
                umma_commit_pair(bar, 3);
            }
            __syncwarp();
        }
        mbar_wait(bar, parity);
    }
    __syncthreads();
    tc_fence_after();
    for (int j = 0; j < N / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            d[((size_t)rank * 128 + threadIdx.x) * N + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tmem_base, 256);
}

}  // namespace

int umma_probe(int gpu, int n_cols, int k_elems, int shift_rows, int layout, int iters, float* max_err,
               double* cycles_per_mma) {
    // layout: 0 = SWIZZLE_NONE, 1 = SWIZZLE_128B, 2 / 3 = the same on a CTA pair (cta_group::2, M = 256),
    //         4 = A operand in tensor memory (B SWIZZLE_NONE), 5 = the same with concurrent shared-memory stores
    const bool ts = layout >= 4;
    const bool pair = layout == 2 || layout == 3;
    const int lay = ts ? 0 : (layout & 1);
    if (n_cols % 32 || n_cols < 32 || n_cols > 256 || k_elems % 64 || k_elems <= 0 || shift_rows < 0 ||
        shift_rows > 64 || iters < 1 || layout < 0 || layout > 6) {
        set_error("umma_probe: bad arguments");
        return NSB_ERR_INVALID;
    }
    if (cudaSetDevice(gpu) != cudaSuccess) return NSB_ERR_NO_DEVICE;
    const int N = n_cols, K = k_elems, M = pair ? 256 : 128;
    const int n_cta = pair ? N / 2 : N;            // B rows each CTA contributes
    const int rows = n_cta + shift_rows + 8;       // B rows each CTA stores
    const int nct = pair ? 2 : 1;
    std::vector<uint16_t> ha((size_t)M * K), hb((size_t)nct * rows * K);
    std::vector<float> fa(ha.size()), fb(hb.size());
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (float)((int)(s % 17) - 8) / 8.0f; };
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); };
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = rnd(); ha[i] = bits(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = rnd(); hb[i] = bits(fb[i]); }
    uint16_t *da = nullptr, *db = nullptr;
    float* dd = nullptr;
    unsigned long long* dc = nullptr;
    cudaError_t e = cudaMalloc(&da, ha.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&db, hb.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&dd, (size_t)M * N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dc, 16);
    if (e == cudaSuccess) e = cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)128 * K * 2 + (size_t)((rows + 7) & ~7) * K * 2 + 4096 + (ts ? 32768 : 0);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ts ? (const void*)umma_ts_probe_kernel
                                    : (pair ? (const void*)umma_pair_probe_kernel : (const void*)umma_probe_kernel),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        if (ts)
            umma_ts_probe_kernel<<<1, 128, smem>>>(da, db, N, K, rows, shift_rows, layout - 4, iters, dd, dc);
        else if (pair)
            umma_pair_probe_kernel<<<2, 128, smem>>>(da, db, N, K, rows, shift_rows, lay, iters, dd, dc);
        else
            umma_probe_kernel<<<1, 128, smem>>>(da, db, N, K, rows, shift_rows, lay, iters, dd, dc);
        e = cudaDeviceSynchronize();
    }
    std::vector<float> hd((size_t)M * N);
    unsigned long long hc[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    if (e != cudaSuccess) {
        set_error("umma_probe: CUDA error: %s", cudaGetErrorString(e));
        return NSB_ERR_CUDA;
    }
    float worst = 0.f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            const int cta = n / n_cta, r = n % n_cta + shift_rows;
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc += fa[(size_t)m * K + k] * fb[((size_t)cta * rows + r) * K + k];
            worst = fmaxf(worst, fabsf(acc - hd[(size_t)m * N + n]));
        }
    if (max_err) *max_err = worst;
    if (cycles_per_mma) *cycles_per_mma = (double)hc[0] / ((double)iters * 16.0 * (K / 16));
    return 0;
}

}  // namespace nsb

// ---- bulk-copy (UBLKCP) stream rate probe: L2-resident source -> shared-memory ring -----------------
namespace nsb {
namespace {
__global__ void __launch_bounds__(64, 1)
bulk_rate_kernel(const uint8_t* __restrict__ src, int src_bytes, int tile_bytes, int stages, int split, int iters,
                 unsigned long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const uint32_t ring = smem_u32(smem);
    const uint32_t bars = ring + (uint32_t)(stages * tile_bytes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bars + 8u * s, 1);
            mbar_init(bars + 8u * (stages + s), 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int tiles = src_bytes / tile_bytes;
    const int part = tile_bytes / split;
    if (threadIdx.x == 0) {            // producer
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it)
            for (int t = 0; t < tiles; ++t) {
                mbar_wait(bars + 8u * (stages + stage), phase ^ 1u);
                mbar_arrive_expect_tx(bars + 8u * stage, (uint32_t)tile_bytes);
                for (int k = 0; k < split; ++k)
                    bulk_g2s(ring + stage * tile_bytes + k * part, src + (size_t)t * tile_bytes + k * part, (uint32_t)part,
                             bars + 8u * stage);
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
            }
    } else if (threadIdx.x == 32) {    // consumer: release each stage as soon as it has landed
        uint32_t stage = 0, phase = 0;
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int t = 0; t < tiles; ++t) {
                mbar_wait(bars + 8u * stage, phase);
                mbar_arrive(bars + 8u * (stages + stage));
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1u; }
            }
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    }
}
}  // namespace

int bulk_rate_probe(int gpu, int ctas, int tile_bytes, int stages, int split, double* bytes_per_cycle) {
    if (cudaSetDevice(gpu) != cudaSuccess) return NSB_ERR_NO_DEVICE;
    if (tile_bytes % (16 * split) || stages < 1 || stages * tile_bytes > 200 * 1024 || ctas < 1) {
        set_error("bulk_rate_probe: bad arguments");
        return NSB_ERR_INVALID;
    }
    const int src_bytes = (288 * 1024 / tile_bytes) * tile_bytes, iters = 20;
    uint8_t* d = nullptr;
    unsigned long long* dc = nullptr;
    cudaError_t e = cudaMalloc(&d, src_bytes);
    if (e == cudaSuccess) e = cudaMemset(d, 1, src_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dc, 8);
    const size_t smem = (size_t)stages * tile_bytes + 16 * stages + 256;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        bulk_rate_kernel<<<ctas, 64, smem>>>(d, src_bytes, tile_bytes, stages, split, iters, dc);
        bulk_rate_kernel<<<ctas, 64, smem>>>(d, src_bytes, tile_bytes, stages, split, iters, dc);
        e = cudaDeviceSynchronize();
    }
    unsigned long long hc = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(dc);
    if (e != cudaSuccess) {
        set_error("bulk_rate_probe: %s", cudaGetErrorString(e));
        return NSB_ERR_CUDA;
    }
    if (bytes_per_cycle) *bytes_per_cycle = (double)src_bytes * iters / (double)hc;
    return 0;
}
}  // namespace nsb
