// Semantics probe: cp.async.bulk .multicast::cluster + mbarrier tx accounting in a cluster of 2 (bring-up of NSB_TRUNK128=mc2).
// Each CTA multicasts its 8 KB slice into both CTAs; a CTA's barrier (expect_tx 16 KB) must complete only when BOTH slices are
// there - also when one rank sends 200,000 cycles late - and the data must be visible to generic loads and to the async proxy
// (read back with a bulk copy, the proxy tcgen05.mma reads through).  Result on B200: 0 early completions, 0 bad reads.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o mcast_probe tools/mcast_probe.cu && ./mcast_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SLICE 8192
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t crank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__global__ void probe(const uint32_t* src, int* out, int delay_rank, int iters, int self_in_mask, uint8_t* dump) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * SLICE);
    const uint32_t r = crank();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    csync();
    int early = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        // clear own buffer (generic), make visible to async proxy, sync cluster so nobody sends before the clear
        if (threadIdx.x == 0) {
            for (int i = 0; i < 2 * SLICE / 4; ++i) reinterpret_cast<volatile uint32_t*>(smem)[i] = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        csync();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(2 * SLICE) : "memory");
            if ((int)r == delay_rank) { long long t0 = clock64(); while (clock64() - t0 < 200000) {} }
            const uint16_t mask = self_in_mask ? 3 : (uint16_t)(1u << (r ^ 1u));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                         ::"r"(s32(smem + r * SLICE)), "l"(src + (it % 4) * (2 * SLICE / 4) + r * (SLICE / 4)), "r"(SLICE), "r"(s32(bar)), "h"(mask) : "memory");
            if (!self_in_mask)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(smem + r * SLICE)), "l"(src + (it % 4) * (2 * SLICE / 4) + r * (SLICE / 4)), "r"(SLICE), "r"(s32(bar)) : "memory");
            long long t0 = clock64();
            while (!try_wait(s32(bar), it & 1)) {}
            long long dt = clock64() - t0;
            // the same check through the async proxy (what tcgen05.mma uses): bulk-copy the buffer out right away
            if (dump) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dump + ((size_t)(blockIdx.x * iters + it)) * 2 * SLICE), "r"(s32(smem)), "r"(2 * SLICE) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            const volatile uint32_t* w = reinterpret_cast<const volatile uint32_t*>(smem);
            const uint32_t* want = src + (it % 4) * (2 * SLICE / 4);
            int miss = 0;
            for (int i = 0; i < 2 * SLICE / 4; i += 64) miss += w[i] != want[i];
            if (miss) ++bad;
            if ((int)r != delay_rank && delay_rank >= 0 && dt < 100000) ++early;  // completed long before the delayed peer sent
        }
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        out[blockIdx.x * 2 + 0] = early;
        out[blockIdx.x * 2 + 1] = bad;
    }
    __syncthreads();
    csync();
}
int main() {
    uint32_t* src; int* out; uint8_t* dump;
    cudaMalloc(&src, 4 * 2 * SLICE); cudaMalloc(&out, 64); cudaMalloc(&dump, (size_t)2 * 200 * 2 * SLICE);
    uint8_t* hd = new uint8_t[(size_t)2 * 200 * 2 * SLICE];
    uint32_t* h = new uint32_t[4 * 2 * SLICE / 4];
    for (int i = 0; i < 4 * 2 * SLICE / 4; ++i) h[i] = 0x9E3779B9u * (i + 1) | 1u;
    cudaMemcpy(src, h, 4 * 2 * SLICE, cudaMemcpyHostToDevice);
    for (int self_in_mask = 1; self_in_mask >= 0; --self_in_mask)
        for (int delay = -1; delay < 2; ++delay) {
            cudaMemset(out, 0xff, 64);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 2 * SLICE + 64;
            cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            cudaMemset(dump, 0, (size_t)2 * 200 * 2 * SLICE);
            cudaError_t e = cudaLaunchKernelEx(&cfg, probe, (const uint32_t*)src, out, delay, 200, self_in_mask, dump);
            cudaError_t e2 = cudaDeviceSynchronize();
            int ho[4]; cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(hd, dump, (size_t)2 * 200 * 2 * SLICE, cudaMemcpyDeviceToHost);
            int abad[2] = {0, 0}, half[2][2] = {{0, 0}, {0, 0}};
            for (int b = 0; b < 2; ++b)
                for (int it = 0; it < 200; ++it) {
                    const uint32_t* got = reinterpret_cast<const uint32_t*>(hd + ((size_t)(b * 200 + it)) * 2 * SLICE);
                    const uint32_t* want = h + (it % 4) * (2 * SLICE / 4);
                    int m0 = 0, m1 = 0;
                    for (int i = 0; i < SLICE / 4; ++i) { m0 += got[i] != want[i]; m1 += got[SLICE / 4 + i] != want[SLICE / 4 + i]; }
                    abad[b] += (m0 + m1) != 0; half[b][0] += m0 != 0; half[b][1] += m1 != 0;
                }
            printf("   async-proxy read-back: rank0 bad iterations %d (lower half %d, upper half %d) | rank1 bad %d (lower %d, upper %d)\n", abad[0], half[0][0], half[0][1], abad[1], half[1][0], half[1][1]);
            printf("self_in_mask %d delayed rank %2d: launch %s sync %s | rank0 early %d bad %d | rank1 early %d bad %d\n", self_in_mask, delay,
                   cudaGetErrorString(e), cudaGetErrorString(e2), ho[0], ho[1], ho[2], ho[3]);
        }
    return 0;
}
