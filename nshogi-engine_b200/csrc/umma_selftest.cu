// umma_selftest.cu — device self-test of the tcgen05 building block the trunk relies on:
// K-major SWIZZLE_NONE shared-memory descriptors whose start address is moved by an arbitrary
// number of 16-byte rows (the zero-copy 3x3 tap shift of trunk_fused.cu).  One CTA computes
// D[128 x N] = A[128 x K] * B[shift .. shift+N)[K]^T on the tensor core and the host compares it
// with a plain fp32 loop, so a wrong descriptor convention shows up here instead of as a silently
// wrong network.  (First B200 run: max error 0.0 for every shape in tests/test_gpu_parity.py.)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "epilogue.cuh"
#include "nsb_internal.h"
#include "umma.cuh"

namespace nsb {
namespace {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint16_t* __restrict__ a /*[128][K]*/, const uint16_t* __restrict__ b /*[rows][K]*/,
                     int N, int K, int rows, int shift, float* __restrict__ d /*[128][N]*/,
                     const float* __restrict__ bias /*[128]*/, const uint16_t* __restrict__ xres /*[16][N][8]*/,
                     uint16_t* __restrict__ e_plain /*[16][N][8]*/, uint16_t* __restrict__ e_res) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const int kch = K / 8;
    const int bp = rows | 1;  // B row pitch per K chunk (odd, like the trunk's SPITCH)
    uint8_t* sa = smem;                       // [kch][128][16 B]
    uint8_t* sb = sa + (size_t)kch * 128 * 16;  // [kch][bp][16 B]
    uint8_t* so = sb + (((size_t)kch * bp * 16 + 15) & ~(size_t)15);  // epilogue output [16][op][16 B]
    const int op = N | 1;
    uint8_t* tail = so + (size_t)16 * op * 16;
    const uint32_t bar = smem_u32(tail);
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(tail + 8);

    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<uint16_t*>(sa + ((size_t)(k >> 3) * 128 + r) * 16 + (k & 7) * 2) = a[i];
    }
    for (int i = threadIdx.x; i < rows * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<uint16_t*>(sb + ((size_t)(k >> 3) * bp + r) * 16 + (k & 7) * 2) = b[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *holder;

    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16_f32(128, N);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            const uint32_t a_addr = smem_u32(sa) + (uint32_t)(k16 * 2 * 128 * 16);
            const uint32_t b_addr = smem_u32(sb) + (uint32_t)((k16 * 2 * bp + shift) * 16);
            const uint64_t ad = make_smem_desc(a_addr, 128 * 16, 128);
            const uint64_t bd = make_smem_desc(b_addr, (uint32_t)bp * 16, 128);
            umma_bf16(tmem_base, ad, bd, idesc, k16 != 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int j = 0; j < N / 32; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) d[(size_t)threadIdx.x * N + j * 32 + i] = __uint_as_float(v[i]);
    }
    // the trunk's epilogue path (epilogue.cuh): 16x256b fragments + stmatrix.trans, plain and residual
    const int lane = threadIdx.x & 31;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < 16 * N * 8; i += blockDim.x) {
            const int c = i / (N * 8), r = (i / 8) % N, e = i % 8;
            *reinterpret_cast<uint16_t*>(so + ((size_t)c * op + r) * 16 + e * 2) = pass ? xres[i] : (uint16_t)0xFFFFu;
        }
        __syncthreads();
        float b4[4];
        for (int k = 0; k < 4; ++k) b4[k] = bias[warp * 32 + 8 * k + (lane >> 2)];
        for (int col0 = 0; col0 < N; col0 += 96) {
            EpilogueMask<3> mask;
            mask.init(col0, lane);
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + col0;
            if (pass)
                epilogue_warp<3, true>(taddr, smem_u32(so), (uint32_t)op * 16, warp * 4, col0, b4, mask, lane);
            else
                epilogue_warp<3, false>(taddr, smem_u32(so), (uint32_t)op * 16, warp * 4, col0, b4, mask, lane);
        }
        __syncthreads();
        uint16_t* dst = pass ? e_res : e_plain;
        for (int i = threadIdx.x; i < 16 * N * 8; i += blockDim.x) {
            const int c = i / (N * 8), r = (i / 8) % N, e = i % 8;
            dst[i] = *reinterpret_cast<uint16_t*>(so + ((size_t)c * op + r) * 16 + e * 2);
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

int umma_selftest(int gpu, int n_cols, int k_elems, int shift_rows, float* max_err, float* epi_err) {
    if (n_cols % 96 || n_cols < 32 || n_cols > 256 || k_elems % 16 || k_elems <= 0 || shift_rows < 0 ||
        shift_rows > 64) {
        set_error("umma_selftest: need N in {96,192}, K%%16==0, 0<=shift<=64");
        return NSB_ERR_INVALID;
    }
    if (cudaSetDevice(gpu) != cudaSuccess) {
        set_error("umma_selftest: cudaSetDevice(%d) failed", gpu);
        return NSB_ERR_NO_DEVICE;
    }
    const int N = n_cols, K = k_elems, rows = N + shift_rows + 8;
    std::vector<uint16_t> ha((size_t)128 * K), hb((size_t)rows * K);
    std::vector<float> fa(ha.size()), fb(hb.size());
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return (float)((int)(s % 17) - 8) / 8.0f;  // exactly representable in bf16
    };
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); };
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = rnd(); ha[i] = bits(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = rnd(); hb[i] = bits(fb[i]); }
    std::vector<float> hbias(128);
    std::vector<uint16_t> hx((size_t)16 * N * 8);
    std::vector<float> fx(hx.size());
    for (auto& b : hbias) b = rnd();
    for (size_t i = 0; i < hx.size(); ++i) { fx[i] = rnd(); hx[i] = bits(fx[i]); }
    uint16_t *da = nullptr, *db = nullptr, *dx = nullptr, *de0 = nullptr, *de1 = nullptr;
    float *dd = nullptr, *dbias = nullptr;
    cudaError_t e = cudaMalloc(&da, ha.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&dx, hx.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&de0, hx.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&de1, hx.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&dbias, 128 * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dbias, hbias.data(), 128 * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&db, hb.size() * 2);
    if (e == cudaSuccess) e = cudaMalloc(&dd, (size_t)128 * N * 4);
    if (e == cudaSuccess) e = cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    const int kch = K / 8, bp = rows | 1;
    const size_t smem = (size_t)kch * 128 * 16 + (size_t)kch * bp * 16 + (size_t)16 * (N | 1) * 16 + 64 + 128 + 16;
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        umma_selftest_kernel<<<1, 128, smem>>>(da, db, N, K, rows, shift_rows, dd, dbias, dx, de0, de1);
        e = cudaDeviceSynchronize();
    }
    std::vector<float> hd((size_t)128 * N);
    if (e == cudaSuccess) e = cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<uint16_t> he0(hx.size()), he1(hx.size());
    if (e == cudaSuccess) e = cudaMemcpy(he0.data(), de0, he0.size() * 2, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(he1.data(), de1, he1.size() * 2, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dx); cudaFree(de0); cudaFree(de1); cudaFree(dbias);
    if (e != cudaSuccess) {
        set_error("umma_selftest: CUDA error: %s", cudaGetErrorString(e));
        return NSB_ERR_CUDA;
    }
    float worst = 0.f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc += fa[(size_t)m * K + k] * fb[(size_t)(n + shift_rows) * K + k];
            worst = fmaxf(worst, fabsf(acc - hd[(size_t)m * N + n]));
        }
    if (max_err) *max_err = worst;
    // epilogue: bf16(relu(acc + bias [+ x])) at real slots, 0 at padding slots, layout [chunk][slot][8]
    auto bf = [](float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u; float r; memcpy(&r, &u, 4); return r; };
    auto unbits = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float r; memcpy(&r, &u, 4); return r; };
    float worst_e = 0.f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc += fa[(size_t)m * K + k] * fb[(size_t)(n + shift_rows) * K + k];
            const size_t idx = ((size_t)(m >> 3) * N + n) * 8 + (m & 7);
            const bool real = is_real_slot(n);
            const float w0 = real ? bf(fmaxf(acc + hbias[m], 0.f)) : 0.f;
            const float w1 = real ? bf(fmaxf(acc + hbias[m] + fx[idx], 0.f)) : 0.f;
            worst_e = fmaxf(worst_e, fabsf(w0 - unbits(he0[idx])));
            worst_e = fmaxf(worst_e, fabsf(w1 - unbits(he1[idx])));
        }
    if (epi_err) *epi_err = worst_e;
    return 0;
}

}  // namespace nsb
