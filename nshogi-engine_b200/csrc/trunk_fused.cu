// trunk_fused.cu — position-stationary ResNet policy/value forward on tcgen05 (sm_100a).
//
// Replaces the reference's TensorRT enqueue (reference src/infer/trt.cc:256-261: extractBits +
// IExecutionContext::enqueueV3) with ONE persistent kernel per batch:
//
//   feature bitboards --(expand in-kernel)--> bf16 activations resident in SHARED MEMORY for the
//   whole network --> policy logits / value / draw (+ optional fused legal-move decode).
//
// Why this shape: leaf positions are independent 9x9 boards, so a CTA can own a
// few positions and run every layer on them without ever touching HBM for activations.  Each
// conv is computed transposed, D^T[Cout x slots] = W[Cout x Cin*9] * X^T, so the UMMA M dimension
// is Cout (128) and N is the number of board slots (192 = 2 positions, any multiple of 16).  The
// activations are stored as the UMMA K-major SWIZZLE_NONE B operand [Cin/8][slot][8] with a uniform
// 16-byte row pitch and a 10x10 zero-padded slot grid per position; a 3x3 tap (dh, dw) is then the
// SAME buffer with the descriptor start address moved by 10*dh + dw rows: im2col costs nothing.
// Weights stream from L2 through a 6-deep ring of 16 KB tiles filled by the bulk-copy engine.
//
// Warp roles (256 threads): warp 0 = weight producer, warp 1 = MMA issuer + TMEM owner,
// warps 4-7 = feature expansion, epilogues (TMEM -> bias/residual/ReLU -> bf16 -> smem), heads.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode_device.cuh"
#include "nsb_internal.h"
#include "umma.cuh"

namespace nsb {

namespace {

constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr uint32_t kEpiBar = 1;  // named barrier id for the 4 epilogue warps

__host__ __device__ constexpr bool is_real_slot(int n) {
    // slot n = 100*pos + 10*row + col; column 9 and row 9 are the permanent zero padding
    return (n % 100) < 90 && ((n % 100) % 10) < 9;
}

__device__ __forceinline__ uint32_t expand_bits(uint4 f, int t) {
    // reference src/cuda/extractbit.cu:20-37 on one (plane, t): fp32 bit pattern or 0
    const uint64_t lo = ((uint64_t)f.y << 32) | f.x, hi = ((uint64_t)f.w << 32) | f.z;
    const int rotate = (int)((hi >> 24) & 1ull);
    const int sq = rotate ? 80 - t : t;
    const int use_hi = sq >= 63;
    const uint64_t word = use_hi ? hi : lo;
    return ((uint32_t)(word >> (sq - 63 * use_hi)) & 1u) * (uint32_t)(hi >> 32);
}

template <int C>
__global__ void __launch_bounds__(kThreads, 1) trunk_fused_kernel(const DeviceNet net, const EvalArgs a) {
    using G = TrunkGeom<C>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bufA = sbase + G::OFF_BUF_A, bufB = sbase + G::OFF_BUF_B;
    const uint32_t ring = sbase + G::OFF_RING;
    float* scratch = reinterpret_cast<float*>(smem + G::OFF_SCRATCH);
    uint4* featS = reinterpret_cast<uint4*>(smem + G::OFF_FEAT);
    float* vbuf = reinterpret_cast<float*>(smem + G::OFF_VBUF);
    float* red = reinterpret_cast<float*>(smem + G::OFF_RED);  // [4][NPOS][2] + wd[NPOS][2]
    const uint32_t bars = sbase + G::OFF_BARS;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (G::NSTAGES + s); };
    const uint32_t bar_act = bars + 8u * (2 * G::NSTAGES);
    const uint32_t bar_acc = bars + 8u * (2 * G::NSTAGES + 1);
    volatile uint32_t* tmem_holder =
        reinterpret_cast<volatile uint32_t*>(smem + G::OFF_BARS + 8 * (2 * G::NSTAGES + 2));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = (a.n + G::NPOS - 1) / G::NPOS;
    const int my_passes =
        (int)blockIdx.x < groups ? (groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int NL = net.num_layers;

    // ---- one-time setup ---------------------------------------------------------------------
    for (int i = threadIdx.x; i < 2 * G::BUF_BYTES / 16; i += kThreads)
        reinterpret_cast<uint4*>(smem + G::OFF_BUF_A)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::NSTAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_act, kEpiThreads);
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), G::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===== weight producer: linear stream of 16 KB tiles, re-walked once per pass ==========
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int p = 0; p < my_passes; ++p) {
                const uint8_t* src = net.tiles;
                for (int s = 0; s < net.stages_per_pass; ++s, src += kStageBytes) {
                    mbar_wait(bar_empty(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(bar_full(stage), kStageBytes);
                    bulk_g2s(ring + stage * kStageBytes, src, kStageBytes, bar_full(stage));
                    if (++stage == G::NSTAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread drives the tensor core ====================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(128, G::NCOLS);
            constexpr uint32_t b_lbo = G::SPITCH * 16;
            uint32_t stage = 0, phase = 0, act_phase = 0;
            for (int p = 0; p < my_passes; ++p) {
                for (int L = 0; L < NL; ++L) {
                    mbar_wait(bar_act, act_phase);
                    act_phase ^= 1u;
                    tc_fence_after();
                    const bool head = (L == NL - 1);
                    const uint32_t in_buf = (L & 1) ? bufA : bufB;
                    const int ntaps = head ? 1 : 9;
                    const int kblocks = (L == 0) ? kStemCin / 64 : G::KC64;
                    const int nh = head ? 1 : G::NHALF;
                    for (int tap = 0; tap < ntaps; ++tap) {
                        const int shift = head ? 0 : (tap / 3 - 1) * 10 + (tap % 3 - 1);
                        for (int kc = 0; kc < kblocks; ++kc) {
                            for (int half = 0; half < nh; ++half) {
                                mbar_wait(bar_full(stage), phase);
                                tc_fence_after();
                                const uint32_t a_base = ring + stage * kStageBytes;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t adesc = make_smem_desc(a_base + k * 4096, 2048, 128);
                                    const uint32_t b_addr =
                                        in_buf + (uint32_t)(((kc * 8 + 2 * k) * G::SPITCH + G::GUARD + shift) * 16);
                                    const uint64_t bdesc = make_smem_desc(b_addr, b_lbo, 128);
                                    umma_bf16(tmem_base + half * G::NCOLS, adesc, bdesc, idesc,
                                              (uint32_t)((tap | kc | k) != 0));
                                }
                                umma_commit(bar_empty(stage));  // frees the ring slot when the MMAs retire
                                if (++stage == G::NSTAGES) { stage = 0; phase ^= 1u; }
                            }
                        }
                    }
                    umma_commit(bar_acc);  // accumulator(s) of layer L complete
                }
            }
        }
    } else if (warp >= 4) {
        // ===== expansion + epilogues + heads =====================================================
        const int et = threadIdx.x - 128;  // 0..127
        const int q = warp - 4;            // TMEM lane quadrant (== warp % 4)
        uint32_t acc_phase = 0;
        for (int p = 0; p < my_passes; ++p) {
            const int b0 = ((int)blockIdx.x + p * (int)gridDim.x) * G::NPOS;

            // -- stage 2 of feature extraction, straight into the stem's B operand (bufB) --------
            for (int i = et; i < G::NPOS * NSB_FEATURE_CHANNELS; i += kEpiThreads) {
                const int pos = i / NSB_FEATURE_CHANNELS, c = i - pos * NSB_FEATURE_CHANNELS;
                const int b = b0 + pos;
                featS[i] = b < a.n ? __ldg(reinterpret_cast<const uint4*>(a.features) +
                                           (size_t)b * NSB_FEATURE_CHANNELS + c)
                                   : make_uint4(0, 0, 0, 0);
            }
            named_bar_sync(kEpiBar, kEpiThreads);
            for (int item = et; item < G::NPOS * (kStemCin / 8) * 81; item += kEpiThreads) {
                const int pos = item / ((kStemCin / 8) * 81);
                const int r = item - pos * ((kStemCin / 8) * 81);
                const int j = r / 81, t = r - j * 81;
                const int slot = pos * 100 + (t / 9) * 10 + (t % 9);
                uint32_t w[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                    uint32_t pk = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = j * 8 + e2 * 2 + h;
                        uint32_t bits = 0;
                        if (c < NSB_FEATURE_CHANNELS)
                            bits = f32_to_bf16_bits(__uint_as_float(expand_bits(featS[pos * NSB_FEATURE_CHANNELS + c], t)));
                        pk |= bits << (16 * h);
                    }
                    w[e2] = pk;
                }
                *reinterpret_cast<uint4*>(smem + G::OFF_BUF_B + (size_t)((j * G::SPITCH + G::GUARD + slot) * 16)) =
                    make_uint4(w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_act);

            // -- conv layers: TMEM -> +bias (+skip) -> ReLU -> bf16 -> next layer's B operand ----
            for (int L = 0; L < NL - 1; ++L) {
                mbar_wait(bar_acc, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                uint8_t* out_buf = smem + ((L & 1) ? G::OFF_BUF_B : G::OFF_BUF_A);
                const bool residual = (L >= 2) && ((L & 1) == 0);
#pragma unroll
                for (int half = 0; half < G::NHALF; ++half) {
                    const int co = half * 128 + et;
                    const float bias = __ldg(net.bias + (size_t)L * C + co);
                    uint8_t* row = out_buf + (size_t)(((co >> 3) * G::SPITCH + G::GUARD) * 16 + (co & 7) * 2);
#pragma unroll
                    for (int j = 0; j < G::NCOLS / 32; ++j) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + half * G::NCOLS + j * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int n = j * 32 + i;
                            if (is_real_slot(n)) {
                                uint16_t* dst = reinterpret_cast<uint16_t*>(row + n * 16);
                                float x = __uint_as_float(v[i]) + bias;
                                if (residual) x += bf16_bits_to_f32(*dst);
                                *dst = f32_to_bf16_bits(fmaxf(x, 0.f));
                            }
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(bar_act);
            }

            // -- heads: rows 0..26 of the accumulator = policy planes, row 27 = value conv --------
            mbar_wait(bar_acc, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            if (q == 0) {
                const float bias = lane <= kPolicyPlanes ? __ldg(net.bias + (size_t)(NL - 1) * C + lane) : 0.f;
#pragma unroll
                for (int j = 0; j < G::NCOLS / 32; ++j) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + j * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = j * 32 + i;
                        if (is_real_slot(n)) {
                            const int pos = n / 100, m = n % 100, t = (m / 10) * 9 + (m % 10);
                            const float x = __uint_as_float(v[i]) + bias;
                            if (lane < kPolicyPlanes)
                                scratch[pos * kPolicySize + lane * 81 + t] = x;  // plane-major logits
                            else if (lane == kPolicyPlanes)
                                vbuf[pos * 81 + t] = fmaxf(x, 0.f);
                        }
                    }
                }
                tc_fence_before();
            }
            named_bar_sync(kEpiBar, kEpiThreads);

            if (a.policy != nullptr) {  // dense logits (the Infer contract, trt.cc:265-267)
                for (int idx = et; idx < G::NPOS * kPolicySize; idx += kEpiThreads) {
                    const int pos = idx / kPolicySize, b = b0 + pos;
                    if (b < a.n) a.policy[(size_t)b * kPolicySize + (idx - pos * kPolicySize)] = scratch[idx];
                }
            }
            {   // value MLP: FC(81 -> H) + ReLU, FC(H -> 2), sigmoid
                const int H = net.hidden;
                float o[G::NPOS][2];
#pragma unroll
                for (int pos = 0; pos < G::NPOS; ++pos) o[pos][0] = o[pos][1] = 0.f;
                for (int h = et; h < H; h += kEpiThreads) {
                    float acc[G::NPOS];
                    const float b1 = __ldg(net.fc1b + h);
#pragma unroll
                    for (int pos = 0; pos < G::NPOS; ++pos) acc[pos] = b1;
                    for (int t = 0; t < 81; ++t) {
                        const float w = __ldg(net.fc1t + (size_t)t * H + h);
#pragma unroll
                        for (int pos = 0; pos < G::NPOS; ++pos) acc[pos] += w * vbuf[pos * 81 + t];
                    }
                    const float w0 = __ldg(net.fc2 + h), w1 = __ldg(net.fc2 + H + h);
#pragma unroll
                    for (int pos = 0; pos < G::NPOS; ++pos) {
                        const float hid = fmaxf(acc[pos], 0.f);
                        o[pos][0] += w0 * hid;
                        o[pos][1] += w1 * hid;
                    }
                }
#pragma unroll
                for (int pos = 0; pos < G::NPOS; ++pos)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float s = warp_sum(o[pos][k]);
                        if (lane == 0) red[(q * G::NPOS + pos) * 2 + k] = s;
                    }
                named_bar_sync(kEpiBar, kEpiThreads);
                if (et < G::NPOS * 2) {
                    const int pos = et >> 1, k = et & 1;
                    float s = __ldg(net.fc2b + k);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) s += red[(qq * G::NPOS + pos) * 2 + k];
                    const float val = 1.0f / (1.0f + expf(-s));
                    red[4 * G::NPOS * 2 + pos * 2 + k] = val;
                    const int b = b0 + pos;
                    if (b < a.n) (k == 0 ? a.win : a.draw)[b] = val;
                }
            }
            if (a.move_off != nullptr) {  // fused decode on logits that never left shared memory
                named_bar_sync(kEpiBar, kEpiThreads);
                if (q < G::NPOS) {
                    const int b = b0 + q;
                    if (b < a.n) {
                        const uint32_t mb = __ldg(a.move_off + b), me = __ldg(a.move_off + b + 1);
                        const bool bad = warp_decode_row(scratch + q * kPolicySize, a.move_idx + mb, (int)(me - mb),
                                                         a.decode_mode, red[4 * G::NPOS * 2 + q * 2 + 0],
                                                         red[4 * G::NPOS * 2 + q * 2 + 1], a.legal_out + mb, lane);
                        if (a.nan_flag && lane == 0) a.nan_flag[b] = bad ? 1 : 0;
                    }
                }
            }
            named_bar_sync(kEpiBar, kEpiThreads);  // scratch / vbuf / red reusable
        }
    }

    // ---- teardown -------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

}  // namespace

int trunk_fused_prepare(int channels) {
    cudaError_t e;
    if (channels == 128)
        e = cudaFuncSetAttribute(trunk_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TrunkGeom<128>::SMEM_BYTES);
    else if (channels == 256)
        e = cudaFuncSetAttribute(trunk_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TrunkGeom<256>::SMEM_BYTES);
    else {
        set_error("trunk: unsupported width %d (128 or 256)", channels);
        return NSB_ERR_INVALID;
    }
    if (e != cudaSuccess) {
        set_error("trunk: cudaFuncSetAttribute failed: %s (is this an sm_100a device?)", cudaGetErrorString(e));
        return NSB_ERR_NO_DEVICE;
    }
    return 0;
}

int launch_trunk_fused(const DeviceNet& net, const EvalArgs& a, int num_sms, cudaStream_t s) {
    if (a.n <= 0) return 0;
    if (net.channels == 128) {
        using G = TrunkGeom<128>;
        const int groups = (a.n + G::NPOS - 1) / G::NPOS;
        const int grid = groups < num_sms ? groups : num_sms;
        trunk_fused_kernel<128><<<grid, kThreads, G::SMEM_BYTES, s>>>(net, a);
    } else {
        using G = TrunkGeom<256>;
        const int groups = (a.n + G::NPOS - 1) / G::NPOS;
        const int grid = groups < num_sms ? groups : num_sms;
        trunk_fused_kernel<256><<<grid, kThreads, G::SMEM_BYTES, s>>>(net, a);
    }
    return 1;
}

}  // namespace nsb
