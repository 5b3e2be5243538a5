// trunk_ts.cu — the position-stationary trunk for 128-channel nets with the WEIGHTS FED THROUGH
// TENSOR MEMORY (sm_100a).  Same contract as trunk_fused.cu (reference src/infer/trt.cc:256-261).
// EXPERIMENTAL: selected with NSB_TRUNK128=ts, parity-tested, but measured 0-7 % slower end to end than
// trunk_fused.cu (per layer 8.8-9.6 k cycles against 8.8-9.3 k; DESIGN.md §6.1 has the numbers and what
// they say about where the time goes).  Kept because the operand path it demonstrates - A from tensor
// memory at the tensor floor, no shared-memory traffic for weights - is the basis for the next step.
//
// Why: with both MMA operands in shared memory the kernel was bound by shared-
// memory bandwidth, not by the tensor pipe - every K = 16 step read 4 KB of weights (A) and 6 KB of
// activations (B) and the weight ring wrote another 4 KB, ~146 B/clk against 128 - and any epilogue
// work overlapped with the MMAs slowed them down by as much as it hid.  tcgen05.mma can take A from
// tensor memory instead.  Four producer warps (one per TMEM lane quadrant) load the weights straight
// from L2 into registers (a thread owns one accumulator row: 32 bytes per K step, coalesced 1 KB per
// warp) and store them into a 4-stage ring of TMEM columns with tcgen05.st; the MMA then reads only
// the activations from shared memory (61 B/clk).  Measured on the probe: N = 96 drops from 80 to 49.7
// cycles per MMA (the tensor floor), N = 192 stays at 97.7 with or without concurrent shared-memory
// stores.  That headroom is what makes the epilogue overlap pay: two accumulators alternate between
// layers, the weights of a layer are ordered K block by K block, and the accumulator rows are
// permuted so that the first half of an epilogue produces exactly K block 0 of the next layer - whose
// MMAs start while the second half is still running.  The shared memory the weight ring occupied
// is simply not used any more.
//
// Warp roles (512 threads, registers re-balanced with setmaxnreg: 128 / 168 / 40): warps 0-3 = weight producers,
// warps 4-11 = expansion / epilogues / heads / tail, warp 12 = MMA issuer + TMEM owner.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "trunk_common.cuh"

namespace nsb {

namespace {

struct TsGeom {
    static constexpr int C = 128;
    static constexpr int KCH = C / 8;
    static constexpr int NPOS = 2;
    static constexpr int NCOLS = NPOS * 96;
    static constexpr int GUARD = 12;
    static constexpr int SPITCH = (GUARD + NCOLS + 11) | 1;
    static constexpr int BUF_BYTES = ((KCH * SPITCH * 16 + 127) / 128) * 128;
    static constexpr int KC64 = C / 64;
    // tensor memory: two accumulators (layer parity) + the A ring
    static constexpr int TMEM_COLS = 512;
    static constexpr int A_COL0 = 2 * NCOLS;     // 384
    static constexpr int A_STAGES = 4;           // ring stages of 4 K steps = 32 columns each
    static constexpr int A_STAGE_COLS = 32;
    static constexpr int THREADS = 512;
    static constexpr int NBARS = 2 * A_STAGES + KC64 + 1;  // a_full[], a_empty[], kb[], acc
    static constexpr int SCRATCH_BYTES = ((NPOS * kPolicySize * 4 + 15) / 16) * 16;
    static constexpr int OFF_BUF_A = 0;
    static constexpr int OFF_BUF_B = OFF_BUF_A + BUF_BYTES;
    static constexpr int OFF_SCRATCH = OFF_BUF_B + BUF_BYTES;
    static constexpr int OFF_FEAT = OFF_SCRATCH + SCRATCH_BYTES;
    static constexpr int OFF_VBUF = OFF_FEAT + NPOS * kMaxInChannels * 16;
    static constexpr int OFF_RED = OFF_VBUF + ((NPOS * 81 * 4 + 15) / 16) * 16;
    static constexpr int OFF_BARS = OFF_RED + ((8 * NPOS * 2 * 4 + NPOS * 2 * 4 + 15) / 16) * 16;
    static constexpr int SMEM_BYTES = OFF_BARS + NBARS * 8 + 16 + 128;
    static_assert(A_COL0 + A_STAGES * A_STAGE_COLS <= TMEM_COLS, "TMEM columns");
};

// K = 16 steps of weight-stream stage (layer L, K block kc, any tap): the stem's second block holds
// input channels 64..95 only
__device__ __forceinline__ int stage_steps(const DeviceNet& net, int L, int kc) { return (L == 0 && kc == 1) ? net.stem_steps - 4 : 4; }

__global__ void __launch_bounds__(TsGeom::THREADS, 1) trunk_ts_kernel(const DeviceNet net, const EvalArgs a) {
    using G = TsGeom;
    constexpr int C = G::C;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bufA = sbase + G::OFF_BUF_A, bufB = sbase + G::OFF_BUF_B;
    float* scratch = reinterpret_cast<float*>(smem + G::OFF_SCRATCH);
    uint4* featS = reinterpret_cast<uint4*>(smem + G::OFF_FEAT);
    float* vbuf = reinterpret_cast<float*>(smem + G::OFF_VBUF);
    float* red = reinterpret_cast<float*>(smem + G::OFF_RED);
    const uint32_t bars = sbase + G::OFF_BARS;
    auto bar_afull = [&](int s) { return bars + 8u * s; };
    auto bar_aempty = [&](int s) { return bars + 8u * (G::A_STAGES + s); };
    auto bar_kb = [&](int kc) { return bars + 8u * (2 * G::A_STAGES + kc); };  // K block kc of the next input is complete
    const uint32_t bar_acc = bars + 8u * (2 * G::A_STAGES + G::KC64);
    volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + G::OFF_BARS + 8 * G::NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_eff = eval_count(a);
    const int groups = (n_eff + G::NPOS - 1) / G::NPOS;
    const int my_passes =
        (int)blockIdx.x < groups ? (groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int NL = net.num_layers;
    const bool stamp = eval_timeline(a) && blockIdx.x == 0;

    if (stamp && threadIdx.x == 128) eval_timeline(a)[4 * NL + 0] = clock64();
    // ---- one-time setup ---------------------------------------------------------------------
    for (int i = threadIdx.x; i < 2 * G::BUF_BYTES / 16; i += G::THREADS)
        reinterpret_cast<uint4*>(smem + G::OFF_BUF_A)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::A_STAGES; ++s) {
            mbar_init(bar_afull(s), 4);   // one arrival per producer warp
            mbar_init(bar_aempty(s), 1);  // tcgen05.commit of the MMAs that read the stage
        }
        for (int kc = 0; kc < G::KC64; ++kc) mbar_init(bar_kb(kc), kEpiWarps);  // one arrival per epilogue warp
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 12) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), G::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (stamp && threadIdx.x == 128) eval_timeline(a)[4 * NL + 1] = clock64();

    if (warp < 4) {
        // ===== weight producers: L2 -> registers -> tensor memory ================================
        // Thread = accumulator row = TMEM lane.  The stream holds, per K = 16 step, 128 rows x 32 B;
        // a stage is the 4 (stem, second block: 2) steps of one (K block, tap).  Loads of the next
        // stage are in flight while this one waits for its ring slot.
        const int row = threadIdx.x;  // 0..127
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + G::A_COL0;
        uint32_t slot = 0, phase = 0;
        struct Cursor {  // a stage of the weight stream: (layer, K block, tap) and where my row of it starts
            int L, kc, tap;
            const uint4* src;
        };
        auto advance = [&](Cursor& c) {
            c.src += (size_t)stage_steps(net, c.L, c.kc) * 256;
            const int ntaps = (c.L == NL - 1) ? 1 : 9;
            if (++c.tap == ntaps) {
                c.tap = 0;
                if (++c.kc == G::KC64) { c.kc = 0; ++c.L; }
            }
        };
        auto load_stage = [&](uint4 (&dst)[8], const Cursor& c) {
            const int nk = stage_steps(net, c.L, c.kc);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nk) {
                    dst[2 * k] = __ldg(c.src + (size_t)k * 256);
                    dst[2 * k + 1] = __ldg(c.src + (size_t)k * 256 + 1);
                }
        };
        auto store_stage = [&](const uint4 (&buf)[8], const Cursor& c) {
            mbar_wait(bar_aempty(slot), phase ^ 1u);
            tc_fence_after();
            if (stage_steps(net, c.L, c.kc) == 4) tmem_st_32x32b_x32(lane_addr + slot * G::A_STAGE_COLS, buf);
            else tmem_st_32x32b_x16(lane_addr + slot * G::A_STAGE_COLS, buf);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull(slot));
            if (++slot == G::A_STAGES) { slot = 0; phase ^= 1u; }
        };
        for (int p = 0; p < my_passes; ++p) {
            // three register stages rotate: two are in flight from L2 while one is stored
            Cursor ld{0, 0, 0, reinterpret_cast<const uint4*>(net.tiles) + (size_t)row * 2}, st = ld;
            uint4 b0[8], b1[8], b2[8];
            load_stage(b0, ld); advance(ld);
            load_stage(b1, ld); advance(ld);
            load_stage(b2, ld); advance(ld);
            for (;;) {
                store_stage(b0, st); advance(st);
                if (ld.L < NL) { load_stage(b0, ld); advance(ld); }
                if (st.L >= NL) break;
                store_stage(b1, st); advance(st);
                if (ld.L < NL) { load_stage(b1, ld); advance(ld); }
                if (st.L >= NL) break;
                store_stage(b2, st); advance(st);
                if (ld.L < NL) { load_stage(b2, ld); advance(ld); }
                if (st.L >= NL) break;
            }
        }
    } else if (warp >= 12) {
        setmaxnreg_dec<40>();
        if (warp == 12) {
            // ===== MMA issuer: warp-uniform loop, one elected lane issues ============================
            // K blocks outermost: block kc of this layer's input is complete when bar_kb[kc] fires, so
            // the MMAs start on block 0 while the previous epilogue is still producing block 1.
            constexpr uint32_t idesc = make_idesc_bf16_f32(128, G::NCOLS);
            constexpr uint32_t b_lbo = G::SPITCH * 16;
            uint32_t slot = 0, phase = 0, kb_phase = 0;
            long long wait_cycles = 0;  // diagnostics: time spent waiting for weights
            for (int p = 0; p < my_passes; ++p) {
                for (int L = 0; L < NL; ++L) {
                    const bool head = (L == NL - 1);
                    const uint32_t in_buf = (L & 1) ? bufA : bufB;
                    const uint32_t acc = tmem_base + (uint32_t)(L & 1) * G::NCOLS;
                    const int ntaps = head ? 1 : 9;
                    for (int kc = 0; kc < G::KC64; ++kc) {
                        mbar_wait(bar_kb(kc), kb_phase);
                        tc_fence_after();
                        if (stamp && p == 0 && lane == 0 && kc == 0) eval_timeline(a)[4 * L + 0] = clock64();
                        const int nk = stage_steps(net, L, kc);
                        for (int tap = 0; tap < ntaps; ++tap) {
                            const int shift = head ? 0 : (tap / 3 - 1) * 10 + (tap % 3 - 1);
                            const uint32_t b_base = in_buf + (uint32_t)((kc * 8 * G::SPITCH + G::GUARD + shift) * 16);
                            const long long w0 = stamp ? clock64() : 0;
                            mbar_wait(bar_afull(slot), phase);
                            if (stamp) wait_cycles += clock64() - w0;
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t a_base = tmem_base + G::A_COL0 + slot * G::A_STAGE_COLS;
                                const uint32_t b_lo = smem_desc_lo(b_base, b_lbo);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (k >= nk) break;
                                    umma_bf16_ts(acc, a_base + k * 8, smem_desc_from(b_lo + (uint32_t)(2 * k * G::SPITCH), 128), idesc,
                                                 (uint32_t)((kc | tap | k) != 0));
                                }
                                umma_commit(bar_aempty(slot));  // frees the ring stage when the MMAs retire
                            }
                            __syncwarp();
                            if (++slot == G::A_STAGES) { slot = 0; phase ^= 1u; }
                        }
                    }
                    kb_phase ^= 1u;
                    if (elect_one()) umma_commit(bar_acc);  // accumulator of layer L complete
                    __syncwarp();
                    if (stamp && p == 0 && lane == 0) eval_timeline(a)[4 * L + 1] = clock64();
                }
            }
            if (stamp && lane == 0) eval_timeline(a)[4 * NL + 12] = (unsigned long long)wait_cycles;
        }
    } else {
        // ===== expansion + epilogues + heads =====================================================
        setmaxnreg_inc<168>();
        const int et = threadIdx.x - 128;  // 0..255
        const int ew = warp - 4;           // 0..7
        const int q = ew & 3;              // TMEM lane quadrant (== warp % 4)
        const int part = ew >> 2;          // column half (= position)
        const int e_col0 = 96 * part;
        if (stamp && et == 0) eval_timeline(a)[4 * NL + 11] = clock64();
        EpilogueMask<3> realmask;
        realmask.init(e_col0, lane);
        uint32_t acc_phase = 0;
        for (int p = 0; p < my_passes; ++p) {
            const int b0 = ((int)blockIdx.x + p * (int)gridDim.x) * G::NPOS;

            // -- stage 2 of feature extraction, straight into the stem's B operand (bufB) --------
            unsigned long long* tl = (stamp && p == 0 && et == 0) ? eval_timeline(a) + 4 * NL : nullptr;
            if (tl) tl[8] = clock64();
            expand_features<G::NPOS, G::SPITCH, G::GUARD>(net, a, n_eff, b0, featS, smem + G::OFF_BUF_B, et, tl);
            fence_proxy_async_smem();
            named_bar_sync(kEpiBar, kEpiThreads);  // the stem input has no block structure: both fire together
            if (lane == 0) {
                mbar_arrive(bar_kb(0));
                mbar_arrive(bar_kb(1));
            }
            if (tl) tl[2] = clock64();

            // -- conv layers: TMEM -> +bias (+skip) -> ReLU -> bf16 -> next layer's B operand ----
            for (int L = 0; L < NL - 1; ++L) {
                float bias[4];  // accumulator row 32q + 16lb + 8h + lane/4 = channel 64lb + 16q + 8h + lane/4
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    bias[k] = __ldg(net.bias + (size_t)L * C + 64 * (k >> 1) + 16 * q + 8 * (k & 1) + (lane >> 2));
                mbar_wait(bar_acc, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                if (stamp && p == 0 && et == 0) eval_timeline(a)[4 * L + 2] = clock64();
                const uint32_t out_buf = ((L & 1) ? bufB : bufA) + G::GUARD * 16;
                const bool residual = (L >= 2) && ((L & 1) == 0);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(L & 1) * G::NCOLS + e_col0;
#pragma unroll
                for (int lb = 0; lb < 2; ++lb) {  // lane block lb = K block lb of the next layer's input
                    if (residual)
                        epilogue_half<3, true, 8>(lb, taddr, out_buf, G::SPITCH * 16, 2 * q, e_col0, bias, realmask, lane);
                    else
                        epilogue_half<3, false, 8>(lb, taddr, out_buf, G::SPITCH * 16, 2 * q, e_col0, bias, realmask, lane);
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_kb(lb));
                }
                if (stamp && p == 0 && et == 0) eval_timeline(a)[4 * L + 3] = clock64();
            }

            // -- heads: accumulator row 32*(h/7) + h%7 holds head channel h (0..26 policy planes,
            //    27 = value conv), i.e. 7 useful lanes in every TMEM quadrant, so all epilogue warps help
            const int hp = 7 * q + lane;
            const float hbias = lane < 7 ? __ldg(net.bias + (size_t)(NL - 1) * C + hp) : 0.f;
            float wpre[kFcPrefetch];
            fc1_prefetch(net, et, wpre);
            mbar_wait(bar_acc, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((NL - 1) & 1) * G::NCOLS;
                if (part == 0)
                    head_read<0>(taddr, hbias, hp, scratch, vbuf, lane);
                else
                    head_read<96>(taddr + 96, hbias, hp, scratch, vbuf, lane);
                tc_fence_before();
            }
            named_bar_sync(kEpiBar, kEpiThreads);
            if (tl) tl[4] = clock64();
            heads_tail<G::NPOS>(net, a, n_eff, b0, scratch, vbuf, red, wpre, et, tl);
        }
    }

    // ---- teardown -------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        __syncwarp();
        tmem_dealloc(tmem_base, G::TMEM_COLS);
    }
}

}  // namespace

int trunk_ts_prepare() {
    cudaError_t e = cudaFuncSetAttribute(trunk_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TsGeom::SMEM_BYTES);
    if (e != cudaSuccess) {
        set_error("trunk ts: cudaFuncSetAttribute failed: %s (is this an sm_100a device?)", cudaGetErrorString(e));
        return NSB_ERR_NO_DEVICE;
    }
    return 0;
}

int launch_trunk_ts(const DeviceNet& net, const EvalArgs& a, int num_sms, cudaStream_t s) {
    if (a.n <= 0) return 0;
    const int groups = (a.n + TsGeom::NPOS - 1) / TsGeom::NPOS;
    const int grid = groups < num_sms ? groups : num_sms;
    trunk_ts_kernel<<<grid, TsGeom::THREADS, TsGeom::SMEM_BYTES, s>>>(net, a);
    return 1;
}

}  // namespace nsb
