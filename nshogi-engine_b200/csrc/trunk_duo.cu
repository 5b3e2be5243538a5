// trunk_duo.cu — 128-channel trunk built for TWO CTAs PER SM (sm_100a).  Same contract as
// trunk_fused.cu (reference src/infer/trt.cc:256-261).
//
// The one-CTA-per-SM kernels leave the tensor pipe idle whenever their CTA is not issuing MMAs:
// accumulator hand-off, epilogue, feature expansion, heads, value MLP, decode - about 30 % of a
// 128-channel launch.  Overlapping the epilogue with the next layer inside one CTA did not pay
// (DESIGN.md §6.1).  Two independent CTAs on the same SM overlap everything for free: while one is in
// an epilogue (or its prologue / tail) the other one's MMAs own the tensor pipe.  What made that
// impossible was the footprint of a CTA: 229 KB of shared memory (activations + a 96 KB weight ring),
// all 512 TMEM columns, 64 K registers.  Feeding the weights through tensor memory (trunk_ts.cu:
// L2 -> registers -> tcgen05.st -> A operand of tcgen05.mma) removes the ring, and the rest is diet:
//   shared memory 108 KB : two activation buffers; feature staging and the logits scratch alias
//                          buffer A while it is dead and are re-zeroed afterwards
//   tensor memory 256 col: one accumulator (192) + a 4-stage ring of 2 K steps (64)
//   registers     30 K   : 384 threads; setmaxnreg: 4 producer warps 72, 4 epilogue warps 128,
//                          MMA warp 40
// Each CTA still owns 2 positions (N = 192: the weight bytes per FLOP that L2 can sustain).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "trunk_common.cuh"

namespace nsb {

namespace {

struct DuoGeom {
    static constexpr int C = 128;
    static constexpr int KCH = C / 8;
    static constexpr int NPOS = 2;
    static constexpr int NCOLS = NPOS * 96;
    static constexpr int GUARD = 12;
    static constexpr int SPITCH = (GUARD + NCOLS + 11) | 1;
    static constexpr int BUF_BYTES = ((KCH * SPITCH * 16 + 127) / 128) * 128;
    static constexpr int KC64 = C / 64;
    static constexpr int TMEM_COLS = 256;
    static constexpr int A_COL0 = NCOLS;         // 192
    static constexpr int A_STAGES = 4;           // ring stages of 2 K steps = 16 columns each
    static constexpr int A_STAGE_COLS = 16;
    static constexpr int THREADS = 384;
    static constexpr int EPI_THREADS = 128;      // warps 4-7
    static constexpr int EPI_WARPS = EPI_THREADS / 32;
    static constexpr int NBARS = 2 * A_STAGES + 2;  // a_full[], a_empty[], act, acc
    static constexpr int SCRATCH_BYTES = ((NPOS * kPolicySize * 4 + 15) / 16) * 16;
    static constexpr int FEAT_BYTES = NPOS * kMaxInChannels * 16;
    // dynamic shared memory map.  The feature staging (prologue) and the logits scratch (tail) alias
    // the front of activation buffer A, which is dead in both phases.
    static constexpr int OFF_BUF_A = 0;
    static constexpr int OFF_BUF_B = OFF_BUF_A + BUF_BYTES;
    static constexpr int OFF_SCRATCH = OFF_BUF_A;
    static constexpr int OFF_FEAT = OFF_BUF_A;
    static constexpr int OFF_VBUF = OFF_BUF_B + BUF_BYTES;
    static constexpr int OFF_RED = OFF_VBUF + ((NPOS * 81 * 4 + 15) / 16) * 16;
    static constexpr int OFF_BARS = OFF_RED + ((EPI_WARPS * NPOS * 2 * 4 + NPOS * 2 * 4 + 15) / 16) * 16;
    static constexpr int SMEM_BYTES = OFF_BARS + NBARS * 8 + 16 + 128;
    static_assert(SCRATCH_BYTES <= BUF_BYTES && FEAT_BYTES <= BUF_BYTES, "aliases fit into buffer A");
    static_assert(A_COL0 + A_STAGES * A_STAGE_COLS <= TMEM_COLS, "TMEM columns");
    static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");
};

// K = 16 steps of (layer L, K block kc, any tap): the stem's second block holds channels 64..95 only
__device__ __forceinline__ int block_steps(const DeviceNet& net, int L, int kc) { return (L == 0 && kc == 1) ? net.stem_steps - 4 : 4; }

// Tail of a pass with NT epilogue threads (the 256-thread version with its FC1 prefetch is
// heads_tail in trunk_common.cuh): dense logits, value MLP, sigmoids, fused decode (+ cache store).
template <int NPOS, int NT>
__device__ __forceinline__ void heads_tail_small(const DeviceNet& net, const EvalArgs& a, int n_eff, int li0,
                                                 float* scratch, const float* vbuf, float* red, int et) {
    constexpr int NW = NT / 32;
    const int ew = et >> 5, lane = et & 31;
    const int H = net.hidden;
    if (a.policy != nullptr) {
        for (int idx = et; idx < NPOS * kPolicySize; idx += NT) {
            const int pos = idx / kPolicySize, b = eval_index(a, li0 + pos, n_eff);
            if (b >= 0) a.policy[(size_t)b * kPolicySize + (idx - pos * kPolicySize)] = scratch[idx];
        }
    }
    float o[NPOS][2];
#pragma unroll
    for (int pos = 0; pos < NPOS; ++pos) o[pos][0] = o[pos][1] = 0.f;
    for (int h = et; h < H; h += NT) {  // FC(81 -> H) + ReLU, folded straight into FC(H -> 2)
        float acc[NPOS];
        const float b1 = __ldg(net.fc1b + h);
#pragma unroll
        for (int pos = 0; pos < NPOS; ++pos) acc[pos] = b1;
#pragma unroll 9
        for (int t = 0; t < 81; ++t) {
            const float w = __ldg(net.fc1t + (size_t)t * H + h);
#pragma unroll
            for (int pos = 0; pos < NPOS; ++pos) acc[pos] += w * vbuf[pos * 81 + t];
        }
        const float w0 = __ldg(net.fc2 + h), w1 = __ldg(net.fc2 + H + h);
#pragma unroll
        for (int pos = 0; pos < NPOS; ++pos) {
            const float hid = fmaxf(acc[pos], 0.f);
            o[pos][0] += w0 * hid;
            o[pos][1] += w1 * hid;
        }
    }
#pragma unroll
    for (int pos = 0; pos < NPOS; ++pos)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float s = warp_sum(o[pos][k]);
            if (lane == 0) red[(ew * NPOS + pos) * 2 + k] = s;
        }
    named_bar_sync(kEpiBar, NT);
    if (et < NPOS * 2) {
        const int pos = et >> 1, k = et & 1;
        float s = __ldg(net.fc2b + k);
#pragma unroll
        for (int qq = 0; qq < NW; ++qq) s += red[(qq * NPOS + pos) * 2 + k];
        const float val = 1.0f / (1.0f + expf(-s));
        red[NW * NPOS * 2 + pos * 2 + k] = val;
        const int b = eval_index(a, li0 + pos, n_eff);
        if (b >= 0) (k == 0 ? a.win : a.draw)[b] = val;
    }
    if (a.move_off != nullptr) {
        named_bar_sync(kEpiBar, NT);
        decode_tail<NPOS, NT>(a, n_eff, li0, scratch, red + NW * NPOS * 2, et);
    }
    named_bar_sync(kEpiBar, NT);
}

__global__ void __launch_bounds__(DuoGeom::THREADS, 2) trunk_duo_kernel(const DeviceNet net, const EvalArgs a) {
    using G = DuoGeom;
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
    const uint32_t bar_act = bars + 8u * (2 * G::A_STAGES);
    const uint32_t bar_acc = bars + 8u * (2 * G::A_STAGES + 1);
    volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(smem + G::OFF_BARS + 8 * G::NBARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_eff = eval_count(a);
    const int groups = (n_eff + G::NPOS - 1) / G::NPOS;
    const int my_passes =
        (int)blockIdx.x < groups ? (groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int NL = net.num_layers;
    const bool stamp = eval_timeline(a) && blockIdx.x == 0;  // diagnostics: tools/timeline.py
    // diagnostics: where and when every CTA ran (co-residency of the two CTAs of an SM: tools/residency.py)
    unsigned long long* cta_rec = (eval_timeline(a) && blockIdx.x < 1024) ? eval_timeline(a) + 4 * NL + 16 + 3 * blockIdx.x : nullptr;
    if (cta_rec && threadIdx.x == 0) {
        unsigned smid;
        unsigned long long t;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        cta_rec[0] = smid;
        cta_rec[1] = t;
    }

    // ---- one-time setup ---------------------------------------------------------------------
    for (int i = threadIdx.x; i < 2 * G::BUF_BYTES / 16; i += G::THREADS)
        reinterpret_cast<uint4*>(smem + G::OFF_BUF_A)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::A_STAGES; ++s) {
            mbar_init(bar_afull(s), 4);   // one arrival per producer warp
            mbar_init(bar_aempty(s), 1);  // tcgen05.commit of the MMAs that read the stage
        }
        mbar_init(bar_act, G::EPI_WARPS);
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 8) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), G::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 4) {
        // ===== weight producers: L2 -> registers -> tensor memory ================================
        // Thread = accumulator row = TMEM lane; the stream holds 128 rows x 32 B per K = 16 step
        // (weights.cc, pack_weights_ts).  A ring stage is two steps; three register stages rotate.
        setmaxnreg_dec<72>();
        const int row = threadIdx.x;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + G::A_COL0;
        const int steps_per_pass = net.stages_per_pass;  // K = 16 steps (always even per (K block, tap))
        uint32_t slot = 0, phase = 0;
        auto load_stage = [&](uint4 (&dst)[4], const uint4* s) {
            dst[0] = __ldg(s);
            dst[1] = __ldg(s + 1);
            dst[2] = __ldg(s + 256);
            dst[3] = __ldg(s + 257);
        };
        auto store_stage = [&](const uint4 (&buf)[4]) {
            mbar_wait(bar_aempty(slot), phase ^ 1u);
            tc_fence_after();
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                ::"r"(lane_addr + slot * G::A_STAGE_COLS), "r"(buf[0].x), "r"(buf[0].y), "r"(buf[0].z), "r"(buf[0].w),
                  "r"(buf[1].x), "r"(buf[1].y), "r"(buf[1].z), "r"(buf[1].w), "r"(buf[2].x), "r"(buf[2].y), "r"(buf[2].z),
                  "r"(buf[2].w), "r"(buf[3].x), "r"(buf[3].y), "r"(buf[3].z), "r"(buf[3].w)
                : "memory");
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull(slot));
            if (++slot == G::A_STAGES) { slot = 0; phase ^= 1u; }
        };
        const int stages_per_pass = steps_per_pass / 2;
        for (int p = 0; p < my_passes; ++p) {
            const uint4* src = reinterpret_cast<const uint4*>(net.tiles) + (size_t)row * 2;
            uint4 b0[4], b1[4], b2[4];
            int ld = 0, st = 0;  // stage cursors: stage s starts at src + s * 512 uint4
            if (ld < stages_per_pass) { load_stage(b0, src + (size_t)ld * 512); ++ld; }
            if (ld < stages_per_pass) { load_stage(b1, src + (size_t)ld * 512); ++ld; }
            if (ld < stages_per_pass) { load_stage(b2, src + (size_t)ld * 512); ++ld; }
            for (;;) {
                store_stage(b0); ++st;
                if (ld < stages_per_pass) { load_stage(b0, src + (size_t)ld * 512); ++ld; }
                if (st >= stages_per_pass) break;
                store_stage(b1); ++st;
                if (ld < stages_per_pass) { load_stage(b1, src + (size_t)ld * 512); ++ld; }
                if (st >= stages_per_pass) break;
                store_stage(b2); ++st;
                if (ld < stages_per_pass) { load_stage(b2, src + (size_t)ld * 512); ++ld; }
                if (st >= stages_per_pass) break;
            }
        }
    } else if (warp >= 8) {
        setmaxnreg_dec<40>();
        if (warp == 8) {
            // ===== MMA issuer: warp-uniform loop, one elected lane issues ============================
            constexpr uint32_t idesc = make_idesc_bf16_f32(128, G::NCOLS);
            constexpr uint32_t b_lbo = G::SPITCH * 16;
            uint32_t slot = 0, phase = 0, act_phase = 0;
            for (int p = 0; p < my_passes; ++p) {
                for (int L = 0; L < NL; ++L) {
                    mbar_wait(bar_act, act_phase);
                    act_phase ^= 1u;
                    tc_fence_after();
                    if (stamp && p == 0 && lane == 0) eval_timeline(a)[4 * L + 0] = clock64();
                    const bool head = (L == NL - 1);
                    const uint32_t in_buf = (L & 1) ? bufA : bufB;
                    const int ntaps = head ? 1 : 9;
                    for (int kc = 0; kc < G::KC64; ++kc) {
                        const int halves = block_steps(net, L, kc) / 2;  // ring stages per (K block, tap)
                        for (int tap = 0; tap < ntaps; ++tap) {
                            const int shift = head ? 0 : (tap / 3 - 1) * 10 + (tap % 3 - 1);
                            const uint32_t b_base = in_buf + (uint32_t)((kc * 8 * G::SPITCH + G::GUARD + shift) * 16);
                            if (halves == 2) {
                                // both ring stages of the tap in one round of the issue protocol (wait, fence,
                                // elect, commit, warp sync cost ~190 cycles a round): 4 MMAs per round
                                const uint32_t s0 = slot, s1 = (slot + 1) % G::A_STAGES;
                                const uint32_t ph1 = s1 == 0 ? phase ^ 1u : phase;
                                mbar_wait(bar_afull(s0), phase);
                                mbar_wait(bar_afull(s1), ph1);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t b_lo = smem_desc_lo(b_base, b_lbo);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint32_t a_base = tmem_base + G::A_COL0 + ((k < 2) ? s0 : s1) * G::A_STAGE_COLS;
                                        umma_bf16_ts(tmem_base, a_base + (k & 1) * 8, smem_desc_from(b_lo + (uint32_t)(2 * k * G::SPITCH), 128),
                                                     idesc, (uint32_t)((kc | tap | k) != 0));
                                        if (k == 1) umma_commit(bar_aempty(s0));
                                    }
                                    umma_commit(bar_aempty(s1));
                                }
                                __syncwarp();
                                slot = (slot + 2) % G::A_STAGES;
                                if (slot < 2) phase ^= 1u;
                                continue;
                            }
                            for (int h = 0; h < halves; ++h) {
                                mbar_wait(bar_afull(slot), phase);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t a_base = tmem_base + G::A_COL0 + slot * G::A_STAGE_COLS;
                                    const uint32_t b_lo = smem_desc_lo(b_base, b_lbo) + (uint32_t)(4 * h * G::SPITCH);
#pragma unroll
                                    for (int k2 = 0; k2 < 2; ++k2)
                                        umma_bf16_ts(tmem_base, a_base + k2 * 8, smem_desc_from(b_lo + (uint32_t)(2 * k2 * G::SPITCH), 128),
                                                     idesc, (uint32_t)((kc | tap | h | k2) != 0));
                                    umma_commit(bar_aempty(slot));
                                }
                                __syncwarp();
                                if (++slot == G::A_STAGES) { slot = 0; phase ^= 1u; }
                            }
                        }
                    }
                    if (elect_one()) umma_commit(bar_acc);
                    __syncwarp();
                    if (stamp && p == 0 && lane == 0) eval_timeline(a)[4 * L + 1] = clock64();
                }
            }
        }
    } else {
        // ===== expansion + epilogues + heads (4 warps: one per TMEM lane quadrant, both positions) =====
        setmaxnreg_inc<128>();
        const int et = threadIdx.x - 128;  // 0..127
        const int q = warp - 4;            // TMEM lane quadrant (== warp % 4)
        EpilogueMask<3> mask0, mask1;
        mask0.init(0, lane);
        mask1.init(96, lane);
        uint32_t acc_phase = 0;
        for (int p = 0; p < my_passes; ++p) {
            const int b0 = ((int)blockIdx.x + p * (int)gridDim.x) * G::NPOS;

            // -- stage 2 of feature extraction, straight into the stem's B operand (bufB); the bit
            //    strings are staged in the (dead) front of buffer A, which is zeroed again afterwards
            expand_features<G::NPOS, G::SPITCH, G::GUARD, G::EPI_THREADS>(net, a, n_eff, b0, featS, smem + G::OFF_BUF_B, et, nullptr);
            named_bar_sync(kEpiBar, G::EPI_THREADS);
            for (int i = et; i < G::FEAT_BYTES / 16; i += G::EPI_THREADS) featS[i] = make_uint4(0, 0, 0, 0);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_act);

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
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int part = 0; part < 2; ++part)
#pragma unroll
                    for (int lb = 0; lb < 2; ++lb) {
                        if (residual)
                            epilogue_half<3, true, 8>(lb, taddr + 96 * part, out_buf, G::SPITCH * 16, 2 * q, 96 * part, bias,
                                                      part ? mask1 : mask0, lane);
                        else
                            epilogue_half<3, false, 8>(lb, taddr + 96 * part, out_buf, G::SPITCH * 16, 2 * q, 96 * part, bias,
                                                       part ? mask1 : mask0, lane);
                    }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_act);
                if (stamp && p == 0 && et == 0) eval_timeline(a)[4 * L + 3] = clock64();
            }

            // -- heads: row 32*(h/7) + h%7 of the accumulator holds head channel h
            const int hp = 7 * q + lane;
            const float hbias = lane < 7 ? __ldg(net.bias + (size_t)(NL - 1) * C + hp) : 0.f;
            mbar_wait(bar_acc, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            {   // the head input (buffer A) is dead now: its front becomes the logits scratch
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
                head_read<0>(taddr, hbias, hp, scratch, vbuf, lane);
                head_read<96>(taddr + 96, hbias, hp, scratch, vbuf, lane);
                tc_fence_before();
            }
            named_bar_sync(kEpiBar, G::EPI_THREADS);
            heads_tail_small<G::NPOS, G::EPI_THREADS>(net, a, n_eff, b0, scratch, vbuf, red, et);
            if (p + 1 < my_passes) {  // the scratch overwrote guard / padding slots of buffer A: zero them again
                for (int i = et; i < G::SCRATCH_BYTES / 16; i += G::EPI_THREADS)
                    reinterpret_cast<uint4*>(smem + G::OFF_SCRATCH)[i] = make_uint4(0, 0, 0, 0);
                named_bar_sync(kEpiBar, G::EPI_THREADS);
            }
        }
    }

    // ---- teardown -------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (cta_rec && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        cta_rec[2] = t;
    }
    if (warp == 8) {
        __syncwarp();
        tmem_dealloc(tmem_base, G::TMEM_COLS);
    }
}

}  // namespace

int trunk_duo_prepare(int* ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(trunk_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DuoGeom::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(trunk_duo_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int nb = 0;
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trunk_duo_kernel, DuoGeom::THREADS, DuoGeom::SMEM_BYTES);
    if (e != cudaSuccess || nb < 1) {
        set_error("trunk duo: cannot configure the kernel: %s (is this an sm_100a device?)", cudaGetErrorString(e));
        return NSB_ERR_NO_DEVICE;
    }
    // The occupancy calculator answers 1 for this kernel, the hardware co-schedules 2 (tools/residency.py: 296 CTAs on
    // 148 SMs, lifetimes overlapping 99 %): the footprint is built for two - 2 x (111 KB + 1 KB) of shared memory,
    // 2 x 256 TMEM columns, 2 x 30 K registers (static_asserts in DuoGeom).  An oversized grid would only cost a
    // second wave, an undersized one halves the kernel's reason to exist.
    if (ctas_per_sm) *ctas_per_sm = nb < 2 ? 2 : nb;
    return 0;
}

int launch_trunk_duo(const DeviceNet& net, const EvalArgs& a, int num_sms, int ctas_per_sm, cudaStream_t s) {
    if (a.n <= 0) return 0;
    const int groups = (a.n + DuoGeom::NPOS - 1) / DuoGeom::NPOS;
    const int cap = num_sms * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    const int grid = groups < cap ? groups : cap;
    trunk_duo_kernel<<<grid, DuoGeom::THREADS, DuoGeom::SMEM_BYTES, s>>>(net, a);
    return 1;
}

}  // namespace nsb
