/*
 * nsb_oracle.h — CPU oracle for the leaf-evaluation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libnsb.so) never links, loads or calls it.
 *
 * Parity status of each function:
 *   expand     : PINNED   — restates reference src/cuda/extractbit.cu:15-39,41-68, whose source
 *                           is the spec; also cross-checked on the GPU against the reference's
 *                           own kernel compiled from /root/reference (oracle/_ref).
 *   random_fill: PINNED   — port of reference src/infer/random.cc:28-42; checked against the
 *                           reference's random.cc compiled here (oracle/_ref).
 *   pack       : UNPINNED — FeatureStackComptime lives in libnshogi (un-vendored, unpinned,
 *                           absent); channel order follows src/evaluate/preset.h:20-66, plane
 *                           semantics follow SURVEY.md App. A.2 and are builder-defined.
 *   decode     : order of operations and NaN semantics pinned by src/mcts/feedworker.cc:56-136
 *                (three pieces: :58-85 value fallback, :100-103 one-move shortcut, :105-127 gather
 *                + logit fallback + softmax) and src/selfplay/frame.cc:93-136; ml::math::softmax_
 *                itself is libnshogi (UNPINNED, assumed max-subtracted exp / sum, T = 1).
 *   cache      : PINNED   — restates src/mcts/evalcache.cc:17-169; checked operation by operation
 *                           against the reference's evalcache.cc compiled here (oracle/_ref).
 *   forward    : UNPINNED — the reference's forward is TensorRT on an external ONNX (neither
 *                           present); the oracle is the fp32 definition of OUR canonical net.
 */
#ifndef NSB_ORACLE_H
#define NSB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/nsb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/cuda/extractbit.cu:15-39 (channels_first) / :41-68 (channels_last). */
void nsb_oracle_expand(const nsb_feature_bitboard* fb, size_t n, int channels, int channels_first,
                       float* planes);

/* position -> 86 feature bitboards, channel order of reference src/evaluate/preset.h:20-66. */
void nsb_oracle_pack(const nsb_position* pos, size_t n, nsb_feature_bitboard* fb);

/* reference src/mcts/feedworker.cc:56-136 (mode PROBS; | NSB_DECODE_NAN_FALLBACK = feedResult<true>, the
 * default is the reference's: off, src/context.h:103) / src/selfplay/frame.cc:93-118 (LOGITS = the gather,
 * BOTH = the whole of setEvaluation<false> before the Dirichlet mix; row_flags: NSB_ROW_SKIP_SOFTMAX). */
void nsb_oracle_decode_ex(const float* policy, const float* win, const float* draw, size_t n,
                          const uint32_t* move_off, const uint16_t* move_idx, int mode, const uint8_t* row_flags,
                          float* legal_out, float* logits_out, uint8_t* nan_flag);
/* feedworker.cc:58-85: NaN win / draw rate replaced from the parent's statistics; returns NaNFound. */
int nsb_oracle_value_fallback(float* win, float* draw, int has_parent, double parent_win_acc,
                              double parent_draw_acc, uint64_t parent_visits);
/* frame.cc:121-133: p = (float)(0.75 * (double)p + 0.25 * noise) at the AlphaZero root of a full search. */
void nsb_oracle_dirichlet_mix(float* probs, const double* noise, uint32_t m);
void nsb_oracle_decode(const float* policy, const float* win, const float* draw, size_t n,
                       const uint32_t* move_off, const uint16_t* move_idx, int mode,
                       float* legal_out, uint8_t* nan_flag);

/* fp32 forward of the canonical net on fp32 NCHW planes [n][in_channels][81].
 * emulate_bf16 != 0 rounds the input planes and every trunk activation to bf16 (RNE) exactly
 * where the device kernel does, for a tight comparison. */
void nsb_oracle_forward(const nsb_net_desc* net, const float* blob, const float* planes, size_t n,
                        int emulate_bf16, float* policy, float* win, float* draw);

/* reference src/infer/random.cc:28-42: one mt19937_64, 2187+2 uniform_real<float>(0,1) draws per
 * sample.  state is opaque (create with seed, reference default seed 0). */
typedef struct nsb_oracle_rng nsb_oracle_rng;
nsb_oracle_rng* nsb_oracle_rng_create(uint64_t seed);
void nsb_oracle_rng_destroy(nsb_oracle_rng* r);
void nsb_oracle_random_fill(nsb_oracle_rng* r, size_t n, float* policy, float* win, float* draw);

/* The reference's CPU leaf-evaluation path for config 1 ("EXECUTOR=random, 8 threads, batch 128"):
 * per thread and per batch: pack (stage 1) [+ expand when with_expand] + Random executor fill +
 * decode(PROBS).  `fill` may be NULL (oracle port) or a pointer to the reference's own compiled
 * Random executor wrapper (oracle/_ref/libnsb_ref_random.so: nsb_ref_random_fill) with one handle
 * per thread made by `mk`.  Runs until every thread has done `batches_per_thread` batches.
 * Returns evaluated samples per second (wall clock, all threads). */
typedef void (*nsb_fill_fn)(void* handle, size_t n, float* policy, float* win, float* draw);
typedef void* (*nsb_fill_make_fn)(uint64_t seed);
double nsb_oracle_cpu_path(const nsb_position* pos, size_t n_pos, const uint32_t* move_off,
                           const uint16_t* move_idx, int batch, int threads,
                           int batches_per_thread, int with_expand, nsb_fill_fn fill,
                           nsb_fill_make_fn mk, double* seconds_out);

/* reference src/mcts/evalcache.{h,cc}: bundles of 3 entries on a recency list, bundle = hash %
 * num_bundles; store (:49-121) and load (:123-169; the caller's move-count check of
 * src/mcts/searchworker.cc:545-556 is applied by `expected_n`).  Single-threaded, so the try-lock
 * never fails.  load returns 1 on a usable hit (row / win / draw filled), 0 otherwise. */
typedef struct nsb_oracle_cache nsb_oracle_cache;
nsb_oracle_cache* nsb_oracle_cache_create(uint64_t num_bundles);
void nsb_oracle_cache_destroy(nsb_oracle_cache* c);
int nsb_oracle_cache_store(nsb_oracle_cache* c, uint64_t hash, uint32_t n, const float* row, float win, float draw);
int nsb_oracle_cache_load(nsb_oracle_cache* c, uint64_t hash, uint32_t expected_n, float* row, float* win, float* draw);

float nsb_oracle_bf16_round(float x);

#ifdef __cplusplus
}
#endif
#endif
