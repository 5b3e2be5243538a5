// selfplay_feed.h — the self-play side of the fused path: what Frame::setEvaluation<false> does with one
// evaluated leaf (reference src/selfplay/frame.cc:93-136), restated over the rows a LeafPipeline slot returns.
//
// The reference, per frame and serially on the evaluation thread (src/selfplay/evaluationworker.cc:106-108):
//   :96-107  gather the legal logits            :110-114 EvalCache->store(raw logits)
//   :116-118 softmax_ unless (Gumbel && root)    :121-133 Dirichlet noise at the AlphaZero root of a full search
//   :135     Node::setEvaluation
// With a NSB_DECODE_BOTH request the executor has done the first three on the GPU - for evaluated rows in the
// trunk kernel's tail, for rows served from the device-resident cache in the probe kernel - with the Gumbel
// root passed as the row flag NSB_ROW_SKIP_SOFTMAX (rowFlags()).  What is left for the host is the noise mix
// (a root is one leaf in a few hundred, its noise is sampled on the CPU at root preparation,
// src/selfplay/worker.cc:166-177) and the hand-over to the node.
#ifndef NSHOGI_ENGINE_SELFPLAY_FEED_B200_H
#define NSHOGI_ENGINE_SELFPLAY_FEED_B200_H

#include <cstddef>
#include <cstdint>

#include "nsb.h"

namespace nshogi {
namespace engine {
namespace selfplay {

// The row flags of a NSB_DECODE_BOTH request for one frame (frame.cc:116: the softmax is skipped exactly at a
// Gumbel frame's root).
inline uint8_t rowFlags(bool IsGumbel, bool IsRoot) {
    return (IsGumbel && IsRoot) ? (uint8_t)NSB_ROW_SKIP_SOFTMAX : (uint8_t)0;
}

// frame.cc:121-133: P[I] = (float)((1 - EPS) * (double)P[I] + EPS * Noise[I]), EPS = 0.25; only at the root of a
// non-Gumbel frame whose current move is a full search (getDidFullSearch().back()).  `Noise` is the frame's
// normalised Dirichlet sample (worker.cc:166-177; 600 entries, frame.cc:26).  In place on the slot's pinned row.
inline void mixDirichletNoise(float* Policy, const double* Noise, std::size_t NumChildren) {
    const double EPS = 0.25;
    for (std::size_t I = 0; I < NumChildren; ++I)
        Policy[I] = (float)((1 - EPS) * (double)Policy[I] + EPS * Noise[I]);
}

// The tail of Frame::setEvaluation for a row decoded by the executor: `Policy` holds what frame.cc has in
// LegalPolicyLogits after :116-118 (probabilities, or raw logits at a Gumbel root).  Returns the pointer the
// node receives (the same row, mixed in place when the noise applies).
template <typename NodeT>
inline void setEvaluationDecoded(NodeT* N, float* Policy, std::size_t NumChildren, float WinRate, float DrawRate,
                                 bool IsGumbel, bool IsRoot, bool DidFullSearch, const double* Noise) {
    if (!IsGumbel && IsRoot && DidFullSearch) mixDirichletNoise(Policy, Noise, NumChildren);  // :121-133
    N->setEvaluation(Policy, WinRate, DrawRate);                                            // :135
}

} // namespace selfplay
} // namespace engine
} // namespace nshogi

#endif
