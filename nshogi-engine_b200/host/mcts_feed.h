// mcts_feed.h — the MCTS side of the fused path: what FeedWorker::feedResult does with one evaluated leaf
// (reference src/mcts/feedworker.cc:56-137), restated over the rows a LeafPipeline slot returns.
//
// The reference, per leaf and on a feed thread: NaN fallback for win / draw rate from the parent's statistics
// (:58-85), gather of the legal logits + softmax (:100-127), Node::setEvaluation, Node::sort() (std::sort of the
// edges by decreasing probability, src/mcts/node.h:163-168), updateAncestors, EvalCache::store (:134-135).  With
// the executor's fused decode the gather, the softmax, the NaN handling of the policy row (:105-118, only with
// NSB_DECODE_NAN_FALLBACK, as in the reference's two builds of feedResult), the cache store and the sort's
// permutation are already done on the GPU; what is left is O(n): write the edges in rank order, apply the
// win / draw fallback (:58-85), back-propagate.
//
// Templated on the node type so that it compiles against the reference's mcts::Node (getNumChildren, getEdge,
// setEvaluation, getParent, getWinRateAccumulated, getDrawRateAccumulated, getVisitsAndVirtualLoss, VisitMask,
// updateAncestors; Edge: getMove / setMove-by-assignment, setProbability) and against the mock in host_unit.cc.
#ifndef NSHOGI_ENGINE_MCTS_FEED_B200_H
#define NSHOGI_ENGINE_MCTS_FEED_B200_H

#include <cstdint>
#include <cstring>

#include "leaf_pipeline.h"

namespace nshogi {
namespace engine {
namespace mcts {

inline bool isNaNBits(float X) {  // reference src/math/math.h:23-39: survives -ffast-math
    uint32_t U;
    std::memcpy(&U, &X, 4);
    return (U & 0x7F800000u) == 0x7F800000u && (U & 0x007FFFFFu) != 0u;
}

// One decoded leaf of a collected slot.
struct LeafRow {
    const float* Legal;     // probabilities in move-generation order
    const uint16_t* Order;  // rank order (LeafPipeline::submit(..., Ranked = true))
    uint16_t NumMoves;
    float WinRate, DrawRate;
};

inline LeafRow leafRow(const evaluate::LeafPipeline::Slot& S, std::size_t I) {
    const uint32_t B = S.MoveOffsets[I], E = S.MoveOffsets[I + 1];
    return LeafRow{S.Legal + B, S.Order + B, (uint16_t)(E - B), S.WinRate[I], S.DrawRate[I]};
}

// feedResult<NaNFallbackEnabled> for a ranked row; the batch must have been submitted with (true) or without
// (false, the reference's default: src/context.h:103) NSB_DECODE_NAN_FALLBACK accordingly.  The policy half of
// the fallback (a NaN logit makes the row uniform, :105-118) has happened on the GPU; the value half (:58-85)
// is here, because it needs the parent node.  Returns NaNFound (the reference then skips the cache store; the
// device cache has skipped it already: the flag of the GPU row covers logits, win and draw rate).
template <bool NaNFallbackEnabled = false, typename NodeT>
bool feedRanked(NodeT* N, const LeafRow& R) {
    float WinRate = R.WinRate, DrawRate = R.DrawRate;
    bool NaNFound = false;
    if (NaNFallbackEnabled && isNaNBits(WinRate)) {  // feedworker.cc:61-72
        NaNFound = true;
        const NodeT* Parent = N->getParent();
        WinRate = Parent == nullptr
                      ? 0.5f
                      : (float)(1.0 - Parent->getWinRateAccumulated() / (double)(Parent->getVisitsAndVirtualLoss() & NodeT::VisitMask));
    }
    if (NaNFallbackEnabled && isNaNBits(DrawRate)) {  // feedworker.cc:73-84
        NaNFound = true;
        const NodeT* Parent = N->getParent();
        DrawRate = Parent == nullptr
                       ? 0.0f
                       : (float)(Parent->getDrawRateAccumulated() / (double)(Parent->getVisitsAndVirtualLoss() & NodeT::VisitMask));
    }
    const uint16_t NumChildren = N->getNumChildren();
    auto* Edges = N->getEdge();
    if (NumChildren == 1) {  // feedworker.cc:101-103 (no sort)
        Edges[0].setProbability(1.0f);
    } else if (NumChildren == R.NumMoves) {
        // setEvaluation() + sort() in one pass: edge r receives the move and probability of row element Order[r].
        // The moves are permuted through a stack copy (a node has at most 593 children).
        decltype(Edges[0].getMove()) Moves[NSB_MAX_LEGAL_MOVES];
        for (uint16_t J = 0; J < NumChildren; ++J) Moves[J] = Edges[J].getMove();
        for (uint16_t Rk = 0; Rk < NumChildren; ++Rk) {
            const uint16_t J = R.Order[Rk];
            Edges[Rk].setMove(Moves[J]);
            Edges[Rk].setProbability(R.Legal[J]);
        }
    }
    N->setEvaluation(nullptr, WinRate, DrawRate);  // node.h:150-160: a null policy leaves the edges alone
    N->updateAncestors(WinRate, DrawRate);
    return NaNFound;
}

} // namespace mcts
} // namespace engine
} // namespace nshogi

#endif
