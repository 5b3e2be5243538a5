// leaf_pipeline.h — pinned, multi-slot batch pipeline for leaf evaluation with the fused decode.
//
// Replaces the host side of reference src/evaluate/evaluator.{h,cc} (pinned batch buffers),
// src/mcts/evaluationworker.cc:124-199 (getBatch memcpy, doInference, second copy of the logits
// into a heap Batch) and src/selfplay/evaluationworker.cc:69-117 (construct features, blocking
// compute, per-frame decode) with a ring of `Slots` page-locked batch slots: while slot k runs on
// the GPU the caller fills slot k+1, and per position only the legal-move rows come back
// (~4 B per move instead of 8,748 B of dense logits).  One pipeline per evaluator thread, like the
// reference's one Infer per EvaluationWorker; all calls from that thread.
#ifndef NSHOGI_ENGINE_EVALUATE_LEAF_PIPELINE_H
#define NSHOGI_ENGINE_EVALUATE_LEAF_PIPELINE_H

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "infer_b200.h"

namespace nshogi {
namespace engine {
namespace evaluate {

class LeafPipeline {
 public:
    struct Slot {
        // inputs, filled by the caller between acquire() and submit()
        nsb_feature_bitboard* Features = nullptr;  // [BatchMax * 86]  (feature-bitboard mode)
        nsb_position* Positions = nullptr;         // [BatchMax]       (packed-position mode)
        uint32_t* MoveOffsets = nullptr;           // [BatchMax + 1]   CSR of legal-move policy indices
        uint16_t* MoveIndices = nullptr;           // [BatchMax * 593] values of ml::getMoveIndex
        uint64_t* Hashes = nullptr;                // [BatchMax]       state hashes (only read when the executor has a cache)
        uint8_t* RowFlags = nullptr;               // [BatchMax]       NSB_ROW_* bits (read by NSB_DECODE_BOTH requests)
        // outputs, valid after collect()
        float* Legal = nullptr;                    // per legal move: probabilities or raw logits
        uint16_t* Order = nullptr;                 // per position the rank order of its row (submit(..., Ranked)):
                                                   // Order[off + r] = row index of the r-th most probable move, the
                                                   // permutation of Node::sort() (src/mcts/node.h:163-168)
        float* WinRate = nullptr;
        float* DrawRate = nullptr;
        uint8_t* NanFlag = nullptr;
        uint8_t* HitFlag = nullptr;                // 1 = served from the device-resident cache
        std::size_t Count = 0;
        bool InFlight = false;
    };

    LeafPipeline(infer::B200* Executor, std::size_t BatchMax)
        : Ex(Executor), BatchSizeMax(BatchMax), Slots(Executor->slots()) {
        for (auto& S : Slots) {
            alloc(S.Features, BatchMax * NSB_FEATURE_CHANNELS);
            alloc(S.Positions, BatchMax);
            alloc(S.MoveOffsets, BatchMax + 1);
            alloc(S.MoveIndices, BatchMax * NSB_MAX_LEGAL_MOVES);
            alloc(S.Legal, BatchMax * NSB_MAX_LEGAL_MOVES);
            alloc(S.Order, BatchMax * NSB_MAX_LEGAL_MOVES);
            alloc(S.WinRate, BatchMax);
            alloc(S.DrawRate, BatchMax);
            alloc(S.NanFlag, BatchMax);
            alloc(S.Hashes, BatchMax);
            alloc(S.HitFlag, BatchMax);
            alloc(S.RowFlags, BatchMax);
            std::memset(S.RowFlags, 0, BatchMax);
        }
    }
    ~LeafPipeline() {
        for (std::size_t I = 0; I < Slots.size(); ++I)
            if (Slots[I].InFlight) nsb_await(Ex->context(), (int)I);
        for (void* P : Pinned) nsb_host_free(P);
    }
    LeafPipeline(const LeafPipeline&) = delete;
    LeafPipeline& operator=(const LeafPipeline&) = delete;

    std::size_t numSlots() const {
        return Slots.size();
    }
    std::size_t batchMax() const {
        return BatchSizeMax;
    }

    // Next slot to fill (round robin).  If it is still in flight its results must be collected
    // first: collect(index) is called for the caller and the slot comes back ready for reuse.
    Slot& acquire(std::size_t* Index) {
        const std::size_t I = Next;
        Next = (Next + 1) % Slots.size();
        if (Slots[I].InFlight) collect(I);
        if (Index) *Index = I;
        return Slots[I];
    }

    // Slot `Index` itself, for callers that keep their own ring order (a slot must have been collected before it is
    // filled again; acquire() does both for the round-robin case).
    Slot& slotAt(std::size_t Index) {
        return Slots[Index];
    }

    // Enqueue stage 1 (if FromPositions) + expansion + forward + fused decode, with the copies around them or, in
    // a one-slot executor, directly on these page-locked arrays (NSB_IO_DIRECT).  With UseCache (executor built
    // with enableCache) the batch goes through the device-resident cache: hits are served from HBM, only the
    // misses are evaluated, evaluated rows are stored.  With Ranked, Order[] receives every row's rank order.
    // DecodeMode: NSB_DECODE_PROBS (MCTS, feedworker.cc:100-136), NSB_DECODE_BOTH (self-play, frame.cc:93-118: the
    // cache keeps raw logits, Legal receives probabilities, RowFlags marks Gumbel roots), optionally
    // | NSB_DECODE_NAN_FALLBACK (Context::isNaNFallbackEnabled()).
    void submit(std::size_t Index, std::size_t Count, bool FromPositions, int DecodeMode = NSB_DECODE_PROBS,
                bool UseCache = false, bool Ranked = false) {
        Slot& S = Slots[Index];
        S.Count = Count;
        if (Count == 0) return;
        nsb_decode_request R{};
        if (FromPositions) R.positions = S.Positions;
        else R.features = S.Features;
        R.n = Count;
        R.hashes = UseCache ? S.Hashes : nullptr;
        R.move_off = S.MoveOffsets;
        R.move_idx = S.MoveIndices;
        R.mode = DecodeMode;
        R.legal_out = S.Legal;
        R.order_out = Ranked ? S.Order : nullptr;
        R.win = S.WinRate;
        R.draw = S.DrawRate;
        R.nan_flag = S.NanFlag;
        R.hit_flag = UseCache ? S.HitFlag : nullptr;
        R.row_flags = (DecodeMode & NSB_DECODE_MODE_MASK) == NSB_DECODE_BOTH ? S.RowFlags : nullptr;
        infer::B200::check(nsb_eval_request_async(Ex->context(), (int)Index, &R), "LeafPipeline::submit");
        S.InFlight = true;
    }

    // Block until the slot's results are in its pinned output arrays.
    Slot& collect(std::size_t Index) {
        Slot& S = Slots[Index];
        if (S.InFlight) {
            infer::B200::check(nsb_await(Ex->context(), (int)Index), "LeafPipeline::collect");
            S.InFlight = false;
        }
        return S;
    }

    bool ready(std::size_t Index) {
        return !Slots[Index].InFlight || nsb_is_computing(Ex->context(), (int)Index) == 0;
    }

    // Drain everything (Worker::stop contract: no results may arrive after await()).
    void drain() {
        for (std::size_t I = 0; I < Slots.size(); ++I) collect(I);
    }

 private:
    template <typename T>
    void alloc(T*& P, std::size_t N) {
        void* Raw = nullptr;
        // page-locked (evaluator.cc:95-106) on the GPU's own NUMA node (evaluator.cc:127-136 numa_alloc_onnode)
        infer::B200::check(nsb_host_alloc_near(&Raw, N * sizeof(T), Ex->gpu()), "nsb_host_alloc_near");
        Pinned.push_back(Raw);
        P = static_cast<T*>(Raw);
    }

    infer::B200* Ex;
    const std::size_t BatchSizeMax;
    std::vector<Slot> Slots;
    std::vector<void*> Pinned;
    std::size_t Next = 0;
};

} // namespace evaluate
} // namespace engine
} // namespace nshogi

#endif
