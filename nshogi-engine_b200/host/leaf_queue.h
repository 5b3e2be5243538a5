// leaf_queue.h — lock-free, copy-free batch assembly for the MCTS side of the fused path.
//
// Replaces reference src/mcts/evaluationqueue.{h,cc} + EvaluationWorker::getBatch
// (src/mcts/evaluationworker.cc:124-154): there a search thread builds the 86 feature bitboards into a local
// stack (evaluationqueue.cc:47), pushes a tuple into a mutex-protected std::queue, the evaluation thread pops up
// to BatchSize tuples into four freshly allocated vectors (evaluationqueue.cc:62-90) and memcpy's 1,376 B per
// leaf into the pinned batch.  Here the open batch IS the pinned slot of a LeafPipeline: a search thread
// reserves a row (and the row's span of the CSR move list) with one compare-and-swap, writes the packed position
// / bitboards, the policy slots of its legal moves and the hash in place, and publishes the row; the evaluation
// thread seals the batch, submits it, and opens the next slot - whose previous batch is collected and fed first.
// No mutex, no allocation, no copy; with direct host I/O the kernel then reads those very bytes.
//
// One 64-bit word holds the open batch's state: [63] sealed, [62:48] generation, [47:32] rows, [31:0] moves.
// Threading: any number of search threads call reserve/publish; ONE evaluation thread calls submitOpen / drain
// (it owns the LeafPipeline, like the reference's one Infer per EvaluationWorker).
#ifndef NSHOGI_ENGINE_EVALUATE_LEAF_QUEUE_H
#define NSHOGI_ENGINE_EVALUATE_LEAF_QUEUE_H

#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

#include "leaf_pipeline.h"

namespace nshogi {
namespace engine {
namespace evaluate {

// PipelineT: LeafPipeline, or anything with its Slot / numSlots / batchMax / acquire / submit / collect surface
// (host_unit.cc runs the protocol on plain memory under ThreadSanitizer).
template <typename PipelineT>
class BasicLeafQueue {
 public:
    using Slot = typename PipelineT::Slot;

    struct Ticket {
        Slot* S = nullptr;
        uint32_t Row = 0;        // fill S->Positions[Row] (or S->Features + 86 * Row) and S->Hashes[Row]
        uint32_t MoveBegin = 0;  // fill S->MoveIndices[MoveBegin .. MoveBegin + NumMoves)
    };

    explicit BasicLeafQueue(PipelineT* P) : Pipe(P), Users(P->numSlots()), Counts(P->numSlots(), 0) {
        for (auto& U : Users) U.resize(P->batchMax(), nullptr);
    }

    // Evaluation thread, once before the search threads start (and implicitly after every submitOpen).
    template <typename Feed>
    void open(Feed&& FeedRow) {
        std::size_t K;
        Slot& S = Pipe->acquire(&K);  // collects the slot's previous batch if it is still in flight
        feedSlot(K, FeedRow);
        OpenIndex = K;
        Published.store(0, std::memory_order_relaxed);
        OpenSlot.store(&S, std::memory_order_release);
        const uint64_t Gen = (Generation = (Generation + 1) & 0x7FFF);
        Cursor.store(Gen << 48, std::memory_order_release);
    }

    // Search threads.  False: the open batch is full or being sealed - retry after yielding.
    bool reserve(uint16_t NumMoves, void* User, Ticket* T) {
        uint64_t C = Cursor.load(std::memory_order_acquire);
        for (;;) {
            const uint32_t Rows = (uint32_t)(C >> 32) & 0xFFFF, Moves = (uint32_t)C;
            if ((C >> 63) || Rows >= Pipe->batchMax() || (uint64_t)Moves + NumMoves > Pipe->batchMax() * (uint64_t)NSB_MAX_LEGAL_MOVES)
                return false;
            Slot* S = OpenSlot.load(std::memory_order_acquire);
            if (Cursor.compare_exchange_weak(C, C + (1ull << 32) + NumMoves, std::memory_order_acq_rel, std::memory_order_acquire)) {
                T->S = S;
                T->Row = Rows;
                T->MoveBegin = Moves;
                S->MoveOffsets[Rows] = Moves;
                Users[OpenIndex][Rows] = User;  // OpenIndex is stable while this generation's cursor is live
                return true;
            }
        }
    }
    void setUser(const Ticket& T, void* User) {  // (the handle may also be set after reserve, before publish)
        Users[OpenIndex][T.Row] = User;
    }
    void publish(const Ticket&) {
        Published.fetch_add(1, std::memory_order_release);
    }

    // Evaluation thread: submitted batches whose rows have not been fed yet.
    std::size_t inFlight() const {
        std::size_t N = 0;
        for (std::size_t K = 0; K < Counts.size(); ++K) N += (K != OpenIndex && Counts[K] > 0) ? 1 : 0;
        return N;
    }

    std::size_t openRows() const {
        return (std::size_t)((Cursor.load(std::memory_order_relaxed) >> 32) & 0xFFFF);
    }

    // Evaluation thread: seal the open batch, wait for the rows still being filled, submit it, open the next slot
    // (FeedRow(Slot&, Row, User) runs for every row of that slot's previous batch).  Returns the rows submitted.
    template <typename Feed>
    std::size_t submitOpen(bool FromPositions, int DecodeMode, bool UseCache, bool Ranked, Feed&& FeedRow) {
        const uint64_t C = Cursor.fetch_or(1ull << 63, std::memory_order_acq_rel);
        const uint32_t Rows = (uint32_t)(C >> 32) & 0xFFFF, Moves = (uint32_t)C;
        while (Published.load(std::memory_order_acquire) != Rows) std::this_thread::yield();
        Slot* S = OpenSlot.load(std::memory_order_relaxed);
        S->MoveOffsets[Rows] = Moves;
        Counts[OpenIndex] = Rows;
        Pipe->submit(OpenIndex, Rows, FromPositions, DecodeMode, UseCache, Ranked);
        open(FeedRow);
        return Rows;
    }

    // Evaluation thread: feed every submitted batch whose results have arrived, without waiting for the ring to come
    // round to its slot.  A tree search needs this - its next leaves depend on the results (the reference's FeedWorkers
    // run as soon as a batch is done, feedworker.cc:29-39); independent leaves (self-play frames) can do without.
    // Returns the number of rows fed.
    template <typename Feed>
    std::size_t pollFeed(Feed&& FeedRow) {
        std::size_t Fed = 0;
        for (std::size_t K = 0; K < Pipe->numSlots(); ++K) {
            if (K == OpenIndex || Counts[K] == 0 || !Pipe->ready(K)) continue;
            Fed += Counts[K];
            feedSlot(K, FeedRow);
        }
        return Fed;
    }

    // Evaluation thread, after the search threads have stopped: submit what is open, collect and feed everything.
    template <typename Feed>
    void drain(bool FromPositions, int DecodeMode, bool UseCache, bool Ranked, Feed&& FeedRow) {
        submitOpen(FromPositions, DecodeMode, UseCache, Ranked, FeedRow);
        for (std::size_t K = 0; K < Pipe->numSlots(); ++K) {
            Pipe->collect(K);
            feedSlot(K, FeedRow);
        }
    }

 private:
    template <typename Feed>
    void feedSlot(std::size_t K, Feed&& FeedRow) {
        Slot& S = Pipe->collect(K);
        for (std::size_t I = 0; I < Counts[K]; ++I) FeedRow(S, I, Users[K][I]);
        Counts[K] = 0;
    }

    PipelineT* Pipe;
    std::vector<std::vector<void*>> Users;  // per slot, per row: the caller's handle (the reference queues Node*)
    std::vector<std::size_t> Counts;        // rows of the batch each slot holds (0 once fed)
    std::size_t OpenIndex = 0;
    uint64_t Generation = 0;
    std::atomic<Slot*> OpenSlot{nullptr};
    std::atomic<uint64_t> Cursor{1ull << 63};  // sealed until open()
    std::atomic<uint32_t> Published{0};
};

using LeafQueue = BasicLeafQueue<LeafPipeline>;

} // namespace evaluate
} // namespace engine
} // namespace nshogi

#endif
