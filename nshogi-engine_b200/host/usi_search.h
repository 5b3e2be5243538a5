// usi_search.h — the USI-style search of host/usi_go_bench.cc as a class over any pipeline (LeafPipeline on the GPU, a
// mock on plain memory in the CPU tests: host_unit.cc --usi-loop, also under ThreadSanitizer): one tree rooted at a
// position, shared without a lock by the search threads and the evaluation thread.
//
// Reference structure: SearchWorker::doTask (src/mcts/searchworker.cc:448-609: collectOneLeaf -> terminal checks ->
// EvaluationQueue::add), EvaluationWorker (src/mcts/evaluationworker.cc:105-199), FeedWorker (src/mcts/feedworker.cc:29-137).
// Here a search thread descends under virtual loss, generates the leaf's moves and writes its row IN PLACE into the open
// pinned batch (host/leaf_queue.h); the evaluation thread seals and submits batches, feeds results back as soon as a
// batch is done and - when there is nothing to send and nothing to feed - collects leaves itself.
#ifndef NSHOGI_ENGINE_B200_USI_SEARCH_H
#define NSHOGI_ENGINE_B200_USI_SEARCH_H

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "leaf_queue.h"
#include "mcts_search.h"
#include "rules/shogi.h"

namespace nshogi {
namespace engine {
namespace b200 {

template <typename PipelineT>
class UsiSearch {
 public:
    using Queue = evaluate::BasicLeafQueue<PipelineT>;
    using Slot = typename PipelineT::Slot;
    using Clock = std::chrono::steady_clock;

    UsiSearch(search::Tree* Tree_, PipelineT* Pipeline, const rules::Position& Root_, uint16_t MaxPly_, int DecodeMode_, bool UseCache_)
        : T(Tree_), Pipe(Pipeline), Q(Pipeline), Root(Root_), History{Root_.Hash}, MaxPly(MaxPly_), DecodeMode(DecodeMode_), UseCache(UseCache_) {}

    // counters (read after run())
    std::atomic<uint64_t> Terminals{0}, Collisions{0}, LegalMoves{0};
    uint64_t Evals = 0, Batches = 0, CacheHits = 0, GpuWaitNs = 0, HelpedLeaves = 0;
    std::atomic<bool> TreeFull{false};

    // Runs the search for Seconds with SearchThreads search threads; the calling thread is the evaluation thread.
    // Help: it also collects leaves between its duties.  Returns the elapsed seconds (drain included).
    double run(double Seconds, int SearchThreads, bool Help) {
        auto feed = [this](Slot& S, std::size_t Row, void* User) { feedRow(S, Row, User); };
        auto EvalScratch = std::make_unique<Scratch>();
        Q.open(feed);
        std::vector<std::thread> Threads;
        const auto T0 = Clock::now();
        auto elapsed = [&]() { return std::chrono::duration<double>(Clock::now() - T0).count(); };
        Running.store(true);
        for (int I = 0; I < SearchThreads; ++I)
            Threads.emplace_back([this]() {
                auto C = std::make_unique<Scratch>();
                while (Running.load(std::memory_order_relaxed) && searchStep(*C, false)) {}
            });
        // EvaluationWorker::doTask (evaluationworker.cc:105-199).  The reference submits "whatever is queued"; with the
        // batch assembled in place the rule is: a full batch goes out at once; a partial one goes out when the GPU would
        // otherwise idle (nothing in flight), or at half size when only one batch is in flight - so the open batch keeps
        // filling while the GPU is busy and its size follows the parallelism the tree offers.
        const std::size_t Batch = Pipe->batchMax(), NS = Pipe->numSlots();
        while (elapsed() < Seconds && !TreeFull.load(std::memory_order_relaxed)) {
            const std::size_t Rows = Q.openRows(), Busy = Q.inFlight();
            if (Rows >= Batch || (Rows > 0 && Busy == 0) || (Rows >= Batch / 2 && Busy == 1 && NS > 2)) {
                submitOpen();
            } else if (Q.pollFeed(feed) == 0) {  // FeedWorker::doTask: results go back into the tree as soon as they exist
                // nothing to send, nothing to feed: collect a leaf like a search thread instead of idling
                if (!Help || !searchStep(*EvalScratch, true)) std::this_thread::yield();
                else ++HelpedLeaves;
            }
        }
        Running.store(false);
        for (auto& Th : Threads) Th.join();
        Q.drain(true, DecodeMode, UseCache, true, feed);
        return elapsed();
    }

 private:
    struct Scratch {
        rules::Move Moves[rules::kMaxMoves];
        uint16_t Slots[rules::kMaxMoves];
        std::vector<uint64_t> Path;
        std::vector<int> Trail;
    };

    // FeedWorker::feedResult (feedworker.cc:56-137) for one row of a collected slot: the gather, the softmax and
    // Node::sort's permutation came back from the executor; setEvaluation + updateAncestors are what is left
    void feedRow(Slot& S, std::size_t Row, void* User) {
        const int Node = (int)(uintptr_t)User - 1;
        const uint32_t B = S.MoveOffsets[Row];
        T->setPriors(Node, S.Legal + B, S.Order + B);
        T->backup(Node, S.WinRate[Row], S.DrawRate[Row]);
        if (UseCache && S.HitFlag[Row]) ++CacheHits;
        ++Evals;
    }

    void submitOpen() {
        const auto W0 = Clock::now();
        Q.submitOpen(/*FromPositions=*/true, DecodeMode, UseCache, /*Ranked=*/true, [this](Slot& S, std::size_t Row, void* User) { feedRow(S, Row, User); });
        GpuWaitNs += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - W0).count();
        ++Batches;
    }

    // SearchWorker::doTask (searchworker.cc:448-609): one leaf.  Returns false when the search is over for this thread.
    // Helper = the evaluation thread between its own duties: it must never wait for a row of the open batch - nobody else
    // would submit the full one - so it submits it itself.
    bool searchStep(Scratch& C, bool Helper) {
        rules::Position Pos = Root;
        C.Path.clear();
        const int Node = T->selectLeaf(Pos, 0.5f, 0.5f, &C.Path, &C.Trail);  // collectOneLeaf; leaves a virtual loss on the path
        if (Node == search::Tree::OutOfMemory) {
            TreeFull.store(true, std::memory_order_relaxed);
            return false;
        }
        if (Node < 0) {  // ran into a leaf that is being evaluated (searchworker.cc:349-357)
            Collisions.fetch_add(1, std::memory_order_relaxed);
            if (!Helper) std::this_thread::yield();
            return true;
        }
        {
            const search::Node& N = T->node(Node);
            if (N.Term == search::Mated) {
                T->backup(Node, 0.0f, 0.0f);
                return true;
            }
            if (N.Term == search::DrawnGame) {
                T->backup(Node, 0.5f, 1.0f);
                return true;
            }
            if (N.Term == search::Declared) {
                T->backup(Node, 1.0f, 0.0f);
                return true;
            }
        }
        if (Node != 0 && Pos.canDeclare()) {  // 27-point declaration: the side to move wins
            T->setTerminal(Node, search::Declared);
            T->backup(Node, 1.0f, 0.0f);
            Terminals.fetch_add(1, std::memory_order_relaxed);
            return true;
        }
        const int NumMoves = Pos.generateLegal(C.Moves);  // expandLeaf, :164-173 - outside any lock
        const bool Mated = NumMoves == 0;
        const bool Drawn = !Mated && Node != 0 && (search::isFourfold(Pos.Hash, History, C.Path) || Pos.Ply >= MaxPly);
        if (Mated || Drawn) {  // terminal checks, :475-538
            T->setTerminal(Node, Mated ? search::Mated : search::DrawnGame);
            T->backup(Node, Mated ? 0.0f : 0.5f, Mated ? 0.0f : 1.0f);
            Terminals.fetch_add(1, std::memory_order_relaxed);
            return true;
        }
        for (int J = 0; J < NumMoves; ++J) C.Slots[J] = (uint16_t)Pos.policyIndex(C.Moves[J]);  // ml::getMoveIndex
        if (!T->expand(Node, C.Moves, NumMoves)) {
            TreeFull.store(true, std::memory_order_relaxed);
            return false;
        }
        typename Queue::Ticket Tk;
        while (!Q.reserve((uint16_t)NumMoves, (void*)(uintptr_t)(Node + 1), &Tk)) {  // EvaluationQueue::add, evaluationqueue.cc:45-60
            if (Helper) {
                submitOpen();  // the open batch is full: sending it is this thread's own job
                continue;
            }
            if (!Running.load(std::memory_order_relaxed)) break;
            std::this_thread::yield();
        }
        if (Tk.S == nullptr) return false;  // shutting down with the leaf unqueued: its virtual loss dies with the tree
        Pos.toRecord(&Tk.S->Positions[Tk.Row], MaxPly, 0.5f, 0.5f);  // stage 1 runs on the GPU
        Tk.S->Hashes[Tk.Row] = Pos.Hash;
        std::memcpy(Tk.S->MoveIndices + Tk.MoveBegin, C.Slots, (std::size_t)NumMoves * sizeof(uint16_t));
        LegalMoves.fetch_add((uint64_t)NumMoves, std::memory_order_relaxed);
        Q.publish(Tk);
        return true;
    }

    search::Tree* T;
    PipelineT* Pipe;
    Queue Q;
    const rules::Position Root;
    const std::vector<uint64_t> History;
    const uint16_t MaxPly;
    const int DecodeMode;
    const bool UseCache;
    std::atomic<bool> Running{false};
};

} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
