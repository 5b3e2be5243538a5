// selfplay_workers.h — the worker side of the self-play harness (host/selfplay_real.cc), in a header so that the CPU unit
// test can run the whole orchestration - search workers, the pipelined evaluation worker, the save worker, start-up and
// wind-down - against a mock pipeline (host_unit.cc --selfplay-loop): reference src/selfplay/{framequeue,worker,
// evaluationworker,saveworker}.{h,cc}.
#ifndef NSHOGI_ENGINE_B200_SELFPLAY_WORKERS_H
#define NSHOGI_ENGINE_B200_SELFPLAY_WORKERS_H

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

#include "evaluation_worker_b200.h"
#include "selfplay_feed.h"
#include "selfplay_game.h"
#include "teacher_io.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace game {

struct HarnessOptions : GameOptions {
    int CacheMiB = 0;
    std::string Out;  // main.cc -o / --out: the teacher file ("" = the records are built and counted, not written)
};

class FrameQueue {  // reference src/selfplay/framequeue.h
 public:
    void add(std::vector<Frame*>& Fs) {
        if (Fs.empty()) return;
        {
            std::lock_guard<std::mutex> L(M);
            for (Frame* F : Fs) Q.push_back(F);
        }
        CV.notify_all();
        Fs.clear();
    }
    void get(std::size_t Max, bool Wait, std::vector<Frame*>& Out) {
        std::unique_lock<std::mutex> L(M);
        if (Wait) CV.wait_for(L, std::chrono::milliseconds(2), [&] { return !Q.empty() || Closed; });
        while (!Q.empty() && Out.size() < Max) {
            Out.push_back(Q.front());
            Q.pop_front();
        }
    }
    void close() {
        {
            std::lock_guard<std::mutex> L(M);
            Closed = true;
        }
        CV.notify_all();
    }

 private:
    std::deque<Frame*> Q;
    std::mutex M;
    std::condition_variable CV;
    bool Closed = false;
};

// reference src/selfplay/saveworker.cc: finished games are replayed and their full-search positions written as teacher
// records (host/teacher_io.h), off the search threads
class SaveQueue {
 public:
    void add(teacher::FinishedGame&& G) {
        {
            std::lock_guard<std::mutex> L(M);
            Q.push_back(std::move(G));
        }
        CV.notify_one();
    }
    bool get(teacher::FinishedGame* G) {
        std::unique_lock<std::mutex> L(M);
        CV.wait_for(L, std::chrono::milliseconds(5), [&] { return !Q.empty() || Closed; });
        if (Q.empty()) return false;
        *G = std::move(Q.front());
        Q.pop_front();
        return true;
    }
    void close() {
        {
            std::lock_guard<std::mutex> L(M);
            Closed = true;
        }
        CV.notify_all();
    }
    bool drained() {
        std::lock_guard<std::mutex> L(M);
        return Q.empty();
    }

 private:
    std::deque<teacher::FinishedGame> Q;
    std::mutex M;
    std::condition_variable CV;
    bool Closed = false;
};

struct SaveStats {
    std::atomic<uint64_t> Games{0}, Records{0}, Winners[3] = {{0}, {0}, {0}};
};

void saveWorker(const HarnessOptions& O, SaveQueue* Queue, SaveStats* Stats, std::atomic<bool>* Running) {
    std::ofstream File;
    if (!O.Out.empty()) {
        File.open(O.Out, std::ios::binary | std::ios::trunc);
        teacher::writeHeader(File);
    }
    teacher::FinishedGame G;
    while (Running->load(std::memory_order_relaxed) || !Queue->drained()) {
        if (!Queue->get(&G)) continue;
        Stats->Records.fetch_add(teacher::saveGame(O.Out.empty() ? nullptr : &File, G), std::memory_order_relaxed);
        Stats->Games.fetch_add(1, std::memory_order_relaxed);
        Stats->Winners[G.Winner].fetch_add(1, std::memory_order_relaxed);  // SaveWorker::updateStatistics
    }
}

// reference src/selfplay/worker.{h,cc}: a search worker is a worker::Worker whose doTask() takes frames off the search
// queue, runs each one's phase machine until it needs the network, and hands them to the evaluation queue
class SearchWorker : public worker::Worker {
 public:
    SearchWorker(const HarnessOptions& Opt, FrameQueue* Search, FrameQueue* Evaluation, SaveQueue* Save, Info* I,
                 const std::atomic<bool>* WindDown)
        : worker::Worker(true), O(Opt), SearchQueue(Search), EvaluationQueue(Evaluation), Saves(Save), SI(I), Closing(WindDown) {
        spawnThread();
    }

 protected:
    bool doTask() override {
        // A worker::Worker is only stopped while its doTask() reports idle (worker.cc:117-134), and a pool of games
        // never runs dry by itself: the reference winds down by no longer re-queueing finished frames
        // (saveworker.cc:69-79); this time-limited harness by no longer taking frames once Closing is set.
        if (Closing->load(std::memory_order_relaxed)) return false;
        In.clear();
        SearchQueue->get(32, true, In);
        if (In.empty()) return false;
        for (Frame* F : In) {
            advance(O, *F, SI, [&](const Frame& Done) { Saves->add(finishedGame(Done)); });
            Out.push_back(F);
        }
        EvaluationQueue->add(Out);
        return true;
    }

 private:
    const HarnessOptions& O;
    FrameQueue* SearchQueue;
    FrameQueue* EvaluationQueue;
    SaveQueue* Saves;
    Info* SI;
    const std::atomic<bool>* Closing;
    std::vector<Frame*> In, Out;
};

// What the pipelined evaluation worker (host/evaluation_worker_b200.h, a worker::Worker) needs to know about a frame:
// the four steps of reference src/selfplay/evaluationworker.cc:69-117 that touch one.
template <typename SlotT>
class FrameClient : public evaluate::EvaluationClient<SlotT> {
 public:
    // Helper: with a save queue the evaluation thread runs search steps while it waits for the GPU (help()).
    // (Like the search workers it stops taking frames once the harness is closing: the pool must run dry.)
    FrameClient(const HarnessOptions& Opt, FrameQueue* Evaluation, FrameQueue* Search, Info* I, SaveQueue* Helper = nullptr,
                const std::atomic<bool>* WindDown = nullptr)
        : O(Opt), EvaluationQueue(Evaluation), SearchQueue(Search), SI(I), Saves(Helper), Closing(WindDown) {}

    bool help() override {  // four frames' worth of a search worker's doTask()
        if (Saves == nullptr || (Closing && Closing->load(std::memory_order_relaxed))) return false;
        HelpIn.clear();
        SearchQueue->get(4, false, HelpIn);
        if (HelpIn.empty()) return false;
        for (Frame* F : HelpIn) advance(O, *F, SI, [&](const Frame& Done) { Saves->add(finishedGame(Done)); });
        EvaluationQueue->add(HelpIn);
        return true;
    }

    void take(std::size_t Max, bool Wait, std::vector<void*>& Out) override {  // :70-81
        Frames.clear();
        EvaluationQueue->get(Max, Wait, Frames);
        for (Frame* F : Frames) Out.push_back(F);
    }
    uint32_t fill(void* Task, SlotT& S, std::size_t Row, uint32_t MoveBegin) override {  // :87-92
        const Frame* F = static_cast<const Frame*>(Task);
        F->Leaf.toRecord(&S.Positions[Row], F->MaxPly, F->BlackDraw, F->WhiteDraw);  // stage 1 runs on the GPU
        S.Hashes[Row] = F->Leaf.Hash;
        S.RowFlags[Row] = nshogi::engine::selfplay::rowFlags(O.Gumbel, F->LeafNode == 0);  // frame.cc:116-118
        std::memcpy(S.MoveIndices + MoveBegin, F->LeafSlots, (std::size_t)F->NumLeafMoves * sizeof(uint16_t));
        SI->LegalMoves.fetch_add((uint64_t)F->NumLeafMoves, std::memory_order_relaxed);
        return (uint32_t)F->NumLeafMoves;
    }
    void deliver(void* Task, SlotT& S, std::size_t Row) override {  // :106-108, frame.cc:93-136
        Frame* F = static_cast<Frame*>(Task);
        const uint32_t B = S.MoveOffsets[Row];
        // gather, cache store of the raw logits and softmax (or its skip at a Gumbel root) happened on the GPU
        // together with the rank order of the row; what is left - priors into the edges in rank order, the Dirichlet mix
        // of a full-search AlphaZero root, back-propagation - is done by the search worker that takes the frame next
        stageEvaluation(*F, S.Legal + B, S.Order + B, S.WinRate[Row], S.DrawRate[Row]);
        if (S.NanFlag[Row]) SI->NanRows.fetch_add(1, std::memory_order_relaxed);
        if (O.CacheMiB > 0 && S.HitFlag[Row]) SI->CacheHits.fetch_add(1, std::memory_order_relaxed);
    }
    void release(std::vector<void*>& Tasks) override {  // :114
        SI->Evals.fetch_add(Tasks.size(), std::memory_order_relaxed);
        SI->Batches.fetch_add(1, std::memory_order_relaxed);
        Frames.clear();
        for (void* T : Tasks) Frames.push_back(static_cast<Frame*>(T));
        SearchQueue->add(Frames);
        Tasks.clear();
    }

 private:
    const HarnessOptions& O;
    FrameQueue* EvaluationQueue;
    FrameQueue* SearchQueue;
    Info* SI;
    SaveQueue* Saves;
    const std::atomic<bool>* Closing;
    std::vector<Frame*> Frames, HelpIn;
};

} // namespace game
} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
