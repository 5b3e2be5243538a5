// evaluation_worker_b200.h — the pipelined evaluation worker as a `worker::Worker`: the shape
// src/selfplay/evaluationworker.{h,cc} (and, with a LeafQueue in front, src/mcts/evaluationworker.{h,cc}) take on this
// executor, inside the reference's own threading contract (src/worker/worker.h: initializationTask() once on the worker
// thread, then doTask() in a loop between start() and stop(); a worker is only stopped while doTask() reports idle).
//
// Reference doTask() (src/selfplay/evaluationworker.cc:69-117): pop up to BatchSize frames, construct 1,376 B of
// features per frame on this thread, computeBlocking(), setEvaluation per frame, hand the frames back - the GPU idles
// while the thread builds features and decodes, the thread idles while the GPU runs.  Here doTask() is one turn of a
// slot ring: take what is queued, write each task's row (108-byte position record or bitboards, policy slots of its
// legal moves, hash, row flag) into the next pinned slot, submit it (stage 1, forward, decode, cache, rank order: one
// launch), and deliver the OLDEST slot's rows only when the ring is full or nothing is queued.  doTask() returns false
// only when nothing is queued AND nothing is in flight, so the Worker contract's "stopped while idle" means "stopped
// while drained": no result can arrive after await() returns.
//
// What a task IS stays with the caller (a selfplay::Frame, a Node* + State, ...): EvaluationClient is the four things the
// worker needs to do with it.  PipelineT = evaluate::LeafPipeline, or anything with its surface (the CPU tests run the
// worker on plain memory, against this repo's Worker stand-in and against the reference's own worker.cc).
#ifndef NSHOGI_ENGINE_EVALUATION_WORKER_B200_H
#define NSHOGI_ENGINE_EVALUATION_WORKER_B200_H

#include <cstddef>
#include <chrono>
#include <cstdint>
#include <deque>
#include <vector>

#include "worker/worker.h"  // the reference's header when built in-tree, host/shim otherwise

namespace nshogi {
namespace engine {
namespace evaluate {

template <typename SlotT>
class EvaluationClient {
 public:
    virtual ~EvaluationClient() = default;
    // Pop up to Max queued tasks into Out (FrameQueue::get, framequeue.cc:48-84).  Wait: block briefly if none is queued.
    virtual void take(std::size_t Max, bool Wait, std::vector<void*>& Out) = 0;
    // Write the task's inputs into row Row of slot S; its legal moves' policy slots go to S.MoveIndices[MoveBegin ..).
    // Returns the number of legal moves.  (evaluationworker.cc:87-92 constructAt + frame.cc:96-107 getMoveIndex)
    virtual uint32_t fill(void* Task, SlotT& S, std::size_t Row, uint32_t MoveBegin) = 0;
    // Row Row of the collected slot S holds the task's result (Frame::setEvaluation, evaluationworker.cc:106-108).
    virtual void deliver(void* Task, SlotT& S, std::size_t Row) = 0;
    // The delivered tasks go back to the search side (SearchQueue->add, evaluationworker.cc:114); clears Tasks.
    virtual void release(std::vector<void*>& Tasks) = 0;
    // Called while the worker would otherwise block on the GPU (the oldest slot is not done yet): do a SMALL piece of
    // useful work - a few microseconds - and return true, or return false to let the worker block.  Default: block.
    virtual bool help() { return false; }
};

template <typename PipelineT>
class PipelinedEvaluationWorker : public worker::Worker {
 public:
    using Slot = typename PipelineT::Slot;

    // The pipeline is created by the caller; Bind (optional) runs once on the worker thread before the first task
    // (the reference binds the executor's device there: selfplay/evaluationworker.cc:62-67).
    // MinFill: while batches are in flight, a batch is only submitted once it has this many rows - fewer tasks stay
    // with the worker and wait for the ones the next delivery sets free (slots are the scarce resource: a slot that
    // carries 20 rows is in flight as long as one that carries 500).  With nothing in flight whatever is there goes out.
    PipelinedEvaluationWorker(PipelineT* Pipeline, EvaluationClient<Slot>* Client_, bool FromPositions, int DecodeMode,
                              bool UseCache, bool Ranked, void (*Bind)(void*) = nullptr, void* BindArg = nullptr,
                              std::size_t MinFill_ = 1)
        : worker::Worker(true), Pipe(Pipeline), Client(Client_), Positions(FromPositions), Mode(DecodeMode), Cache(UseCache),
          Rank(Ranked), BindFn(Bind), BindArgument(BindArg), MinFill(MinFill_ < 1 ? 1 : MinFill_), SlotTasks(Pipeline->numSlots()) {
        spawnThread();
    }

    uint64_t batches() const { return Batches; }
    uint64_t rows() const { return Rows; }
    // where the worker thread's time went (seconds; read after await()): writing rows, waiting for the oldest slot,
    // handing rows back, waiting for work
    double secondsFilling() const { return TFill; }
    double secondsCollecting() const { return TCollect; }
    double secondsDelivering() const { return TDeliver; }
    double secondsTaking() const { return TTake; }
    double secondsHelping() const { return THelp; }  // part of secondsCollecting(): spent in EvaluationClient::help()

 protected:
    void initializationTask() override {
        if (BindFn) BindFn(BindArgument);
    }

    bool doTask() override {
        const auto T0 = Clock::now();
        if (Tasks.size() < Pipe->batchMax()) Client->take(Pipe->batchMax() - Tasks.size(), InFlight.empty() && Tasks.empty(), Tasks);
        TTake += seconds(T0);
        if (Tasks.empty()) {
            if (InFlight.empty()) return false;  // idle AND drained: the only state in which the worker can be stopped
            // nothing to submit, a slot free, the oldest batch still running: a search step may produce work for that slot
            if (InFlight.size() < Pipe->numSlots() && !Pipe->ready(InFlight.front())) {
                const auto TH = Clock::now();
                const bool Helped = Client->help();
                THelp += seconds(TH);
                TCollect += seconds(TH);
                if (Helped) return true;
            }
            deliverOldest();
            return true;
        }
        if (Tasks.size() < MinFill && !InFlight.empty()) {  // wait for a fuller batch: the next delivery frees more tasks
            deliverOldest();
            return true;
        }
        if (InFlight.size() == Pipe->numSlots()) deliverOldest();  // ring full: acquire() hands out the oldest slot
        std::size_t K;
        const auto T1 = Clock::now();
        Slot& S = Pipe->acquire(&K);
        uint32_t Off = 0;
        for (std::size_t I = 0; I < Tasks.size(); ++I) {
            S.MoveOffsets[I] = Off;
            Off += Client->fill(Tasks[I], S, I, Off);
        }
        S.MoveOffsets[Tasks.size()] = Off;
        SlotTasks[K].swap(Tasks);
        Tasks.clear();
        Pipe->submit(K, SlotTasks[K].size(), Positions, Mode, Cache, Rank);
        InFlight.push_back(K);
        TFill += seconds(T1);
        return true;
    }

 private:
    void deliverOldest() {
        const std::size_t K = InFlight.front();
        InFlight.pop_front();
        const auto T0 = Clock::now();
        const auto TH = Clock::now();
        while (!Pipe->ready(K) && Client->help()) {}  // a thread that waits for the GPU can run a search step meanwhile
        THelp += seconds(TH);
        Slot& S = Pipe->collect(K);
        TCollect += seconds(T0);
        const auto T1 = Clock::now();
        std::vector<void*>& Ts = SlotTasks[K];
        for (std::size_t I = 0; I < Ts.size(); ++I) Client->deliver(Ts[I], S, I);
        Rows += Ts.size();
        ++Batches;
        Client->release(Ts);
        Ts.clear();
        TDeliver += seconds(T1);
    }

    using Clock = std::chrono::steady_clock;
    static double seconds(Clock::time_point Since) { return std::chrono::duration<double>(Clock::now() - Since).count(); }

    PipelineT* Pipe;
    EvaluationClient<Slot>* Client;
    const bool Positions;
    const int Mode;
    const bool Cache, Rank;
    void (*BindFn)(void*);
    void* BindArgument;
    const std::size_t MinFill;
    std::vector<std::vector<void*>> SlotTasks;
    std::deque<std::size_t> InFlight;
    std::vector<void*> Tasks;
    uint64_t Batches = 0, Rows = 0;
    double TFill = 0, TCollect = 0, TDeliver = 0, TTake = 0, THelp = 0;
};

} // namespace evaluate
} // namespace engine
} // namespace nshogi

#endif
