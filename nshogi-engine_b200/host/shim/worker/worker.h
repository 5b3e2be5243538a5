// host/shim/worker/worker.h - build-time stand-in for the reference's src/worker/worker.{h,cc} (the generic worker-thread
// state machine every worker of the engine derives from: Uninitialized -> Idle <-> Running -> Exiting -> Exit), for
// building this repo's harnesses where the reference tree is not on the include path.  Same class name, same public and
// protected surface, same observable contract, own implementation:
//   - a child calls spawnThread() once in its constructor; it returns after initializationTask() has run on the new thread
//   - start(): Idle -> Running; the thread calls doTask() - once, or in a loop if LoopTask
//   - in a loop, doTask() == true means "call me again at once"; after every 4th false the thread looks at the stop
//     request (reference src/worker/worker.cc:117-134) - a worker is therefore only ever stopped while it is idle
//   - stop(): request the loop to end; await(): block until the thread is Idle again; the destructor ends the thread
// With the reference on the include path (-I<reference>/src before this directory) its own header and worker.cc are used
// instead; tests/test_host_cpp.py builds the pipelined evaluation worker against both.
#ifndef NSB_HOST_SHIM_WORKER_WORKER_H
#define NSB_HOST_SHIM_WORKER_WORKER_H

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>

namespace nshogi {
namespace engine {
namespace worker {

class Worker {
 public:
    explicit Worker(bool LoopTask) : Loop(LoopTask) {}
    virtual ~Worker() {
        {
            std::lock_guard<std::mutex> L(M);
            Phase = Quit;
        }
        CV.notify_all();
        if (T.joinable()) T.join();
    }

    virtual void start() {
        {
            std::lock_guard<std::mutex> L(M);
            StopRequested.store(false, std::memory_order_relaxed);
            Phase = Busy;
        }
        CV.notify_all();
    }
    virtual void stop() { StopRequested.store(true, std::memory_order_release); }
    virtual void await() {
        std::unique_lock<std::mutex> L(M);
        CV.wait(L, [this] { return Phase == Idle || Phase == Gone; });
    }
    bool isRunning() {
        std::lock_guard<std::mutex> L(M);
        return Phase == Busy;
    }

 protected:
    void spawnThread() {
        T = std::thread([this] { run(); });
        std::unique_lock<std::mutex> L(M);
        CV.wait(L, [this] { return Phase != Fresh; });
    }
    virtual void initializationTask() {}
    virtual bool doTask() = 0;

 private:
    enum State { Fresh, Idle, Busy, Quit, Gone };

    void run() {
        {
            std::lock_guard<std::mutex> L(M);
            initializationTask();
            Phase = Idle;
        }
        CV.notify_all();
        for (;;) {
            {
                std::unique_lock<std::mutex> L(M);
                CV.wait(L, [this] { return Phase == Busy || Phase == Quit; });
                if (Phase == Quit) break;
            }
            for (uint64_t Idles = 0;;) {
                const bool Again = doTask();
                if (!Loop) break;
                if (Again) continue;
                if (++Idles == 4) {
                    if (StopRequested.load(std::memory_order_acquire)) break;
                    Idles = 0;
                }
            }
            {
                std::lock_guard<std::mutex> L(M);
                if (Phase == Busy) Phase = Idle;
            }
            CV.notify_all();
        }
        {
            std::lock_guard<std::mutex> L(M);
            Phase = Gone;
        }
        CV.notify_all();
    }

    const bool Loop;
    State Phase = Fresh;
    std::atomic<bool> StopRequested{false};
    std::thread T;
    std::mutex M;
    std::condition_variable CV;
};

} // namespace worker
} // namespace engine
} // namespace nshogi

#endif
