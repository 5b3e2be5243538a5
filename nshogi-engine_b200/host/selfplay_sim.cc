// selfplay_sim.cc — the reference's self-play data-generation loop (reference src/selfplay/:
// main.cc frame pool + search workers + evaluation worker + save worker, phases of phase.h) driven
// against infer::B200 through the pinned multi-slot LeafPipeline, to measure "self-play
// positions/sec" (BASELINE.json metric, configs[3]) on the B200 path.
//
// What is real: the executor and everything below it (H2D of packed positions, stage-1 pack kernel,
// fused trunk, fused legal-move decode, D2H of legal-move rows), the pinned slot ring, the
// frame-pool / two-queue / worker-thread structure and the playout budget per move
// (--num-playouts 200, --full-search-ratio 0.25: main.cc defaults; a reduced search uses a quarter
// of the playouts, worker.cc:184-197).
// What is SYNTHETIC, because libnshogi (rules, move generation, repetition, declaration) is not
// available in this build: tree descent is replaced by a few random piece relocations of the root
// position, legal moves by a random set of distinct policy slots (count ~ N(80, 35) clipped to
// [1, 593], a function of the position), game length by U[80, 200] plies.  --descent-ns adds a busy
// wait per leaf to model the CPU cost of a real descent; --revisit-ratio makes that fraction of the
// descents end in a recently visited leaf (transpositions) and --cache-mb puts the device-resident
// evaluation cache (Frame::setEvaluationCache, frame.cc:89-114) in front of the network.  Records are counted, not written (saveworker.cc writes one teacher
// record per played position).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <deque>
#include <limits>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "infer_b200.h"
#include "leaf_pipeline.h"
#include "selfplay_feed.h"

using namespace nshogi::engine;
using Clock = std::chrono::steady_clock;

namespace {

enum class Phase : uint8_t { LeafSelection, Evaluation, Backpropagation };  // the phases of phase.h this loop needs

struct Frame {  // reference src/selfplay/frame.h: one game in flight
    nsb_position Root, Leaf;
    uint16_t MoveIdx[NSB_MAX_LEGAL_MOVES];
    uint32_t NumMoves = 0;
    uint64_t Rng = 0;
    uint32_t PlayoutsLeft = 0, Ply = 0, GameLen = 0;
    float Win = 0.f, Draw = 0.f, PolicyMass = 0.f;
    Phase P = Phase::LeafSelection;
    uint64_t Hash = 0;         // state hash of the leaf (key of the evaluation cache)
    bool AtRoot = true;        // the next leaf is the root itself (first evaluation after a move, worker.cc:159-160)
    bool FullSearch = true;    // getDidFullSearch().back() (worker.cc:184-197)
    double Noise[600];         // Dirichlet (or Gumbel) noise sampled at root preparation (frame.cc:26, worker.cc:166-177)
    nsb_position Recent[4];    // leaves visited lately: --revisit-ratio re-descends to one of them
    uint32_t RecentCount = 0;
    bool CacheHit = false;
};

inline uint64_t next(uint64_t& S) {  // xorshift64*
    S ^= S >> 12; S ^= S << 25; S ^= S >> 27;
    return S * 0x2545F4914F6CDD1Dull;
}

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
    // up to Max frames; blocks (until something arrives or the queue is closed) only if Wait
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

struct Info {  // reference src/selfplay/selfplayinfo.h
    std::atomic<uint64_t> Evals{0}, Batches{0}, Records{0}, Games{0}, NanRows{0}, CacheHits{0};
};

inline uint64_t hashPosition(const nsb_position& P) {  // stand-in for core::State::getHash(): FNV-1a of the record
    const unsigned char* B = reinterpret_cast<const unsigned char*>(&P);
    uint64_t H = 0xCBF29CE484222325ull;
    for (std::size_t I = 0; I < sizeof(nsb_position); ++I) H = (H ^ B[I]) * 0x100000001B3ull;
    return H;
}

nsb_position startpos() {  // hirate; squares s = 9 * (file - 1) + (rank - 1)
    nsb_position P;
    std::memset(&P, 0, sizeof P);
    auto put = [&](int File, int Rank, int Pt, int Colour) { P.board[9 * (File - 1) + (Rank - 1)] = (uint8_t)(1 + Pt + 14 * Colour); };
    const int Back[9] = {1, 2, 3, 4, 5, 4, 3, 2, 1};
    for (int F = 1; F <= 9; ++F) {
        put(F, 9, Back[F - 1], 0);
        put(F, 1, Back[9 - F], 1);
        put(F, 7, 0, 0);
        put(F, 3, 0, 1);
    }
    put(8, 8, 6, 0); put(2, 8, 7, 0); put(2, 2, 6, 1); put(8, 2, 7, 1);
    P.max_ply = 320;
    P.black_draw_value = P.white_draw_value = 0.5f;
    return P;
}

// Stand-in for one move: a random piece of the side to move relocates to a random empty square.
void relocate(nsb_position& P, uint64_t& Rng) {
    const int Side = P.side & 1;
    for (int Try = 0; Try < 16; ++Try) {
        const int From = (int)(next(Rng) % 81), To = (int)(next(Rng) % 81);
        const int Code = P.board[From];
        if (Code == 0 || (Code - 1) / 14 != Side || P.board[To] != 0) continue;
        P.board[To] = (uint8_t)Code;
        P.board[From] = 0;
        break;
    }
    P.side = (uint8_t)(Side ^ 1);
    P.ply = (uint16_t)(P.ply + 1);
}

void newGame(Frame& F, uint32_t Playouts) {
    F.Root = startpos();
    F.Ply = 0;
    F.GameLen = 80 + (uint32_t)(next(F.Rng) % 121);
    F.PlayoutsLeft = Playouts;
}

struct Options {
    int Channels = 256, Blocks = 20, Batch = 512, Frames = 1024, SearchWorkers = 4, Slots = 3, GPU = 0;
    int Playouts = 200, DescentNs = 0, CacheMiB = 0;
    bool Gumbel = false;  // main.cc --gumbel
    double FullSearchRatio = 0.25, Seconds = 5.0, Warmup = 1.0, RevisitRatio = 0.0;
};

uint32_t playoutsFor(const Options& O, uint64_t& Rng, bool* FullSearch) {  // worker.cc:184-197
    const double U = (double)(next(Rng) >> 11) * (1.0 / 9007199254740992.0);
    *FullSearch = U < O.FullSearchRatio;
    return *FullSearch ? (uint32_t)O.Playouts : (uint32_t)std::max(1, O.Playouts / 4);
}

// Worker::prepareRoot, worker.cc:166-177 + sampleNoise :640-655: 600 noise values per root - Gumbel
// -log(-log U), or Gamma(0.15, 1) normalised to a Dirichlet sample for the AlphaZero style.
void prepareRoot(const Options& O, Frame& F, std::mt19937_64& MT) {
    F.AtRoot = true;
    if (O.Gumbel) {
        std::uniform_real_distribution<double> D(std::numeric_limits<double>::min(), 1.0);
        for (double& X : F.Noise) X = -std::log(-std::log(D(MT)));
    } else {
        std::gamma_distribution<double> D(0.15, 1.0);
        double Sum = 0.0;
        for (double& X : F.Noise) Sum += (X = D(MT));
        for (double& X : F.Noise) X /= Sum;
    }
}

// reference src/selfplay/worker.cc: LeafSelection / Backpropagation / Transition on the CPU
void searchWorker(const Options& O, FrameQueue* SearchQueue, FrameQueue* EvaluationQueue, Info* SI,
                  std::atomic<bool>* Running) {
    std::vector<Frame*> In, Out;
    std::mt19937_64 MT(0x5EED5EEDull + (uint64_t)(uintptr_t)&In);  // worker.cc: one generator per search worker
    while (Running->load(std::memory_order_relaxed)) {
        In.clear();
        SearchQueue->get(64, true, In);
        for (Frame* F : In) {
            if (F->P == Phase::Backpropagation) {
                if (--F->PlayoutsLeft == 0) {  // Transition: play a move, one teacher record
                    relocate(F->Root, F->Rng);
                    SI->Records.fetch_add(1, std::memory_order_relaxed);
                    if (++F->Ply >= F->GameLen) {
                        SI->Games.fetch_add(1, std::memory_order_relaxed);
                        newGame(*F, playoutsFor(O, F->Rng, &F->FullSearch));
                    } else {
                        F->PlayoutsLeft = playoutsFor(O, F->Rng, &F->FullSearch);
                    }
                    prepareRoot(O, *F, MT);
                }
            }
            // LeafSelection: descend (synthetic) - or, with --revisit-ratio, reach a leaf seen lately
            // (a transposition) - and list the leaf's legal moves as policy slots.  The move list is a
            // function of the position, as it is with a real move generator.
            const double U = (double)(next(F->Rng) >> 11) * (1.0 / 9007199254740992.0);
            if (F->AtRoot) {  // RootPreparation: the first leaf after a move is the root position itself
                F->Leaf = F->Root;
            } else if (F->RecentCount > 0 && U < O.RevisitRatio) {
                F->Leaf = F->Recent[next(F->Rng) % F->RecentCount];
            } else {
                F->Leaf = F->Root;
                const int Depth = 1 + (int)(next(F->Rng) % 6);
                for (int D = 0; D < Depth; ++D) relocate(F->Leaf, F->Rng);
                F->Recent[F->RecentCount < 4 ? F->RecentCount++ : next(F->Rng) % 4] = F->Leaf;
            }
            F->Hash = hashPosition(F->Leaf);
            uint64_t MoveRng = F->Hash | 1ull;
            // n ~ N(80, 35) by the sum of 4 uniforms, clipped to [1, 593]
            double Z = 0.0;
            for (int K = 0; K < 4; ++K) Z += (double)(next(MoveRng) >> 11) * (1.0 / 9007199254740992.0);
            int N = (int)(80.0 + 35.0 * (Z - 2.0) * 1.7320508);
            N = N < 1 ? 1 : (N > NSB_MAX_LEGAL_MOVES ? NSB_MAX_LEGAL_MOVES : N);
            const uint32_t Start = (uint32_t)(next(MoveRng) % NSB_POLICY_SIZE);
            uint32_t Stride = 1 + (uint32_t)(next(MoveRng) % (NSB_POLICY_SIZE - 1));
            if (Stride % 3 == 0) ++Stride;  // 2187 = 3^7: any stride not divisible by 3 visits distinct slots
            for (int J = 0; J < N; ++J) F->MoveIdx[J] = (uint16_t)((Start + (uint64_t)J * Stride) % NSB_POLICY_SIZE);
            F->NumMoves = (uint32_t)N;
            if (O.DescentNs > 0) {
                const auto T0 = Clock::now();
                while (std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - T0).count() < O.DescentNs) {
                }
            }
            F->P = Phase::Evaluation;
            Out.push_back(F);
        }
        EvaluationQueue->add(Out);
    }
}

// reference src/selfplay/evaluationworker.cc:69-117, restructured: instead of constructing features
// on this thread and blocking on one batch, it copies 108-byte positions + move slots into the next
// pinned slot, submits, and only waits for the OLDEST batch when the ring is full or starved.
void evaluationWorker(const Options& O, infer::B200* Exec, FrameQueue* EvaluationQueue, FrameQueue* SearchQueue,
                      Info* SI, std::atomic<bool>* Running) {
    Exec->resetGPU();
    Exec->bindThreadToGpuNode();  // evaluator.cc:39-83
    evaluate::LeafPipeline Pipe(Exec, (std::size_t)O.Batch);
    const std::size_t NS = Pipe.numSlots();
    std::vector<std::vector<Frame*>> SlotTasks(NS);
    std::deque<std::size_t> InFlight;
    std::vector<Frame*> Tasks;
    auto deliver = [&](std::size_t Idx) {
        evaluate::LeafPipeline::Slot& S = Pipe.collect(Idx);
        std::vector<Frame*>& Fs = SlotTasks[Idx];
        for (std::size_t I = 0; I < Fs.size(); ++I) {  // frame.cc:93-136 consumer side
            Frame* F = Fs[I];
            F->Win = S.WinRate[I];
            F->Draw = S.DrawRate[I];
            float Mass = 0.f;
            float* Row = S.Legal + S.MoveOffsets[I];
            // the gather, the cache store of the raw logits and the softmax (or its skip at a Gumbel root) happened
            // on the GPU (NSB_DECODE_BOTH); the Dirichlet mix of an AlphaZero root of a full search is left
            if (!O.Gumbel && F->AtRoot && F->FullSearch) selfplay::mixDirichletNoise(Row, F->Noise, F->NumMoves);  // frame.cc:121-133
            F->AtRoot = false;
            for (uint32_t J = 0; J < F->NumMoves; ++J) Mass += Row[J];
            F->PolicyMass = Mass;
            if (S.NanFlag[I]) SI->NanRows.fetch_add(1, std::memory_order_relaxed);
            if (O.CacheMiB > 0 && S.HitFlag[I]) SI->CacheHits.fetch_add(1, std::memory_order_relaxed);
            F->P = Phase::Backpropagation;
        }
        SI->Evals.fetch_add(Fs.size(), std::memory_order_relaxed);
        SI->Batches.fetch_add(1, std::memory_order_relaxed);
        SearchQueue->add(Fs);
    };
    while (Running->load(std::memory_order_relaxed)) {
        Tasks.clear();
        EvaluationQueue->get((std::size_t)O.Batch, InFlight.empty(), Tasks);
        if (Tasks.empty()) {
            if (!InFlight.empty()) {
                deliver(InFlight.front());
                InFlight.pop_front();
            }
            continue;
        }
        if (InFlight.size() == NS) {  // ring full: the slot acquire() hands out is the oldest one
            deliver(InFlight.front());
            InFlight.pop_front();
        }
        std::size_t Idx;
        evaluate::LeafPipeline::Slot& S = Pipe.acquire(&Idx);
        uint32_t Off = 0;
        for (std::size_t I = 0; I < Tasks.size(); ++I) {
            const Frame* F = Tasks[I];
            S.Positions[I] = F->Leaf;
            S.Hashes[I] = F->Hash;
            S.RowFlags[I] = selfplay::rowFlags(O.Gumbel, F->AtRoot);  // frame.cc:116-118
            S.MoveOffsets[I] = Off;
            std::memcpy(S.MoveIndices + Off, F->MoveIdx, F->NumMoves * sizeof(uint16_t));
            Off += F->NumMoves;
        }
        S.MoveOffsets[Tasks.size()] = Off;
        SlotTasks[Idx].swap(Tasks);
        // Frame::setEvaluation<false> in one launch: the cache keeps raw logits (frame.cc:110-114), the frames
        // receive probabilities (frame.cc:116-118) - for rows served from the cache too
        Pipe.submit(Idx, SlotTasks[Idx].size(), /*FromPositions=*/true, NSB_DECODE_BOTH, /*UseCache=*/O.CacheMiB > 0);
        InFlight.push_back(Idx);
    }
    while (!InFlight.empty()) {  // Worker::stop contract: drain before returning
        deliver(InFlight.front());
        InFlight.pop_front();
    }
}

}  // namespace

int main(int argc, char** argv) {
    Options O;
    for (int I = 1; I < argc; ++I) {
        const std::string A = argv[I];
        auto nextI = [&]() { return I + 1 < argc ? std::atoi(argv[++I]) : 0; };
        auto nextD = [&]() { return I + 1 < argc ? std::atof(argv[++I]) : 0.0; };
        if (A == "--channels") O.Channels = nextI();
        else if (A == "--blocks") O.Blocks = nextI();
        else if (A == "--batch-size") O.Batch = nextI();
        else if (A == "--frame-pool-size") O.Frames = nextI();
        else if (A == "--num-search-workers") O.SearchWorkers = nextI();
        else if (A == "--slots") O.Slots = nextI();
        else if (A == "--gpu") O.GPU = nextI();
        else if (A == "--num-playouts") O.Playouts = nextI();
        else if (A == "--full-search-ratio") O.FullSearchRatio = nextD();
        else if (A == "--descent-ns") O.DescentNs = nextI();
        else if (A == "--cache-mb") O.CacheMiB = nextI();
        else if (A == "--gumbel") O.Gumbel = true;
        else if (A == "--revisit-ratio") O.RevisitRatio = nextD();
        else if (A == "--seconds") O.Seconds = nextD();
        else if (A == "--warmup") O.Warmup = nextD();
        else {
            std::fprintf(stderr, "unknown option %s\n", A.c_str());
            return 2;
        }
    }
    if (nsb_device_count() <= O.GPU) {
        std::fprintf(stderr, "nsb_selfplay_sim: no CUDA device %d; infer::B200 has no CPU fallback\n", O.GPU);
        return 2;
    }
    infer::B200 Exec(O.GPU, (uint16_t)O.Batch, NSB_FEATURE_CHANNELS, O.Channels, O.Blocks, O.Slots);
    Exec.load("");
    if (O.CacheMiB > 0) Exec.enableCache((std::size_t)O.CacheMiB);  // Frame::setEvaluationCache, frame.cc:89

    std::vector<Frame> Pool((std::size_t)O.Frames);
    FrameQueue SearchQueue, EvaluationQueue;
    Info SI;
    std::vector<Frame*> Init;
    for (std::size_t I = 0; I < Pool.size(); ++I) {
        Pool[I].Rng = 0x9E3779B97F4A7C15ull * (I + 1) + (uint64_t)O.GPU * 0xD1B54A32D192ED03ull;
        newGame(Pool[I], (uint32_t)O.Playouts);
        {
            std::mt19937_64 MT(Pool[I].Rng);
            prepareRoot(O, Pool[I], MT);
        }
        Pool[I].Ply = (uint32_t)(next(Pool[I].Rng) % Pool[I].GameLen);  // games start staggered
        Init.push_back(&Pool[I]);
    }
    SearchQueue.add(Init);

    std::atomic<bool> Running{true};
    std::vector<std::thread> Threads;
    Threads.emplace_back(evaluationWorker, std::cref(O), &Exec, &EvaluationQueue, &SearchQueue, &SI, &Running);
    for (int W = 0; W < O.SearchWorkers; ++W)
        Threads.emplace_back(searchWorker, std::cref(O), &SearchQueue, &EvaluationQueue, &SI, &Running);

    std::this_thread::sleep_for(std::chrono::duration<double>(O.Warmup));
    const uint64_t E0 = SI.Evals.load(), B0 = SI.Batches.load(), R0 = SI.Records.load(), G0 = SI.Games.load();
    const uint64_t H0 = SI.CacheHits.load();
    const auto T0 = Clock::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(O.Seconds));
    const uint64_t E1 = SI.Evals.load(), B1 = SI.Batches.load(), R1 = SI.Records.load(), G1 = SI.Games.load();
    const uint64_t H1 = SI.CacheHits.load();
    const double Sec = std::chrono::duration<double>(Clock::now() - T0).count();
    Running.store(false);
    SearchQueue.close();
    EvaluationQueue.close();
    for (auto& T : Threads) T.join();

    const double Evals = (double)(E1 - E0), Batches = (double)(B1 - B0);
    std::printf("{\"metric\": \"selfplay_positions_per_sec\", \"value\": %.1f, \"unit\": \"positions/s\", "
                "\"leaf_evals_per_sec\": %.1f, \"games_per_sec\": %.3f, \"avg_batch\": %.1f, \"seconds\": %.3f, "
                "\"records\": %llu, \"evals\": %llu, \"batches\": %llu, \"games\": %llu, "
                "\"cache_mb\": %d, \"revisit_ratio\": %.2f, \"cache_hit_rate\": %.4f, "
                "\"net\": \"%dx%d\", \"batch_size\": %d, \"frame_pool\": %d, \"search_workers\": %d, \"slots\": %d, "
                "\"num_playouts\": %d, \"full_search_ratio\": %.2f, \"descent_ns\": %d, \"nan_rows\": %llu, "
                "\"decode\": \"NSB_DECODE_BOTH (logits cached, probabilities out; %s)\", "
                "\"rules\": \"synthetic (libnshogi absent): random relocations, random legal-move slots\"}\n",
                (double)(R1 - R0) / Sec, Evals / Sec, (double)(G1 - G0) / Sec, Batches > 0 ? Evals / Batches : 0.0, Sec,
                (unsigned long long)(R1 - R0), (unsigned long long)(E1 - E0), (unsigned long long)(B1 - B0),
                (unsigned long long)(G1 - G0), O.CacheMiB, O.RevisitRatio, Evals > 0 ? (double)(H1 - H0) / Evals : 0.0,
                O.Blocks, O.Channels, O.Batch, O.Frames, O.SearchWorkers, O.Slots, O.Playouts, O.FullSearchRatio,
                O.DescentNs, (unsigned long long)SI.NanRows.load(),
                O.Gumbel ? "Gumbel roots skip the softmax" : "Dirichlet mix at full-search roots on the host");
    return 0;
}
