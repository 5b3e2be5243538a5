// selfplay_real.cc — self-play data generation with REAL shogi rules and a REAL MCTS, driving infer::B200 through the
// pinned multi-slot LeafPipeline: BASELINE.json configs[3] ("self-play data generation, 20x256 ResNet, 1024 concurrent
// games per GPU") and its metric "self-play positions/sec" (SURVEY.md §8 f1).
//
// Structure = the reference's src/selfplay/: a pool of frames (games, main.cc:100-110), search workers that run the
// phase machine of worker.cc:55-110 on one frame at a time (initialize -> prepareRoot -> selectLeaf -> checkTerminal ->
// [evaluation] -> backpropagate -> transition -> judge), ONE evaluation worker per GPU (evaluationworker.cc:69-117), two
// frame queues between them (framequeue.cc).  What differs from the reference, on purpose:
//   - rules, move generation, repetition: host/rules/shogi.h (libnshogi is not available; perft-pinned);
//   - tree: host/mcts_search.h (PUCT with the reference's constants; no tree reuse between moves); the 3-ply mate search
//     the reference asks at every leaf is optional here (--leaf-mate-plies 3: rules/shogi.h mateIn3), its df-pn solver absent;
//   - the evaluation worker (host/evaluation_worker_b200.h, a worker::Worker like the reference's) does not build
//     features and does not block per batch: it copies 108-byte position records and the legal moves' policy slots into
//     the next pinned slot, submits (stage 1, forward, gather, cache store of the raw logits, softmax and the edges' rank
//     order all happen in ONE launch: NSB_DECODE_BOTH + order_out), waits only for the oldest batch and leaves each
//     decoded row with its frame; the search worker that takes the frame next writes the priors and back-propagates
//     (host/selfplay_workers.h, host/selfplay_game.h);
//   - finished games are written by a save worker as NSBT teacher records (host/teacher_io.h; --out FILE), not in
//     libnshogi's io::file::simple_teacher layout (saveworker.cc:160-182), which is defined outside the reference tree.
// Games start from hirate; per game MaxPly ~ U[224, 640] and the draw values of worker.cc:135-150; per move a full
// search (--num-playouts) with probability --full-search-ratio, else a quarter of it (worker.cc:184-197); AlphaZero
// style: Dirichlet(0.15) noise mixed into the root priors of full searches (frame.cc:121-133), the most visited move is
// played (worker.cc:555-590).  A game ends by mate, 27-point declaration, four-fold repetition (a draw, or lost by the
// side whose every move of the cycle gave check) or at MaxPly (draw).
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "infer_b200.h"
#include "evaluation_worker_b200.h"
#include "leaf_pipeline.h"
#include "selfplay_workers.h"

using namespace nshogi::engine;
using namespace nshogi::engine::b200;
using Clock = std::chrono::steady_clock;

namespace {

struct Options : b200::game::HarnessOptions {
    int Channels = 256, Blocks = 20, Batch = 256, Frames = 1024, SearchWorkers = 4, Slots = 4, GPU = 0, MinFill = 1;
    double Seconds = 5.0, Warmup = 1.0;
    uint64_t Seed = 1234;
};
using namespace b200::game;

}  // namespace

int main(int argc, char** argv) {
    Options O;
    bool Verbose = false, NoHelp = false;
    auto stage = [&](const char* What) {
        if (Verbose) std::fprintf(stderr, "nsb_selfplay_real: %s\n", What);
    };
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
        else if (A == "--min-fill") O.MinFill = nextI();
        else if (A == "--gpu") O.GPU = nextI();
        else if (A == "--num-playouts") O.Playouts = nextI();
        else if (A == "--full-search-ratio") O.FullSearchRatio = nextD();
        else if (A == "--cache-mb") O.CacheMiB = nextI();
        else if (A == "--gumbel") O.Gumbel = true;
        else if (A == "--leaf-mate-plies") O.MatePlies = nextI();
        else if (A == "--seconds") O.Seconds = nextD();
        else if (A == "--warmup") O.Warmup = nextD();
        else if (A == "--seed") O.Seed = (uint64_t)nextI();
        else if (A == "--out" || A == "-o") O.Out = I + 1 < argc ? argv[++I] : "";
        else if (A == "--verbose") Verbose = true;
        else if (A == "--no-help") NoHelp = true;  // the evaluation thread blocks on the GPU instead of running search steps meanwhile
        else {
            std::fprintf(stderr, "unknown option %s\n", A.c_str());
            return 2;
        }
    }
    if (nsb_device_count() <= O.GPU) {
        std::fprintf(stderr, "nsb_selfplay_real: no CUDA device %d; infer::B200 has no CPU fallback\n", O.GPU);
        return 2;
    }
    infer::B200 Exec(O.GPU, (uint16_t)O.Batch, NSB_FEATURE_CHANNELS, O.Channels, O.Blocks, O.Slots);
    Exec.load("", O.Seed);
    if (O.CacheMiB > 0) Exec.enableCache((std::size_t)O.CacheMiB);  // Frame::setEvaluationCache, frame.cc:89

    std::vector<Frame> Pool((std::size_t)O.Frames);
    FrameQueue SearchQueue, EvaluationQueue;
    Info SI;
    std::vector<Frame*> Init;
    for (std::size_t I = 0; I < Pool.size(); ++I) {
        Pool[I].MT.seed(0x9E3779B97F4A7C15ull * (I + 1) + (uint64_t)O.GPU * 0xD1B54A32D192ED03ull);
        b200::game::newGame(O, Pool[I]);
        b200::game::prepareRoot(O, Pool[I]);
        Init.push_back(&Pool[I]);
    }
    SearchQueue.add(Init);

    // reference src/selfplay/main.cc:181-211: workers are constructed (their threads initialise and wait), then started
    SaveQueue Saves;
    SaveStats Saved;
    std::atomic<bool> Saving{true};
    std::thread Saver(saveWorker, std::cref(O), &Saves, &Saved, &Saving);
    evaluate::LeafPipeline Pipe(&Exec, (std::size_t)O.Batch);
    std::atomic<bool> Closing{false};
    FrameClient<evaluate::LeafPipeline::Slot> Client(O, &EvaluationQueue, &SearchQueue, &SI, NoHelp ? nullptr : &Saves, &Closing);
    evaluate::PipelinedEvaluationWorker<evaluate::LeafPipeline> Evaluation(
        &Pipe, &Client, /*FromPositions=*/true, NSB_DECODE_BOTH, /*UseCache=*/O.CacheMiB > 0, /*Ranked=*/true,
        [](void* E) {  // on the worker thread, as selfplay/evaluationworker.cc:62-67 binds its executor
            static_cast<infer::B200*>(E)->resetGPU();
            static_cast<infer::B200*>(E)->bindThreadToGpuNode();  // evaluator.cc:39-83
        },
        &Exec, (std::size_t)O.MinFill);
    std::vector<std::unique_ptr<SearchWorker>> Searchers;
    for (int W = 0; W < O.SearchWorkers; ++W)
        Searchers.push_back(std::make_unique<SearchWorker>(O, &SearchQueue, &EvaluationQueue, &Saves, &SI, &Closing));
    stage("workers constructed");
    Evaluation.start();
    for (auto& W : Searchers) W->start();
    stage("workers started");

    std::this_thread::sleep_for(std::chrono::duration<double>(O.Warmup));
    const uint64_t E0 = SI.Evals.load(), B0 = SI.Batches.load(), R0 = SI.Records.load(), G0 = SI.Games.load();
    const uint64_t H0 = SI.CacheHits.load(), T0n = SI.Terminals.load(), L0 = SI.LegalMoves.load();
    const auto T0 = Clock::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(O.Seconds));
    const uint64_t E1 = SI.Evals.load(), B1 = SI.Batches.load(), R1 = SI.Records.load(), G1 = SI.Games.load();
    const uint64_t H1 = SI.CacheHits.load(), T1n = SI.Terminals.load(), L1 = SI.LegalMoves.load();
    const double Sec = std::chrono::duration<double>(Clock::now() - T0).count();
    // stop the search side first, then the evaluation worker: it leaves its loop only when nothing is queued and
    // nothing is in flight (worker::Worker stops a worker while doTask() reports idle)
    stage("measured; winding down");
    Closing.store(true);
    for (auto& W : Searchers) W->stop();
    for (auto& W : Searchers) W->await();
    stage("search workers idle");
    Evaluation.stop();
    Evaluation.await();
    stage("evaluation worker drained");
    Searchers.clear();
    Saving.store(false);
    Saves.close();
    Saver.join();

    const double EvalTotal = std::max(1e-9, Evaluation.secondsFilling() + Evaluation.secondsDelivering() + Evaluation.secondsCollecting() +
                                                Evaluation.secondsTaking());
    uint64_t MaxDepthPly = 0;
    for (const Frame& F : Pool) MaxDepthPly = std::max<uint64_t>(MaxDepthPly, F.Root.Ply);
    const double Evals = (double)(E1 - E0), Batches = (double)(B1 - B0);
    std::printf("{\"metric\": \"selfplay_positions_per_sec\", \"value\": %.1f, \"unit\": \"positions/s\", "
                "\"leaf_evals_per_sec\": %.1f, \"games_per_sec\": %.3f, \"avg_batch\": %.1f, \"seconds\": %.3f, "
                "\"records\": %llu, \"evals\": %llu, \"batches\": %llu, \"games\": %llu, "
                "\"cache_mb\": %d, \"cache_hit_rate\": %.4f, \"terminal_leaves_per_eval\": %.4f, \"avg_legal_moves\": %.1f, "
                "\"games_ended\": {\"mate\": %llu, \"repetition\": %llu, \"of_which_perpetual_check\": %llu, \"declaration\": %llu, \"max_ply\": %llu, \"mate_found_by_search\": %llu}, \"leaf_mate_plies\": %d, \"leaf_mates_by_search\": %llu, \"deepest_game_ply\": %llu, "
                "\"teacher\": {\"games_saved\": %llu, \"records_saved\": %llu, \"black_wins\": %llu, \"white_wins\": %llu, \"draws\": %llu, "
                "\"file\": \"%s\", \"what\": \"full-search positions of finished games (saveworker.cc:160-182), NSBT format\"}, "
                "\"net\": \"%dx%d\", \"batch_size\": %d, \"frame_pool\": %d, \"search_workers\": %d, \"slots\": %d, \"min_fill\": %d, "
                "\"num_playouts\": %d, \"full_search_ratio\": %.2f, \"nan_rows\": %llu, "
                "\"evaluation_worker\": {\"us_per_row_filling\": %.3f, \"us_per_row_delivering\": %.3f, \"share_waiting_for_gpu\": %.3f, "
                "\"share_waiting_for_frames\": %.3f, \"share_helping_search\": %.3f}, "
                "\"decode\": \"NSB_DECODE_BOTH + order_out (logits cached, probabilities and rank order out; %s)\", "
                "\"rules\": \"real: host/rules/shogi.h (perft-pinned move generation, mate, four-fold repetition, "
                "perpetual check, 27-point declaration, max ply), PUCT tree host/mcts_search.h; shallow mate search optional (--leaf-mate-plies)\"}\n",
                (double)(R1 - R0) / Sec, Evals / Sec, (double)(G1 - G0) / Sec, Batches > 0 ? Evals / Batches : 0.0, Sec,
                (unsigned long long)(R1 - R0), (unsigned long long)(E1 - E0), (unsigned long long)(B1 - B0),
                (unsigned long long)(G1 - G0), O.CacheMiB, Evals > 0 ? (double)(H1 - H0) / Evals : 0.0,
                Evals > 0 ? (double)(T1n - T0n) / Evals : 0.0, Evals > 0 ? (double)(L1 - L0) / Evals : 0.0,
                (unsigned long long)SI.Mates.load(), (unsigned long long)SI.Repetitions.load(), (unsigned long long)SI.PerpetualChecks.load(),
                (unsigned long long)SI.Declarations.load(), (unsigned long long)SI.MaxPlies.load(), (unsigned long long)SI.MatesBySearch.load(),
                O.MatePlies, (unsigned long long)SI.LeafMatesBySearch.load(),
                (unsigned long long)MaxDepthPly, (unsigned long long)Saved.Games.load(), (unsigned long long)Saved.Records.load(),
                (unsigned long long)Saved.Winners[0].load(), (unsigned long long)Saved.Winners[1].load(),
                (unsigned long long)Saved.Winners[2].load(), O.Out.c_str(), O.Blocks, O.Channels, O.Batch, O.Frames, O.SearchWorkers, O.Slots, O.MinFill, O.Playouts,
                O.FullSearchRatio, (unsigned long long)SI.NanRows.load(),
                1e6 * Evaluation.secondsFilling() / std::max<double>(1.0, (double)Evaluation.rows()),
                1e6 * Evaluation.secondsDelivering() / std::max<double>(1.0, (double)Evaluation.rows()),
                (Evaluation.secondsCollecting() - Evaluation.secondsHelping()) / EvalTotal, Evaluation.secondsTaking() / EvalTotal,
                Evaluation.secondsHelping() / EvalTotal,
                O.Gumbel ? "Gumbel roots skip the softmax" : "Dirichlet mix at full-search roots on the host");
    return 0;
}
