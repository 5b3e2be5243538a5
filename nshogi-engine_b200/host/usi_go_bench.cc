// usi_go_bench.cc — "USI go from hirate startpos, 20-block 256-ch ResNet, batch 512, 1xB200 nodes/sec" (BASELINE.json
// configs[2]) with real rules: ONE search tree rooted at the start position, leaves collected in batches of up to
// --batch-size under virtual loss, evaluated by infer::B200 through the pinned multi-slot LeafPipeline (MCTS flavour of
// the fused decode: NSB_DECODE_PROBS + order_out = FeedWorker::feedResult's gather + softmax + Node::sort, optional
// device-resident cache = the Manager's EvalCache), results fed back as setEvaluation + updateAncestors.  Prints what
// the reference's USI logger prints after a search - nodes, time, nps (src/protocol/usilogger.cc:29-65: nps = visited
// nodes * 1000 / elapsed ms) - and the principal variation by most-visited edges.
//
// Reference structure: SearchWorker::doTask (src/mcts/searchworker.cc:448-609: collectOneLeaf -> terminal checks ->
// EvalCache load -> EvaluationQueue::add), EvaluationWorker (src/mcts/evaluationworker.cc:105-199), FeedWorker
// (src/mcts/feedworker.cc:29-137).  Here one collector thread does the three roles around the slot ring: while a batch is
// on the GPU it descends the tree for the next one (the virtual losses of the batch in flight steer it elsewhere).  Rules:
// host/rules/shogi.h; tree: host/mcts_search.h (no df-pn, no declaration win, no tree-parallel search threads: the
// number printed is one collector thread's, and it says whether the CPU or the GPU was the limit).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "infer_b200.h"
#include "leaf_pipeline.h"
#include "mcts_search.h"
#include "rules/shogi.h"

using namespace nshogi::engine;
using namespace nshogi::engine::b200;
using Clock = std::chrono::steady_clock;

int main(int argc, char** argv) {
    int Channels = 256, Blocks = 20, Batch = 512, Slots = 3, GPU = 0, CacheMiB = 0, MaxCollisions = 64;
    double Seconds = 5.0;
    uint64_t Seed = 1234;
    for (int I = 1; I < argc; ++I) {
        const std::string A = argv[I];
        auto nextI = [&]() { return I + 1 < argc ? std::atoi(argv[++I]) : 0; };
        if (A == "--channels") Channels = nextI();
        else if (A == "--blocks") Blocks = nextI();
        else if (A == "--batch-size") Batch = nextI();
        else if (A == "--slots") Slots = nextI();
        else if (A == "--gpu") GPU = nextI();
        else if (A == "--cache-mb") CacheMiB = nextI();
        else if (A == "--max-collisions") MaxCollisions = nextI();
        else if (A == "--seed") Seed = (uint64_t)nextI();
        else if (A == "--seconds") Seconds = I + 1 < argc ? std::atof(argv[++I]) : 0.0;
        else {
            std::fprintf(stderr, "unknown option %s\n", A.c_str());
            return 2;
        }
    }
    if (nsb_device_count() <= GPU) {
        std::fprintf(stderr, "nsb_usi_go_bench: no CUDA device %d; infer::B200 has no CPU fallback\n", GPU);
        return 2;
    }
    infer::B200 Exec(GPU, (uint16_t)Batch, NSB_FEATURE_CHANNELS, Channels, Blocks, Slots);
    Exec.load("", Seed);
    if (CacheMiB > 0) Exec.enableCache((std::size_t)CacheMiB);  // Manager's EvalCache, manager.cc:202-206
    Exec.resetGPU();
    Exec.bindThreadToGpuNode();
    evaluate::LeafPipeline Pipe(&Exec, (std::size_t)Batch);
    const std::size_t NS = Pipe.numSlots();

    rules::Position Root;  // hirate
    std::vector<uint64_t> History{Root.Hash}, Path;
    search::Tree T;
    T.reset();
    T.Nodes.reserve(1u << 21);
    T.Edges.reserve(1u << 25);
    const uint16_t MaxPly = 320;  // StateConfig default of the USI front-end
    std::vector<std::vector<int>> SlotNodes(NS);
    std::deque<std::size_t> InFlight;
    uint64_t Evals = 0, Batches = 0, Terminals = 0, Collisions = 0, CacheHits = 0, LegalMoves = 0, GpuWaitNs = 0;

    auto deliver = [&](std::size_t Idx) {
        const auto W0 = Clock::now();
        evaluate::LeafPipeline::Slot& S = Pipe.collect(Idx);
        GpuWaitNs += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - W0).count();
        const std::vector<int>& Ns = SlotNodes[Idx];
        for (std::size_t I = 0; I < Ns.size(); ++I) {  // FeedWorker::feedResult, feedworker.cc:56-137
            const uint32_t B = S.MoveOffsets[I];
            T.setPriors(Ns[I], S.Legal + B, S.Order + B);
            T.backup(Ns[I], S.WinRate[I], S.DrawRate[I]);
            if (CacheMiB > 0 && S.HitFlag[I]) ++CacheHits;
        }
        Evals += Ns.size();
        ++Batches;
        SlotNodes[Idx].clear();
    };

    rules::Move Moves[rules::kMaxMoves];
    const auto T0 = Clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(Clock::now() - T0).count(); };
    std::size_t NextSlot = 0;  // the ring advances only when a batch is submitted
    while (elapsed() < Seconds) {
        if (InFlight.size() == NS) {
            deliver(InFlight.front());
            InFlight.pop_front();
        }
        const std::size_t Idx = NextSlot;
        evaluate::LeafPipeline::Slot& S = Pipe.slotAt(Idx);  // (free: at most NS - 1 batches are in flight here)
        std::vector<int>& Ns = SlotNodes[Idx];
        uint32_t Off = 0;
        int Failed = 0;
        while ((int)Ns.size() < Batch && Failed < MaxCollisions) {
            rules::Position Pos = Root;
            Path.clear();
            const int Node = T.selectLeaf(Pos, 0.5f, 0.5f, &Path);  // SearchWorker::collectOneLeaf
            if (Node < 0) {                                         // ran into a leaf that is being evaluated
                ++Failed;
                ++Collisions;
                continue;
            }
            search::Node& N = T.Nodes[(std::size_t)Node];
            if (N.Term == search::Mated) {
                T.backup(Node, 0.0f, 0.0f);
                continue;
            }
            if (N.Term == search::DrawnGame) {
                T.backup(Node, 0.5f, 1.0f);
                continue;
            }
            const int NumMoves = Pos.generateLegal(Moves);          // expandLeaf, searchworker.cc:164-173
            if (NumMoves == 0) {                                    // terminal checks, :475-538
                N.Term = search::Mated;
                N.Evaluated = true;
                ++Terminals;
                T.backup(Node, 0.0f, 0.0f);
                continue;
            }
            if (Node != 0 && (search::isFourfold(Pos.Hash, History, Path) || Pos.Ply >= MaxPly)) {
                N.Term = search::DrawnGame;
                N.Evaluated = true;
                ++Terminals;
                T.backup(Node, 0.5f, 1.0f);
                continue;
            }
            T.expand(Node, Moves, NumMoves);
            const std::size_t Row = Ns.size();
            Pos.toRecord(&S.Positions[Row], MaxPly, 0.5f, 0.5f);    // stage 1 runs on the GPU
            S.Hashes[Row] = Pos.Hash;
            S.MoveOffsets[Row] = Off;
            for (int J = 0; J < NumMoves; ++J) S.MoveIndices[Off + (uint32_t)J] = (uint16_t)Pos.policyIndex(Moves[J]);
            Off += (uint32_t)NumMoves;
            LegalMoves += (uint64_t)NumMoves;
            Ns.push_back(Node);
        }
        if (Ns.empty()) {  // everything reachable is in flight: wait for the oldest batch
            if (!InFlight.empty()) {
                deliver(InFlight.front());
                InFlight.pop_front();
            }
            continue;
        }
        S.MoveOffsets[Ns.size()] = Off;
        Pipe.submit(Idx, Ns.size(), /*FromPositions=*/true, NSB_DECODE_PROBS, /*UseCache=*/CacheMiB > 0, /*Ranked=*/true);
        InFlight.push_back(Idx);
        NextSlot = (NextSlot + 1) % NS;
    }
    while (!InFlight.empty()) {
        deliver(InFlight.front());
        InFlight.pop_front();
    }
    const double Sec = elapsed();

    // usilogger.cc:29-65: nodes = visits of the root, nps, pv by most-visited edges
    const uint64_t Nodes = T.Nodes[0].Visits;
    std::string PV;
    int Cur = 0, Depth = 0;
    while (Depth < 12) {
        const search::Node& N = T.Nodes[(std::size_t)Cur];
        if (!N.Evaluated || N.NumEdges == 0) break;
        int Best = -1;
        uint32_t BestV = 0;
        for (int I = 0; I < N.NumEdges; ++I) {
            const search::Edge& E = T.Edges[(std::size_t)N.EdgeBegin + (std::size_t)I];
            const uint32_t V = E.Child >= 0 ? T.Nodes[(std::size_t)E.Child].Visits : 0;
            if (V > BestV) {
                BestV = V;
                Best = I;
            }
        }
        if (Best < 0) break;
        const rules::Move& M = T.Edges[(std::size_t)N.EdgeBegin + (std::size_t)Best].M;
        char Buf[16];
        if (M.isDrop())
            std::snprintf(Buf, sizeof Buf, "%c*%d%c ", "PLNSGBR"[M.dropSlot()], M.To / 9 + 1, 'a' + M.To % 9);
        else
            std::snprintf(Buf, sizeof Buf, "%d%c%d%c%s ", M.From / 9 + 1, 'a' + M.From % 9, M.To / 9 + 1, 'a' + M.To % 9, M.Promote ? "+" : "");
        PV += Buf;
        Cur = T.Edges[(std::size_t)N.EdgeBegin + (std::size_t)Best].Child;
        ++Depth;
    }
    if (!PV.empty()) PV.pop_back();
    const double RootWin = Nodes ? T.Nodes[0].WinAcc / (double)Nodes : 0.0;
    std::printf("{\"metric\": \"usi_go_nodes_per_sec\", \"value\": %.1f, \"unit\": \"nodes/s\", \"nodes\": %llu, \"time_ms\": %.0f, "
                "\"leaf_evals_per_sec\": %.1f, \"avg_batch\": %.1f, \"batches\": %llu, \"terminal_leaves\": %llu, \"collisions\": %llu, "
                "\"cache_mb\": %d, \"cache_hit_rate\": %.4f, \"avg_legal_moves\": %.1f, \"tree_nodes\": %zu, \"gpu_wait_fraction\": %.3f, "
                "\"root_win_rate\": %.4f, \"pv\": \"%s\", \"net\": \"%dx%d\", \"batch_size\": %d, \"slots\": %d, "
                "\"position\": \"hirate startpos\", \"search_threads\": 1, "
                "\"rules\": \"real: host/rules/shogi.h + host/mcts_search.h (PUCT, virtual loss); no df-pn, no declaration win\"}\n",
                (double)Nodes / Sec, (unsigned long long)Nodes, Sec * 1e3, (double)Evals / Sec, Batches ? (double)Evals / (double)Batches : 0.0,
                (unsigned long long)Batches, (unsigned long long)Terminals, (unsigned long long)Collisions, CacheMiB,
                Evals ? (double)CacheHits / (double)Evals : 0.0, Evals ? (double)LegalMoves / (double)Evals : 0.0, T.Nodes.size(),
                (double)GpuWaitNs * 1e-9 / Sec, RootWin, PV.c_str(), Blocks, Channels, Batch, (int)NS);
    return 0;
}
