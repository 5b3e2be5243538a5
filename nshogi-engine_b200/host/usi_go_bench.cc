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
// (src/mcts/feedworker.cc:29-137).  Here: --num-search-threads search threads (default 2, context.h:74) descend the
// shared tree under virtual loss, generate the leaf's moves and write its row IN PLACE into the open pinned batch
// (host/leaf_queue.h: no queue of tuples, no copy); one evaluation thread seals and submits batches and feeds the
// results as soon as a batch is done.  Rules: host/rules/shogi.h; tree: host/mcts_search.h, shared lock-free (27-point declaration;
// no mate solver).
#include <atomic>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "infer_b200.h"
#include "leaf_pipeline.h"
#include "usi_search.h"

using namespace nshogi::engine;
using namespace nshogi::engine::b200;
using Clock = std::chrono::steady_clock;

int main(int argc, char** argv) {
    int Channels = 256, Blocks = 20, Batch = 512, Slots = 3, GPU = 0, CacheMiB = 0, SearchThreads = 2;
    bool NoHelp = false;
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
        else if (A == "--num-search-threads") SearchThreads = nextI();  // context.h:74 default 2
        else if (A == "--seed") Seed = (uint64_t)nextI();
        else if (A == "--no-help") NoHelp = true;  // the evaluation thread idles instead of collecting leaves between its duties
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
    if (SearchThreads < 1) SearchThreads = 1;
    infer::B200 Exec(GPU, (uint16_t)Batch, NSB_FEATURE_CHANNELS, Channels, Blocks, Slots);
    Exec.load("", Seed);
    if (CacheMiB > 0) Exec.enableCache((std::size_t)CacheMiB);  // Manager's EvalCache, manager.cc:202-206
    Exec.resetGPU();
    Exec.bindThreadToGpuNode();
    evaluate::LeafPipeline Pipe(&Exec, (std::size_t)Batch);  // search threads assemble batches in place in its pinned slots (host/leaf_queue.h)
    const std::size_t NS = Pipe.numSlots();

    const rules::Position Root;  // hirate
    // shared by the search threads and the feeding thread without a lock, like the reference's tree (node.h:59-100);
    // arenas sized for the run: ~350 k nodes/s at ~40 legal moves each
    const std::size_t MaxNodes = (std::size_t)(6.0e5 * (Seconds + 1.0)) + (1u << 16);
    search::Tree T(MaxNodes, MaxNodes * 48);
    const uint16_t MaxPly = 320;  // StateConfig default of the USI front-end
    // MCTS flavour of the decode: NSB_DECODE_PROBS + order_out (host/usi_search.h holds the search itself)
    UsiSearch<evaluate::LeafPipeline> Search(&T, &Pipe, Root, MaxPly, NSB_DECODE_PROBS, CacheMiB > 0);
    const double Sec = Search.run(Seconds, SearchThreads, !NoHelp);
    const uint64_t Evals = Search.Evals, Batches = Search.Batches, CacheHits = Search.CacheHits, GpuWaitNs = Search.GpuWaitNs;
    const uint64_t HelpedLeaves = Search.HelpedLeaves;
    const std::atomic<uint64_t>&Terminals = Search.Terminals, &Collisions = Search.Collisions, &LegalMoves = Search.LegalMoves;

    // usilogger.cc:29-65: nodes = visits of the root, nps, pv by most-visited edges
    const uint64_t Nodes = T.node(0).Visits;
    std::string PV;
    int Cur = 0, Depth = 0;
    while (Depth < 12) {
        const search::Node& N = T.node(Cur);
        if (!N.evaluated() || N.NumEdges == 0 || N.Term != search::Open) break;
        const int Best = T.bestEdge(Cur);
        if (Best < 0 || T.edgesOf(Cur)[Best].CVisits == 0) break;
        const rules::Move& M = T.edgesOf(Cur)[Best].M;
        char Buf[16];
        if (M.isDrop())
            std::snprintf(Buf, sizeof Buf, "%c*%d%c ", "PLNSGBR"[M.dropSlot()], M.To / 9 + 1, 'a' + M.To % 9);
        else
            std::snprintf(Buf, sizeof Buf, "%d%c%d%c%s ", M.From / 9 + 1, 'a' + M.From % 9, M.To / 9 + 1, 'a' + M.To % 9, M.Promote ? "+" : "");
        PV += Buf;
        Cur = T.edgesOf(Cur)[Best].Child;
        ++Depth;
    }
    if (!PV.empty()) PV.pop_back();
    const double RootWin = Nodes ? T.node(0).WinAcc / (double)Nodes : 0.0;
    std::printf("{\"metric\": \"usi_go_nodes_per_sec\", \"value\": %.1f, \"unit\": \"nodes/s\", \"nodes\": %llu, \"time_ms\": %.0f, "
                "\"leaf_evals_per_sec\": %.1f, \"avg_batch\": %.1f, \"batches\": %llu, \"terminal_leaves\": %llu, \"collisions\": %llu, "
                "\"cache_mb\": %d, \"cache_hit_rate\": %.4f, \"avg_legal_moves\": %.1f, \"tree_nodes\": %zu, \"submit_and_feed_fraction\": %.3f, "
                "\"root_win_rate\": %.4f, \"pv\": \"%s\", \"net\": \"%dx%d\", \"batch_size\": %d, \"slots\": %d, "
                "\"position\": \"hirate startpos\", \"search_threads\": %d, \"leaves_collected_by_the_evaluation_thread\": %llu, "
                "\"rules\": \"real: host/rules/shogi.h + host/mcts_search.h (PUCT, virtual loss, 27-point declaration); no mate solver\"}\n",
                (double)Nodes / Sec, (unsigned long long)Nodes, Sec * 1e3, (double)Evals / Sec, Batches ? (double)Evals / (double)Batches : 0.0,
                (unsigned long long)Batches, (unsigned long long)Terminals.load(), (unsigned long long)Collisions.load(), CacheMiB,
                Evals ? (double)CacheHits / (double)Evals : 0.0, Evals ? (double)LegalMoves.load() / (double)Evals : 0.0, T.numNodes(),
                (double)GpuWaitNs * 1e-9 / Sec, RootWin, PV.c_str(), Blocks, Channels, Batch, (int)NS, SearchThreads, (unsigned long long)HelpedLeaves);
    return 0;
}
