// host_bench.cc — the reference's micro-benchmark shape (reference src/bench/batchsize.cc:32-82:
// startpos x B, 4 warm-ups, Repeat x Evaluator.computeBlocking(B), prints "B, ms, evals/s") run
// against infer::B200, plus the pipelined fused-decode path and a self-check that the two agree.
// Pure host C++ over the C ABI: what a maintainer of the reference would link.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "eval_cache.h"
#include "infer_b200.h"
#include "leaf_pipeline.h"
#include "leaf_queue.h"
#include "move_index.h"

using namespace nshogi::engine;

static nsb_position startpos() {  // hirate, squares s = 9*(file-1) + (rank-1)
    nsb_position P;
    std::memset(&P, 0, sizeof P);
    auto put = [&](int file, int rank, int pt, int colour) { P.board[9 * (file - 1) + (rank - 1)] = (uint8_t)(1 + pt + 14 * colour); };
    const int back[9] = {1, 2, 3, 4, 5, 4, 3, 2, 1};  // L N S G K G S N L
    for (int f = 1; f <= 9; ++f) {
        put(f, 9, back[f - 1], 0);
        put(f, 1, back[9 - f], 1);
        put(f, 7, 0, 0);
        put(f, 3, 0, 1);
    }
    put(8, 8, 6, 0); put(2, 8, 7, 0); put(2, 2, 6, 1); put(8, 2, 7, 1);
    P.max_ply = 320;
    P.black_draw_value = P.white_draw_value = 0.5f;
    return P;
}

// The 30 legal moves of the start position through the move-index adaptor (pawn pushes, lance,
// silver, gold, king, rook and bishop-side moves); enough structure for a realistic decode row.
static std::vector<uint16_t> startposMoveIndices() {
    std::vector<uint16_t> Idx;
    auto sq = [](int file, int rank) { return 9 * (file - 1) + (rank - 1); };
    auto add = [&](int ff, int fr, int tf, int tr) {
        const int I = b200::getMoveIndex(0, b200::MoveSpec{sq(ff, fr), sq(tf, tr), false, -1});
        if (I >= 0) Idx.push_back((uint16_t)I);
    };
    for (int f = 1; f <= 9; ++f) add(f, 7, f, 6);                         // 9 pawn pushes
    add(1, 9, 1, 8); add(9, 9, 9, 8);                                       // lances
    add(3, 9, 3, 8); add(3, 9, 4, 8); add(7, 9, 7, 8); add(7, 9, 6, 8);     // silvers
    add(4, 9, 3, 8); add(4, 9, 4, 8); add(4, 9, 5, 8);                      // golds
    add(6, 9, 5, 8); add(6, 9, 6, 8); add(6, 9, 7, 8);
    add(5, 9, 4, 8); add(5, 9, 5, 8); add(5, 9, 6, 8);                      // king
    for (int f = 3; f <= 7; ++f) add(2, 8, f, 8);                           // rook slides
    add(2, 8, 1, 8);
    return Idx;
}

template <typename T>
static T* pinned(size_t N) {
    void* P = nullptr;
    infer::B200::check(nsb_host_alloc(&P, N * sizeof(T)), "nsb_host_alloc");
    return static_cast<T*>(P);
}

int main(int argc, char** argv) {
    int Channels = 128, Blocks = 10, B = 256, Repeat = 300, Slots = 4, QueueThreads = 4;
    bool SelfCheck = false, MallocBuffers = false;
    std::string Weights;  // NSBW file (nshogi-engine_b200/weights_io.py); empty = seeded random-init net
    for (int I = 1; I < argc; ++I) {
        const std::string A = argv[I];
        auto next = [&]() { return I + 1 < argc ? std::atoi(argv[++I]) : 0; };
        if (A == "--channels") Channels = next();
        else if (A == "--blocks") Blocks = next();
        else if (A == "--batch") B = next();
        else if (A == "--repeat") Repeat = next();
        else if (A == "--slots") Slots = next();
        else if (A == "--queue-threads") QueueThreads = next();
        else if (A == "--selfcheck") SelfCheck = true;
        else if (A == "--malloc-buffers") MallocBuffers = true;  // plain page-aligned memory: infer::B200 page-locks it itself
        else if (A == "--weights" && I + 1 < argc) Weights = argv[++I];
    }
    if (nsb_device_count() < 1) {
        std::fprintf(stderr, "nsb_host_bench: no CUDA device; infer::B200 has no CPU fallback\n");
        return 2;
    }
    infer::B200 Exec(0, (uint16_t)B, NSB_FEATURE_CHANNELS, Channels, Blocks, Slots);
    Exec.load(Weights);  // "" = seeded random-init net (the reference ships no model)
    Exec.resetGPU();

    // --- Evaluator-style pinned batch buffers (reference src/evaluate/evaluator.cc:85-106) -----------
    auto buffer = [&](auto* Tag, std::size_t N) {
        using T = std::remove_pointer_t<decltype(Tag)>;
        if (!MallocBuffers) return pinned<T>(N);
        return static_cast<T*>(std::aligned_alloc(4096, ((N * sizeof(T) + 4095) / 4096) * 4096));
    };
    auto* Features = buffer((nshogi::ml::FeatureBitboard*)nullptr, (size_t)B * NSB_FEATURE_CHANNELS);
    float* Policy = buffer((float*)nullptr, (size_t)B * NSB_POLICY_SIZE);
    float* Win = buffer((float*)nullptr, (size_t)B);
    float* Draw = buffer((float*)nullptr, (size_t)B);
    {   // startpos x B (batchsize.cc:47-59); stage 1 runs on the device, features come back once
        std::vector<nsb_position> Pos(B, startpos());
        void *DPos = nullptr, *DFeat = nullptr;
        infer::B200::check(nsb_device_alloc(&DPos, Pos.size() * sizeof(nsb_position)), "alloc");
        infer::B200::check(nsb_device_alloc(&DFeat, (size_t)B * NSB_FEATURE_CHANNELS * 16), "alloc");
        infer::B200::check(nsb_memcpy_h2d(DPos, Pos.data(), Pos.size() * sizeof(nsb_position)), "h2d");
        infer::B200::check(nsb_pack_positions_device(Exec.context(), 0, static_cast<nsb_position*>(DPos), B,
                                                     static_cast<nsb_feature_bitboard*>(DFeat)), "pack");
        infer::B200::check(nsb_await(Exec.context(), 0), "await");
        infer::B200::check(nsb_memcpy_d2h(Features, DFeat, (size_t)B * NSB_FEATURE_CHANNELS * 16), "d2h");
        nsb_device_free(DPos);
        nsb_device_free(DFeat);
    }
    for (int I = 0; I < 4; ++I) Exec.computeBlocking(Features, B, Policy, Win, Draw);  // batchsize.cc:61-63
    auto T0 = std::chrono::steady_clock::now();
    for (int I = 0; I < Repeat; ++I) Exec.computeBlocking(Features, B, Policy, Win, Draw);
    double Ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
    const double BlockingRate = (double)B * Repeat / Ms * 1000.0;
    std::printf("%d, %.3f, %.1f\n", B, Ms, BlockingRate);  // "B, ms, evals/s" (batchsize.cc:77-79)

    // --- pipelined fused-decode path: positions in, legal-move rows out ---------------------------------
    evaluate::LeafPipeline Pipe(&Exec, B);
    mcts::EvalCacheB200 Cache(64);
    const std::vector<uint16_t> Moves = startposMoveIndices();
    std::vector<std::vector<uint64_t>> SlotHashes(Pipe.numSlots(), std::vector<uint64_t>(B));
    auto fill = [&](evaluate::LeafPipeline::Slot& S, std::vector<uint64_t>& Hashes, uint64_t Salt) {
        const nsb_position P = startpos();
        for (int I = 0; I < B; ++I) {
            S.Positions[I] = P;
            S.MoveOffsets[I] = (uint32_t)(I * Moves.size());
            std::memcpy(S.MoveIndices + I * Moves.size(), Moves.data(), Moves.size() * sizeof(uint16_t));
            Hashes[I] = (Salt * 0x9E3779B97F4A7C15ull) ^ (uint64_t)I * 0xD1B54A32D192ED03ull;
        }
        S.MoveOffsets[B] = (uint32_t)(B * Moves.size());
    };
    size_t Stored = 0;
    auto run = [&](int Steps) {
        for (int I = 0; I < Steps; ++I) {
            size_t Idx;
            evaluate::LeafPipeline::Slot& S = Pipe.acquire(&Idx);  // collects the slot's previous batch
            if (S.Count)  // the slot's previous batch has landed: decode rows -> evaluation cache
                Stored += Cache.feed(SlotHashes[Idx].data(), S.Count, S.MoveOffsets, S.Legal, S.WinRate, S.DrawRate);
            fill(S, SlotHashes[Idx], (uint64_t)I);
            Pipe.submit(Idx, B, /*FromPositions=*/true, NSB_DECODE_PROBS, /*UseCache=*/false, /*Ranked=*/true);
        }
        Pipe.drain();
        for (size_t Idx = 0; Idx < Pipe.numSlots(); ++Idx) {
            evaluate::LeafPipeline::Slot& S = Pipe.collect(Idx);
            if (S.Count) Stored += Cache.feed(SlotHashes[Idx].data(), S.Count, S.MoveOffsets, S.Legal, S.WinRate, S.DrawRate);
            S.Count = 0;
        }
    };
    run(8);
    T0 = std::chrono::steady_clock::now();
    run(Repeat);
    Ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
    const double PipeRate = (double)B * Repeat / Ms * 1000.0;

    // --- the MCTS-side assembly: QueueThreads search threads write their leaves straight into the open pinned
    //     batch (leaf_queue.h: one CAS per leaf, no mutex / allocation / copy; replaces evaluationqueue.cc +
    //     EvaluationWorker::getBatch), this thread seals, submits and feeds ----------------------------------------
    double QueueRate = 0.0;
    bool QueueOk = true;
    std::size_t QueueFed = 0;
    if (QueueThreads > 0) {
        evaluate::LeafQueue Queue(&Pipe);
        const std::size_t Target = (std::size_t)B * Repeat;
        std::atomic<std::size_t> Pushed{0};
        std::atomic<bool> Stop{false};
        const nsb_position P0 = startpos();
        auto feed = [&](evaluate::LeafPipeline::Slot& S, std::size_t Row, void* User) {
            const uint32_t Bg = S.MoveOffsets[Row], En = S.MoveOffsets[Row + 1];
            // every leaf is the start position: the most probable move's row index must be the same everywhere,
            // and the handle must be the one pushed with the row
            QueueOk = QueueOk && En - Bg == Moves.size() && S.Order[Bg] == S.Order[0] && User == (void*)(uintptr_t)(Row + 1);
            ++QueueFed;
        };
        Queue.open(feed);
        std::vector<std::thread> Search;
        for (int T = 0; T < QueueThreads; ++T)
            Search.emplace_back([&]() {
                evaluate::LeafQueue::Ticket Tk;
                while (!Stop.load(std::memory_order_relaxed)) {
                    if (Pushed.load(std::memory_order_relaxed) >= Target) break;
                    if (!Queue.reserve((uint16_t)Moves.size(), nullptr, &Tk)) {
                        std::this_thread::yield();
                        continue;
                    }
                    Tk.S->Positions[Tk.Row] = P0;
                    std::memcpy(Tk.S->MoveIndices + Tk.MoveBegin, Moves.data(), Moves.size() * sizeof(uint16_t));
                    Tk.S->Hashes[Tk.Row] = (uint64_t)Tk.Row;
                    Queue.setUser(Tk, (void*)(uintptr_t)(Tk.Row + 1));
                    Queue.publish(Tk);
                    Pushed.fetch_add(1, std::memory_order_relaxed);
                }
            });
        T0 = std::chrono::steady_clock::now();
        std::size_t Submitted = 0;
        while (Submitted < Target) {
            if (Queue.openRows() < (std::size_t)B && Pushed.load(std::memory_order_relaxed) < Target) {
                std::this_thread::yield();
                continue;
            }
            Submitted += Queue.submitOpen(/*FromPositions=*/true, NSB_DECODE_PROBS, /*UseCache=*/false, /*Ranked=*/true, feed);
        }
        Stop.store(true);
        for (auto& Th : Search) Th.join();
        Queue.drain(true, NSB_DECODE_PROBS, false, true, feed);
        Ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
        QueueRate = (double)QueueFed / Ms * 1000.0;
        QueueOk = QueueOk && QueueFed >= Target;
    }

    // --- self-check: fused probabilities == softmax of the Infer contract's logits at the same slots -----
    double MaxDiff = 0.0, SumErr = 0.0;
    bool RowsEqual = true, OrderOk = true;
    {
        evaluate::LeafPipeline::Slot& S = Pipe.collect(0);
        const size_t M = Moves.size();
        S.Count = 0;
        std::vector<double> Ref(M);
        double Mx = -1e30, Sum = 0.0;
        for (size_t J = 0; J < M; ++J) Mx = std::fmax(Mx, (double)Policy[Moves[J]]);
        for (size_t J = 0; J < M; ++J) Sum += (Ref[J] = std::exp((double)Policy[Moves[J]] - Mx));
        double RowSum = 0.0;
        for (size_t J = 0; J < M; ++J) {
            MaxDiff = std::fmax(MaxDiff, std::fabs(Ref[J] / Sum - (double)S.Legal[J]));
            RowSum += S.Legal[J];
        }
        SumErr = std::fabs(RowSum - 1.0);
        for (int I = 1; I < B; ++I)
            RowsEqual = RowsEqual && std::memcmp(S.Legal, S.Legal + I * M, M * sizeof(float)) == 0 &&
                        std::memcmp(Policy, Policy + (size_t)I * NSB_POLICY_SIZE, NSB_POLICY_SIZE * sizeof(float)) == 0;
        MaxDiff = std::fmax(MaxDiff, std::fabs((double)S.WinRate[0] - (double)Win[0]));
        // the rank order replaces Node::sort() (src/mcts/node.h:163-168): a permutation with non-increasing values
        std::vector<int> Seen(M, 0);
        for (size_t R = 0; R < M; ++R) {
            const uint16_t J = S.Order[R];
            OrderOk = OrderOk && J < M && !Seen[J]++ && (R == 0 || S.Legal[S.Order[R - 1]] >= S.Legal[J]);
        }
    }
    mcts::EvalCacheB200::EvalInfo Info;
    const bool CacheHit = Cache.load(SlotHashes[0][B / 2], &Info) && Info.NumMoves == Moves.size();
    const bool Ok = MaxDiff < 1e-5 && SumErr < 1e-5 && RowsEqual && CacheHit && OrderOk && QueueOk;
    std::printf("{\"batch\": %d, \"net\": \"%dx%d\", \"infer_blocking_evals_per_s\": %.1f, \"pipeline_evals_per_s\": %.1f, "
                "\"queue_evals_per_s\": %.1f, \"queue_threads\": %d, \"queue_ok\": %s, \"io_mode\": \"%s\", \"legal_moves\": %zu, \"max_prob_diff\": %.3g, \"row_sum_err\": %.3g, \"rows_identical\": %s, "
                "\"cache_rows_stored\": %zu, \"cache_hit\": %s, \"order_ok\": %s, \"ok\": %s}\n",
                B, Exec.net().blocks, Exec.net().channels, BlockingRate, PipeRate, QueueRate, QueueThreads, QueueOk ? "true" : "false", nsb_io_mode(Exec.context()) == NSB_IO_DIRECT ? "direct" : "staged", Moves.size(), MaxDiff, SumErr, RowsEqual ? "true" : "false",
                Stored, CacheHit ? "true" : "false", OrderOk ? "true" : "false", Ok ? "true" : "false");
    if (!MallocBuffers) {
        nsb_host_free(Features); nsb_host_free(Policy); nsb_host_free(Win); nsb_host_free(Draw);
    }  // (malloc'ed buffers are released at exit, after the executor has unlocked them in its destructor)
    return (SelfCheck && !Ok) ? 1 : 0;
}
