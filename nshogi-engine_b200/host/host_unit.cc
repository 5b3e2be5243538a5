// host_unit.cc — CPU-only unit checks of the host-side pieces (no GPU, no CUDA calls):
// EvalCacheB200 behaviour (reference src/mcts/evalcache.cc) and the move-index adaptor.
// `--cache-trace NUM_BUNDLES_MIB` replays "s hash n win draw" / "l hash n" lines from stdin through
// EvalCacheB200 and prints one result line per operation (tests/test_evalcache.py compares them
// with the oracle and with the reference's own evalcache.cc).
// `--onnx-blob IN.onnx OUT.bin` runs the model-file reader of infer::B200::load (onnx_import.h) and writes
// i32 in_channels, channels, blocks, hidden + the canonical fp32 blob (tests/test_onnx_io.py compares it with
// the Python reader bit for bit); a graph it rejects prints the reason and exits 3.
#include <cstdio>
#include <chrono>
#include <mutex>
#include <deque>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>
#include <vector>

#include <fstream>
#include <functional>
#include <iterator>

#include <algorithm>
#include <cmath>
#include <random>

#include <atomic>
#include <thread>

#include "eval_cache.h"
#include "evaluation_worker_b200.h"
#include "leaf_queue.h"
#include "mcts_feed.h"
#include "selfplay_feed.h"
#include "move_index.h"
#include "onnx_import.h"
#include "selfplay_game.h"
#include "selfplay_workers.h"
#include "usi_search.h"

using namespace nshogi::engine;

#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int cacheTrace(std::size_t MiB) {
    mcts::EvalCacheB200 C(MiB);
    std::printf("bundles %zu\n", C.numBundles());
    char Op;
    unsigned long long H;
    unsigned N;
    while (std::scanf(" %c %llu %u", &Op, &H, &N) == 3) {
        if (Op == 's') {
            float W, D;
            if (std::scanf("%f %f", &W, &D) != 2) return 1;
            std::vector<float> Row(N);
            for (unsigned J = 0; J < N; ++J) Row[J] = W + (float)J;  // the row is a function of (win, j)
            std::printf("%d\n", C.store(H, (uint16_t)N, Row.data(), W, D) ? 1 : 0);
        } else {
            mcts::EvalCacheB200::EvalInfo E;
            const bool Hit = C.load(H, &E) && E.NumMoves == N;  // searchworker.cc:545-556
            if (Hit) std::printf("1 %.9g %.9g %.9g\n", E.WinRate, E.DrawRate, N ? E.Policy[N - 1] : 0.f);
            else std::printf("0\n");
        }
    }
    return 0;
}

static int onnxBlob(const char* InPath, const char* OutPath) {
    std::ifstream In(InPath, std::ios::binary);
    if (!In) return 2;
    std::vector<uint8_t> Bytes((std::istreambuf_iterator<char>(In)), std::istreambuf_iterator<char>());
    std::vector<float> Blob;
    infer::onnx::NetShape S;
    try {
        S = infer::onnx::importModel(Bytes, &Blob);
    } catch (const infer::onnx::Error& E) {
        std::printf("rejected: %s\n", E.what());
        return 3;
    }
    const int32_t Head[4] = {S.InChannels, S.Channels, S.Blocks, S.Hidden};
    std::ofstream Out(OutPath, std::ios::binary);
    Out.write(reinterpret_cast<const char*>(Head), sizeof Head);
    Out.write(reinterpret_cast<const char*>(Blob.data()), (std::streamsize)(Blob.size() * sizeof(float)));
    std::printf("ok %d %d %d %d %zu\n", S.InChannels, S.Channels, S.Blocks, S.Hidden, Blob.size());
    return Out ? 0 : 2;
}

// ---- mock of the reference's mcts::Node / Edge surface that feedRanked touches (src/mcts/node.h, edge.h) ----
struct MockEdge {
    uint16_t Move = 0;
    float Probability = 0.f;
    uint16_t getMove() const { return Move; }
    void setMove(uint16_t M) { Move = M; }
    void setProbability(float P) { Probability = P; }
    float getProbability() const { return Probability; }
};
struct MockNode {
    static constexpr uint64_t VisitMask = (1ull << 32) - 1;
    std::vector<MockEdge> Edges;
    MockNode* Parent = nullptr;
    double WinAcc = 0.0, DrawAcc = 0.0;
    uint64_t Visits = 0;
    float WinPred = -1.f, DrawPred = -1.f, BackedWin = -1.f, BackedDraw = -1.f;
    uint16_t getNumChildren() const { return (uint16_t)Edges.size(); }
    MockEdge* getEdge() { return Edges.data(); }
    const MockNode* getParent() const { return Parent; }
    double getWinRateAccumulated() const { return WinAcc; }
    double getDrawRateAccumulated() const { return DrawAcc; }
    uint64_t getVisitsAndVirtualLoss() const { return Visits; }
    void setEvaluation(const float* Policy, float W, float D) {  // node.h:150-160
        if (Policy) for (size_t I = 0; I < Edges.size(); ++I) Edges[I].setProbability(Policy[I]);
        WinPred = W;
        DrawPred = D;
    }
    void sort() {  // node.h:163-168 (stable here: the tie order the GPU rank defines)
        std::stable_sort(Edges.begin(), Edges.end(), [](const MockEdge& A, const MockEdge& B) { return A.Probability > B.Probability; });
    }
    void updateAncestors(float W, float D) { BackedWin = W; BackedDraw = D; }
};

// ---- LeafQueue protocol on plain memory: `--queue-stress THREADS LEAVES` (built with -fsanitize=thread by
//      tests/test_host_cpp.py).  Every pushed leaf must come back exactly once, in a batch whose CSR is consistent.
struct FakePipeline {
    struct Slot {
        std::vector<uint32_t> MoveOffsets;
        std::vector<uint16_t> MoveIndices;
        std::vector<uint64_t> Hashes;
        std::size_t Count = 0;
    };
    std::vector<Slot> Slots;
    std::size_t BatchMax, Next = 0, Batches = 0;
    FakePipeline(std::size_t NumSlots, std::size_t B) : Slots(NumSlots), BatchMax(B) {
        for (auto& S : Slots) {
            S.MoveOffsets.resize(B + 1);
            S.MoveIndices.resize(B * NSB_MAX_LEGAL_MOVES);
            S.Hashes.resize(B);
        }
    }
    std::size_t numSlots() const { return Slots.size(); }
    std::size_t batchMax() const { return BatchMax; }
    Slot& acquire(std::size_t* K) { *K = Next; Next = (Next + 1) % Slots.size(); return Slots[*K]; }
    void submit(std::size_t K, std::size_t Rows, bool, int, bool, bool) { Slots[K].Count = Rows; ++Batches; }
    Slot& collect(std::size_t K) { return Slots[K]; }
    bool ready(std::size_t) const { return true; }
};

static int queueStress(int Threads, std::size_t Leaves) {
    FakePipeline Pipe(3, 64);
    evaluate::BasicLeafQueue<FakePipeline> Queue(&Pipe);
    std::vector<uint8_t> SeenLeaf(Leaves, 0);
    std::size_t Fed = 0;
    bool Ok = true;
    auto feed = [&](FakePipeline::Slot& S, std::size_t Row, void* User) {
        const std::size_t Id = (std::size_t)(uintptr_t)User - 1;
        const uint32_t Bg = S.MoveOffsets[Row], En = S.MoveOffsets[Row + 1];
        Ok = Ok && Id < Leaves && !SeenLeaf[Id] && S.Hashes[Row] == Id * 2654435761ull && En - Bg == 1 + Id % 7;
        for (uint32_t J = Bg; Ok && J < En; ++J) Ok = S.MoveIndices[J] == (uint16_t)(Id + J - Bg);
        if (Id < Leaves) SeenLeaf[Id] = 1;
        ++Fed;
    };
    Queue.open(feed);
    std::atomic<std::size_t> NextLeaf{0};
    std::vector<std::thread> Search;
    for (int T = 0; T < Threads; ++T)
        Search.emplace_back([&]() {
            for (;;) {
                const std::size_t Id = NextLeaf.fetch_add(1);
                if (Id >= Leaves) return;
                const uint16_t N = (uint16_t)(1 + Id % 7);
                evaluate::BasicLeafQueue<FakePipeline>::Ticket Tk;
                while (!Queue.reserve(N, (void*)(uintptr_t)(Id + 1), &Tk)) std::this_thread::yield();
                for (uint16_t J = 0; J < N; ++J) Tk.S->MoveIndices[Tk.MoveBegin + J] = (uint16_t)(Id + J);
                Tk.S->Hashes[Tk.Row] = Id * 2654435761ull;
                Queue.publish(Tk);
            }
        });
    std::size_t Submitted = 0, Polls = 0, PolledRows = 0;
    while (Submitted < Leaves) {
        if (Queue.openRows() < 48 && NextLeaf.load() < Leaves + (std::size_t)Threads) {  // partial batches too
            // results are fed as soon as they exist (pollFeed), not only when the ring comes round: what a tree search needs
            if ((++Polls & 3) == 0) PolledRows += Queue.pollFeed(feed);
            std::this_thread::yield();
            continue;
        }
        Ok = Ok && Queue.inFlight() < Pipe.numSlots();
        Submitted += Queue.submitOpen(true, 0, false, true, feed);
    }
    for (auto& Th : Search) Th.join();
    Queue.drain(true, 0, false, true, feed);
    std::printf("queue stress: %zu leaves fed in %zu batches by %d threads (%zu rows fed early by pollFeed): %s\n", Fed, Pipe.Batches,
                Threads, PolledRows, Ok && Fed == Leaves ? "ok" : "FAIL");
    return Ok && Fed == Leaves ? 0 : 1;
}

static int feedChecks() {
    std::mt19937_64 Rng(7);
    for (int Trial = 0; Trial < 200; ++Trial) {
        const uint16_t N = (uint16_t)(Trial < 3 ? Trial + 1 : 2 + Rng() % 300);
        std::vector<float> Legal(N);
        double Sum = 0.0;
        for (auto& P : Legal) Sum += (P = (float)((Rng() % 1000) + 1));
        for (auto& P : Legal) P = (float)(P / Sum);
        if (Trial % 5 == 0 && N > 4) Legal[1] = Legal[3] = Legal[0];  // exact ties
        std::vector<uint16_t> Order(N);
        for (uint16_t I = 0; I < N; ++I) Order[I] = I;
        std::stable_sort(Order.begin(), Order.end(), [&](uint16_t A, uint16_t B) { return Legal[A] > Legal[B]; });  // == the GPU's rank order
        MockNode Parent, A, B;
        Parent.WinAcc = 30.0; Parent.DrawAcc = 5.0; Parent.Visits = 100;
        A.Edges.resize(N);
        for (uint16_t I = 0; I < N; ++I) A.Edges[I].Move = (uint16_t)(1000 + I);
        B = A;
        A.Parent = B.Parent = (Trial % 2) ? &Parent : nullptr;
        const bool NanW = Trial % 7 == 0, NanD = Trial % 11 == 0;
        const float W = NanW ? std::nanf("") : 0.625f, D = NanD ? std::nanf("") : 0.125f;
        // the reference's sequence (feedworker.cc:58-131) on node A
        float RW = W, RD = D;
        if (NanW) RW = A.Parent ? (float)(1.0 - 30.0 / 100.0) : 0.5f;
        if (NanD) RD = A.Parent ? (float)(5.0 / 100.0) : 0.0f;
        if (N == 1) { float One = 1.0f; A.setEvaluation(&One, RW, RD); }
        else { A.setEvaluation(Legal.data(), RW, RD); A.sort(); }
        A.updateAncestors(RW, RD);
        // feedRanked<true> on node B
        const mcts::LeafRow Row{Legal.data(), Order.data(), N, W, D};
        const bool Nan = mcts::feedRanked<true>(&B, Row);
        CHECK(Nan == (NanW || NanD));
        CHECK(B.WinPred == A.WinPred && B.DrawPred == A.DrawPred && B.BackedWin == A.BackedWin && B.BackedDraw == A.BackedDraw);
        for (uint16_t I = 0; I < N; ++I) CHECK(B.Edges[I].Move == A.Edges[I].Move && B.Edges[I].Probability == A.Edges[I].Probability);
        // feedResult<false>, the reference's default build (context.h:103): nothing is replaced, nothing is found
        MockNode C2;
        C2.Edges.resize(N);
        C2.Parent = A.Parent;
        const bool Nan2 = mcts::feedRanked<false>(&C2, Row);
        CHECK(!Nan2);
        CHECK(NanW ? std::isnan(C2.WinPred) : C2.WinPred == W);
        CHECK(NanD ? std::isnan(C2.DrawPred) : C2.DrawPred == D);
    }
    return 0;
}

// selfplay_feed.h == the tail of Frame::setEvaluation (frame.cc:116-135) written out literally
static int selfplayFeedChecks() {
    std::mt19937_64 Rng(11);
    std::gamma_distribution<double> Gamma(0.15, 1.0);
    for (int Trial = 0; Trial < 64; ++Trial) {
        const std::size_t N = 1 + Rng() % 400;
        std::vector<float> P(N);
        double Sum = 0.0;
        for (auto& X : P) Sum += (X = (float)((Rng() % 1000) + 1));
        for (auto& X : P) X = (float)(X / Sum);
        std::vector<double> Noise(600);
        double NS = 0.0;
        for (auto& X : Noise) NS += (X = Gamma(Rng));
        for (auto& X : Noise) X /= NS;
        const bool Gumbel = Trial & 1, Root = Trial & 2, Full = Trial & 4;
        std::vector<float> Want = P;
        if (!Gumbel && Root && Full)  // frame.cc:121-133
            for (std::size_t I = 0; I < N; ++I) Want[I] = (float)((1 - 0.25) * (double)Want[I] + 0.25 * Noise[I]);
        MockNode Nd;
        Nd.Edges.resize(N);
        std::vector<float> Row = P;
        selfplay::setEvaluationDecoded(&Nd, Row.data(), N, 0.5f, 0.25f, Gumbel, Root, Full, Noise.data());
        for (std::size_t I = 0; I < N; ++I) CHECK(std::memcmp(&Nd.Edges[I].Probability, &Want[I], 4) == 0);
        CHECK(Nd.WinPred == 0.5f && Nd.DrawPred == 0.25f);
        CHECK(selfplay::rowFlags(Gumbel, Root) == ((Gumbel && Root) ? NSB_ROW_SKIP_SOFTMAX : 0));
    }
    return 0;
}


// ---- rules/shogi.h: known-answer perft counts and one hand-made position per special rule ------------------------
static int rulesChecks(int PerftDepth) {
    using namespace b200::rules;
    {
        Position P;
        const unsigned long long Want[6] = {1ull, 30ull, 900ull, 25470ull, 719731ull, 19861490ull};  // public perft values of hirate
        for (int D = 1; D <= PerftDepth && D <= 5; ++D) CHECK(P.perft(D) == Want[D]);
        const uint64_t H0 = P.Hash;
        Move Ms[kMaxMoves];
        const int N = P.generateLegal(Ms);
        CHECK(N == 30);
        for (int I = 0; I < N; ++I) {  // make / unmake restores everything, the hash is incremental == recomputed
            Position Q = P;
            Position::Undo U;
            Q.make(Ms[I], &U);
            const uint64_t Inc = Q.Hash;
            Q.rehash();
            CHECK(Inc == Q.Hash && Q.Hash != H0);
            Q.unmake(Ms[I], U);
            CHECK(Q.Hash == H0 && std::memcmp(Q.Board, P.Board, 81) == 0 && Q.Side == 0 && Q.Ply == 0);
            const int Slot = P.policyIndex(Ms[I]);
            CHECK(Slot >= 0 && Slot < NSB_POLICY_SIZE);
        }
        std::set<int> Slots;                 // distinct policy slots for distinct legal moves
        for (int I = 0; I < N; ++I) Slots.insert(P.policyIndex(Ms[I]));
        CHECK((int)Slots.size() == N);
    }
    auto has = [](Position& P, int From, int To, int Promote) {
        Move Ms[kMaxMoves];
        const int N = P.generateLegal(Ms);
        for (int I = 0; I < N; ++I)
            if (Ms[I].From == From && Ms[I].To == To && Ms[I].Promote == Promote) return true;
        return false;
    };
    auto sq = [](int F, int R) { return 9 * (F - 1) + (R - 1); };
    {   // the position with the most legal moves known in shogi: 593 (sfen R8/2K1S1SSk/4B4/9/9/9/9/9/1L1L1L3 b RBGSNLP3g3n17p 1),
        // the bound NSB_MAX_LEGAL_MOVES comes from; 469 of them are drops, 52 promotions
        Position P;
        P.clear();
        P.put(9, 1, Rook, 0);
        P.put(7, 2, King, 0); P.put(5, 2, Silver, 0); P.put(3, 2, Silver, 0); P.put(2, 2, Silver, 0); P.put(1, 2, King, 1);
        P.put(5, 3, Bishop, 0);
        P.put(8, 9, Lance, 0); P.put(6, 9, Lance, 0); P.put(4, 9, Lance, 0);
        for (int K = 0; K < 7; ++K) P.Hands[0][K] = 1;
        P.Hands[1][0] = 17; P.Hands[1][2] = 3; P.Hands[1][4] = 3;
        P.rehash();
        Move Ms[kMaxMoves];
        const int N = P.generateLegal(Ms);
        int Drops = 0, Promotions = 0;
        std::set<int> Slots;
        for (int I = 0; I < N; ++I) {
            Drops += Ms[I].isDrop();
            Promotions += Ms[I].Promote;
            Slots.insert(P.policyIndex(Ms[I]));
        }
        CHECK(N == NSB_MAX_LEGAL_MOVES && Drops == 469 && Promotions == 52);
        CHECK((int)Slots.size() == N);   // the move-index adaptor gives every legal move its own policy slot
    }
    {   // nifu, drop ranks, mandatory promotion
        Position P;
        P.clear();
        P.put(5, 9, King, 0); P.put(5, 1, King, 1);
        P.put(3, 5, Pawn, 0);                 // black pawn on file 3
        P.put(7, 2, Pawn, 0);                 // a pawn one step from the far rank
        P.put(8, 3, Knight, 0);               // a knight that can only jump to the far rank
        P.Hands[0][0] = 1; P.Hands[0][1] = 1; P.Hands[0][2] = 1;   // pawn, lance, knight in hand
        P.rehash();
        CHECK(!has(P, 81 + 0, sq(3, 4), 0) && has(P, 81 + 0, sq(4, 4), 0));      // nifu on file 3 only
        CHECK(!has(P, 81 + 0, sq(4, 1), 0) && !has(P, 81 + 1, sq(4, 1), 0));     // pawn / lance never on the far rank
        CHECK(has(P, 81 + 1, sq(4, 2), 0));
        CHECK(!has(P, 81 + 2, sq(4, 2), 0) && !has(P, 81 + 2, sq(4, 1), 0) && has(P, 81 + 2, sq(4, 3), 0));  // knight: not on the far two
        CHECK(has(P, sq(7, 2), sq(7, 1), 1) && !has(P, sq(7, 2), sq(7, 1), 0));  // the pawn must promote
        CHECK(has(P, sq(8, 3), sq(7, 1), 1) && !has(P, sq(8, 3), sq(7, 1), 0) && has(P, sq(8, 3), sq(9, 1), 1));
    }
    {   // pawn-drop mate is illegal; the same drop is legal as soon as the king has a flight square
        Position P;
        P.clear();
        P.put(5, 9, King, 0); P.put(1, 1, King, 1);
        P.put(2, 1, Knight, 1); P.put(2, 2, Pawn, 1);   // the king's neighbours are its own pieces
        P.put(1, 9, Lance, 0);                            // guards the whole first file
        P.Hands[0][0] = 1;
        P.rehash();
        CHECK(!has(P, 81 + 0, sq(1, 2), 0));
        CHECK(has(P, 81 + 0, sq(1, 3), 0));               // a pawn drop that is no check is fine
        P.Board[sq(2, 2)] = 0;                            // open a flight square
        P.rehash();
        CHECK(has(P, 81 + 0, sq(1, 2), 0));
    }
    {   // pins and check evasions
        Position P;
        P.clear();
        P.put(5, 9, King, 0); P.put(5, 8, Gold, 0); P.put(5, 1, Rook, 1); P.put(1, 1, King, 1);
        P.rehash();
        CHECK(has(P, sq(5, 8), sq(5, 7), 0) && !has(P, sq(5, 8), sq(4, 8), 0) && !has(P, sq(5, 8), sq(6, 7), 0));   // pinned on the file
        P.Board[sq(5, 8)] = 0;
        P.put(9, 9, Gold, 0);
        P.rehash();
        CHECK(P.inCheck(0));
        Move Ms[kMaxMoves];
        const int N = P.generateLegal(Ms);
        for (int I = 0; I < N; ++I) CHECK(Ms[I].Piece == King && fileOf(Ms[I].To) != 4);   // only king moves off the file
        CHECK(N == 4);                                                                       // 4i, 6i, 4h, 6h ... (5h stays attacked)
    }
    {   // the pin-aware generator against the brute-force definition (make every pseudo-legal move, test the king) on
        // 30,000 positions of random playouts - captures, drops, promotions, checks, pins, perpetuals
        std::mt19937_64 Rng(2024);
        std::size_t Positions = 0, InCheck = 0, WithPins = 0;
        for (int Game = 0; Game < 300; ++Game) {
            Position P;
            for (int Ply = 0; Ply < 100; ++Ply) {
                Move Fast[kMaxMoves], Pseudo[kMaxMoves + 64];
                const int NF = P.generateLegal(Fast);
                const int NP = P.generatePseudoLegal(Pseudo);
                int NS = 0;
                const int Me = P.Side;
                for (int I = 0; I < NP; ++I) {
                    Position::Undo U;
                    P.make(Pseudo[I], &U);
                    bool Ok = !P.inCheck(Me);
                    if (Ok && Pseudo[I].isDrop() && Pseudo[I].Piece == Pawn && P.inCheck(Me ^ 1)) Ok = P.hasLegalMove();
                    P.unmake(Pseudo[I], U);
                    if (Ok) {
                        CHECK(NS < NF && Fast[NS] == Pseudo[I]);   // same moves in the same order
                        ++NS;
                    }
                }
                CHECK(NS == NF);
                ++Positions;
                const Position::KingSafety KS = P.kingSafety(Me);
                InCheck += KS.InCheck;
                WithPins += KS.Pinned.any();
                if (NF == 0) break;
                Position::Undo U;
                P.make(Fast[Rng() % (uint64_t)NF], &U);
            }
        }
        std::printf("rules: %zu random-playout positions, %zu in check, %zu with pinned pieces: generators agree\n", Positions, InCheck, WithPins);
        CHECK(Positions >= 20000 && InCheck > 300 && WithPins > 300);
    }
    {   // declaration win, 27-point rule: king in the camp, not in check, 10 other pieces there, 28 points (black) / 27 (white)
        Position P;
        P.clear();
        P.put(5, 2, King, 0); P.put(5, 9, King, 1);
        P.put(9, 1, ProRook, 0); P.put(8, 1, Bishop, 0);                       // 5 + 5
        for (int F = 1; F <= 7; ++F) P.put(F, 3, ProPawn, 0);                  // 7
        P.put(1, 1, Gold, 0);                                                  // 1: 10 pieces, 18 points on the board
        P.Hands[0][0] = 5; P.Hands[0][6] = 1;                                  // 5 pawns + a rook in hand: 10
        P.rehash();
        CHECK(P.canDeclare());
        P.Hands[0][0] = 4;
        CHECK(!P.canDeclare());                                                // 27 points are not enough for black
        P.Hands[0][0] = 5;
        P.Board[sq(1, 1)] = 0;
        P.Hands[0][4] = 1;                                                     // the gold in hand instead: 28 points, 9 pieces
        CHECK(!P.canDeclare());
        P.Hands[0][4] = 0;
        P.put(1, 1, Gold, 0);
        P.put(5, 1, Gold, 1);                                                  // a white gold in front of the king
        CHECK(P.inCheck(0) && !P.canDeclare());                                // not while in check
        P.Board[sq(5, 1)] = 0;
        P.Board[sq(5, 2)] = 0; P.put(5, 4, King, 0);
        CHECK(!P.canDeclare());                                                // the king must stand inside the camp
        // white: the mirrored position needs 27
        Position W;
        W.clear();
        W.put(5, 8, King, 1); W.put(5, 1, King, 0);
        W.put(1, 9, ProRook, 1); W.put(2, 9, Bishop, 1);
        for (int F = 3; F <= 9; ++F) W.put(F, 7, ProPawn, 1);
        W.put(9, 9, Gold, 1);
        W.Hands[1][0] = 4; W.Hands[1][6] = 1;                                  // 18 + 9 = 27
        W.Side = 1;
        W.rehash();
        CHECK(W.canDeclare());
        W.Hands[1][0] = 3;
        CHECK(!W.canDeclare());
        W.Hands[1][0] = 4;
        W.Side = 0;
        CHECK(!W.canDeclare());                                                // black to move: black's king is not in white's camp
    }
    {   // four-fold repetition: a draw, but lost by a side that checked with every move of the cycle
        using b200::game::Repetition;
        using b200::game::repetitionStatus;
        Position P;
        P.clear();
        P.put(5, 9, King, 0); P.put(1, 1, King, 1); P.put(1, 5, Rook, 0); P.put(9, 9, Gold, 1);
        P.Side = 1;                                   // white to move, in check from the rook on its file
        P.Ply = 1;
        P.rehash();
        std::vector<uint64_t> History{0x1234u, P.Hash};                       // (ply 0: anything)
        std::vector<uint8_t> InCheck{0, (uint8_t)P.inCheck(1)};
        CHECK(InCheck[1] == 1);
        auto play = [&](int From, int To) {
            Move Ms[kMaxMoves];
            const int N = P.generateLegal(Ms);
            for (int I = 0; I < N; ++I)
                if (Ms[I].From == From && Ms[I].To == To && !Ms[I].Promote) {
                    Position::Undo U;
                    P.make(Ms[I], &U);
                    History.push_back(P.Hash);
                    InCheck.push_back(P.inCheck(P.Side) ? 1 : 0);
                    return true;
                }
            return false;
        };
        Repetition R = Repetition::None;
        for (int Cycle = 0; Cycle < 3; ++Cycle) {      // king 1a-2a, rook 1e-2e+, king 2a-1a, rook 2e-1e+ : black checks every time
            CHECK(R == Repetition::None);
            CHECK(play(sq(1, 1), sq(2, 1)) && repetitionStatus(History, InCheck) == Repetition::None);
            CHECK(play(sq(1, 5), sq(2, 5)) && repetitionStatus(History, InCheck) == Repetition::None);
            CHECK(play(sq(2, 1), sq(1, 1)) && repetitionStatus(History, InCheck) == Repetition::None);
            CHECK(play(sq(2, 5), sq(1, 5)));
            R = repetitionStatus(History, InCheck);
        }
        CHECK(R == Repetition::BlackLoses);            // the fourth occurrence of the start of the cycle
        // the same cycle with one quiet move in it is an ordinary draw
        std::vector<uint8_t> Quiet = InCheck;
        Quiet[5] = 0;
        CHECK(repetitionStatus(History, Quiet) == Repetition::Draw);
        // and with the colours swapped (white's replies are the checks) white loses
        std::vector<uint64_t> H2{1, 2, 3, 4, 5, 2, 3, 4, 5, 2, 3, 4, 5, 2};  // position "2" (ply 1, 5, 9, 13)
        std::vector<uint8_t> C2(H2.size(), 0);
        for (std::size_t I = 2; I < H2.size(); I += 2) C2[I] = 1;              // black to move at even plies, always in check
        CHECK(repetitionStatus(H2, C2) == Repetition::WhiteLoses);
        C2.assign(H2.size(), 0);
        CHECK(repetitionStatus(H2, C2) == Repetition::Draw);
        H2.pop_back();
        CHECK(repetitionStatus(H2, C2) == Repetition::None);
    }
    {   // shallow mate search against the exhaustive definition on random-playout positions
        std::function<bool(Position&, int)> brute = [&](Position& P, int Plies) -> bool {   // side to move mates within Plies (odd)
            Move Ms[kMaxMoves];
            const int N = P.generateLegal(Ms);
            const int Me = P.Side;
            for (int I = 0; I < N; ++I) {
                Position::Undo U;
                P.make(Ms[I], &U);
                bool Wins = false;
                if (P.inCheck(Me ^ 1)) {
                    Move Ev[kMaxMoves];
                    const int NE = P.generateLegal(Ev);
                    if (NE == 0) Wins = true;
                    else if (Plies >= 3) {
                        Wins = true;
                        for (int J = 0; J < NE && Wins; ++J) {
                            Position::Undo V;
                            P.make(Ev[J], &V);
                            Wins = brute(P, Plies - 2);
                            P.unmake(Ev[J], V);
                        }
                    }
                }
                P.unmake(Ms[I], U);
                if (Wins) return true;
            }
            return false;
        };
        std::mt19937_64 Rng(31);
        std::size_t Checked = 0, Mates1 = 0, Mates3 = 0;
        for (int Game = 0; Game < 400 && Checked < 12000; ++Game) {
            Position P;
            for (int Ply = 0; Ply < 160; ++Ply) {
                Move Ms[kMaxMoves];
                const int N = P.generateLegal(Ms);
                if (N == 0) break;
                if (Ply >= 20) {
                    Position Q = P;
                    const bool B1 = brute(Q, 1), B3 = brute(Q, 3);
                    Move M1, M3;
                    const bool F1 = Q.mateIn1(&M1), F3 = Q.mateIn3(&M3);
                    CHECK(B1 == F1 && B3 == F3 && (!F1 || F3));
                    CHECK(Q.Hash == P.Hash && std::memcmp(Q.Board, P.Board, 81) == 0);   // the search leaves the position as it was
                    if (F1) {   // the move it reports mates
                        Position::Undo U;
                        Q.make(M1, &U);
                        CHECK(Q.inCheck(Q.Side) && !Q.hasLegalMove());
                        Q.unmake(M1, U);
                    }
                    Mates1 += F1;
                    Mates3 += F3;
                    ++Checked;
                }
                Position::Undo U;
                P.make(Ms[Rng() % (uint64_t)N], &U);
            }
        }
        std::printf("rules: mate search == exhaustive search on %zu random-playout positions (%zu mates in 1, %zu within 3 plies)\n", Checked, Mates1, Mates3);
        CHECK(Checked >= 5000 && Mates1 > 20 && Mates3 > Mates1);
    }
    {   // promoted sliders keep sliding and gain the king's other steps; captures go to the hand unpromoted
        Position P;
        P.clear();
        P.put(5, 9, King, 0); P.put(1, 1, King, 1); P.put(5, 5, ProBishop, 0); P.put(7, 3, ProRook, 1);
        P.rehash();
        CHECK(has(P, sq(5, 5), sq(5, 4), 0) && has(P, sq(5, 5), sq(1, 9), 0) && has(P, sq(5, 5), sq(7, 3), 0) && !has(P, sq(5, 5), sq(9, 1), 0));
        Move Ms[kMaxMoves];
        const int N = P.generateLegal(Ms);
        for (int I = 0; I < N; ++I)
            if (Ms[I].From == sq(5, 5) && Ms[I].To == sq(7, 3)) {
                Position::Undo U;
                P.make(Ms[I], &U);
                CHECK(P.Hands[0][6] == 1 && P.Board[sq(7, 3)] == Position::code(ProBishop, 0));   // a ROOK in hand
                P.unmake(Ms[I], U);
                CHECK(P.Hands[0][6] == 0 && P.Board[sq(7, 3)] == Position::code(ProRook, 1));
            }
    }
    return 0;
}

// ---- selfplay_game.h + mcts_search.h: whole games against a mock evaluator (no GPU) -----------------------------------
static int selfplayMock(int Games, int Playouts, int MatePlies = 0) {
    using namespace b200;
    game::GameOptions O;
    O.Playouts = Playouts;
    O.FullSearchRatio = 0.5;
    O.MatePlies = MatePlies;
    game::Info SI;
    game::Frame F;
    F.MT.seed(42);
    game::newGame(O, F);
    F.MaxPly = 60;
    game::prepareRoot(O, F);
    uint64_t Evals = 0;
    std::size_t SavedRecords = 0;
    bool SaveOk = true;
    while ((int)SI.Games.load() < Games) {
        const uint64_t GamesBefore = SI.Games.load(), RecordsBefore = SI.Records.load();
        const uint32_t RootVisitsBefore = F.Tree.node(0).Visits;
        game::advance(O, F, &SI, [&](const game::Frame& Done) {   // SelfplayPhase::Save: the finished game as teacher records
            const teacher::FinishedGame G = game::finishedGame(Done);
            std::ostringstream Out;
            teacher::writeHeader(Out);
            const std::size_t N = teacher::saveGame(&Out, G);
            SavedRecords += N;
            const std::string Bytes = Out.str();
            SaveOk = SaveOk && Bytes.size() == 12 + N * sizeof(teacher::TeacherRecord) && G.Moves.size() == G.DidFullSearch.size();
            std::size_t Full = 0;
            for (uint8_t B : G.DidFullSearch) Full += B;
            SaveOk = SaveOk && Full == N && G.Winner <= teacher::WinnerNone;
            rules::Position Replay;   // every record holds the position BEFORE its move, and the move is legal there
            std::size_t K = 0;
            for (std::size_t Ply = 0; Ply < G.Moves.size(); ++Ply) {
                rules::Move Ms[rules::kMaxMoves];
                const int NM = Replay.generateLegal(Ms);
                bool Legal = false;
                for (int J = 0; J < NM; ++J) Legal = Legal || Ms[J] == G.Moves[Ply];
                SaveOk = SaveOk && Legal;
                if (G.DidFullSearch[Ply]) {
                    teacher::TeacherRecord R;
                    std::memcpy(&R, Bytes.data() + 12 + K * sizeof R, sizeof R);
                    nsb_position Want;
                    Replay.toRecord(&Want, G.MaxPly, G.BlackDraw, G.WhiteDraw);
                    SaveOk = SaveOk && std::memcmp(&R.Position, &Want, sizeof Want) == 0 && R.From == G.Moves[Ply].From &&
                             R.To == G.Moves[Ply].To && R.Promote == G.Moves[Ply].Promote && R.Winner == G.Winner;
                    ++K;
                }
                rules::Position::Undo U;
                Replay.make(G.Moves[Ply], &U);
            }
            SaveOk = SaveOk && K == N && Replay.Hash == Done.Root.Hash;
        });
        if (SI.Games.load() != GamesBefore) F.MaxPly = 60;  // (newGame drew a long one)
        if (SI.Records.load() == RecordsBefore && F.LeafNode != 0) CHECK(F.Tree.node(0).Visits >= RootVisitsBefore);
        // the frame now waits for an evaluation of F.Leaf with F.NumLeafMoves legal moves
        CHECK(F.NumLeafMoves >= 1 && F.NumLeafMoves <= rules::kMaxMoves && F.LeafNode >= 0);
        rules::Move Check[rules::kMaxMoves];
        rules::Position L = F.Leaf;
        CHECK(L.generateLegal(Check) == F.NumLeafMoves);
        std::vector<float> Row((size_t)F.NumLeafMoves);
        double Sum = 0.0;
        uint64_t H = F.Leaf.Hash;
        for (int J = 0; J < F.NumLeafMoves; ++J) {
            CHECK(F.LeafSlots[J] < NSB_POLICY_SIZE);
            H = H * 6364136223846793005ull + F.LeafSlots[J];
            Sum += (Row[(size_t)J] = 1.0f + (float)((H >> 40) % 1000) / 250.0f);
        }
        for (auto& X : Row) X = (float)(X / Sum);
        std::vector<uint16_t> Order((size_t)F.NumLeafMoves);
        for (int J = 0; J < F.NumLeafMoves; ++J) Order[(size_t)J] = (uint16_t)J;
        std::stable_sort(Order.begin(), Order.end(), [&](uint16_t A, uint16_t B) { return Row[A] > Row[B]; });
        const float Win = 0.3f + 0.4f * (float)((F.Leaf.Hash >> 20) % 1000) / 1000.0f, Draw = 0.05f;
        game::applyEvaluation(O, F, Row.data(), Order.data(), Win, Draw);
        ++Evals;
        // tree invariants after the back-propagation
        const search::Node& R = F.Tree.node(0);
        uint64_t ChildVisits = 0;
        for (int I = 0; I < R.NumEdges; ++I) {
            const search::Edge& E = F.Tree.edgesOf(0)[I];
            if (I > 0) CHECK(E.P <= F.Tree.edgesOf(0)[I - 1].P);   // sorted by prior
            if (E.Child >= 0) {
                CHECK(E.CVisits == F.Tree.node(E.Child).Visits && E.CVirtualLoss == 0);   // the edge mirrors its child
                ChildVisits += F.Tree.node(E.Child).Visits;
            }
        }
        CHECK(R.evaluated() && R.Visits == ChildVisits + 1 && R.VirtualLoss == 0);
        CHECK(R.WinAcc >= 0.0 && R.WinAcc <= (double)R.Visits);
        CHECK(R.Visits <= (uint32_t)Playouts + 1);
    }
    std::printf("selfplay mock: %d games, %llu records, %llu evaluations, %llu terminal leaves; ended by mate %llu / repetition %llu / max ply %llu\n",
                Games, (unsigned long long)SI.Records.load(), (unsigned long long)Evals, (unsigned long long)SI.Terminals.load(),
                (unsigned long long)SI.Mates.load(), (unsigned long long)SI.Repetitions.load(), (unsigned long long)SI.MaxPlies.load());
    CHECK(SI.Records.load() >= (uint64_t)Games * 10 && Evals > SI.Records.load());
    CHECK(SaveOk && SavedRecords > 0 && SavedRecords < SI.Records.load());   // teacher records: full-search plies only
    return 0;
}

// ---- mcts_search.h shared tree: `--tree-stress THREADS NODES` (also built with -fsanitize=thread by the tests) ------
// Search threads descend, claim, expand, "evaluate" (a hash of the position), publish and back-propagate concurrently on
// one lock-free tree of real shogi positions.  Afterwards: no virtual loss is left anywhere, the root's visits are its
// children's + 1, every edge mirrors its child exactly, every evaluated node has sorted priors.
static int treeStress(int Threads, std::size_t Nodes) {
    using namespace b200;
    const std::size_t Cap = Nodes + Nodes / 2 + 64 * (std::size_t)Threads + 1024;  // (terminal nodes and children created twice take slots too)
    search::Tree T(Cap, Cap * 64);
    const rules::Position Root;
    const std::vector<uint64_t> History{Root.Hash};
    std::atomic<uint64_t> Done{0}, Collisions{0};
    std::atomic<bool> Full{false};
    std::vector<std::thread> Th;
    for (int Id = 0; Id < Threads; ++Id)
        Th.emplace_back([&]() {
            rules::Move Moves[rules::kMaxMoves];
            std::vector<uint64_t> Path;
            std::vector<int> Trail;
            std::vector<float> Row(rules::kMaxMoves);
            std::vector<uint16_t> Order(rules::kMaxMoves);
            while (Done.load(std::memory_order_relaxed) < Nodes && !Full.load(std::memory_order_relaxed)) {
                rules::Position Pos = Root;
                Path.clear();
                const int Node = T.selectLeaf(Pos, 0.5f, 0.5f, &Path, &Trail);
                if (Node == search::Tree::OutOfMemory) { Full.store(true); break; }
                if (Node < 0) { Collisions.fetch_add(1, std::memory_order_relaxed); std::this_thread::yield(); continue; }
                if (T.node(Node).Term != search::Open) { T.backup(Node, T.node(Node).Term == search::Mated ? 0.f : 0.5f, T.node(Node).Term == search::Mated ? 0.f : 1.f); continue; }
                const int N = Pos.generateLegal(Moves);
                if (N == 0 || (Node != 0 && (search::isFourfold(Pos.Hash, History, Path) || Pos.Ply >= 60))) {
                    T.setTerminal(Node, N == 0 ? search::Mated : search::DrawnGame);
                    T.backup(Node, N == 0 ? 0.f : 0.5f, N == 0 ? 0.f : 1.f);
                    continue;
                }
                if (!T.expand(Node, Moves, N)) { Full.store(true); break; }
                double Sum = 0.0;
                uint64_t H = Pos.Hash;
                for (int J = 0; J < N; ++J) {
                    H = H * 6364136223846793005ull + (uint64_t)J;
                    Sum += (Row[(size_t)J] = 1.0f + (float)((H >> 40) % 1000) / 250.0f);
                    Order[(size_t)J] = (uint16_t)J;
                }
                for (int J = 0; J < N; ++J) Row[(size_t)J] = (float)(Row[(size_t)J] / Sum);
                std::stable_sort(Order.begin(), Order.begin() + N, [&](uint16_t A, uint16_t B) { return Row[A] > Row[B]; });
                T.setPriors(Node, Row.data(), Order.data());
                T.backup(Node, 0.3f + 0.4f * (float)((Pos.Hash >> 20) % 1000) / 1000.0f, 0.05f);
                Done.fetch_add(1, std::memory_order_relaxed);
            }
        });
    for (auto& X : Th) X.join();
    bool Ok = !Full.load();
    uint64_t Evaluated = 0, ChildVisits = 0;
    for (std::size_t I = 0; I < T.numNodes() && Ok; ++I) {
        const search::Node& N = T.node((int)I);
        Ok = Ok && N.VirtualLoss == 0;
        if (!N.evaluated()) { Ok = Ok && N.Visits == 0; continue; }   // a child another thread created first and nobody used
        ++Evaluated;
        uint64_t Sum = 0;
        for (int J = 0; J < N.NumEdges && Ok; ++J) {
            const search::Edge& E = T.edgesOf((int)I)[J];
            if (J > 0) Ok = Ok && E.P <= T.edgesOf((int)I)[J - 1].P;
            if (E.Child >= 0) {
                const search::Node& C = T.node(E.Child);
                Ok = Ok && C.Parent == (int)I && E.CVisits == C.Visits && E.CVirtualLoss == 0 && std::fabs((double)E.CWinAcc - C.WinAcc) < 1e-2 * (1.0 + C.WinAcc);
                Sum += C.Visits;
            } else {
                Ok = Ok && E.CVisits == 0;
            }
        }
        if (N.Term == search::Open) Ok = Ok && N.Visits == Sum + 1;
        if (I == 0) ChildVisits = Sum;
    }
    std::printf("tree stress: %llu evaluations by %d threads, %zu nodes (%llu evaluated), root visits %u = %llu + 1, %llu collisions: %s\n",
                (unsigned long long)Done.load(), Threads, T.numNodes(), (unsigned long long)Evaluated, T.node(0).Visits,
                (unsigned long long)ChildVisits, (unsigned long long)Collisions.load(), Ok ? "ok" : "FAIL");
    return Ok ? 0 : 1;
}

// ---- evaluation_worker_b200.h: the pipelined evaluation worker inside the worker::Worker contract, on plain memory ----
// `--worker-cycles TASKS`: built against this repo's Worker stand-in (host/shim/worker) by the Makefile and against the
// REFERENCE's own src/worker/worker.{h,cc} by tests/test_host_cpp.py.  Three start() / stop() / await() cycles; after
// every await() each task pushed so far has been filled once and delivered once (a stopped worker is a drained worker).
struct CycleTask {
    uint32_t Id = 0;
    std::atomic<int> Filled{0}, Delivered{0};
    bool RowOk = true;
};
struct CycleClient : evaluate::EvaluationClient<FakePipeline::Slot> {
    std::mutex M;
    std::deque<CycleTask*> Queue;
    std::atomic<uint64_t> Released{0};
    void push(CycleTask* T) {
        std::lock_guard<std::mutex> L(M);
        Queue.push_back(T);
    }
    void take(std::size_t Max, bool Wait, std::vector<void*>& Out) override {
        {
            std::lock_guard<std::mutex> L(M);
            while (!Queue.empty() && Out.size() < Max) {
                Out.push_back(Queue.front());
                Queue.pop_front();
            }
        }
        if (Out.empty() && Wait) std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    uint32_t fill(void* Task, FakePipeline::Slot& S, std::size_t Row, uint32_t MoveBegin) override {
        CycleTask* T = static_cast<CycleTask*>(Task);
        const uint32_t N = 1 + T->Id % 5;
        for (uint32_t J = 0; J < N; ++J) S.MoveIndices[MoveBegin + J] = (uint16_t)(T->Id + J);
        S.Hashes[Row] = T->Id * 2654435761ull;
        T->Filled.fetch_add(1);
        return N;
    }
    void deliver(void* Task, FakePipeline::Slot& S, std::size_t Row) override {
        CycleTask* T = static_cast<CycleTask*>(Task);
        const uint32_t B = S.MoveOffsets[Row], E = S.MoveOffsets[Row + 1];
        T->RowOk = S.Hashes[Row] == T->Id * 2654435761ull && E - B == 1 + T->Id % 5 && S.MoveIndices[B] == (uint16_t)T->Id;
        T->Delivered.fetch_add(1);
    }
    void release(std::vector<void*>& Tasks) override {
        Released.fetch_add(Tasks.size());
        Tasks.clear();
    }
};

static int workerCycles(std::size_t Tasks, std::size_t MinFill = 1) {
    FakePipeline Pipe(3, 64);
    CycleClient Client;
    std::vector<CycleTask> Pool(Tasks * 3);
    for (std::size_t I = 0; I < Pool.size(); ++I) Pool[I].Id = (uint32_t)I;
    int Bound = 0;
    {
        evaluate::PipelinedEvaluationWorker<FakePipeline> W(&Pipe, &Client, true, 0, false, true,
                                                            [](void* P) { ++*static_cast<int*>(P); }, &Bound, MinFill);
        CHECK(Bound == 1 && !W.isRunning());   // initializationTask ran on the worker thread before spawnThread returned
        std::size_t Pushed = 0;
        for (int Cycle = 0; Cycle < 3; ++Cycle) {
            W.start();
            std::thread Producer([&]() {
                for (std::size_t I = 0; I < Tasks; ++I) {
                    Client.push(&Pool[Pushed + I]);
                    if ((I & 255) == 255) std::this_thread::yield();
                }
            });
            Producer.join();
            Pushed += Tasks;
            while (Client.Released.load() < Pushed) std::this_thread::yield();
            W.stop();
            W.await();
            CHECK(!W.isRunning());
            for (std::size_t I = 0; I < Pushed; ++I) CHECK(Pool[I].Filled.load() == 1 && Pool[I].Delivered.load() == 1 && Pool[I].RowOk);
            for (std::size_t I = Pushed; I < Pool.size(); ++I) CHECK(Pool[I].Filled.load() == 0);
            CHECK(W.rows() == Pushed);
        }
        std::printf("worker cycles (min fill %zu): %zu tasks in %llu batches over 3 start/stop/await cycles: ok\n", MinFill, Pushed, (unsigned long long)W.batches());
    }   // ~Worker joins the thread
    return 0;
}

// ---- the whole self-play harness (host/selfplay_workers.h: search workers, pipelined evaluation worker, save worker;
//      start-up and wind-down exactly as host/selfplay_real.cc runs them) against a mock pipeline whose collect() invents
//      the evaluations: `--selfplay-loop WORKERS FRAMES MILLISECONDS`.  The program must END - a worker::Worker is only
//      stopped while it reports idle - and every frame must be back in one of the two queues afterwards.
struct MockEvalPipeline {
    struct Slot {
        std::vector<nsb_position> PositionsV;
        std::vector<uint32_t> MoveOffsetsV;
        std::vector<uint16_t> MoveIndicesV, OrderV;
        std::vector<uint64_t> HashesV;
        std::vector<uint8_t> RowFlagsV, NanFlagV, HitFlagV;
        std::vector<float> LegalV, WinRateV, DrawRateV;
        nsb_position* Positions;
        uint32_t* MoveOffsets;
        uint16_t *MoveIndices, *Order;
        uint64_t* Hashes;
        uint8_t *RowFlags, *NanFlag, *HitFlag;
        float *Legal, *WinRate, *DrawRate;
        std::size_t Count = 0;
    };
    std::vector<Slot> Slots;
    std::size_t BatchMax, Next = 0;
    MockEvalPipeline(std::size_t NumSlots, std::size_t B) : Slots(NumSlots), BatchMax(B) {
        for (auto& S : Slots) {
            S.PositionsV.resize(B); S.MoveOffsetsV.resize(B + 1); S.MoveIndicesV.resize(B * NSB_MAX_LEGAL_MOVES);
            S.OrderV.resize(B * NSB_MAX_LEGAL_MOVES); S.HashesV.resize(B); S.RowFlagsV.resize(B); S.NanFlagV.assign(B, 0);
            S.HitFlagV.assign(B, 0); S.LegalV.resize(B * NSB_MAX_LEGAL_MOVES); S.WinRateV.resize(B); S.DrawRateV.resize(B);
            S.Positions = S.PositionsV.data(); S.MoveOffsets = S.MoveOffsetsV.data(); S.MoveIndices = S.MoveIndicesV.data();
            S.Order = S.OrderV.data(); S.Hashes = S.HashesV.data(); S.RowFlags = S.RowFlagsV.data(); S.NanFlag = S.NanFlagV.data();
            S.HitFlag = S.HitFlagV.data(); S.Legal = S.LegalV.data(); S.WinRate = S.WinRateV.data(); S.DrawRate = S.DrawRateV.data();
        }
    }
    std::size_t numSlots() const { return Slots.size(); }
    std::size_t batchMax() const { return BatchMax; }
    Slot& acquire(std::size_t* K) { *K = Next; Next = (Next + 1) % Slots.size(); return Slots[*K]; }
    void submit(std::size_t K, std::size_t Rows, bool, int, bool, bool) { Slots[K].Count = Rows; Polls = 0; }
    bool ready(std::size_t) { return ++Polls > 2; }   // (a batch is "done" on the third look: the worker's help() path runs)
    std::size_t Polls = 0;
    Slot& collect(std::size_t K) {
        Slot& S = Slots[K];
        for (std::size_t R = 0; R < S.Count; ++R) {
            const uint32_t B = S.MoveOffsets[R], N = S.MoveOffsets[R + 1] - B;
            uint64_t H = S.Hashes[R];
            double Sum = 0.0;
            for (uint32_t J = 0; J < N; ++J) {
                H = H * 6364136223846793005ull + S.MoveIndices[B + J];
                Sum += (S.Legal[B + J] = 1.0f + (float)((H >> 40) % 1000) / 250.0f);
            }
            for (uint32_t J = 0; J < N; ++J) {
                S.Legal[B + J] = (float)(S.Legal[B + J] / Sum);
                S.Order[B + J] = (uint16_t)J;
            }
            std::stable_sort(S.Order + B, S.Order + B + N, [&](uint16_t X, uint16_t Y) { return S.Legal[B + X] > S.Legal[B + Y]; });
            S.WinRate[R] = 0.3f + 0.4f * (float)((S.Hashes[R] >> 20) % 1000) / 1000.0f;
            S.DrawRate[R] = 0.05f;
        }
        return S;
    }
};

static int selfplayLoop(int Workers, std::size_t Frames, int Milliseconds) {
    using namespace b200::game;
    HarnessOptions O;
    O.Playouts = 12;
    O.FullSearchRatio = 0.5;
    std::vector<Frame> Pool(Frames);
    FrameQueue SearchQueue, EvaluationQueue;
    Info SI;
    std::vector<Frame*> Init;
    for (std::size_t I = 0; I < Pool.size(); ++I) {
        Pool[I].MT.seed(77 * (I + 1));
        newGame(O, Pool[I]);
        Pool[I].MaxPly = 40;   // short games: the save path runs too
        prepareRoot(O, Pool[I]);
        Init.push_back(&Pool[I]);
    }
    SearchQueue.add(Init);
    SaveQueue Saves;
    SaveStats Saved;
    std::atomic<bool> Saving{true};
    std::thread Saver(saveWorker, std::cref(O), &Saves, &Saved, &Saving);
    MockEvalPipeline Pipe(3, 32);
    std::atomic<bool> Closing{false};
    FrameClient<MockEvalPipeline::Slot> Client(O, &EvaluationQueue, &SearchQueue, &SI, &Saves, &Closing);
    {
        evaluate::PipelinedEvaluationWorker<MockEvalPipeline> Evaluation(&Pipe, &Client, true, 2, false, true);
        std::vector<std::unique_ptr<SearchWorker>> Searchers;
        for (int W = 0; W < Workers; ++W)
            Searchers.push_back(std::make_unique<SearchWorker>(O, &SearchQueue, &EvaluationQueue, &Saves, &SI, &Closing));
        Evaluation.start();
        for (auto& W : Searchers) W->start();
        std::this_thread::sleep_for(std::chrono::milliseconds(Milliseconds));
        Closing.store(true);
        for (auto& W : Searchers) W->stop();
        for (auto& W : Searchers) W->await();
        Evaluation.stop();
        Evaluation.await();   // drained: nothing queued for evaluation, nothing in flight
        CHECK(Evaluation.rows() == SI.Evals.load());
        std::printf("evaluation worker: %.2f us per row filling, %.2f delivering; %.0f %% of its time waiting for frames\n",
                    1e6 * Evaluation.secondsFilling() / (double)Evaluation.rows(), 1e6 * Evaluation.secondsDelivering() / (double)Evaluation.rows(),
                    100.0 * Evaluation.secondsTaking() / (Evaluation.secondsTaking() + Evaluation.secondsFilling() + Evaluation.secondsDelivering() + Evaluation.secondsCollecting()));
    }
    Saving.store(false);
    Saves.close();
    Saver.join();
    // every frame is in the search queue again (the evaluation queue is empty), none lost, none twice
    std::vector<Frame*> Left, LeftEval;
    SearchQueue.get(Frames + 1, false, Left);
    EvaluationQueue.get(Frames + 1, false, LeftEval);
    CHECK(LeftEval.empty() && Left.size() == Frames);
    std::set<Frame*> Distinct(Left.begin(), Left.end());
    CHECK(Distinct.size() == Frames);
    CHECK(SI.Evals.load() > Frames && SI.Records.load() > 0);
    CHECK(Saved.Games.load() == SI.Games.load() && Saves.drained());
    std::printf("selfplay loop: %d search workers, %zu frames, %d ms: %llu evals in %llu batches, %llu positions, %llu games saved "
                "(%llu records): wound down, all frames accounted for: ok\n",
                Workers, Frames, Milliseconds, (unsigned long long)SI.Evals.load(), (unsigned long long)SI.Batches.load(),
                (unsigned long long)SI.Records.load(), (unsigned long long)Saved.Games.load(), (unsigned long long)Saved.Records.load());
    return 0;
}

// ---- host cost of one self-play leaf, single thread, no GPU: `--host-cost FRAMES LEAVES`.  The three things a leaf costs
//      (`--host-cost FRAMES LEAVES WARMUP_LEAVES` measures after the games have left the opening.)  The host in nsb_selfplay_real - advance() on a search worker, fill() and deliver() on the evaluation worker - timed
//      around a free mock evaluation over a pool of games as large as the harness's (cache footprint included).
static int hostCost(std::size_t Frames, std::size_t Leaves, std::size_t Warmup) {
    using namespace b200::game;
    using Clk = std::chrono::steady_clock;
    HarnessOptions O;
    std::vector<Frame> Pool(Frames);
    Info SI;
    for (std::size_t I = 0; I < Pool.size(); ++I) {
        Pool[I].MT.seed(77 * (I + 1));
        newGame(O, Pool[I]);
        prepareRoot(O, Pool[I]);
    }
    // the search worker's pattern (selfplay_workers.h): frames in groups of 32; per
    // frame advance() - which first applies the evaluation staged on the last visit - and the row hand-over
    std::vector<float> Row(NSB_MAX_LEGAL_MOVES);
    std::vector<uint16_t> Order(NSB_MAX_LEGAL_MOVES), Slots(NSB_MAX_LEGAL_MOVES);
    nsb_position Rec;
    double TSearch = 0, TFill = 0;
    uint64_t Moves = 0, Sink = 0;
    constexpr std::size_t Group = 32;
    for (std::size_t L = 0; L < Leaves + Warmup; L += Group) {
        if (L >= Warmup && L < Warmup + Group) {  // (the first plies of every game are opening positions: few drops, no checks)
            TSearch = TFill = 0;
            Moves = 0;
        }
        const std::size_t First = L % Frames;
        const auto T0 = Clk::now();
        for (std::size_t G = 0; G < Group; ++G) advance(O, Pool[(First + G) % Frames], &SI);
        const auto T1 = Clk::now();
        TSearch += std::chrono::duration<double>(T1 - T0).count();
        for (std::size_t G = 0; G < Group; ++G) {
            Frame& F = Pool[(First + G) % Frames];
            const auto T2 = Clk::now();
            F.Leaf.toRecord(&Rec, F.MaxPly, F.BlackDraw, F.WhiteDraw);
            std::memcpy(Slots.data(), F.LeafSlots, (std::size_t)F.NumLeafMoves * sizeof(uint16_t));
            Sink += Rec.board[40] + Slots[0];
            const int N = F.NumLeafMoves;
            uint64_t H = F.Leaf.Hash;
            float Sum = 0.f;
            for (int J = 0; J < N; ++J) {
                H = H * 6364136223846793005ull + 1442695040888963407ull;
                Sum += (Row[(size_t)J] = 1.0f + (float)(H >> 54) / 256.0f);
            }
            for (int J = 0; J < N; ++J) {
                Row[(size_t)J] /= Sum;
                Order[(size_t)J] = (uint16_t)J;
            }
            const auto T3 = Clk::now();   // (the rank order is the executor's work)
            std::stable_sort(Order.begin(), Order.begin() + N, [&](uint16_t A, uint16_t B) { return Row[A] > Row[B]; });
            const auto T4 = Clk::now();
            stageEvaluation(F, Row.data(), Order.data(), 0.3f + 0.4f * (float)((F.Leaf.Hash >> 20) % 1000) / 1000.0f, 0.05f);
            TFill += std::chrono::duration<double>(T3 - T2).count() + std::chrono::duration<double>(Clk::now() - T4).count();
            Moves += (uint64_t)N;
        }
    }
    std::printf("{\"host_cost_us_per_leaf\": {\"search_worker\": %.3f, \"evaluation_worker_incl_mock_network\": %.3f}, "
                "\"frames\": %zu, \"leaves\": %zu, \"avg_legal_moves\": %.1f, \"positions_played\": %llu, \"games\": %llu, "
                "\"terminals\": %llu, \"sink\": %llu}\n",
                1e6 * TSearch / (double)Leaves, 1e6 * TFill / (double)Leaves, Frames, Leaves,
                (double)Moves / (double)Leaves, (unsigned long long)SI.Records.load(), (unsigned long long)SI.Games.load(),
                (unsigned long long)SI.Terminals.load(), (unsigned long long)(Sink & 1));
#ifdef NSB_PHASE_TIMING
    const PhaseCycles& PC = phaseCycles();   // (includes the warm-up leaves)
    const double Per = 1.0 / (double)(Leaves + Warmup);
    std::printf("cycles per leaf: select %.0f, generate %.0f, terminal checks %.0f, expand %.0f, policy slots %.0f, transition + prepareRoot %.0f, apply staged evaluation %.0f\n",
                PC.Select * Per, PC.Generate * Per, PC.Terminal * Per, PC.Expand * Per, PC.Slots * Per, PC.Root * Per, PC.Apply * Per);
#endif
    return 0;
}

// ---- `--movegen-bench`: ns per generateLegal over mid-game positions of random playouts (drops, checks, pins) -------
static int movegenBench() {
    using namespace b200::rules;
    std::mt19937_64 Rng(7);
    std::vector<Position> Sample;
    while (Sample.size() < 4000) {
        Position P;
        for (int Ply = 0; Ply < 140; ++Ply) {
            Move Ms[kMaxMoves];
            const int N = P.generateLegal(Ms);
            if (N == 0) break;
            if (Ply >= 30 && Ply % 5 == 0) Sample.push_back(P);
            Position::Undo U;
            P.make(Ms[Rng() % (uint64_t)N], &U);
        }
    }
    uint64_t Moves = 0, InCheck = 0;
    double Best = 1e30;
    for (int Rep = 0; Rep < 7; ++Rep) {
        Moves = InCheck = 0;
        const auto T0 = std::chrono::steady_clock::now();
        for (int K = 0; K < 25; ++K)
            for (Position& P : Sample) {
                Move Ms[kMaxMoves];
                Moves += (uint64_t)P.generateLegal(Ms);
            }
        Best = std::min(Best, std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
    }
    double BestMate = 1e30;
    uint64_t Found = 0;
    for (int Rep = 0; Rep < 5; ++Rep) {
        Found = 0;
        const auto T0 = std::chrono::steady_clock::now();
        for (Position& P : Sample) Found += P.mateIn3() ? 1 : 0;
        BestMate = std::min(BestMate, std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
    }
    std::printf("{\"mate_in_3_ns_per_position\": %.1f, \"found\": %llu, \"positions\": %zu}\n", 1e9 * BestMate / (double)Sample.size(),
                (unsigned long long)Found, Sample.size());
    for (Position& P : Sample) InCheck += P.inCheck(P.Side);
    const double Calls = 25.0 * (double)Sample.size();
    std::printf("{\"movegen_ns_per_position\": %.1f, \"ns_per_legal_move\": %.2f, \"avg_legal_moves\": %.1f, \"positions\": %zu, \"in_check\": %.3f}\n",
                1e9 * Best / Calls, 1e9 * Best / (double)Moves, (double)Moves / Calls, Sample.size(), (double)InCheck / (double)Sample.size());
    return 0;
}

// ---- the USI-style search (host/usi_search.h: search threads filling pinned batches in place, the evaluation thread
//      submitting, feeding and collecting leaves between its duties) on the mock pipeline: `--usi-loop THREADS MILLISECONDS
//      [nohelp]`.  Must END; afterwards the tree's invariants hold (no virtual loss left but the abandoned leaves', every edge mirrors its child).
static int usiLoop(int Threads, int Milliseconds, bool Help) {
    using namespace b200;
    MockEvalPipeline Pipe(3, 64);
    const std::size_t Cap = 400000;
    search::Tree T(Cap, Cap * 48);
    const rules::Position Root;
    UsiSearch<MockEvalPipeline> Search(&T, &Pipe, Root, 320, 0, false);
    const double Sec = Search.run(Milliseconds * 1e-3, Threads, Help);
    const search::Node& R = T.node(0);
    // (a search thread that is stopped while it waits for a row abandons its claimed leaf: at most one virtual loss per
    //  search thread is left on the paths, usi_search.h searchStep)
    const uint32_t Slack = (uint32_t)Threads;
    CHECK(R.evaluated() && R.Visits > 100 && R.VirtualLoss <= Slack);
    CHECK(Search.Evals > 100 && Search.Batches > 2);
    if (Help && Threads <= 1) CHECK(Search.HelpedLeaves > 0);
    if (!Help) CHECK(Search.HelpedLeaves == 0);
    bool Ok = true;
    uint64_t Evaluated = 0;
    for (std::size_t I = 0; I < T.numNodes(); ++I) {
        const search::Node& N = T.node((int)I);
        Ok = Ok && N.VirtualLoss <= Slack;
        if (!N.evaluated()) continue;
        ++Evaluated;
        uint64_t Sum = 0;
        for (int J = 0; J < N.NumEdges; ++J) {
            const search::Edge& E = T.edgesOf((int)I)[J];
            if (J > 0 && N.Term == search::Open) Ok = Ok && E.P <= T.edgesOf((int)I)[J - 1].P;
            if (E.Child >= 0) {
                Ok = Ok && E.CVisits == T.node(E.Child).Visits && E.CVirtualLoss <= Slack;
                Sum += T.node(E.Child).Visits;
            }
        }
        if (N.Term == search::Open) Ok = Ok && N.Visits == Sum + 1;
    }
    CHECK(Ok);
    std::printf("usi loop: %d search threads%s, %.0f ms: %u root visits, %llu evaluations in %llu batches, %llu leaves collected by the evaluation "
                "thread, %llu collisions, %llu evaluated nodes: invariants hold: ok\n",
                Threads, Help ? " + helping evaluation thread" : "", Sec * 1e3, R.Visits, (unsigned long long)Search.Evals,
                (unsigned long long)Search.Batches, (unsigned long long)Search.HelpedLeaves, (unsigned long long)Search.Collisions.load(),
                (unsigned long long)Evaluated);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 4 && std::strcmp(argv[1], "--usi-loop") == 0)
        return usiLoop(std::atoi(argv[2]), std::atoi(argv[3]), !(argc >= 5 && std::strcmp(argv[4], "nohelp") == 0));
    if (argc >= 2 && std::strcmp(argv[1], "--movegen-bench") == 0) return movegenBench();
    if (argc >= 4 && std::strcmp(argv[1], "--host-cost") == 0)
        return hostCost((std::size_t)std::atol(argv[2]), (std::size_t)std::atol(argv[3]), argc >= 5 ? (std::size_t)std::atol(argv[4]) : 0);
    if (argc >= 5 && std::strcmp(argv[1], "--selfplay-loop") == 0)
        return selfplayLoop(std::atoi(argv[2]), (std::size_t)std::atol(argv[3]), std::atoi(argv[4]));
    if (argc >= 3 && std::strcmp(argv[1], "--worker-cycles") == 0)
        return workerCycles((std::size_t)std::atol(argv[2])) || workerCycles((std::size_t)std::atol(argv[2]), 48);
    if (argc >= 4 && std::strcmp(argv[1], "--tree-stress") == 0) return treeStress(std::atoi(argv[2]), (std::size_t)std::atol(argv[3]));
    if (argc >= 3 && std::strcmp(argv[1], "--perft") == 0) return rulesChecks(std::atoi(argv[2]));
    if (argc >= 4 && std::strcmp(argv[1], "--queue-stress") == 0) return queueStress(std::atoi(argv[2]), (std::size_t)std::atol(argv[3]));
    if (argc >= 4 && std::strcmp(argv[1], "--onnx-blob") == 0) return onnxBlob(argv[2], argv[3]);
    if (argc >= 3 && std::strcmp(argv[1], "--cache-trace") == 0) return cacheTrace((std::size_t)std::atoi(argv[2]));
    {   // cache: store/load, 164-move cap, refresh-only on duplicate, the reference's replacement order
        mcts::EvalCacheB200 C(1);
        const size_t NB = C.numBundles();
        CHECK(NB > 0);
        float P[200];
        for (int I = 0; I < 200; ++I) P[I] = (float)I;
        mcts::EvalCacheB200::EvalInfo E;
        CHECK(!C.load(42, &E));
        CHECK(!C.store(42, 165, P, 0.5f, 0.1f));            // evalcache.cc:51-53
        CHECK(C.store(42, 164, P, 0.5f, 0.1f));
        CHECK(C.load(42, &E) && E.NumMoves == 164 && E.Policy[163] == 163.f && E.WinRate == 0.5f && E.DrawRate == 0.1f);
        P[0] = 99.f;
        CHECK(C.store(42, 164, P, 0.9f, 0.9f));             // same (hash, n): recency only, no overwrite (:73-91)
        CHECK(C.load(42, &E) && E.Policy[0] == 0.f && E.WinRate == 0.5f);
        const uint64_t H1 = 42 + NB, H2 = 42 + 2 * NB, H3 = 42 + 3 * NB;  // same bundle
        CHECK(C.store(H1, 3, P, 0.1f, 0.f) && C.store(H2, 3, P, 0.2f, 0.f));
        // list is now H2, H1, 42 and both H1 and 42 have been heads: their Prev is null, so this hit
        // does NOT move 42 to the front (evalcache.cc:146 guard) ...
        CHECK(C.load(42, &E));
        CHECK(C.store(H3, 3, P, 0.3f, 0.f));                 // ... and the full bundle overwrites its last entry: 42
        CHECK(!C.load(42, &E) && C.load(H1, &E) && C.load(H2, &E) && C.load(H3, &E) && E.WinRate == 0.3f);
        // feed(): CSR rows, > 164 moves skipped
        const uint32_t Off[4] = {0, 2, 2 + 170, 2 + 170 + 1};
        std::vector<float> Legal(Off[3], 0.25f);
        const uint64_t Hs[3] = {1000, 1001, 1002};
        const float W[3] = {0.1f, 0.2f, 0.3f}, D[3] = {0.f, 0.f, 0.f};
        CHECK(C.feed(Hs, 3, Off, Legal.data(), W, D) == 2);
        CHECK(C.load(1000, &E) && E.NumMoves == 2 && !C.load(1001, &E) && C.load(1002, &E) && E.NumMoves == 1);
    }
    {   // move index: range, planes, mirroring, no collisions among moves that can be legal together
        using b200::MoveSpec;
        auto sq = [](int f, int r) { return 9 * (f - 1) + (r - 1); };
        CHECK(b200::getMoveIndex(0, MoveSpec{sq(7, 7), sq(7, 6), false, -1}) == 0 * 81 + sq(7, 6));   // N
        CHECK(b200::getMoveIndex(0, MoveSpec{sq(7, 7), sq(7, 6), true, -1}) == 10 * 81 + sq(7, 6));
        CHECK(b200::getMoveIndex(0, MoveSpec{sq(2, 9), sq(1, 7), false, -1}) / 81 == 8);               // knight
        CHECK(b200::getMoveIndex(0, MoveSpec{sq(2, 9), sq(3, 7), false, -1}) / 81 == 9);
        CHECK(b200::getMoveIndex(0, MoveSpec{0, sq(5, 5), false, 6}) == 26 * 81 + sq(5, 5));            // rook drop
        // white's mirrored move lands on the mirrored slot of the same plane
        CHECK(b200::getMoveIndex(1, MoveSpec{80 - sq(7, 7), 80 - sq(7, 6), false, -1}) == 0 * 81 + sq(7, 6));
        std::set<int> Seen;
        for (int From = 0; From < 81; ++From)
            for (int To = 0; To < 81; ++To) {
                if (From == To) continue;
                const int Df = To / 9 - From / 9, Dr = To % 9 - From % 9;
                const bool Line = Df == 0 || Dr == 0 || Df == Dr || Df == -Dr, Knight = Dr == -2 && (Df == 1 || Df == -1);
                if (!Line && !Knight) continue;
                for (int Pr = 0; Pr < 2; ++Pr) {
                    const int I = b200::getMoveIndex(0, MoveSpec{From, To, Pr == 1, -1});
                    CHECK(I >= 0 && I < 20 * 81 && I % 81 == To);
                    Seen.insert(I);
                }
            }
        for (int Pc = 0; Pc < 7; ++Pc)
            for (int To = 0; To < 81; ++To) Seen.insert(b200::getMoveIndex(0, MoveSpec{0, To, false, Pc}));
        std::printf("reachable policy slots: %zu\n", Seen.size());
        CHECK(*Seen.rbegin() < 2187 && Seen.size() > 1500);  // most slots are reachable
    }
    if (feedChecks()) return 1;  // mcts_feed.h == setEvaluation + sort + updateAncestors of the reference
    if (selfplayFeedChecks()) return 1;
    if (rulesChecks(4)) return 1;          // shogi rules: perft(1..4) of hirate + the special rules
    if (selfplayMock(2, 24)) return 1;     // two whole games of the self-play loop against a mock evaluator
    if (selfplayMock(1, 24, 3)) return 1;  // one more with the shallow mate search at leaves and roots
    if (workerCycles(5000) || workerCycles(5000, 48)) return 1;  // the pipelined evaluation worker inside the worker::Worker contract
    if (selfplayLoop(2, 48, 300)) return 1; // the whole harness: starts, plays, winds down
    if (usiLoop(2, 300, true) || usiLoop(1, 200, true) || usiLoop(2, 200, false)) return 1;  // the USI-style search on the mock pipeline
    std::printf("host_unit ok\n");
    return 0;
}
