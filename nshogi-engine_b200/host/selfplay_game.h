// selfplay_game.h — one self-play game ("frame") and the search side of its phase machine, shared by the GPU harness
// (selfplay_real.cc) and the CPU unit test that plays whole games against a mock evaluator (host_unit.cc --selfplay).
// Reference: src/selfplay/frame.h (Frame), src/selfplay/worker.cc:55-110 (phases), :112-156 (initialize), :159-206
// (prepareRoot), :330-372 (terminal checks), :415-430 (playout budget), :520-640 (transition), :476-518 (judge),
// src/selfplay/frame.cc:93-136 (setEvaluation).  Rules: rules/shogi.h; tree: mcts_search.h.
#ifndef NSHOGI_ENGINE_B200_SELFPLAY_GAME_H
#define NSHOGI_ENGINE_B200_SELFPLAY_GAME_H

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "mcts_search.h"
#include "rules/shogi.h"
#include "teacher_io.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace game {

struct GameOptions {
    int Playouts = 200;
    bool Gumbel = false;
    double FullSearchRatio = 0.25;
    // 0 / 1 / 3: shallow mate search at every new leaf and at the root of a position just reached (the reference asks
    // libnshogi's solver::dfs::solve(State, 3) at leaves, worker.cc:352-362, and its df-pn solver in judge, :517).  Off by
    // default: rules/shogi.h's search costs 4-7x a move generation; a build against libnshogi uses its solvers.
    int MatePlies = 0;
};

struct Frame {  // reference src/selfplay/frame.h: one game in flight
    rules::Position Root, Leaf;
    std::vector<uint64_t> History, Path;
    std::vector<uint8_t> InCheck;            // per History entry: is the side to move in check there?
    search::Tree Tree;
    int LeafNode = -1;
    rules::Move LeafMoves[rules::kMaxMoves];
    uint16_t LeafSlots[rules::kMaxMoves];
    int NumLeafMoves = 0;
    uint16_t MaxPly = 320;
    float BlackDraw = 0.5f, WhiteDraw = 0.5f;
    uint32_t Playouts = 0;
    bool FullSearch = true;
    double Noise[600];
    std::mt19937_64 MT;
    std::vector<rules::Move> GameMoves;      // the game so far (State::getHistoryMove)
    std::vector<uint8_t> DidFullSearch;      // per ply (Frame::getDidFullSearch, frame.h)
    uint8_t Winner = teacher::WinnerNone;
    // the leaf's evaluation as the evaluation worker left it (stageEvaluation), applied by the search worker that takes
    // the frame next (SelfplayPhase::Backpropagation runs on the search workers in the reference too, worker.cc:94-96)
    bool EvalPending = false;
    float EvalWin = 0.f, EvalDraw = 0.f;
    float EvalRow[rules::kMaxMoves];
    uint16_t EvalOrder[rules::kMaxMoves];
};

struct Info {  // reference src/selfplay/selfplayinfo.h
    std::atomic<uint64_t> Evals{0}, Batches{0}, Records{0}, Games{0}, NanRows{0}, CacheHits{0};
    std::atomic<uint64_t> Mates{0}, Repetitions{0}, MaxPlies{0}, Terminals{0}, PliesPlayed{0}, LegalMoves{0};
    std::atomic<uint64_t> Declarations{0}, PerpetualChecks{0}, MatesBySearch{0}, LeafMatesBySearch{0};
};

// Worker::initialize, worker.cc:112-156
inline void newGame(const GameOptions& O, Frame& F) {
    F.Root.setHirate();
    F.History.clear();
    F.History.push_back(F.Root.Hash);
    F.InCheck.assign(1, 0);
    F.GameMoves.clear();
    F.DidFullSearch.clear();
    F.Winner = teacher::WinnerNone;
    std::uniform_int_distribution<int> MaxPlyD(160 + 64, 512 + 128);
    std::uniform_real_distribution<float> DrawD(0.0f, 1.0f);
    F.MaxPly = (uint16_t)MaxPlyD(F.MT);
    if (F.MT() % 4 < 2) {
        F.BlackDraw = F.WhiteDraw = 0.5f;
    } else {
        F.BlackDraw = DrawD(F.MT);
        F.WhiteDraw = 1.0f - F.BlackDraw;
    }
    (void)O;
}

// Worker::prepareRoot, worker.cc:159-206
inline void prepareRoot(const GameOptions& O, Frame& F) {
    F.Tree.reset();
    if (O.Gumbel) {
        std::uniform_real_distribution<double> D(std::numeric_limits<double>::min(), 1.0);
        for (double& X : F.Noise) X = -std::log(-std::log(D(F.MT)));
    }
    // (the Dirichlet noise of an AlphaZero root is drawn when the root's evaluation arrives: drawRootNoise)
    std::uniform_real_distribution<double> U(0.0, 1.0);
    F.FullSearch = U(F.MT) <= O.FullSearchRatio;
    F.Playouts = F.FullSearch ? (uint32_t)O.Playouts : (uint32_t)std::max(1, O.Playouts / 4);
}

// judge, worker.cc:477-488 (State::getRepetitionStatus is libnshogi's): the position that was just reached - the last
// entry of History - has now occurred four times: a draw, unless one side gave check with every move since the first
// occurrence (RepetitionStatus::WinRepetition / LossRepetition): that side loses.  History[I] is the position at ply I of
// a game from hirate, so its side to move is I & 1, and InCheck[I] says whether that side is in check; "X always checked"
// == every position after the first occurrence with X's opponent to move is a check.
enum class Repetition { None, Draw, BlackLoses, WhiteLoses };
inline Repetition repetitionStatus(const std::vector<uint64_t>& History, const std::vector<uint8_t>& InCheck) {
    const uint64_t Hash = History.back();
    int Seen = 0;
    std::size_t First = History.size();
    for (std::size_t I = 0; I < History.size(); ++I)
        if (History[I] == Hash) {
            ++Seen;
            First = std::min(First, I);
        }
    if (Seen < 4) return Repetition::None;
    for (int X = 0; X < 2; ++X) {
        int Replies = 0, Checked = 0;
        for (std::size_t I = First + 1; I < History.size(); ++I)
            if ((int)(I & 1) != X) {
                ++Replies;
                Checked += InCheck[I];
            }
        if (Replies > 0 && Checked == Replies) return X == 0 ? Repetition::BlackLoses : Repetition::WhiteLoses;
    }
    return Repetition::Draw;
}

// Worker::transition + judge, worker.cc:520-640,476-518: play the most visited move; true when the game is over.
inline bool transition(const GameOptions& O, Frame& F, Info* SI) {
    const int Best = F.Tree.bestRootEdge();
    const rules::Move M = F.Tree.edgesOf(0)[Best].M;
    rules::Position::Undo U;
    F.Root.make(M, &U);
    F.History.push_back(F.Root.Hash);
    F.GameMoves.push_back(M);
    F.DidFullSearch.push_back(F.FullSearch ? 1 : 0);      // Frame::pushDidFullSearch
    SI->Records.fetch_add(1, std::memory_order_relaxed);  // positions played (teacher records: full-search plies only, teacher_io.h)
    SI->PliesPlayed.fetch_add(1, std::memory_order_relaxed);
    F.InCheck.push_back(F.Root.inCheck(F.Root.Side) ? 1 : 0);
    const Repetition Rep = repetitionStatus(F.History, F.InCheck);
    if (Rep != Repetition::None) {
        if (Rep != Repetition::Draw) {
            F.Winner = Rep == Repetition::BlackLoses ? teacher::WinnerWhite : teacher::WinnerBlack;
            SI->PerpetualChecks.fetch_add(1, std::memory_order_relaxed);
        }
        SI->Repetitions.fetch_add(1, std::memory_order_relaxed);
        return true;
    }
    if (F.Root.canDeclare()) {  // worker.cc:490-494
        F.Winner = F.Root.Side == 0 ? teacher::WinnerBlack : teacher::WinnerWhite;
        SI->Declarations.fetch_add(1, std::memory_order_relaxed);
        return true;
    }
    if (F.Root.Ply >= F.MaxPly) {
        SI->MaxPlies.fetch_add(1, std::memory_order_relaxed);
        return true;
    }
    if (!F.Root.hasLegalMove()) {
        SI->Mates.fetch_add(1, std::memory_order_relaxed);
        F.Winner = F.Root.Side == 0 ? teacher::WinnerWhite : teacher::WinnerBlack;  // the side to move is mated
        return true;
    }
    if (O.MatePlies > 0 && (O.MatePlies >= 3 ? F.Root.mateIn3() : F.Root.mateIn1())) {  // judge, worker.cc:517-523: a forced mate ends the game
        SI->MatesBySearch.fetch_add(1, std::memory_order_relaxed);
        F.Winner = F.Root.Side == 0 ? teacher::WinnerBlack : teacher::WinnerWhite;
        return true;
    }
    return false;
}

// A finished game for the save worker (SelfplayPhase::Save, worker.cc:98-99; saveworker.cc:56-88).
inline teacher::FinishedGame finishedGame(const Frame& F) {
    teacher::FinishedGame G;
    G.Moves = F.GameMoves;
    G.DidFullSearch = F.DidFullSearch;
    G.MaxPly = F.MaxPly;
    G.BlackDraw = F.BlackDraw;
    G.WhiteDraw = F.WhiteDraw;
    G.Winner = F.Winner;
    return G;
}

// -DNSB_PHASE_TIMING: cycles per phase of advance() (host_unit --host-cost prints them); compiled out otherwise.
#ifdef NSB_PHASE_TIMING
struct PhaseCycles {
    uint64_t Select = 0, Generate = 0, Terminal = 0, Expand = 0, Slots = 0, Root = 0, Apply = 0;
};
inline PhaseCycles& phaseCycles() {
    static thread_local PhaseCycles P;
    return P;
}
#define NSB_PHASE(Field, Stmt)                                      \
    do {                                                            \
        const uint64_t T_ = __builtin_ia32_rdtsc();                 \
        Stmt;                                                       \
        phaseCycles().Field += __builtin_ia32_rdtsc() - T_;         \
    } while (0)
#else
#define NSB_PHASE(Field, Stmt) \
    do {                       \
        Stmt;                  \
    } while (0)
#endif

// One frame until it needs the network: selectLeaf / checkTerminal / backpropagate / transition (worker.cc:82-106).
// OnGameEnd(const Frame&) is called when a game is over, before the frame starts its next one.
inline void applyEvaluation(const GameOptions& O, Frame& F, const float* Row, const uint16_t* Order, float WinRate, float DrawRate);

template <typename GameEnd>
inline void advance(const GameOptions& O, Frame& F, Info* SI, GameEnd&& OnGameEnd) {
    if (F.EvalPending) {
        F.EvalPending = false;
        NSB_PHASE(Apply, applyEvaluation(O, F, F.EvalRow, F.EvalOrder, F.EvalWin, F.EvalDraw));
    }
    for (;;) {
        const search::Node& Root = F.Tree.node(0);
        if (Root.evaluated() && (Root.NumEdges == 1 || Root.Visits >= F.Playouts + 1)) {  // worker.cc:415-430 (+1: the root's own evaluation)
            NSB_PHASE(Root, {
                if (transition(O, F, SI)) {
                    SI->Games.fetch_add(1, std::memory_order_relaxed);
                    OnGameEnd(F);
                    newGame(O, F);
                }
                prepareRoot(O, F);
            });
            continue;
        }
        int Node;
        NSB_PHASE(Select, {
            F.Leaf = F.Root;
            F.Path.clear();
            Node = F.Tree.selectLeaf(F.Leaf, F.BlackDraw, F.WhiteDraw, &F.Path);
        });
        const search::Node& N = F.Tree.node(Node);
        if (N.Term == search::Mated) {
            F.Tree.backup(Node, 0.0f, 0.0f);
            continue;
        }
        if (N.Term == search::DrawnGame) {
            F.Tree.backup(Node, 0.5f, 1.0f);
            continue;
        }
        if (N.Term == search::Declared) {
            F.Tree.backup(Node, 1.0f, 0.0f);
            continue;
        }
        // a new leaf: terminal checks first (searchworker.cc:475-538, selfplay/worker.cc:270-372)
        bool Declares;
        NSB_PHASE(Terminal, Declares = Node != 0 && F.Leaf.canDeclare());
        if (Declares) {  // worker.cc:299-317: the side to move declares and wins
            F.Tree.setTerminal(Node, search::Declared);
            SI->Terminals.fetch_add(1, std::memory_order_relaxed);
            F.Tree.backup(Node, 1.0f, 0.0f);
            continue;
        }
        int NumMoves;
        NSB_PHASE(Generate, NumMoves = F.Leaf.generateLegal(F.LeafMoves));
        if (NumMoves == 0) {
            F.Tree.setTerminal(Node, search::Mated);
            SI->Terminals.fetch_add(1, std::memory_order_relaxed);
            F.Tree.backup(Node, 0.0f, 0.0f);
            continue;
        }
        bool Drawn;
        NSB_PHASE(Terminal, Drawn = Node != 0 && (search::isFourfold(F.Leaf.Hash, F.History, F.Path) || F.Leaf.Ply >= F.MaxPly));
        if (Drawn) {
            F.Tree.setTerminal(Node, search::DrawnGame);
            SI->Terminals.fetch_add(1, std::memory_order_relaxed);
            F.Tree.backup(Node, 0.5f, 1.0f);
            continue;
        }
        if (O.MatePlies > 0 && Node != 0 && (O.MatePlies >= 3 ? F.Leaf.mateIn3() : F.Leaf.mateIn1())) {  // worker.cc:352-362
            F.Tree.setTerminal(Node, search::Declared);  // (the same terminal kind: the side to move wins)
            SI->Terminals.fetch_add(1, std::memory_order_relaxed);
            SI->LeafMatesBySearch.fetch_add(1, std::memory_order_relaxed);
            F.Tree.backup(Node, 1.0f, 0.0f);
            continue;
        }
        NSB_PHASE(Expand, F.Tree.expand(Node, F.LeafMoves, NumMoves));
        NSB_PHASE(Slots, for (int J = 0; J < NumMoves; ++J) F.LeafSlots[J] = (uint16_t)F.Leaf.policyIndex(F.LeafMoves[J]));  // ml::getMoveIndex
        F.NumLeafMoves = NumMoves;
        F.LeafNode = Node;
        return;
    }
}

inline void advance(const GameOptions& O, Frame& F, Info* SI) {
    advance(O, F, SI, [](const Frame&) {});
}

// The reference draws 600 Gamma(0.15) samples at every root and divides them by their sum (worker.cc:165-177), of which
// the root's NumEdges first are used - and only at a full-search root (frame.cc:121).  Same joint distribution, drawn
// when needed: NumEdges samples + one Gamma(0.15 * (600 - NumEdges)) sample for the sum of the ones nobody looks at.
inline void drawRootNoise(Frame& F, int NumEdges) {
    constexpr int Total = (int)(sizeof(F.Noise) / sizeof(F.Noise[0]));
    std::gamma_distribution<double> D(0.15, 1.0);
    double Sum = 0.0;
    for (int J = 0; J < NumEdges; ++J) Sum += (F.Noise[J] = D(F.MT));
    if (NumEdges < Total) Sum += std::gamma_distribution<double>(0.15 * (double)(Total - NumEdges), 1.0)(F.MT);
    for (int J = 0; J < NumEdges; ++J) F.Noise[J] /= Sum;
}

// Frame::setEvaluation's consumer side (frame.cc:93-136) for a row the executor decoded (NSB_DECODE_BOTH + order_out):
// priors in rank order (Node::setEvaluation + Node::sort in one pass), the Dirichlet mix of a full-search AlphaZero
// root (frame.cc:121-133; the noise is i.i.d., so mixing it after the sort draws from the same distribution), then
// Node::updateAncestors.
inline void applyEvaluation(const GameOptions& O, Frame& F, const float* Row, const uint16_t* Order, float WinRate, float DrawRate) {
    F.Tree.setPriors(F.LeafNode, Row, Order, /*Publish=*/false);
    if (!O.Gumbel && F.LeafNode == 0 && F.FullSearch) {
        search::Edge* E = F.Tree.edgesOf(0);
        const int NumEdges = F.Tree.node(0).NumEdges;
        const double EPS = 0.25;
        drawRootNoise(F, NumEdges);
        for (int J = 0; J < NumEdges; ++J) E[J].P = (float)((1 - EPS) * (double)E[J].P + EPS * F.Noise[J]);
        F.Tree.sortEdges(0);
    }
    F.Tree.publish(F.LeafNode);
    F.Tree.backup(F.LeafNode, WinRate, DrawRate);
}

// What the evaluation worker does with a decoded row: leave it with the frame.  Walking the tree (setPriors over the
// leaf's edges, backup along ~15 ancestors: dependent cache misses in memory another core wrote last) is the search
// workers' job - many of them, one evaluation thread.
inline void stageEvaluation(Frame& F, const float* Row, const uint16_t* Order, float WinRate, float DrawRate) {
    std::memcpy(F.EvalRow, Row, (std::size_t)F.NumLeafMoves * sizeof(float));
    std::memcpy(F.EvalOrder, Order, (std::size_t)F.NumLeafMoves * sizeof(uint16_t));
    F.EvalWin = WinRate;
    F.EvalDraw = DrawRate;
    F.EvalPending = true;
}

} // namespace game
} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
