// mcts_search.h — a compact PUCT tree for the callers of the leaf-evaluation path: the part of the
// reference's search that decides WHICH position needs evaluating and consumes (policy over legal moves, win rate,
// draw rate), restated over rules/shogi.h so that the self-play and USI-style harnesses of this repo drive the
// executor with real positions, real legal-move lists and a real visit distribution.
//
// What follows the reference (src/mcts/searchworker.cc, src/mcts/node.h):
//   - selection: UCB(child) = value(child) + Const * P / (1 + child virtual visits),
//     Const = (log((Nv + CBase) / CBase) + CInit) * sqrt(Nv), CBase = 19652, CInit = 1.25 (searchworker.h:46-47,
//     searchworker.cc:283-288,404-408); value(child) = DrawRate * DrawValue + (1 - DrawRate) * WinRate with
//     WinRate = (visits - WinAcc) / virtual visits (computeWinRateOfChild, :432-446); an unvisited child scores
//     Const * P (:309-320); edges are kept sorted by prior (Node::sort, node.h:163-168 - here written in rank order
//     from the executor's order_out), so the first unvisited edge ends the scan (:331-341);
//   - virtual loss per node while a leaf is in flight (node.h:59-100), children being evaluated are skipped (:349-357);
//   - back-propagation: updateAncestors, the win rate flips at every level, the draw rate does not (node.h:170-202).
// What it leaves out: mate-distance propagation / df-pn (searchworker.cc:220-240,361-424), tree reuse between moves
// and the garbage collector (tree.cc:31-94), lock-free sharing of one tree by several threads (a tree here belongs
// to one thread at a time - a self-play frame - or is shared under one mutex: the USI-style harness, whose search
// threads hold it for the descent and the back-propagation only, not for move generation).
#ifndef NSHOGI_ENGINE_B200_MCTS_SEARCH_H
#define NSHOGI_ENGINE_B200_MCTS_SEARCH_H

#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "rules/shogi.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace search {

constexpr int32_t CBase = 19652;  // searchworker.h:46
constexpr double CInit = 1.25;    // searchworker.h:47

enum Terminal : uint8_t { Open = 0, Mated = 1, DrawnGame = 2 };

struct Edge {
    rules::Move M;
    float P = 0.f;
    int32_t Child = -1;
};

struct Node {
    int32_t Parent = -1;
    int32_t EdgeBegin = 0;
    uint16_t NumEdges = 0;
    uint8_t Term = Open;
    bool Evaluated = false;  // setEvaluation() has happened (priors and predicted rates are valid)
    uint32_t Visits = 0;
    uint32_t VirtualLoss = 0;
    double WinAcc = 0.0, DrawAcc = 0.0;  // from the point of view of the side to move AT this node
};

class Tree {
 public:
    std::vector<Node> Nodes;
    std::vector<Edge> Edges;

    void reset() {
        Nodes.clear();
        Edges.clear();
        Nodes.emplace_back();
    }
    Node& root() { return Nodes[0]; }

    // Descend from the root, applying the chosen moves to Pos and appending the hashes of the positions passed to
    // Path, until a node that has not been evaluated yet.  Every node on the way gets a virtual loss.  Returns the
    // leaf's index, or -1 when the descent ran into a node that is being evaluated (collision: nothing was changed).
    int selectLeaf(rules::Position& Pos, float BlackDraw, float WhiteDraw, std::vector<uint64_t>* Path) {
        Trail.clear();
        int Cur = 0;
        for (;;) {
            Node& N = Nodes[Cur];
            if (!N.Evaluated || N.Term != Open) {
                if (!N.Evaluated && N.VirtualLoss > 0) return abandon();  // its evaluation is in flight (the root's too)
                break;
            }
            const float DrawValue = Pos.Side == 0 ? BlackDraw : WhiteDraw;
            const int E = selectEdge(N, DrawValue);
            if (E < 0) return abandon();
            Trail.push_back(Cur);
            Edge& Ed = Edges[N.EdgeBegin + E];
            rules::Position::Undo U;
            Pos.make(Ed.M, &U);
            if (Path) Path->push_back(Pos.Hash);
            if (Ed.Child < 0) {
                Ed.Child = (int32_t)Nodes.size();
                Node Fresh;
                Fresh.Parent = Cur;
                Nodes.push_back(Fresh);  // (invalidates N / Ed)
            }
            Cur = Edges[Nodes[Cur].EdgeBegin + E].Child;
        }
        Trail.push_back(Cur);
        for (int I : Trail) ++Nodes[I].VirtualLoss;
        return Cur;
    }

    // expandLeaf (searchworker.cc:164-173): the leaf's legal moves become its edges (priors follow with setPriors).
    void expand(int Leaf, const rules::Move* Moves, int N) {
        Node& L = Nodes[Leaf];
        L.EdgeBegin = (int32_t)Edges.size();
        L.NumEdges = (uint16_t)N;
        for (int I = 0; I < N; ++I) {
            Edge E;
            E.M = Moves[I];
            Edges.push_back(E);
        }
    }

    // Node::setEvaluation + Node::sort in one pass (== host/mcts_feed.h feedRanked): edge r receives the move and
    // probability of row element Order[r]; Order == nullptr keeps the generation order.
    void setPriors(int Leaf, const float* Probs, const uint16_t* Order) {
        Node& L = Nodes[Leaf];
        Edge* E = Edges.data() + L.EdgeBegin;
        if (Order != nullptr && L.NumEdges > 1) {
            rules::Move Tmp[rules::kMaxMoves];
            for (int J = 0; J < L.NumEdges; ++J) Tmp[J] = E[J].M;
            for (int R = 0; R < L.NumEdges; ++R) {
                E[R].M = Tmp[Order[R]];
                E[R].P = Probs[Order[R]];
            }
        } else {
            for (int J = 0; J < L.NumEdges; ++J) E[J].P = Probs[J];
        }
        L.Evaluated = true;
    }
    // std::sort of the edges by decreasing prior, for rows whose values changed after the executor ranked them
    // (the Dirichlet mix at a self-play root, frame.cc:121-133)
    void sortEdges(int NodeIdx) {
        Node& L = Nodes[NodeIdx];
        Edge* E = Edges.data() + L.EdgeBegin;
        for (int I = 1; I < L.NumEdges; ++I) {  // insertion sort: stable, and the row is nearly sorted already
            Edge X = E[I];
            int J = I - 1;
            while (J >= 0 && E[J].P < X.P) {
                E[J + 1] = E[J];
                --J;
            }
            E[J + 1] = X;
        }
    }

    // updateAncestors (node.h:170-202): removes the virtual losses selectLeaf left.
    void backup(int Leaf, float WinRate, float DrawRate) {
        float W = WinRate;
        for (int Cur = Leaf; Cur >= 0; Cur = Nodes[Cur].Parent) {
            Node& N = Nodes[Cur];
            N.WinAcc += W;
            N.DrawAcc += DrawRate;
            ++N.Visits;
            if (N.VirtualLoss > 0) --N.VirtualLoss;
            W = 1.0f - W;
        }
    }

    // The most visited root move (selfplay/worker.cc:555-590; ties and unvisited edges by prior).
    int bestRootEdge() const {
        const Node& R = Nodes[0];
        int Best = -1;
        uint32_t BestVisits = 0;
        for (int I = 0; I < R.NumEdges; ++I) {
            const Edge& E = Edges[R.EdgeBegin + I];
            const uint32_t V = E.Child >= 0 ? Nodes[E.Child].Visits : 0;
            if (Best < 0 || V > BestVisits || (V == BestVisits && E.P > Edges[R.EdgeBegin + Best].P)) {
                Best = I;
                BestVisits = V;
            }
        }
        return Best;
    }

 private:
    std::vector<int> Trail;

    int abandon() {
        Trail.clear();
        return -1;
    }

    // searchworker.cc:242-430 without the solved-node bookkeeping.  Returns the edge index or -1 (every candidate is
    // being evaluated).
    int selectEdge(const Node& N, float DrawValue) const {
        const uint64_t Nv = (uint64_t)N.Visits + N.VirtualLoss;
        const double Const = (std::log((double)(Nv + CBase) / (double)CBase) + CInit) * std::sqrt((double)(Nv ? Nv : 1));
        int Best = -1;
        double BestValue = std::numeric_limits<double>::lowest();
        for (int I = 0; I < N.NumEdges; ++I) {
            const Edge& E = Edges[N.EdgeBegin + I];
            if (E.Child < 0) {  // not visited yet: the edges are sorted by prior, no later unvisited edge can beat it
                const double U = Const * E.P;
                if (U > BestValue) {
                    BestValue = U;
                    Best = I;
                }
                break;
            }
            const Node& C = Nodes[E.Child];
            if (C.Visits == 0) continue;  // being evaluated (:349-357)
            const uint64_t Cvv = (uint64_t)C.Visits + C.VirtualLoss;
            const double WinRate = ((double)C.Visits - C.WinAcc) / (double)Cvv;
            const double DrawRate = C.DrawAcc / (double)C.Visits;
            const double Value = DrawRate * DrawValue + (1.0 - DrawRate) * WinRate;
            const double U = Value + Const * E.P / (double)(1 + Cvv);
            if (U > BestValue) {
                BestValue = U;
                Best = I;
            }
        }
        return Best;
    }
};

// Four-fold repetition: has `Hash` occurred three times before in the game history + the search path?
inline bool isFourfold(uint64_t Hash, const std::vector<uint64_t>& History, const std::vector<uint64_t>& Path) {
    int Seen = 0;
    for (uint64_t H : History) Seen += H == Hash;
    for (std::size_t I = 0; I + 1 < Path.size(); ++I) Seen += Path[I] == Hash;  // (the last entry is the position itself)
    return Seen >= 3;
}

} // namespace search
} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
