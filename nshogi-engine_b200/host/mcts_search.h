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
//   - virtual loss per node while a leaf is in flight, children being evaluated are skipped (:349-357);
//   - back-propagation: updateAncestors, the win rate flips at every level, the draw rate does not (node.h:170-202);
//   - one tree shared by several search threads WITHOUT a lock (node.h:59-100): counters are relaxed atomics, a leaf is
//     claimed with one compare-and-swap, a node's edges are published with a release store of its state.
// What it leaves out: mate-distance propagation / df-pn (searchworker.cc:220-240,361-424), tree reuse between moves
// and the garbage collector (tree.cc:31-94).
//
// Layout: nodes and edges live in two arenas indexed by 32-bit ints.  An edge carries a MIRROR of its child's counters,
// so that selection scans one contiguous array per level instead of dereferencing every child node (a cache miss each
// in a tree of a million nodes).  A tree is either private to one thread (self-play frame: the arenas grow on demand) or
// shared (USI-style search: fixed capacity given at construction, allocation by atomic bump).
#ifndef NSHOGI_ENGINE_B200_MCTS_SEARCH_H
#define NSHOGI_ENGINE_B200_MCTS_SEARCH_H

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "rules/shogi.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace search {

constexpr int32_t CBase = 19652;  // searchworker.h:46
constexpr double CInit = 1.25;    // searchworker.h:47

enum Terminal : uint8_t { Open = 0, Mated = 1, DrawnGame = 2, Declared = 3 };  // Declared: the side to move wins (27-point rule)
enum NodeState : uint8_t { Fresh = 0, Claimed = 1, Ready = 2 };  // Claimed: its evaluation is in flight

struct Edge {
    rules::Move M;
    float P;
    int32_t Child;  // -1: not created yet
    uint32_t CVisits, CVirtualLoss;
    float CWinAcc, CDrawAcc;
    int32_t ChildEdges;  // hint: EdgeBegin of the child once it has been expanded (-1 before); only ever used to prefetch
};

struct Node {
    int32_t Parent;
    int32_t ParentEdge;  // index of the edge that leads here, -1 for the root
    int32_t EdgeBegin;
    uint16_t NumEdges;
    uint8_t Term;
    uint8_t State;
    uint32_t Visits;
    uint32_t VirtualLoss;
    double WinAcc, DrawAcc;  // from the point of view of the side to move AT this node
    bool evaluated() const { return std::atomic_ref<const uint8_t>(State).load(std::memory_order_acquire) == Ready; }
};

template <typename T>
inline T relaxed(const T& X) {
    return std::atomic_ref<const T>(X).load(std::memory_order_relaxed);
}

class Tree {
 public:
    // Private tree (one thread): Tree().  Shared tree: Tree(maxNodes, maxEdges) - fixed arenas, lock-free operations.
    Tree() = default;
    Tree(std::size_t MaxNodes, std::size_t MaxEdges) : Shared(true) {
        NodeCap = MaxNodes;
        EdgeCap = MaxEdges;
        NodesP = static_cast<Node*>(std::malloc(NodeCap * sizeof(Node)));  // pages are touched as the tree grows
        EdgesP = static_cast<Edge*>(std::malloc(EdgeCap * sizeof(Edge)));
        if (!NodesP || !EdgesP) throw std::bad_alloc();
        reset();
    }
    ~Tree() {
        std::free(NodesP);
        std::free(EdgesP);
    }
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;

    void reset() {
        if (!Shared && NodeCap == 0) grow(64, 2048);
        NumNodes.store(1, std::memory_order_relaxed);
        NumEdgesUsed.store(0, std::memory_order_relaxed);
        initNode(0, -1, -1);
    }
    Node& node(int I) { return NodesP[I]; }
    const Node& node(int I) const { return NodesP[I]; }
    Edge& edge(std::size_t I) { return EdgesP[I]; }
    const Edge& edge(std::size_t I) const { return EdgesP[I]; }
    Edge* edgesOf(int NodeIdx) { return EdgesP + NodesP[NodeIdx].EdgeBegin; }
    const Edge* edgesOf(int NodeIdx) const { return EdgesP + NodesP[NodeIdx].EdgeBegin; }
    std::size_t numNodes() const { return NumNodes.load(std::memory_order_relaxed); }
    std::size_t numEdges() const { return NumEdgesUsed.load(std::memory_order_relaxed); }

    static constexpr int Collision = -1;    // ran into a leaf whose evaluation is in flight
    static constexpr int OutOfMemory = -2;  // shared tree: an arena is full

    // Descend from the root, applying the chosen moves to Pos and appending the hashes of the positions passed to
    // Path, until a node that has not been evaluated yet - which is CLAIMED for the caller - or a terminal node.
    // Every node on the way gets a virtual loss.  Returns the leaf's index, Collision (nothing was changed) or
    // OutOfMemory.  `Trail` is the caller's scratch (one per thread).
    int selectLeaf(rules::Position& Pos, float BlackDraw, float WhiteDraw, std::vector<uint64_t>* Path, std::vector<int>* Trail) {
        Trail->clear();
        int Cur = 0;
        for (;;) {
            Node& N = NodesP[Cur];
            std::atomic_ref<uint8_t> St(N.State);
            uint8_t S = St.load(std::memory_order_acquire);
            if (S != Ready) {
                if (S == Fresh && St.compare_exchange_strong(S, Claimed, std::memory_order_acq_rel)) break;  // ours to evaluate
                return Collision;
            }
            if (N.Term != Open) break;
            const float DrawValue = Pos.Side == 0 ? BlackDraw : WhiteDraw;
            const int E = selectEdge(N, DrawValue);
            if (E < 0) return Collision;
            Trail->push_back(Cur);
            const int32_t EdgeIdx = N.EdgeBegin + E;
            std::atomic_ref<int32_t> ChildRef(EdgesP[EdgeIdx].Child);
            int32_t Child = ChildRef.load(std::memory_order_acquire);
            if (Child < 0) {
                const int32_t NewIdx = allocNode();  // (a private tree may move its arenas here: no references are held across)
                if (NewIdx < 0) return OutOfMemory;
                initNode(NewIdx, Cur, EdgeIdx);
                std::atomic_ref<int32_t> Ref2(EdgesP[EdgeIdx].Child);
                if (Ref2.compare_exchange_strong(Child, NewIdx, std::memory_order_acq_rel)) Child = NewIdx;
                // else: another thread created the child first (Child now holds its index); ours stays unused
            }
            // A descent is a chain of dependent cache misses - node, its edges, the next node ... - through trees that
            // do not fit any cache (1,024 games per GPU): ask for the child's node and its edges together, now.
            __builtin_prefetch(&NodesP[Child]);
            const int32_t Hint = relaxed(EdgesP[EdgeIdx].ChildEdges);
            if (Hint >= 0) {
                __builtin_prefetch(&EdgesP[Hint]);
                __builtin_prefetch(reinterpret_cast<const char*>(&EdgesP[Hint]) + 64);
            }
            rules::Position::Undo U;
            Pos.make(EdgesP[EdgeIdx].M, &U);
            if (Path) Path->push_back(Pos.Hash);
            Cur = Child;
        }
        Trail->push_back(Cur);
        for (int I : *Trail) {
            Node& X = NodesP[I];
            add(X.VirtualLoss, 1u);
            if (X.ParentEdge >= 0) add(EdgesP[X.ParentEdge].CVirtualLoss, 1u);
        }
        return Cur;
    }
    int selectLeaf(rules::Position& Pos, float BlackDraw, float WhiteDraw, std::vector<uint64_t>* Path) {  // private tree
        return selectLeaf(Pos, BlackDraw, WhiteDraw, Path, &OwnTrail);
    }

    // expandLeaf (searchworker.cc:164-173): the claimed leaf's legal moves become its edges (priors follow with
    // setPriors).  False: the edge arena is full.
    bool expand(int Leaf, const rules::Move* Moves, int N) {
        const int64_t Begin = allocEdges((std::size_t)N);
        if (Begin < 0) return false;
        Node& L = NodesP[Leaf];
        L.EdgeBegin = (int32_t)Begin;
        L.NumEdges = (uint16_t)N;
        Edge* E = EdgesP + Begin;
        for (int I = 0; I < N; ++I) E[I] = Edge{Moves[I], 0.f, -1, 0u, 0u, 0.f, 0.f, -1};
        if (L.ParentEdge >= 0) std::atomic_ref<int32_t>(EdgesP[L.ParentEdge].ChildEdges).store((int32_t)Begin, std::memory_order_relaxed);
        return true;
    }

    // Node::setEvaluation + Node::sort in one pass (== host/mcts_feed.h feedRanked): edge r receives the move and
    // probability of row element Order[r]; Order == nullptr keeps the generation order.  With Publish the node becomes
    // visible to selection (release); without it the caller may still edit the priors (sortEdges) and then publish().
    void setPriors(int Leaf, const float* Probs, const uint16_t* Order, bool Publish = true) {
        Node& L = NodesP[Leaf];
        Edge* E = EdgesP + L.EdgeBegin;
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
        if (Publish) publish(Leaf);
    }
    void publish(int NodeIdx) { std::atomic_ref<uint8_t>(NodesP[NodeIdx].State).store(Ready, std::memory_order_release); }
    // A claimed leaf that turned out to be terminal (no legal moves, repetition, max ply).
    void setTerminal(int NodeIdx, Terminal T) {
        NodesP[NodeIdx].Term = T;
        publish(NodeIdx);
    }
    // std::sort of the edges by decreasing prior, for rows whose values changed after the executor ranked them (the
    // Dirichlet mix at a self-play root, frame.cc:121-133).  Before publish() only.
    void sortEdges(int NodeIdx) {
        Node& L = NodesP[NodeIdx];
        Edge* E = EdgesP + L.EdgeBegin;
        for (int I = 1; I < L.NumEdges; ++I) {  // insertion sort: stable, and the row is nearly sorted already
            const Edge X = E[I];
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
        for (int Cur = Leaf; Cur >= 0; Cur = NodesP[Cur].Parent) {
            Node& N = NodesP[Cur];
            add(N.WinAcc, (double)W);
            add(N.DrawAcc, (double)DrawRate);
            add(N.Visits, 1u);
            add(N.VirtualLoss, ~0u);  // - 1
            if (N.ParentEdge >= 0) {
                Edge& E = EdgesP[N.ParentEdge];
                add(E.CWinAcc, W);
                add(E.CDrawAcc, DrawRate);
                add(E.CVirtualLoss, ~0u);
                if (Shared) std::atomic_ref<uint32_t>(E.CVisits).fetch_add(1, std::memory_order_release);  // last: the sums are there
                else ++E.CVisits;
            }
            W = 1.0f - W;
        }
    }

    // The most visited move of a node (selfplay/worker.cc:555-590; ties and unvisited edges by prior).
    int bestEdge(int NodeIdx = 0) const {
        const Node& R = NodesP[NodeIdx];
        int Best = -1;
        uint32_t BestVisits = 0;
        for (int I = 0; I < R.NumEdges; ++I) {
            const Edge& E = EdgesP[R.EdgeBegin + I];
            const uint32_t V = relaxed(E.CVisits);
            if (Best < 0 || V > BestVisits || (V == BestVisits && E.P > EdgesP[R.EdgeBegin + Best].P)) {
                Best = I;
                BestVisits = V;
            }
        }
        return Best;
    }
    int bestRootEdge() const { return bestEdge(0); }

 private:
    bool Shared = false;
    Node* NodesP = nullptr;
    Edge* EdgesP = nullptr;
    std::size_t NodeCap = 0, EdgeCap = 0;
    std::atomic<uint32_t> NumNodes{0};
    std::atomic<uint64_t> NumEdgesUsed{0};
    std::vector<int> OwnTrail;

    // counters: relaxed atomic read-modify-writes on a shared tree, plain arithmetic on a private one
    template <typename T>
    void add(T& X, T V) {
        if (Shared) std::atomic_ref<T>(X).fetch_add(V, std::memory_order_relaxed);
        else X += V;
    }

    void initNode(int I, int Parent, int ParentEdge) {
        NodesP[I] = Node{Parent, ParentEdge, 0, 0, Open, Fresh, 0u, 0u, 0.0, 0.0};
    }
    void grow(std::size_t Nodes_, std::size_t Edges_) {  // private trees only
        if (Nodes_ > NodeCap) {
            NodesP = static_cast<Node*>(std::realloc(NodesP, Nodes_ * sizeof(Node)));
            NodeCap = Nodes_;
        }
        if (Edges_ > EdgeCap) {
            EdgesP = static_cast<Edge*>(std::realloc(EdgesP, Edges_ * sizeof(Edge)));
            EdgeCap = Edges_;
        }
        if (!NodesP || !EdgesP) throw std::bad_alloc();
    }
    int32_t allocNode() {
        const uint32_t I = NumNodes.fetch_add(1, std::memory_order_relaxed);
        if (I >= NodeCap) {
            if (Shared) {
                NumNodes.fetch_sub(1, std::memory_order_relaxed);
                return -1;
            }
            grow(NodeCap * 2, EdgeCap);
        }
        return (int32_t)I;
    }
    int64_t allocEdges(std::size_t N) {
        const uint64_t B = NumEdgesUsed.fetch_add(N, std::memory_order_relaxed);
        if (B + N > EdgeCap) {
            if (Shared || B + N > (uint64_t)std::numeric_limits<int32_t>::max()) {
                NumEdgesUsed.fetch_sub(N, std::memory_order_relaxed);
                return -1;
            }
            grow(NodeCap, std::max<std::size_t>(EdgeCap * 2, B + N));
        }
        return (int64_t)B;
    }

    // searchworker.cc:242-430 without the solved-node bookkeeping.  Returns the edge index or -1 (every candidate is
    // being evaluated).
    int selectEdge(const Node& N, float DrawValue) const {
        const uint64_t Nv = (uint64_t)relaxed(N.Visits) + relaxed(N.VirtualLoss);
        const double Const = (std::log((double)(Nv + CBase) / (double)CBase) + CInit) * std::sqrt((double)(Nv ? Nv : 1));
        const float ConstF = (float)Const;
        int Best = -1;
        float BestValue = -std::numeric_limits<float>::max();
        const Edge* E = EdgesP + N.EdgeBegin;
        for (int I = 0; I < N.NumEdges; ++I) {
            if (relaxed(E[I].Child) < 0) {  // not visited yet: the edges are sorted by prior, no later unvisited edge can beat it
                const float U = ConstF * E[I].P;
                if (U > BestValue) {
                    BestValue = U;
                    Best = I;
                }
                break;
            }
            const uint32_t CV = std::atomic_ref<const uint32_t>(E[I].CVisits).load(std::memory_order_acquire);
            if (CV == 0) continue;  // being evaluated (:349-357)
            const float Cvv = (float)(CV + relaxed(E[I].CVirtualLoss));
            const float WinRate = ((float)CV - relaxed(E[I].CWinAcc)) / Cvv;
            const float DrawRate = relaxed(E[I].CDrawAcc) / (float)CV;
            const float Value = DrawRate * DrawValue + (1.0f - DrawRate) * WinRate;
            const float U = Value + ConstF * E[I].P / (1.0f + Cvv);
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
