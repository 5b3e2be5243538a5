// rules/shogi.h — a self-contained shogi rules core for the CALLERS of the leaf-evaluation path:
// position, legal move generation, make / unmake, check / mate / repetition, Zobrist hash, hirate start position.
//
// In the reference all of this is libnshogi (core::State, core::MoveGenerator, core::Position; linked from outside
// the tree): src/mcts/searchworker.cc:164-173 (expandLeaf -> generateLegalMoves), :475-538 (terminal /
// repetition / max-ply checks), src/selfplay/worker.cc:82-110 (phases), :349-358.  The library is not available to
// this build, so the search and self-play harnesses of this repo (host/mcts_search.h, host/selfplay_real.cc,
// host/usi_go_bench.cc) run on this restatement of the RULES OF SHOGI instead.  What pins it: the perft counts of the
// start position (30, 900, 25470, 719731, 19861490 - public known answers for shogi move generators), the 593 legal
// moves of the known maximum position, hand-made positions for every special rule, and the brute-force legality filter
// on random playouts (nsb_host_unit).  Mate search: mateIn3() stands in for the 3-ply solver::dfs::solve the reference asks at every
// self-play leaf (selfplay/worker.cc:352-362); libnshogi's df-pn solver (100,000 nodes at a finished game's root, :517) has no equivalent.  Games end by mate, by declaration (27-point rule),
// by four-fold repetition (draw; lost by a side whose every move of the cycle gave check) or at max ply (draw).
//
// Squares, piece codes and hand order are those of nsb_position (include/nsb.h), so a position is handed to stage 1 of
// the executor by copying the board: square s = 9 * (file - 1) + (rank - 1); board[s] = 0 or 1 + type + 14 * colour,
// type in {P, L, N, S, G, K, B, R, +P, +L, +N, +S, +B, +R}; hands[colour][{P, L, N, S, G, B, R}].  Black (colour 0,
// sente) moves toward rank 1.
#ifndef NSHOGI_ENGINE_B200_RULES_SHOGI_H
#define NSHOGI_ENGINE_B200_RULES_SHOGI_H

#include <cstdint>
#include <cstring>
#include <vector>

#include "../move_index.h"
#include "nsb.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace rules {

enum PieceType : uint8_t {
    Pawn = 0, Lance, Knight, Silver, Gold, King, Bishop, Rook,
    ProPawn, ProLance, ProKnight, ProSilver, ProBishop, ProRook, NumPieceTypes
};
constexpr int kMaxMoves = NSB_MAX_LEGAL_MOVES;  // 593

// hand slot {P, L, N, S, G, B, R} of a (possibly promoted) piece type; -1 for the king
constexpr int kHandSlot[NumPieceTypes] = {0, 1, 2, 3, 4, -1, 5, 6, 0, 1, 2, 3, 5, 6};
constexpr PieceType kHandPiece[7] = {Pawn, Lance, Knight, Silver, Gold, Bishop, Rook};
constexpr int kPromoted[NumPieceTypes] = {ProPawn, ProLance, ProKnight, ProSilver, -1, -1, ProBishop, ProRook, -1, -1, -1, -1, -1, -1};
constexpr int kDemoted[NumPieceTypes] = {Pawn, Lance, Knight, Silver, Gold, King, Bishop, Rook, Pawn, Lance, Knight, Silver, Bishop, Rook};

struct Move {
    uint8_t From;     // 0..80; 81 + hand slot for a drop
    uint8_t To;       // 0..80
    uint8_t Promote;  // 1: the piece promotes on this move
    uint8_t Piece;    // type of the moving piece before the move (a drop: the dropped type)
    bool isDrop() const { return From >= 81; }
    int dropSlot() const { return From - 81; }
    bool operator==(const Move& O) const { return From == O.From && To == O.To && Promote == O.Promote; }
};

struct SquareTables {
    uint8_t File[81], Rank[81];
    uint8_t Type[32], Colour[32];  // of a board code 1 + type + 14 * colour (entry 0: empty)
};
constexpr SquareTables makeSquareTables() {
    SquareTables T{};
    for (int S = 0; S < 81; ++S) {
        T.File[S] = (uint8_t)(S / 9);
        T.Rank[S] = (uint8_t)(S % 9);
    }
    for (int C = 0; C < 32; ++C) {
        T.Type[C] = C ? (uint8_t)((C - 1) % 14) : 0;
        T.Colour[C] = C ? (uint8_t)((C - 1) / 14) : 0;
    }
    return T;
}
constexpr SquareTables kSq = makeSquareTables();
inline int fileOf(int S) { return kSq.File[S]; }  // 0..8 (file - 1)
inline int rankOf(int S) { return kSq.Rank[S]; }  // 0..8 (rank - 1); black's camp is ranks 7..9 (6..8 here)
inline bool onBoard(int F, int R) { return F >= 0 && F < 9 && R >= 0 && R < 9; }

struct ZobristKeys {
    uint64_t Piece[81][29];
    uint64_t Hand[2][7][19];
    uint64_t Side;
    ZobristKeys() {
        uint64_t S = 0x9E3779B97F4A7C15ull;
        auto next = [&S]() {  // splitmix64
            uint64_t Z = (S += 0x9E3779B97F4A7C15ull);
            Z = (Z ^ (Z >> 30)) * 0xBF58476D1CE4E5B9ull;
            Z = (Z ^ (Z >> 27)) * 0x94D049BB133111EBull;
            return Z ^ (Z >> 31);
        };
        for (auto& Sq : Piece)
            for (auto& K : Sq) K = next();
        for (auto& C : Hand)
            for (auto& P : C)
                for (auto& K : P) K = next();
        Side = next();
        for (auto& Sq : Piece) Sq[0] = 0;       // empty square
        for (auto& C : Hand)
            for (auto& P : C) P[0] = 0;         // nothing in hand
    }
};
inline const ZobristKeys& zobrist() {
    static const ZobristKeys K;
    return K;
}

// step / slide tables in BLACK's orientation (forward = rank - 1); white's are the same with the rank step negated
constexpr bool stepsTo(int Type, int Df, int Dr) {  // can a black piece of Type step by (Df, Dr)?
    const bool Fwd = Dr == -1, Back = Dr == 1, Side_ = Dr == 0;
    const int Af = Df < 0 ? -Df : Df;
    switch (Type) {
    case Pawn: return Df == 0 && Fwd;
    case Knight: return Af == 1 && Dr == -2;
    case Silver: return (Fwd && Af <= 1) || (Back && Af == 1);
    case Gold: case ProPawn: case ProLance: case ProKnight: case ProSilver:
        return (Fwd && Af <= 1) || (Side_ && Af == 1) || (Back && Df == 0);
    case King: return Af <= 1 && Dr >= -1 && Dr <= 1 && (Df != 0 || Dr != 0);
    case ProBishop: return Af + (Dr < 0 ? -Dr : Dr) == 1;  // orthogonal king steps (the diagonals slide)
    case ProRook: return Af == 1 && (Dr == 1 || Dr == -1);  // diagonal king steps (the orthogonals slide)
    default: return false;
    }
}
constexpr bool slidesAlong(int Type, int Df, int Dr) {  // black orientation, unit direction
    switch (Type) {
    case Lance: return Df == 0 && Dr == -1;
    case Bishop: case ProBishop: return Df != 0 && Dr != 0;
    case Rook: case ProRook: return (Df == 0) != (Dr == 0);
    default: return false;
    }
}

// The 10 steps (8 neighbours + the two knight jumps) and the 8 slide directions, black's orientation; which of them
// a piece type uses is a bit mask computed once from stepsTo / slidesAlong.
constexpr int8_t kSteps[10][2] = {{0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {-1, -2}, {1, -2}};
constexpr int8_t kDirs[8][2] = {{0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}};
struct TypeDirs {
    uint16_t Step[NumPieceTypes];
    uint8_t Slide[NumPieceTypes];
};
constexpr TypeDirs makeTypeDirs() {
    TypeDirs T{};
    for (int Type = 0; Type < NumPieceTypes; ++Type) {
        T.Step[Type] = 0;
        T.Slide[Type] = 0;
        for (int D = 0; D < 10; ++D)
            if (stepsTo(Type, kSteps[D][0], kSteps[D][1])) T.Step[Type] |= (uint16_t)(1u << D);
        for (int D = 0; D < 8; ++D)
            if (slidesAlong(Type, kDirs[D][0], kDirs[D][1])) T.Slide[Type] |= (uint8_t)(1u << D);
    }
    return T;
}
constexpr TypeDirs kTypeDirs = makeTypeDirs();

// Move-generation tables, per colour (white's directions are black's with the rank step negated): the step targets of
// every (type, square) in direction order, the length of every ray, which codes are the mover's own pieces, and for a
// move of a type from rank Rf to rank Rt whether it may promote (bit 0) and whether it may stay unpromoted (bit 1).
struct StepList {
    uint8_t N;
    uint8_t To[8];
};
struct MoveTables {
    StepList Steps[2][NumPieceTypes][81];
    uint8_t RayLen[2][81][8];
    int8_t RayStep[2][8];
    uint8_t Own[2][32];
    uint8_t Flags[2][NumPieceTypes][9][9];
    // for "is S attacked by colour By": the squares a By piece would step FROM to reach S with the step it would use,
    // and per board code the steps / slides a piece of colour By has (0 for the other colour and for empty)
    struct StepSource {
        uint8_t N;
        uint8_t From[10], D[10];
    } StepFrom[2][81];
    uint16_t StepsOf[2][32];
    uint8_t SlidesOf[2][32];
};
constexpr MoveTables makeMoveTables() {
    MoveTables T{};
    for (int Me = 0; Me < 2; ++Me) {
        const int Sign = Me == 0 ? 1 : -1;
        for (int Type = 0; Type < NumPieceTypes; ++Type)
            for (int S = 0; S < 81; ++S) {
                StepList& L = T.Steps[Me][Type][S];
                L.N = 0;
                for (int D = 0; D < 10; ++D) {
                    if (!((kTypeDirs.Step[Type] >> D) & 1u)) continue;
                    const int Tf = S / 9 + kSteps[D][0], Tr = S % 9 + Sign * kSteps[D][1];
                    if (Tf < 0 || Tf > 8 || Tr < 0 || Tr > 8) continue;
                    L.To[L.N++] = (uint8_t)(9 * Tf + Tr);
                }
            }
        for (int D = 0; D < 8; ++D) T.RayStep[Me][D] = (int8_t)(9 * kDirs[D][0] + Sign * kDirs[D][1]);
        for (int S = 0; S < 81; ++S)
            for (int D = 0; D < 8; ++D) {
                int Tf = S / 9 + kDirs[D][0], Tr = S % 9 + Sign * kDirs[D][1], Len = 0;
                while (Tf >= 0 && Tf <= 8 && Tr >= 0 && Tr <= 8) {
                    ++Len;
                    Tf += kDirs[D][0];
                    Tr += Sign * kDirs[D][1];
                }
                T.RayLen[Me][S][D] = (uint8_t)Len;
            }
        for (int C = 0; C < 32; ++C) {
            const bool Mine = C >= 1 && C <= 28 && (C - 1) / 14 == Me;
            T.Own[Me][C] = Mine ? 1 : 0;
            T.StepsOf[Me][C] = Mine ? kTypeDirs.Step[(C - 1) % 14] : (uint16_t)0;
            T.SlidesOf[Me][C] = Mine ? kTypeDirs.Slide[(C - 1) % 14] : (uint8_t)0;
        }
        for (int S = 0; S < 81; ++S) {
            MoveTables::StepSource& A = T.StepFrom[Me][S];
            A.N = 0;
            for (int D = 0; D < 10; ++D) {
                const int Af = S / 9 - kSteps[D][0], Ar = S % 9 - Sign * kSteps[D][1];
                if (Af < 0 || Af > 8 || Ar < 0 || Ar > 8) continue;
                A.From[A.N] = (uint8_t)(9 * Af + Ar);
                A.D[A.N++] = (uint8_t)D;
            }
        }
        for (int Type = 0; Type < NumPieceTypes; ++Type)
            for (int Rf = 0; Rf < 9; ++Rf)
                for (int Rt = 0; Rt < 9; ++Rt) {
                    const bool ZoneF = Me == 0 ? Rf <= 2 : Rf >= 6, ZoneT = Me == 0 ? Rt <= 2 : Rt >= 6;
                    const int Rel = Me == 0 ? Rt : 8 - Rt;  // 0 = the far rank
                    const bool Stay = (Type == Pawn || Type == Lance) ? Rel >= 1 : Type == Knight ? Rel >= 2 : true;
                    T.Flags[Me][Type][Rf][Rt] = (uint8_t)(((kPromoted[Type] >= 0 && (ZoneF || ZoneT)) ? 1 : 0) | (Stay ? 2 : 0));
                }
    }
    return T;
}
inline const MoveTables& moveTables() {
    static const MoveTables T = makeMoveTables();
    return T;
}

// Could a piece on square A give check to a king on square B at all?  True when B is a neighbour of A, a knight's jump
// away (either colour) or on the same file, rank or diagonal - the cheap filter in front of "make the move and look".
struct LineTable {
    uint8_t Aligned[81][81];
};
inline const LineTable& lineTable() {
    static const LineTable T = [] {
        LineTable L{};
        for (int A = 0; A < 81; ++A)
            for (int B = 0; B < 81; ++B) {
                const int Df = B / 9 - A / 9, Dr = B % 9 - A % 9;
                const int Af = Df < 0 ? -Df : Df, Ar = Dr < 0 ? -Dr : Dr;
                L.Aligned[A][B] = (A != B) && (Df == 0 || Dr == 0 || Af == Ar || (Af == 1 && Ar == 2));
            }
        return L;
    }();
    return T;
}

class Position {
 public:
    uint8_t Board[81];
    uint8_t Hands[2][7];
    uint8_t Side = 0;     // 0 = black to move
    uint16_t Ply = 0;
    uint8_t KingSq[2] = {0, 0};
    uint64_t Hash = 0;

    static int code(int Type, int Colour) { return 1 + Type + 14 * Colour; }
    static int typeOf(int Code) { return kSq.Type[Code]; }
    static int colourOf(int Code) { return kSq.Colour[Code]; }

    Position() { setHirate(); }

    void clear() {
        std::memset(Board, 0, sizeof Board);
        std::memset(Hands, 0, sizeof Hands);
        Side = 0;
        Ply = 0;
        KingSq[0] = KingSq[1] = 255;
        Hash = 0;
    }
    void put(int File, int Rank, int Type, int Colour) {  // 1-based file / rank
        const int S = 9 * (File - 1) + (Rank - 1);
        Board[S] = (uint8_t)code(Type, Colour);
        if (Type == King) KingSq[Colour] = (uint8_t)S;
    }
    void setHirate() {
        clear();
        const int Back[9] = {Lance, Knight, Silver, Gold, King, Gold, Silver, Knight, Lance};
        for (int F = 1; F <= 9; ++F) {
            put(F, 9, Back[F - 1], 0);
            put(F, 1, Back[9 - F], 1);
            put(F, 7, Pawn, 0);
            put(F, 3, Pawn, 1);
        }
        put(8, 8, Bishop, 0);
        put(2, 8, Rook, 0);
        put(2, 2, Bishop, 1);
        put(8, 2, Rook, 1);
        rehash();
    }
    void rehash() {
        const ZobristKeys& Z = zobrist();
        Hash = Side ? Z.Side : 0;
        for (int S = 0; S < 81; ++S) Hash ^= Z.Piece[S][Board[S]];
        for (int C = 0; C < 2; ++C)
            for (int K = 0; K < 7; ++K) Hash ^= Z.Hand[C][K][Hands[C][K]];
        for (int S = 0; S < 81; ++S)
            if (Board[S] && typeOf(Board[S]) == King) KingSq[colourOf(Board[S])] = (uint8_t)S;
    }

    // ---- attacks -------------------------------------------------------------------------------------------------
    // Is square S attacked by a piece of colour By?
    bool attacked(int S, int By) const {
        const MoveTables& T = moveTables();
        const MoveTables::StepSource& A = T.StepFrom[By][S];
        unsigned Hit = 0;
        for (int K = 0; K < A.N; ++K) Hit |= (unsigned)(T.StepsOf[By][Board[A.From[K]]] >> A.D[K]) & 1u;
        if (Hit) return true;
        for (int D = 0; D < 8; ++D) {  // a By slider moving along D comes from the opposite direction
            const int Back = (D + 4) & 7, Step = T.RayStep[By][Back];
            int Sq = S;
            for (int Len = T.RayLen[By][S][Back]; Len > 0; --Len) {
                Sq += Step;
                const int C = Board[Sq];
                if (C) {
                    if ((T.SlidesOf[By][C] >> D) & 1u) return true;
                    break;
                }
            }
        }
        return false;
    }
    bool inCheck(int Colour) const { return KingSq[Colour] != 255 && attacked(KingSq[Colour], Colour ^ 1); }

    // ---- make / unmake -------------------------------------------------------------------------------------------
    struct Undo {
        uint8_t Captured;
        uint64_t Hash;
    };
    void make(const Move& M, Undo* U) {
        const ZobristKeys& Z = zobrist();
        U->Hash = Hash;
        const int Me = Side;
        if (M.isDrop()) {
            const int K = M.dropSlot();
            U->Captured = 0;
            Hash ^= Z.Hand[Me][K][Hands[Me][K]];
            --Hands[Me][K];
            Hash ^= Z.Hand[Me][K][Hands[Me][K]];
            Board[M.To] = (uint8_t)code(kHandPiece[K], Me);
            Hash ^= Z.Piece[M.To][Board[M.To]];
        } else {
            const int Cap = Board[M.To];
            U->Captured = (uint8_t)Cap;
            if (Cap) {
                Hash ^= Z.Piece[M.To][Cap];
                const int K = kHandSlot[typeOf(Cap)];
                if (K >= 0) {
                    Hash ^= Z.Hand[Me][K][Hands[Me][K]];
                    ++Hands[Me][K];
                    Hash ^= Z.Hand[Me][K][Hands[Me][K]];
                } else {
                    KingSq[Me ^ 1] = 255;  // (never happens in legal play)
                }
            }
            const int Moving = Board[M.From];
            Hash ^= Z.Piece[M.From][Moving];
            Board[M.From] = 0;
            const int NewType = M.Promote ? kPromoted[typeOf(Moving)] : typeOf(Moving);
            Board[M.To] = (uint8_t)code(NewType, Me);
            Hash ^= Z.Piece[M.To][Board[M.To]];
            if (NewType == King) KingSq[Me] = M.To;
        }
        Side ^= 1;
        Hash ^= Z.Side;
        ++Ply;
    }
    void unmake(const Move& M, const Undo& U) {
        Side ^= 1;
        --Ply;
        const int Me = Side;
        if (M.isDrop()) {
            Board[M.To] = 0;
            ++Hands[Me][M.dropSlot()];
        } else {
            Board[M.From] = (uint8_t)code(M.Piece, Me);
            Board[M.To] = U.Captured;
            if (U.Captured) {
                const int K = kHandSlot[typeOf(U.Captured)];
                if (K >= 0) --Hands[Me][K];
                else KingSq[Me ^ 1] = M.To;
            }
            if (M.Piece == King) KingSq[Me] = M.From;
        }
        Hash = U.Hash;
    }

    // ---- move generation -----------------------------------------------------------------------------------------
    static bool inPromotionZone(int Colour, int R) { return Colour == 0 ? R <= 2 : R >= 6; }
    static bool canStay(int Type, int Colour, int R) {  // may an UNPROMOTED piece of Type stand on rank R?
        const int Rel = Colour == 0 ? R : 8 - R;  // 0 = the far rank
        if (Type == Pawn || Type == Lance) return Rel >= 1;
        if (Type == Knight) return Rel >= 2;
        return true;
    }

    // Pseudo-legal moves of the side to move (king safety not yet checked).  Returns the count.
    int generatePseudoLegal(Move* Out) const {
        static_assert(sizeof(Move) == 4, "a move is stored as one 32-bit word: From | To << 8 | Promote << 16 | Piece << 24");
        const MoveTables& T = moveTables();
        int N = 0;
        const int Me = Side;
        const uint8_t* Own = T.Own[Me];
        // Both forms of a move are always written; N only advances over the ones that exist (no branch on the board).
        auto emit = [&](uint32_t Word, unsigned Flags, unsigned Ok) {
            const uint32_t Promoting = Word | (1u << 16);
            std::memcpy(static_cast<void*>(Out + N), &Promoting, 4);
            N += (int)(Ok & Flags & 1u);
            std::memcpy(static_cast<void*>(Out + N), &Word, 4);
            N += (int)(Ok & (Flags >> 1) & 1u);
        };
        bool PawnOnFile[9] = {false, false, false, false, false, false, false, false, false};
        uint8_t Empty[81];
        int NumEmpty = 0;
        for (int F = 0, S = 0; F < 9; ++F)
            for (int R = 0; R < 9; ++R, ++S) {
                const int C = Board[S];
                if (!C) {
                    Empty[NumEmpty++] = (uint8_t)S;
                    continue;
                }
                if (!Own[C]) continue;
                const int Type = typeOf(C);
                if (Type == Pawn) PawnOnFile[F] = true;
                const uint32_t Base = (uint32_t)S | ((uint32_t)Type << 24);
                const uint8_t* FlagsFrom = T.Flags[Me][Type][R];
                const StepList& L = T.Steps[Me][Type][S];
                for (int K = 0; K < L.N; ++K) {
                    const int To = L.To[K];
                    emit(Base | ((uint32_t)To << 8), FlagsFrom[rankOf(To)], 1u ^ Own[Board[To]]);
                }
                for (unsigned Mask = kTypeDirs.Slide[Type]; Mask; Mask &= Mask - 1) {
                    const int D = __builtin_ctz(Mask);
                    const int Step = T.RayStep[Me][D];
                    int To = S;
                    for (int Len = T.RayLen[Me][S][D]; Len > 0; --Len) {
                        To += Step;
                        const int Target = Board[To];
                        if (Own[Target]) break;
                        emit(Base | ((uint32_t)To << 8), FlagsFrom[rankOf(To)], 1u);
                        if (Target) break;
                    }
                }
            }
        for (int K = 0; K < 7; ++K) {
            if (!Hands[Me][K]) continue;
            const int Type = kHandPiece[K];
            const int MinRel = Type == Knight ? 2 : (Type == Pawn || Type == Lance) ? 1 : 0;  // canStay
            const uint32_t Base = (uint32_t)(81 + K) | ((uint32_t)Type << 24);
            for (int J = 0; J < NumEmpty; ++J) {
                const int S = Empty[J], R = rankOf(S);
                const unsigned Ok = (unsigned)((Me == 0 ? R : 8 - R) >= MinRel) & (unsigned)!(Type == Pawn && PawnOnFile[fileOf(S)]);  // nifu
                const uint32_t Word = Base | ((uint32_t)S << 8);
                std::memcpy(static_cast<void*>(Out + N), &Word, 4);
                N += (int)Ok;
            }
        }
        return N;
    }

    // King safety of one side: the squares of its pieces that are pinned to its king (a piece is pinned when it is the
    // only piece between its king and an enemy slider that moves along that ray), the pieces that give check, and - for
    // a single check - the squares on which a move of another piece answers it: the checker's own square and, for a
    // slider, the squares between it and the king.
    struct SquareSet {
        uint64_t Lo = 0;   // squares 0..63
        uint32_t Hi = 0;   // squares 64..80
        void add(int S) {
            if (S < 64) Lo |= 1ull << S;
            else Hi |= 1u << (S - 64);
        }
        bool has(int S) const { return S < 64 ? (Lo >> S) & 1u : (Hi >> (S - 64)) & 1u; }
        bool any() const { return (Lo | Hi) != 0; }
    };
    struct KingSafety {
        SquareSet Pinned, Answer;
        int NumCheckers = 0;
        bool InCheck = false;
        bool pinned(int S) const { return Pinned.has(S); }
    };
    KingSafety kingSafety(int Colour) const {
        KingSafety K;
        const int Ks = KingSq[Colour];
        if (Ks == 255) return K;
        const MoveTables& T = moveTables();
        const int Enemy = Colour ^ 1;
        const MoveTables::StepSource& A = T.StepFrom[Enemy][Ks];
        for (int I = 0; I < A.N; ++I)
            if ((T.StepsOf[Enemy][Board[A.From[I]]] >> A.D[I]) & 1u) {
                if (K.NumCheckers++ == 0) K.Answer.add(A.From[I]);
            }
        for (int D = 0; D < 8; ++D) {  // an enemy slider moving along D towards the king comes from the opposite direction
            const int Back = (D + 4) & 7, Step = T.RayStep[Enemy][Back];
            int Sq = Ks, Own = -1;
            SquareSet Ray;
            for (int Len = T.RayLen[Enemy][Ks][Back]; Len > 0; --Len) {
                Sq += Step;
                const int C = Board[Sq];
                Ray.add(Sq);
                if (!C) continue;
                if (T.Own[Colour][C]) {
                    if (Own >= 0) break;  // two own pieces in the way
                    Own = Sq;
                } else {
                    if ((T.SlidesOf[Enemy][C] >> D) & 1u) {
                        if (Own >= 0) K.Pinned.add(Own);
                        else if (K.NumCheckers++ == 0) K.Answer = Ray;
                    }
                    break;
                }
            }
        }
        K.InCheck = K.NumCheckers > 0;
        return K;
    }

    // Legal moves: the mover's king is not left in check; a pawn drop that mates is illegal (uchifuzume).  Most moves
    // are decided without touching the board: not in check, a move of a piece that is neither the king nor pinned, and
    // any drop, cannot expose the king; in a single check a non-king move must land on an answering square (capture the
    // checker or interpose), in a double check only the king moves.  A king move asks whether its target is attacked with
    // the king off the board; moves of pinned pieces and the one pawn drop that gives check (the square in front of the
    // enemy king) are verified by making them.
    template <bool FirstOnly>
    int legalMoves(Move* Out) {
        Move Tmp[kMaxMoves + 64];
        const int NP = generatePseudoLegal(Tmp);
        const int Me = Side;
        const KingSafety KS = kingSafety(Me);
        const int EnemyKing = KingSq[Me ^ 1];
        // the square from which a pawn of the mover attacks the enemy king (same file, one rank behind it from the mover's view)
        const int PawnCheckSq = EnemyKing == 255 ? -1
                                : Me == 0 ? (rankOf(EnemyKing) < 8 ? EnemyKing + 1 : -1)
                                          : (rankOf(EnemyKing) > 0 ? EnemyKing - 1 : -1);
        int N = 0;
        for (int I = 0; I < NP; ++I) {
            const Move& M = Tmp[I];
            const bool Drop = M.isDrop(), KingMove = !Drop && M.Piece == King;
            if (KS.InCheck && !KingMove && (KS.NumCheckers > 1 || !KS.Answer.has(M.To))) continue;
            const bool MatingDrop = Drop && M.Piece == Pawn && M.To == PawnCheckSq;
            bool Ok = true;
            if (KingMove) {  // is the target attacked once the king has left its square (it may have shielded the target)?
                const uint8_t KingCode = Board[M.From];
                Board[M.From] = 0;
                Ok = !attacked(M.To, Me ^ 1);
                Board[M.From] = KingCode;
            } else if (MatingDrop || (!Drop && KS.pinned(M.From))) {
                Undo U;
                make(M, &U);
                Ok = !inCheck(Me);
                if (Ok && MatingDrop) Ok = hasLegalMove();  // uchifuzume
                unmake(M, U);
            }
            if (!Ok) continue;
            if (FirstOnly) return 1;
            if (N < kMaxMoves) Out[N++] = M;
        }
        return N;
    }
    int generateLegal(Move* Out) { return legalMoves<false>(Out); }
    bool hasLegalMove() { return legalMoves<true>(nullptr) != 0; }

    // Declaration win, 27-point rule (the reference plays self-play with core::EndingRule::ER_Declare27,
    // src/selfplay/worker.cc:143,299-317; State::canDeclare is libnshogi's).  The side to move may declare when its
    // king stands in the opponent's camp and is not in check, at least 10 of its other pieces stand there too, and those
    // pieces plus its hand are worth 28 points (black) / 27 (white) with rook and bishop - promoted or not - 5 and
    // everything else 1.
    bool canDeclare() const {
        const int Me = Side;
        const int K = KingSq[Me];
        if (K == 255 || !inPromotionZone(Me, rankOf(K))) return false;
        int Pieces = 0, Points = 0;
        for (int F = 0; F < 9; ++F)
            for (int R = Me == 0 ? 0 : 6; R < (Me == 0 ? 3 : 9); ++R) {
                const int C = Board[9 * F + R];
                if (C == 0 || colourOf(C) != Me) continue;
                const int T = typeOf(C);
                if (T == King) continue;
                ++Pieces;
                Points += (kDemoted[T] == Bishop || kDemoted[T] == Rook) ? 5 : 1;
            }
        if (Pieces < 10) return false;
        for (int Slot = 0; Slot < 7; ++Slot) Points += Hands[Me][Slot] * (Slot >= 5 ? 5 : 1);
        if (Points < (Me == 0 ? 28 : 27)) return false;
        return !inCheck(Me);
    }

    // ---- shallow mate search (the reference asks libnshogi's solver::dfs::solve(State, 3) at every self-play leaf,
    //      src/selfplay/worker.cc:352-362) ------------------------------------------------------------------------------
    // A move can only give check if its target is aligned with the enemy king (direct check) or its origin is (a
    // discovered check); only those are made and looked at.
    bool mayGiveCheck(const Move& M) const {
        const int K = KingSq[Side ^ 1];
        if (K == 255) return false;
        const LineTable& L = lineTable();
        return L.Aligned[M.To][K] || (!M.isDrop() && L.Aligned[M.From][K]);
    }
    // Has the side to move a move that checkmates?  (A mating pawn drop is not a legal move, so it is never found.)
    bool mateIn1(Move* Mate = nullptr) {
        Move Ms[kMaxMoves];
        const int N = generateLegal(Ms);
        const int Me = Side;
        for (int I = 0; I < N; ++I) {
            if (!mayGiveCheck(Ms[I])) continue;
            Undo U;
            make(Ms[I], &U);
            const bool Mated = inCheck(Me ^ 1) && !hasLegalMove();
            unmake(Ms[I], U);
            if (Mated) {
                if (Mate) *Mate = Ms[I];
                return true;
            }
        }
        return false;
    }
    // Mate in at most three plies: a check to which every answer allows a mate in one (or to which there is none).
    bool mateIn3(Move* First = nullptr) {
        Move Ms[kMaxMoves];
        const int N = generateLegal(Ms);
        const int Me = Side;
        for (int I = 0; I < N; ++I) {
            if (!mayGiveCheck(Ms[I])) continue;
            Undo U;
            make(Ms[I], &U);
            bool Wins = false;
            if (inCheck(Me ^ 1)) {
                Move Ev[kMaxMoves];
                const int NE = generateLegal(Ev);
                Wins = true;  // no evasion: mate in one
                for (int J = 0; J < NE && Wins; ++J) {
                    Undo V;
                    make(Ev[J], &V);
                    Wins = mateIn1();
                    unmake(Ev[J], V);
                }
            }
            unmake(Ms[I], U);
            if (Wins) {
                if (First) *First = Ms[I];
                return true;
            }
        }
        return false;
    }

    uint64_t perft(int Depth) {
        if (Depth == 0) return 1;
        Move Ms[kMaxMoves];
        const int N = generateLegal(Ms);
        if (Depth == 1) return (uint64_t)N;
        uint64_t Sum = 0;
        for (int I = 0; I < N; ++I) {
            Undo U;
            make(Ms[I], &U);
            Sum += perft(Depth - 1);
            unmake(Ms[I], U);
        }
        return Sum;
    }

    // ---- hand-over to the executor ---------------------------------------------------------------------------------
    // Stage-1 record of this position (include/nsb.h nsb_position; the board codes are the same by construction).
    void toRecord(nsb_position* P, uint16_t MaxPly, float BlackDraw, float WhiteDraw) const {
        std::memcpy(P->board, Board, 81);
        P->side = Side;
        std::memcpy(P->hands, Hands, 14);
        P->ply = Ply;
        P->max_ply = MaxPly;
        P->black_draw_value = BlackDraw;
        P->white_draw_value = WhiteDraw;
    }
    // Policy slot of a legal move of the side to move (host/move_index.h == ml::getMoveIndex's role).
    int policyIndex(const Move& M) const {
        return getMoveIndex(Side, MoveSpec{M.isDrop() ? 0 : (int)M.From, (int)M.To, M.Promote != 0, M.isDrop() ? M.dropSlot() : -1});
    }
};

} // namespace rules
} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
