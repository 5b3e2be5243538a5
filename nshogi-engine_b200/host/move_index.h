// move_index.h — legal move -> policy slot in [0, 2187), the value the decode kernels gather at.
//
// In the reference this is libnshogi's ml::getMoveIndex<ChannelsFirst>(Color, Move16) (called at
// src/mcts/feedworker.cc:108,122 and src/selfplay/frame.cc:103); the library is not vendored and
// not available here, so this restatement is BUILDER-DEFINED and parity of
// the numbering is UNPINNED.  Only the range (27 planes x 81 squares, plane-major,
// src/mcts/evaluationworker.cc:166) and the side-relative mirroring are taken from the reference.
// The GPU never derives indices itself — it consumes whatever the host adaptor supplies — so when
// built against the real library this header is replaced by a call to ml::getMoveIndex.
#ifndef NSHOGI_ENGINE_B200_MOVE_INDEX_H
#define NSHOGI_ENGINE_B200_MOVE_INDEX_H

#include <cstdint>

namespace nshogi {
namespace engine {
namespace b200 {

// Squares: s = 9 * (file - 1) + (rank - 1), file 1..9, rank 1..9.
// Planes : 0..9  = move along direction d without promotion, 10..19 = with promotion,
//          20..26 = drop of {P, L, N, S, G, B, R}.
// Directions (mover's view, "forward" = decreasing rank): 0 N, 1 NE, 2 E, 3 SE, 4 S, 5 SW, 6 W,
//          7 NW, 8 knight-left, 9 knight-right.
struct MoveSpec {
    int From;       // 0..80, ignored for drops
    int To;         // 0..80
    bool Promote;
    int DropPiece;  // -1 for board moves, else 0..6
};

inline int moveDirection(int From, int To) {
    const int Df = To / 9 - From / 9, Dr = To % 9 - From % 9;
    if (Dr == -2 && (Df == 1 || Df == -1)) return Df < 0 ? 8 : 9;
    const int Sx = (Df > 0) - (Df < 0), Sy = (Dr > 0) - (Dr < 0);
    static constexpr int Table[3][3] = {/* Sx=-1 */ {7, 6, 5}, /* Sx=0 */ {0, -1, 4}, /* Sx=+1 */ {1, 2, 3}};
    return Table[Sx + 1][Sy + 1];
}

// Color: 0 = black, 1 = white.  White moves are mirrored (s -> 80 - s) so the mover always plays "up".
inline int getMoveIndex(int Color, const MoveSpec& M) {
    const int To = Color ? 80 - M.To : M.To;
    if (M.DropPiece >= 0) return (20 + M.DropPiece) * 81 + To;
    const int From = Color ? 80 - M.From : M.From;
    const int Dir = moveDirection(From, To);
    if (Dir < 0) return -1;
    return (Dir + (M.Promote ? 10 : 0)) * 81 + To;
}

} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
