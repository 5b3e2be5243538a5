// teacher_io.h — teacher records of finished self-play games: what the reference's SaveWorker::save writes
// (src/selfplay/saveworker.cc:160-182) - the game is replayed from its initial position and, for every ply at which a
// FULL search was conducted (:171-176), one record {state, next move, winner, state config} is appended to the output.
//
// The reference's record layout is libnshogi's (ml::SimpleTeacher + io::file::simple_teacher::save), which is not
// available to this build, so the byte format here is BUILDER-DEFINED ("NSBT", below) and parity with libnshogi's file
// format is UNPINNED; which plies are saved, and what a record holds, follow the reference.
//
//   file   : "NSBT", u32 version = 1, u32 record_bytes = sizeof(TeacherRecord), then records
//   record : nsb_position (108 B: board, side to move, hands, ply, max ply, draw values - include/nsb.h), the move
//            played {from (81 + hand slot for a drop), to, promote, piece type}, the winner (0 black, 1 white, 2 draw)
#ifndef NSHOGI_ENGINE_B200_TEACHER_IO_H
#define NSHOGI_ENGINE_B200_TEACHER_IO_H

#include <cstdint>
#include <cstring>
#include <ostream>
#include <vector>

#include "nsb.h"
#include "rules/shogi.h"

namespace nshogi {
namespace engine {
namespace b200 {
namespace teacher {

constexpr uint8_t WinnerBlack = 0, WinnerWhite = 1, WinnerNone = 2;

struct TeacherRecord {
    nsb_position Position;
    uint8_t From, To, Promote, Piece;
    uint8_t Winner;
    uint8_t Pad[3];
};
static_assert(sizeof(TeacherRecord) == 116, "NSBT record is 116 bytes");

// A finished game as the save worker receives it (the reference hands over the whole Frame).
struct FinishedGame {
    std::vector<rules::Move> Moves;       // from the start position (hirate)
    std::vector<uint8_t> DidFullSearch;   // per ply (Frame::getDidFullSearch)
    uint16_t MaxPly = 320;
    float BlackDraw = 0.5f, WhiteDraw = 0.5f;
    uint8_t Winner = WinnerNone;
};

inline void writeHeader(std::ostream& Out) {
    const uint32_t Head[2] = {1u, (uint32_t)sizeof(TeacherRecord)};
    Out.write("NSBT", 4);
    Out.write(reinterpret_cast<const char*>(Head), sizeof Head);
}

// saveworker.cc:160-182.  Out may be null (count only).  Returns the number of records of this game.
inline std::size_t saveGame(std::ostream* Out, const FinishedGame& G) {
    rules::Position Replay;  // hirate
    std::size_t Records = 0;
    for (std::size_t Ply = 0; Ply < G.Moves.size(); ++Ply) {
        if (G.DidFullSearch[Ply]) {  // only positions where the full search was conducted
            if (Out != nullptr) {
                TeacherRecord R;
                std::memset(&R, 0, sizeof R);
                Replay.toRecord(&R.Position, G.MaxPly, G.BlackDraw, G.WhiteDraw);
                R.From = G.Moves[Ply].From;
                R.To = G.Moves[Ply].To;
                R.Promote = G.Moves[Ply].Promote;
                R.Piece = G.Moves[Ply].Piece;
                R.Winner = G.Winner;
                Out->write(reinterpret_cast<const char*>(&R), sizeof R);
            }
            ++Records;
        }
        rules::Position::Undo U;
        Replay.make(G.Moves[Ply], &U);
    }
    return Records;
}

} // namespace teacher
} // namespace b200
} // namespace engine
} // namespace nshogi

#endif
