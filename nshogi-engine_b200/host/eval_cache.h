// eval_cache.h — the consumer of the decode step: an evaluation cache with the reference's
// observable behaviour (reference src/mcts/evalcache.{h,cc}): rows of at most 164 legal-move
// values + win rate + draw rate keyed by the 64-bit state hash; `Hash % NumBundle` picks a bundle
// of 3 entries kept on a recency list; store() and load() give up (return false) when the bundle
// is busy instead of waiting (try_lock, evalcache.cc:60-66,135-139); store() of a (hash, move-count)
// pair that is already present only refreshes its recency (:73-91); a full bundle overwrites the
// LAST entry of the list (:90-121).  Whether the row holds probabilities (MCTS,
// src/mcts/feedworker.cc:135) or raw logits (self-play, src/selfplay/frame.cc:110-114) is the
// caller's choice of decode mode.
//
// The recency list is the reference's, quirk included: its "move to front" (evalcache.cc:75-86)
// never repairs the old head's Prev pointer, so an element that has been the head keeps
// Prev == nullptr and the guard `Prev != nullptr` then refuses to move it again until its
// predecessor is moved away.  Layout differs (flat bundles, an order permutation and one "Prev is
// null" flag per list position instead of linked CacheData nodes); behaviour does not - it is
// checked operation by operation against the reference's own evalcache.cc compiled in place
// (tests/test_evalcache.py).
#ifndef NSHOGI_ENGINE_MCTS_EVAL_CACHE_B200_H
#define NSHOGI_ENGINE_MCTS_EVAL_CACHE_B200_H

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

namespace nshogi {
namespace engine {
namespace mcts {

class EvalCacheB200 {
 public:
    static constexpr std::size_t MAX_CACHE_MOVES_COUNT = 164;  // evalcache.h:26
    static constexpr int BUNDLE_WAYS = 3;                      // evalcache.h:43

    struct EvalInfo {  // evalcache.h:28-38
        uint16_t NumMoves;
        float Policy[MAX_CACHE_MOVES_COUNT];
        float WinRate;
        float DrawRate;
    };

    explicit EvalCacheB200(std::size_t MemoryMiB)
        : NumBundle(MemoryMiB * 1024ull * 1024ull / sizeof(Bundle) ? MemoryMiB * 1024ull * 1024ull / sizeof(Bundle) : 1)
        , Bundles(new Bundle[NumBundle]) {
    }

    std::size_t numBundles() const {
        return NumBundle;
    }

    bool store(uint64_t Hash, uint16_t NumM, const float* P, float WR, float D) {
        if (NumM > MAX_CACHE_MOVES_COUNT) return false;
        Bundle& B = Bundles[Hash % NumBundle];
        if (B.Busy.test_and_set(std::memory_order_acquire)) return false;
        int Pos = 0;  // position in MRU order of the way to (re)use
        for (; Pos < BUNDLE_WAYS; ++Pos) {
            const Way& W = B.Ways[B.Order[Pos]];
            if (!W.Used) break;
            if (W.Hash == Hash && W.Info.NumMoves == NumM) {  // already cached: refresh recency only
                touch(B, Pos);
                B.Busy.clear(std::memory_order_release);
                return true;
            }
        }
        if (Pos == BUNDLE_WAYS) Pos = BUNDLE_WAYS - 1;  // full: evict the least recently used
        Way& W = B.Ways[B.Order[Pos]];
        touch(B, Pos);
        W.Used = true;
        W.Hash = Hash;
        W.Info.NumMoves = NumM;
        std::memcpy(W.Info.Policy, P, sizeof(float) * NumM);
        W.Info.WinRate = WR;
        W.Info.DrawRate = D;
        B.Busy.clear(std::memory_order_release);
        return true;
    }

    bool load(uint64_t Hash, EvalInfo* Out) {
        Bundle& B = Bundles[Hash % NumBundle];
        if (B.Busy.test_and_set(std::memory_order_acquire)) return false;
        bool Hit = false;
        for (int Pos = 0; Pos < BUNDLE_WAYS; ++Pos) {
            const Way& W = B.Ways[B.Order[Pos]];
            if (!W.Used) break;
            if (W.Hash == Hash) {
                *Out = W.Info;
                touch(B, Pos);
                Hit = true;
                break;
            }
        }
        B.Busy.clear(std::memory_order_release);
        return Hit;
    }

    // Feed one decoded batch (CSR rows as produced by nsb_eval_decode_async) into the cache.
    // Rows with more than 164 moves are skipped, as in the reference.  Returns rows stored.
    std::size_t feed(const uint64_t* Hashes, std::size_t N, const uint32_t* MoveOffsets, const float* Legal,
                     const float* WinRate, const float* DrawRate) {
        std::size_t Stored = 0;
        for (std::size_t I = 0; I < N; ++I) {
            const uint32_t M = MoveOffsets[I + 1] - MoveOffsets[I];
            if (M <= MAX_CACHE_MOVES_COUNT &&
                store(Hashes[I], (uint16_t)M, Legal + MoveOffsets[I], WinRate[I], DrawRate[I]))
                ++Stored;
        }
        return Stored;
    }

 private:
    struct Way {
        bool Used = false;
        uint64_t Hash = 0;
        EvalInfo Info;
    };
    struct Bundle {
        std::atomic_flag Busy = ATOMIC_FLAG_INIT;
        uint8_t Order[BUNDLE_WAYS] = {0, 1, 2};  // way indices in list order (head first)
        bool PrevNull[BUNDLE_WAYS] = {true, false, false};  // by list position: the element's Prev == nullptr
        Way Ways[BUNDLE_WAYS];
    };
    // evalcache.cc:75-86 on the list position Pos: nothing happens when the element's Prev is null;
    // otherwise it becomes the head, the old head keeps its null Prev (now at position 1), and the
    // element that followed the moved one inherits the moved one's (valid) Prev.
    static void touch(Bundle& B, int Pos) {
        if (B.PrevNull[Pos]) return;
        const uint8_t O0 = B.Order[0], O1 = B.Order[1], O2 = B.Order[2];
        if (Pos == 1) {
            B.Order[0] = O1; B.Order[1] = O0; B.Order[2] = O2;
            B.PrevNull[1] = true;
            B.PrevNull[2] = false;
        } else {
            B.Order[0] = O2; B.Order[1] = O0; B.Order[2] = O1;
            B.PrevNull[2] = B.PrevNull[1];
            B.PrevNull[1] = true;
        }
    }

    const std::size_t NumBundle;
    std::unique_ptr<Bundle[]> Bundles;
};

} // namespace mcts
} // namespace engine
} // namespace nshogi

#endif
