// cache_device.cuh — device-resident evaluation cache, warp-level operations.
//
// Restates reference src/mcts/evalcache.{h,cc} for HBM: rows of at most 164 legal-move values +
// win + draw keyed by the 64-bit state hash, bundles of 3 entries kept in recency order, bundle =
// hash % NumBundle, try-lock per bundle (a busy bundle drops the operation: evalcache.cc:58-62,
// 127-131).  The reference threads its 3 entries on a doubly linked list; here a 32-bit word per
// bundle holds the lock bit, the 3 slots in list order, their used bits and two "Prev is null"
// flags, so that the reference's reorder is one register permutation and one atomic store.
//
// The flags restate a quirk of the reference that decides which entry a full bundle loses: its
// "move to front" (evalcache.cc:75-86,96-107,146-157) never repairs the old head's Prev pointer, so
// an element that has been the head keeps Prev == nullptr, and the guard `Prev != nullptr` of every
// reorder then leaves it where it is until its predecessor is moved away (which repairs it).  The
// behaviour is checked operation by operation against the reference's own evalcache.cc compiled in
// place (oracle/_ref/libnsb_ref_evalcache.so, tests/test_evalcache.py).
//
//   store (evalcache.cc:49-121): n > 164 -> false; walk from the front: first unused entry, or an
//       entry with the same hash AND the same move count (-> only moved to the front, data kept),
//       or the last entry; the chosen entry is moved to the front and overwritten.
//   load  (evalcache.cc:123-169): walk the used entries from the front; first entry with the same
//       hash is copied out and moved to the front (the caller compares the move count,
//       src/mcts/searchworker.cc:545-556).
// One warp serves one position; lanes 0..2 inspect the three entries, all lanes copy the row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsb_internal.h"

namespace nsb {

constexpr uint32_t kCacheLockBit = 0x80000000u;
constexpr uint32_t kCacheInitMeta = 0u | (1u << 2) | (2u << 4);  // list order = slots 0, 1, 2; nothing used, Prev valid

__device__ __forceinline__ uint32_t cache_order_slot(uint32_t meta, int i) { return (meta >> (2 * i)) & 3u; }
__device__ __forceinline__ uint32_t cache_used(uint32_t meta, uint32_t slot) { return (meta >> (8 + slot)) & 1u; }

constexpr uint32_t kCachePrevNull1 = 1u << 6;  // the element at list position 1 has Prev == nullptr
constexpr uint32_t kCachePrevNull2 = 1u << 7;  // same for list position 2

// The reference's reorder of the element at list position i (used bits untouched): a no-op when its
// Prev is null; otherwise it becomes the head, the old head (Prev still null) drops to position 1,
// and the element behind the moved one inherits the moved one's valid Prev.
__device__ __forceinline__ uint32_t cache_move_to_front(uint32_t meta, int i) {
    const uint32_t s0 = cache_order_slot(meta, 0), s1 = cache_order_slot(meta, 1), s2 = cache_order_slot(meta, 2);
    if (i == 1 && !(meta & kCachePrevNull1))
        return (meta & ~0xFFu) | s1 | (s0 << 2) | (s2 << 4) | kCachePrevNull1;
    if (i == 2 && !(meta & kCachePrevNull2))
        return (meta & ~0xFFu) | s2 | (s0 << 2) | (s1 << 4) | kCachePrevNull1 |
               ((meta & kCachePrevNull1) ? kCachePrevNull2 : 0u);
    return meta;
}

// Warp-uniform try-lock of the bundle word; on success *meta is its content without the lock bit.
// `tries` = 1 is the reference's try_lock (a busy bundle drops the operation).  The batch probe uses
// a short bounded retry instead: batches of concurrent streams tend to touch the same bundles within
// the same microsecond, and a dropped load costs a whole network evaluation, while the holder (one
// warp, a few hundred nanoseconds, never waiting on anything) is certain to release.
__device__ __forceinline__ bool cache_try_lock_warp(uint32_t* word, uint32_t* meta, int lane, int tries = 1) {
    uint32_t old = kCacheLockBit;
    for (int t = 0; t < tries; ++t) {
        if (lane == 0) old = atomicOr(word, kCacheLockBit);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (!(old & kCacheLockBit)) break;
        if (t + 1 < tries) __nanosleep(200);
    }
    if (old & kCacheLockBit) return false;
    __threadfence();  // acquire: entry reads below must not be satisfied before the lock is held
    *meta = old;
    return true;
}
__device__ __forceinline__ void cache_unlock_warp(uint32_t* word, uint32_t meta, int lane) {
    __threadfence();  // release: entry writes of every lane are visible before the word changes
    __syncwarp();
    if (lane == 0) atomicExch(word, meta & ~kCacheLockBit);
}

// evalcache.cc:49-121.  The row's n values come from `get(k)` = element lane + 32 k (k < 6: n <= 164), so that
// a decode can store from its registers.  Returns true when the entry is present afterwards (stored or
// refreshed), false when dropped (n > 164 or bundle busy).
constexpr int kCacheRowPerLane = (NSB_CACHE_MAX_MOVES + 31) / 32;  // 6 row elements per lane at most

template <typename Get>
__device__ __forceinline__ bool cache_store_warp_from(const DeviceCache& c, uint64_t hash, int n, Get get, float win,
                                                      float draw, int lane) {
    if (n > NSB_CACHE_MAX_MOVES) return false;
    const unsigned long long bundle = hash % c.num_bundles;
    uint32_t* word = c.meta + bundle;
    uint32_t meta;
    if (!cache_try_lock_warp(word, &meta, lane)) return false;
    CacheEntry* base = c.entries + bundle * 3;
    bool used = false, same = false;
    if (lane < 3) {
        const uint32_t slot = cache_order_slot(meta, lane);
        used = cache_used(meta, slot) != 0;
        const CacheEntry* e = base + slot;
        same = used && __ldcg(&e->hash) == hash && __ldcg(&e->n) == (uint32_t)n;
    }
    const uint32_t same_mask = __ballot_sync(0xffffffffu, same);
    const uint32_t free_mask = __ballot_sync(0xffffffffu, lane < 3 && !used);
    if (same_mask) {  // already there: refresh its recency, keep its data (evalcache.cc:71-88)
        cache_unlock_warp(word, cache_move_to_front(meta, __ffs(same_mask) - 1), lane);
        return true;
    }
    const int target = free_mask ? __ffs(free_mask) - 1 : 2;  // first unused entry, else the last one of the list
    const uint32_t slot = cache_order_slot(meta, target);
    CacheEntry* e = base + slot;
#pragma unroll
    for (int k = 0; k < kCacheRowPerLane; ++k) {
        const int j = lane + 32 * k;
        if (j < n) e->policy[j] = get(k);
    }
    if (lane == 0) {
        e->hash = hash;
        e->n = (uint32_t)n;
        e->win = win;
        e->draw = draw;
    }
    cache_unlock_warp(word, cache_move_to_front(meta | (1u << (8 + slot)), target), lane);
    return true;
}

// the same with the row in (global or shared) memory
__device__ __forceinline__ bool cache_store_warp(const DeviceCache& c, uint64_t hash, int n, const float* row, float win,
                                                 float draw, int lane) {
    return cache_store_warp_from(c, hash, n, [&](int k) { return row[lane + 32 * k]; }, win, draw, lane);
}

// evalcache.cc:123-169 + the caller's move-count check (searchworker.cc:545-556): returns true and
// fills row[0..expected_n), *win, *draw when an entry with this hash exists AND has expected_n moves.
// A hash match with a different move count still refreshes the entry's recency, as in the reference.
constexpr int kCacheProbeTries = 64;

// The row comes back in registers: lane l holds element l + 32 k in vals[k] (0 beyond expected_n).
__device__ __forceinline__ bool cache_load_warp(const DeviceCache& c, uint64_t hash, int expected_n, float (&vals)[kCacheRowPerLane],
                                                float* win, float* draw, int lane) {
    const unsigned long long bundle = hash % c.num_bundles;
    uint32_t* word = c.meta + bundle;
    uint32_t meta;
    if (!cache_try_lock_warp(word, &meta, lane, kCacheProbeTries)) return false;
    const CacheEntry* base = c.entries + bundle * 3;
    bool match = false;
    uint32_t n_e = 0, slot = 0;
    if (lane < 3) {
        slot = cache_order_slot(meta, lane);
        const CacheEntry* e = base + slot;
        match = cache_used(meta, slot) != 0 && __ldcg(&e->hash) == hash;
        n_e = __ldcg(&e->n);
    }
    const uint32_t match_mask = __ballot_sync(0xffffffffu, match);
    if (!match_mask) {
        cache_unlock_warp(word, meta, lane);
        return false;
    }
    const int at = __ffs(match_mask) - 1;
    n_e = __shfl_sync(0xffffffffu, n_e, at);
    slot = __shfl_sync(0xffffffffu, slot, at);
    const bool ok = (int)n_e == expected_n;
    if (ok) {
        const CacheEntry* e = base + slot;
#pragma unroll
        for (int k = 0; k < kCacheRowPerLane; ++k) {
            const int j = lane + 32 * k;
            vals[k] = j < expected_n ? __ldcg(&e->policy[j]) : 0.f;
        }
        *win = __ldcg(&e->win);
        *draw = __ldcg(&e->draw);
    }
    cache_unlock_warp(word, cache_move_to_front(meta, at), lane);
    return ok;
}

}  // namespace nsb
