// extern "C" handle around the REFERENCE's own EvalCache (compiled from
// /root/reference/src/mcts/evalcache.cc in place; see oracle/Makefile).  Test infrastructure: pins
// nsb_oracle_cache_* (and through it the host and device caches) to the reference's behaviour.
#include <cstddef>
#include <cstdint>
#include <cstring>
#define private public  // read NumBundle (a private const) - this file is a test probe, not product code
#include "mcts/evalcache.h"
#undef private
using nshogi::engine::mcts::EvalCache;
extern "C" {
void* nsb_ref_evalcache_make(size_t memory_mb) { return new EvalCache(memory_mb); }
void nsb_ref_evalcache_free(void* h) { delete static_cast<EvalCache*>(h); }
uint64_t nsb_ref_evalcache_num_bundles(void* h) { return static_cast<EvalCache*>(h)->NumBundle; }
int nsb_ref_evalcache_store(void* h, uint64_t hash, uint32_t n, const float* row, float win, float draw) {
    return static_cast<EvalCache*>(h)->store(hash, (uint16_t)n, row, win, draw) ? 1 : 0;
}
// EvalCache::load + the caller's move-count check (reference src/mcts/searchworker.cc:545-556)
int nsb_ref_evalcache_load(void* h, uint64_t hash, uint32_t expected_n, float* row, float* win, float* draw) {
    EvalCache::EvalInfo info;
    const nshogi::core::State st(hash);
    if (!static_cast<EvalCache*>(h)->load(st, &info)) return 0;
    if (info.NumMoves != expected_n) return 0;
    std::memcpy(row, info.Policy, sizeof(float) * info.NumMoves);
    *win = info.WinRate;
    *draw = info.DrawRate;
    return 1;
}
}
