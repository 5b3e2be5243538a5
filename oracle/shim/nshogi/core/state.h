// Minimal stand-in for libnshogi's <nshogi/core/state.h>: the one member the reference's
// EvalCache::load uses (reference src/mcts/evalcache.cc:124: St.getHash()).  Test infrastructure.
#ifndef NSB_SHIM_NSHOGI_CORE_STATE_H
#define NSB_SHIM_NSHOGI_CORE_STATE_H
#include <cstdint>
namespace nshogi {
namespace core {
class State {
 public:
    explicit State(uint64_t Hash) : Hash_(Hash) {}
    uint64_t getHash() const { return Hash_; }

 private:
    uint64_t Hash_;
};
} // namespace core
} // namespace nshogi
#endif
