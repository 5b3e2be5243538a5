// extern "C" handle around the REFERENCE's own Random executor (compiled from
// /root/reference/src/infer/random.cc in place; see oracle/Makefile).  Test infrastructure.
#include <cstddef>
#include <cstdint>
#include "infer/random.h"
extern "C" {
void* nsb_ref_random_make(uint64_t seed) { return new nshogi::engine::infer::Random(seed); }
void nsb_ref_random_free(void* h) { delete static_cast<nshogi::engine::infer::Random*>(h); }
void nsb_ref_random_fill(void* h, size_t n, float* policy, float* win, float* draw) {
    static_cast<nshogi::engine::infer::Random*>(h)->computeBlocking(nullptr, n, policy, win, draw);
}
}
