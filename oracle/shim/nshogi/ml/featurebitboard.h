// Minimal stand-in for libnshogi's <nshogi/ml/featurebitboard.h> (library is un-vendored and
// absent here).  Only what the leaf-evaluation path needs: the 16-byte POD whose
// bit layout is pinned by reference src/cuda/extractbit.cu:20-37 and its 2 x uint64 view
// (src/test/test_extractbit.cc:40-45, src/infer/trt.cc:57-58).  TEST/INTEGRATION SHIM ONLY: when
// the real library is installed this directory is simply left off the include path.
#ifndef NSB_SHIM_NSHOGI_ML_FEATUREBITBOARD_H
#define NSB_SHIM_NSHOGI_ML_FEATUREBITBOARD_H
#include <cstddef>
#include <cstdint>
namespace nshogi {
namespace ml {
struct alignas(16) FeatureBitboard {
    uint64_t Lo;
    uint64_t Hi;
};
static_assert(sizeof(FeatureBitboard) == 16, "FeatureBitboard must be 16 bytes");
} // namespace ml
} // namespace nshogi
#endif
