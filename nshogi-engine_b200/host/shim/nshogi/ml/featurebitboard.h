// host/shim/nshogi/ml/featurebitboard.h - build-time stand-in for libnshogi's <nshogi/ml/featurebitboard.h>
// when the library is not installed (it is un-vendored: SURVEY.md section 8c).  infer::Infer's signature
// (reference src/infer/infer.h:24-29) names ml::FeatureBitboard; the executor only needs its size and
// alignment - 16 bytes, two 64-bit words, the layout src/cuda/extractbit.cu:20-37 reads.  With the real
// library on the include path this directory is simply left off it.
#ifndef NSB_HOST_SHIM_NSHOGI_ML_FEATUREBITBOARD_H
#define NSB_HOST_SHIM_NSHOGI_ML_FEATUREBITBOARD_H
#include <cstdint>
namespace nshogi {
namespace ml {
struct alignas(16) FeatureBitboard {
    uint64_t Words[2];
};
static_assert(sizeof(FeatureBitboard) == 16 && alignof(FeatureBitboard) == 16, "16-byte packed plane");
} // namespace ml
} // namespace nshogi
#endif
