// Stand-in for the reference's src/infer/infer.h (reference src/infer/infer.h:19-32) so that the
// host mirror compiles without the reference tree.  When building INSIDE the reference (the
// intended use) this directory is left off the include path and the reference's
// own header is picked up instead; the interface below is the contract both sides agree on.
#ifndef NSB_SHIM_INFER_INFER_H
#define NSB_SHIM_INFER_INFER_H
#include <cstddef>

#include <nshogi/ml/featurebitboard.h>

namespace nshogi {
namespace engine {
namespace infer {

class Infer {
 public:
    virtual ~Infer() = default;
    virtual void computeNonBlocking(const ml::FeatureBitboard* Features, std::size_t BatchSize,
                                    float* DstPolicy, float* DstWinRate, float* DstDrawRate) = 0;
    virtual void computeBlocking(const ml::FeatureBitboard* Features, std::size_t BatchSize,
                                 float* DstPolicy, float* DstWinRate, float* DstDrawRate) = 0;
    virtual void await() = 0;
    virtual bool isComputing() = 0;
};

} // namespace infer
} // namespace engine
} // namespace nshogi
#endif
