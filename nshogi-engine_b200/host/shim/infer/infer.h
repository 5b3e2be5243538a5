// Stand-in for the reference's executor interface so that the host mirror (infer_b200.h, leaf_pipeline.h, the
// benches) compiles without the reference tree.  When building INSIDE the reference - the intended use, see
// INTEGRATION.md - this directory is left off the include path and the reference's own src/infer/infer.h is picked
// up instead; tests/test_host_cpp.py compiles infer_b200.h against that header when the tree is present.
//
// The contract (reference src/infer/infer.h:19-32, callers src/evaluate/evaluator.h:32-48):
//   * one instance per evaluator thread, never shared; all four calls from that thread;
//   * the caller owns the buffers, sized for the executor's maximum batch; sample i lives at Features + 86 i,
//     DstPolicy + 2187 i (raw logits, plane-major), DstWinRate / DstDrawRate hold probabilities;
//   * computeNonBlocking only enqueues; inputs stay untouched and outputs are undefined until await().
#ifndef NSB_SHIM_INFER_INFER_H
#define NSB_SHIM_INFER_INFER_H
#include <cstddef>

#include <nshogi/ml/featurebitboard.h>

namespace nshogi::engine::infer {

class Infer {
 public:
    using Features = const ml::FeatureBitboard*;

    virtual ~Infer() = default;

    // enqueue one batch; returns at once
    virtual void computeNonBlocking(Features In, std::size_t Batch, float* Policy, float* WinRate, float* DrawRate) = 0;
    // the same, returning when the results are in the output arrays
    virtual void computeBlocking(Features In, std::size_t Batch, float* Policy, float* WinRate, float* DrawRate) = 0;
    // wait for the batch enqueued last
    virtual void await() = 0;
    // true while that batch is still running
    virtual bool isComputing() = 0;
};

}  // namespace nshogi::engine::infer
#endif
