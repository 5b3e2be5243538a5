// Minimal stand-in for libnshogi's <nshogi/ml/common.h>: the policy size the engine uses
// (reference src/infer/trt.cc:205, src/mcts/evaluationworker.cc:166 uses 27 * NumSquares).
#ifndef NSB_SHIM_NSHOGI_ML_COMMON_H
#define NSB_SHIM_NSHOGI_ML_COMMON_H
#include <cstddef>
namespace nshogi {
namespace core {
constexpr std::size_t NumSquares = 81;
}
namespace ml {
constexpr std::size_t MoveIndexMax = 27 * 81;
}
} // namespace nshogi
#endif
