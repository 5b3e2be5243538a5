// host/shim/nshogi/ml/common.h - build-time stand-in for libnshogi's <nshogi/ml/common.h>: the policy size
// (reference src/infer/trt.cc:205; 27 move planes x 81 squares, src/mcts/evaluationworker.cc:166).
#ifndef NSB_HOST_SHIM_NSHOGI_ML_COMMON_H
#define NSB_HOST_SHIM_NSHOGI_ML_COMMON_H
#include <cstddef>
namespace nshogi {
namespace core {
constexpr std::size_t NumSquares = 81;
} // namespace core
namespace ml {
constexpr std::size_t MoveIndexMax = 27 * core::NumSquares;
} // namespace ml
} // namespace nshogi
#endif
