// Link-time stub for src/treelearner/linear_tree_learner.cpp, which needs Eigen (an un-vendored,
// empty submodule in /root/reference).  Linear trees are out of scope; the factory
// in tree_learner.cpp still references the two instantiations, so give them Fatal bodies.
#include <LightGBM/utils/log.h>
#include "linear_tree_learner.h"  // found via -I/root/reference/src/treelearner

namespace LightGBM {
#define LT_FATAL() Log::Fatal("linear_tree is not available in this oracle build (Eigen absent)")
template <typename T> void LinearTreeLearner<T>::Init(const Dataset*, bool) { LT_FATAL(); }
template <typename T> void LinearTreeLearner<T>::InitLinear(const Dataset*, const int) { LT_FATAL(); }
template <typename T> Tree* LinearTreeLearner<T>::Train(const score_t*, const score_t*, bool) { LT_FATAL(); return nullptr; }
template <typename T> void LinearTreeLearner<T>::GetLeafMap(Tree*) const { LT_FATAL(); }
template <typename T> Tree* LinearTreeLearner<T>::FitByExistingTree(const Tree*, const score_t*, const score_t*) const { LT_FATAL(); return nullptr; }
template <typename T> Tree* LinearTreeLearner<T>::FitByExistingTree(const Tree*, const std::vector<int>&, const score_t*, const score_t*) const { LT_FATAL(); return nullptr; }
template class LinearTreeLearner<SerialTreeLearner>;
template class LinearTreeLearner<GPUTreeLearner>;
}  // namespace LightGBM
