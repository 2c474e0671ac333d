// tree_learner_factory.cpp — link-time replacement for reference src/treelearner/tree_learner.cpp (the
// factory, 55 lines): identical dispatch for cpu/gpu learners, but ("serial","cuda") returns the
// B200-native learner instead of CUDASingleGPUTreeLearner.  In an in-tree integration this is a
// ONE-LINE change of tree_learner.cpp:49.
#include <LightGBM/tree_learner.h>

#include <string>

#include "b200_tree_learner.hpp"
#include "gpu_tree_learner.h"
#include "linear_tree_learner.h"
#include "parallel_tree_learner.h"
#include "serial_tree_learner.h"

namespace LightGBM {

TreeLearner* TreeLearner::CreateTreeLearner(const std::string& learner_type, const std::string& device_type,
                                            const Config* config, const bool boosting_on_cuda) {
  if (device_type == "cuda") {
    if (learner_type == "serial") return new B200TreeLearner(config, boosting_on_cuda);
    Log::Fatal("lgbm_b200 supports tree_learner=serial on a single machine (multi-GPU is configured through num_gpu).");
  }
  const bool gpu = device_type == "gpu";
  if (device_type == "cpu" || gpu) {
    if (learner_type == "serial") {
      if (config->linear_tree) return gpu ? static_cast<TreeLearner*>(new LinearTreeLearner<GPUTreeLearner>(config))
                                          : static_cast<TreeLearner*>(new LinearTreeLearner<SerialTreeLearner>(config));
      return gpu ? static_cast<TreeLearner*>(new GPUTreeLearner(config)) : static_cast<TreeLearner*>(new SerialTreeLearner(config));
    }
    if (learner_type == "feature") return gpu ? static_cast<TreeLearner*>(new FeatureParallelTreeLearner<GPUTreeLearner>(config))
                                              : static_cast<TreeLearner*>(new FeatureParallelTreeLearner<SerialTreeLearner>(config));
    if (learner_type == "data") return gpu ? static_cast<TreeLearner*>(new DataParallelTreeLearner<GPUTreeLearner>(config))
                                           : static_cast<TreeLearner*>(new DataParallelTreeLearner<SerialTreeLearner>(config));
    if (learner_type == "voting") return gpu ? static_cast<TreeLearner*>(new VotingParallelTreeLearner<GPUTreeLearner>(config))
                                             : static_cast<TreeLearner*>(new VotingParallelTreeLearner<SerialTreeLearner>(config));
  }
  return nullptr;
}

}  // namespace LightGBM
