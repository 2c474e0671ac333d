// b200_tree_learner.hpp — the reference-side binding a LightGBM maintainer adds to use liblgbm_b200.so
// under device_type=cuda.  It derives from the reference's own plug-in interface `TreeLearner`
// (reference include/LightGBM/tree_learner.h:27-114) and forwards every virtual to the C-ABI of
// include/lgbm_b200.h.  This file INCLUDES reference headers (it is compiled against /root/reference,
// like any in-tree learner would be) but contains no reference code.
#ifndef INTEGRATION_B200_TREE_LEARNER_HPP_
#define INTEGRATION_B200_TREE_LEARNER_HPP_

#include <LightGBM/config.h>
#include <LightGBM/dataset.h>
#include <LightGBM/objective_function.h>
#include <LightGBM/tree.h>
#include <LightGBM/tree_learner.h>

#include <memory>
#include <vector>

#include "../include/lgbm_b200.h"
#include "col_sampler.hpp"   // reference src/treelearner/col_sampler.hpp (header-only), found via -I

namespace LightGBM {

class B200TreeLearner : public TreeLearner {
 public:
  explicit B200TreeLearner(const Config* config, bool boosting_on_cuda = false);
  ~B200TreeLearner() override;

  void Init(const Dataset* train_data, bool is_constant_hessian) override;
  void ResetIsConstantHessian(bool is_constant_hessian) override;
  void ResetBoostingOnGPU(const bool boosting_on_gpu) override { on_device_ = boosting_on_gpu; }
  void ResetTrainingData(const Dataset* train_data, bool is_constant_hessian) override;
  void ResetConfig(const Config* config) override;
  void SetForcedSplit(const Json* forced_split_json) override;
  Tree* Train(const score_t* gradients, const score_t* hessians, bool is_first_tree) override;
  Tree* FitByExistingTree(const Tree* old_tree, const score_t* gradients, const score_t* hessians) const override;
  Tree* FitByExistingTree(const Tree* old_tree, const std::vector<int>& leaf_pred, const score_t* gradients,
                          const score_t* hessians) const override;
  void SetBaggingData(const Dataset* subset, const data_size_t* used_indices, data_size_t num_data) override;
  void AddPredictionToScore(const Tree* tree, double* out_score) const override;
  void RenewTreeOutput(Tree* tree, const ObjectiveFunction* obj, std::function<double(const label_t*, int)> residual_getter,
                       data_size_t total_num_data, const data_size_t* bag_indices, data_size_t bag_cnt,
                       const double* train_score) const override;

 private:
  static LGBMB200_Config ToB200Config(const Config* config);
  static void CheckSupported(const Config* config);
  void Check(int ret) const;

  // num_gpu > 1 (config.h:1126; gpu_device_id_list :1132): features are sharded over the GPUs of this box, one
  // library learner per device inside this process, driven by one host thread each (the shape of the reference's
  // own multi-GPU mode, cuda_nccl_topology.hpp:177-188).  handle_ is rank 0's learner.
  std::vector<LGBMB200_LearnerHandle> handles_;
  std::vector<int> devices_;
  std::vector<int> feat_begin_;          // inner feature range [feat_begin_[r], feat_begin_[r + 1]) of rank r
  template <typename F> void ForEachRank(F&& f) const;

  const Config* config_;
  const Dataset* train_data_ = nullptr;
  LGBMB200_LearnerHandle handle_ = nullptr;
  ColSampler col_sampler_;
  int num_data_ = 0;
  int num_features_ = 0;
  int last_num_leaves_ = 0;
  bool mask_set_ = false;
  // boosting_on_gpu_ of GBDT (gbdt.cpp:110-135; only ever true in a -DUSE_CUDA build of the reference host code, whose
  // CUDA objectives and CUDAScoreUpdater then hand DEVICE gradients / scores to Train / AddPredictionToScore)
  bool on_device_ = false;
};

}  // namespace LightGBM
#endif  // INTEGRATION_B200_TREE_LEARNER_HPP_
