// b200_tree_learner.cpp — see b200_tree_learner.hpp.  ~200 lines: Dataset -> LGBMB200_Layout, Config ->
// LGBMB200_Config, split records -> Tree::Split replay.  Everything numeric happens in liblgbm_b200.so.
#include "b200_tree_learner.hpp"

#include <LightGBM/bin.h>
#include <LightGBM/utils/log.h>

#include <LightGBM/utils/openmp_wrapper.h>

#include <algorithm>
#include <cmath>

namespace LightGBM {

LGBMB200_Config B200TreeLearner::ToB200Config(const Config* c) {
  LGBMB200_Config o;
  o.num_leaves = c->num_leaves;
  o.max_depth = c->max_depth;
  o.min_data_in_leaf = c->min_data_in_leaf;
  o.gpu_device_id = c->gpu_device_id;
  o.min_sum_hessian_in_leaf = c->min_sum_hessian_in_leaf;
  o.lambda_l1 = c->lambda_l1;
  o.lambda_l2 = c->lambda_l2;
  o.min_gain_to_split = c->min_gain_to_split;
  o.max_delta_step = c->max_delta_step;
  o.path_smooth = c->path_smooth;
  o.use_cuda_graph = 1;
  o.reserved = 0;
  // quantized-gradient training (config.h:626-651)
  o.use_quantized_grad = c->use_quantized_grad ? 1 : 0;
  o.num_grad_quant_bins = c->num_grad_quant_bins;
  o.quant_train_renew_leaf = c->quant_train_renew_leaf ? 1 : 0;
  o.stochastic_rounding = c->stochastic_rounding ? 1 : 0;
  o.seed = c->seed;
  o.pad_ = 0;
  return o;
}

void B200TreeLearner::Check(int ret) const {
  if (ret != 0) Log::Fatal("lgbm_b200: %s", LGBMB200_GetLastError());
}

// Config fields that change the reference learner's split choice but are outside this hot-path library (SURVEY.md §8,
// DESIGN.md §7): refuse them loudly instead of silently training a different model than device_type=cpu would.
void B200TreeLearner::CheckSupported(const Config* config) {
  if (!config->monotone_constraints.empty()) Log::Fatal("lgbm_b200: monotone constraints are not supported");
  if (config->extra_trees) Log::Fatal("lgbm_b200: extra_trees is not supported");
  if (config->feature_fraction_bynode < 1.0) Log::Fatal("lgbm_b200: feature_fraction_bynode is not supported");
  if (config->cegb_tradeoff < 1.0 || config->cegb_penalty_split > 0.0) Log::Fatal("lgbm_b200: CEGB is not supported");
  for (double v : config->feature_contri)
    if (v != 1.0) Log::Fatal("lgbm_b200: feature_contri (per-feature gain penalty, feature_histogram.hpp:174) is not supported");
  if (!config->interaction_constraints.empty()) Log::Fatal("lgbm_b200: interaction_constraints are not supported");
  if (config->linear_tree) Log::Fatal("lgbm_b200: linear_tree is not supported");
  if (config->num_machines > 1) Log::Fatal("lgbm_b200: num_machines > 1 is not supported (multi-GPU runs inside one machine)");
  if (config->max_bin > 255) Log::Fatal("lgbm_b200: max_bin > 255 is not supported");
}

B200TreeLearner::B200TreeLearner(const Config* config) : config_(config), col_sampler_(config) {
  CheckSupported(config);
  LGBMB200_Config c = ToB200Config(config);
  Check(LGBMB200_LearnerCreate(&c, &handle_));
}

B200TreeLearner::~B200TreeLearner() {
  if (handle_) LGBMB200_LearnerFree(handle_);
}

void B200TreeLearner::Init(const Dataset* train_data, bool is_constant_hessian) {
  train_data_ = train_data;
  num_data_ = train_data->num_data();
  num_features_ = train_data->num_features();
  const int C = train_data->num_feature_groups();
  std::vector<int32_t> col(num_features_), lo(num_features_), nbin(num_features_), mfb(num_features_),
      dbin(num_features_), miss(num_features_), real(num_features_);
  for (int f = 0; f < num_features_; ++f) {
    const BinMapper* bm = train_data->FeatureBinMapper(f);
    const int g = train_data->Feature2Group(f);
    if (bm->bin_type() != BinType::NumericalBin) Log::Fatal("lgbm_b200: categorical features are not supported");
    if (train_data->IsMultiGroup(g)) Log::Fatal("lgbm_b200: multi-value (sparse row-wise) groups are not supported");
    col[f] = g;
    lo[f] = static_cast<int32_t>(train_data->feature_min_bin(f));
    nbin[f] = bm->num_bin();
    mfb[f] = static_cast<int32_t>(bm->GetMostFreqBin());
    dbin[f] = static_cast<int32_t>(bm->GetDefaultBin());
    miss[f] = static_cast<int32_t>(bm->missing_type());
    real[f] = train_data->RealFeatureIndex(f);
  }
  // stored group values, row-major [num_data x num_groups] (what FeatureGroup::PushData wrote).  Row blocks are
  // dealt to OpenMP threads; every thread owns one iterator per group, so a 256-row x C tile is written while it is
  // cache-resident (one thread walking a whole column would touch a new cache line of the matrix per byte).
  for (int g = 0; g < C; ++g)
    if (train_data->FeatureGroupNumBin(g) > 256) Log::Fatal("lgbm_b200: a feature group has more than 256 bins (use max_bin <= 255)");
  std::vector<uint8_t> bins(static_cast<size_t>(num_data_) * C);
  constexpr int kTileRows = 256;
  const int num_tiles = (num_data_ + kTileRows - 1) / kTileRows;
  bool iter_failed = false;
#pragma omp parallel num_threads(OMP_NUM_THREADS())
  {
    std::vector<std::unique_ptr<BinIterator>> its(C);
    for (int g = 0; g < C; ++g) {
      its[g].reset(train_data->FeatureGroupIterator(g));
      if (!its[g]) {
#pragma omp critical
        iter_failed = true;
      } else {
        its[g]->Reset(0);
      }
    }
#pragma omp barrier
    if (!iter_failed) {
#pragma omp for schedule(static)
      for (int tile = 0; tile < num_tiles; ++tile) {
        const int r0 = tile * kTileRows, r1 = std::min(num_data_, r0 + kTileRows);
        for (int g = 0; g < C; ++g) {
          BinIterator* it = its[g].get();
          for (int i = r0; i < r1; ++i) bins[static_cast<size_t>(i) * C + g] = static_cast<uint8_t>(it->RawGet(i));
        }
      }
    }
  }
  if (iter_failed) Log::Fatal("lgbm_b200: cannot iterate a feature group");
  LGBMB200_Layout lay;
  lay.num_data = num_data_; lay.num_columns = C; lay.num_features = num_features_;
  lay.feat_column = col.data(); lay.feat_lo = lo.data(); lay.feat_num_bin = nbin.data();
  lay.feat_most_freq_bin = mfb.data(); lay.feat_default_bin = dbin.data(); lay.feat_missing_type = miss.data();
  lay.feat_real_index = real.data();
  Check(LGBMB200_LearnerInit(handle_, &lay, bins.data(), is_constant_hessian ? 1 : 0));
  col_sampler_.SetTrainingData(train_data);
}

void B200TreeLearner::ResetTrainingData(const Dataset* train_data, bool is_constant_hessian) {
  Init(train_data, is_constant_hessian);
}

void B200TreeLearner::ResetIsConstantHessian(bool is_constant_hessian) {
  Check(LGBMB200_LearnerSetConstantHessian(handle_, is_constant_hessian ? 1 : 0));
}

void B200TreeLearner::ResetConfig(const Config* config) {
  CheckSupported(config);
  config_ = config;
  LGBMB200_Config c = ToB200Config(config);
  Check(LGBMB200_LearnerResetConfig(handle_, &c));
  if (train_data_ != nullptr) col_sampler_.SetConfig(config);
}

void B200TreeLearner::SetForcedSplit(const Json* forced_split_json) {
  if (forced_split_json != nullptr && !forced_split_json->is_null()) Log::Fatal("lgbm_b200: forced splits are not supported");
}

Tree* B200TreeLearner::Train(const score_t* gradients, const score_t* hessians, bool /*is_first_tree*/) {
  // ColSampler by tree (serial_tree_learner.cpp:297)
  col_sampler_.ResetByTree();
  const std::vector<int8_t>& used = col_sampler_.is_feature_used_bytree();
  if (config_->feature_fraction < 1.0) {
    Check(LGBMB200_LearnerSetFeatureMask(handle_, reinterpret_cast<const uint8_t*>(used.data())));
    mask_set_ = true;
  } else if (mask_set_) {
    Check(LGBMB200_LearnerSetFeatureMask(handle_, nullptr));
    mask_set_ = false;
  }
  const int NL = config_->num_leaves;
  std::vector<LGBMB200_Split> splits(NL);
  std::vector<double> leaf_value(NL), leaf_weight(NL);
  std::vector<int32_t> leaf_count(NL), leaf_depth(NL);
  LGBMB200_Tree t;
  t.num_leaves = 0; t.splits = splits.data(); t.leaf_value = leaf_value.data(); t.leaf_weight = leaf_weight.data();
  t.leaf_count = leaf_count.data(); t.leaf_depth = leaf_depth.data();
  // host gradients (boosting_on_gpu_ == false in a non-USE_CUDA build of GBDT, gbdt.cpp:110-135)
  Check(LGBMB200_LearnerTrain(handle_, gradients, hessians, /*on_device=*/0, &t));
  last_num_leaves_ = t.num_leaves;

  // replay through the reference's own Tree::Split (tree.cpp:65-79) so model text / predict are unchanged
  std::unique_ptr<Tree> tree(new Tree(NL, false, false));
  tree->SetLeafOutput(0, t.num_leaves > 1 ? 0.0 : leaf_value[0]);
  if (t.num_leaves == 1) tree->SetLeafOutput(0, leaf_value[0]);
  for (int i = 0; i < t.num_leaves - 1; ++i) {
    const LGBMB200_Split& s = splits[i];
    const int real_f = train_data_->RealFeatureIndex(s.feature);
    const double thr = train_data_->RealThreshold(s.feature, static_cast<uint32_t>(s.threshold));
    tree->Split(s.leaf, s.feature, real_f, static_cast<uint32_t>(s.threshold), thr, s.left_output, s.right_output,
                s.left_count, s.right_count, s.left_sum_hessian, s.right_sum_hessian,
                static_cast<float>(s.gain + config_->min_gain_to_split),
                train_data_->FeatureBinMapper(s.feature)->missing_type(), s.default_left != 0);
  }
  if (config_->use_quantized_grad && config_->quant_train_renew_leaf) {
    // RenewIntGradTreeOutput (gradient_discretizer.cpp:236-259) ran on the device: Tree::SetLeafOutput per leaf
    for (int i = 0; i < t.num_leaves; ++i) tree->SetLeafOutput(i, leaf_value[i]);
  }
  return tree.release();
}

Tree* B200TreeLearner::FitByExistingTree(const Tree*, const score_t*, const score_t*) const {
  Log::Fatal("lgbm_b200: refit (FitByExistingTree) is not supported");
  return nullptr;
}
Tree* B200TreeLearner::FitByExistingTree(const Tree*, const std::vector<int>&, const score_t*, const score_t*) const {
  Log::Fatal("lgbm_b200: refit (FitByExistingTree) is not supported");
  return nullptr;
}

void B200TreeLearner::SetBaggingData(const Dataset* subset, const data_size_t* used_indices, data_size_t num_data) {
  if (subset != nullptr) Log::Fatal("lgbm_b200: bagging with a subset Dataset is not supported (set bagging_fraction >= 0.5 or use GOSS)");
  Check(LGBMB200_LearnerSetBaggingData(handle_, used_indices, num_data, /*on_device=*/0));
}

void B200TreeLearner::AddPredictionToScore(const Tree* tree, double* out_score) const {
  if (tree->num_leaves() <= 1) return;
  std::vector<double> lv(tree->num_leaves());
  for (int i = 0; i < tree->num_leaves(); ++i) lv[i] = tree->LeafOutput(i);
  Check(LGBMB200_LearnerAddPredictionToScore(handle_, lv.data(), tree->num_leaves(), out_score, /*on_device=*/0));
}

void B200TreeLearner::RenewTreeOutput(Tree* tree, const ObjectiveFunction* obj,
                                      std::function<double(const label_t*, int)> residual_getter,
                                      data_size_t total_num_data, const data_size_t* bag_indices, data_size_t bag_cnt,
                                      const double* /*train_score*/) const {
  // same contract as SerialTreeLearner::RenewTreeOutput (serial_tree_learner.cpp:927-965), single machine
  if (obj == nullptr || !obj->IsRenewTreeOutput()) return;
  const int nl = tree->num_leaves();
  std::vector<int32_t> begin(nl), count(nl), indices(num_data_);
  Check(LGBMB200_LearnerGetPartition(handle_, begin.data(), count.data(), indices.data()));
  const data_size_t* bag_mapper = (total_num_data != num_data_) ? bag_indices : nullptr;
  (void)bag_cnt;
  for (int i = 0; i < nl; ++i) {
    if (count[i] <= 0) continue;
    const double out = obj->RenewTreeOutput(tree->LeafOutput(i), residual_getter, indices.data() + begin[i], bag_mapper, count[i]);
    tree->SetLeafOutput(i, out);
  }
}

}  // namespace LightGBM
