// b200_tree_learner.cpp — see b200_tree_learner.hpp.  ~200 lines: Dataset -> LGBMB200_Layout, Config ->
// LGBMB200_Config, split records -> Tree::Split replay.  Everything numeric happens in liblgbm_b200.so.
#include "b200_tree_learner.hpp"

#include <LightGBM/bin.h>
#include <LightGBM/utils/log.h>

#include <LightGBM/utils/openmp_wrapper.h>

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <thread>

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

B200TreeLearner::B200TreeLearner(const Config* config, bool boosting_on_cuda)
    : config_(config), col_sampler_(config), on_device_(boosting_on_cuda) {
  CheckSupported(config);
  // devices: gpu_device_id_list ("0,1,2,...") if given, else gpu_device_id .. gpu_device_id + num_gpu - 1
  const int W = std::max(1, config->num_gpu);
  if (W > 8) Log::Fatal("lgbm_b200: num_gpu > 8 is not supported");
  if (!config->gpu_device_id_list.empty()) {
    for (const std::string& tok : Common::Split(config->gpu_device_id_list.c_str(), ',')) devices_.push_back(std::atoi(tok.c_str()));
    if (static_cast<int>(devices_.size()) != W) Log::Fatal("lgbm_b200: gpu_device_id_list must name num_gpu devices");
  } else {
    for (int r = 0; r < W; ++r) devices_.push_back(std::max(0, config->gpu_device_id) + r);
  }
  handles_.assign(W, nullptr);
  for (int r = 0; r < W; ++r) {
    LGBMB200_Config c = ToB200Config(config);
    if (W > 1 || config->gpu_device_id >= 0) c.gpu_device_id = devices_[r];
    Check(LGBMB200_LearnerCreate(&c, &handles_[r]));
  }
  handle_ = handles_[0];
}

B200TreeLearner::~B200TreeLearner() {
  for (LGBMB200_LearnerHandle h : handles_) if (h) LGBMB200_LearnerFree(h);
}

// run f(rank) for every rank: concurrently (one host thread per device) when there are several — Train blocks inside
// the in-kernel NVLink exchange until every rank has joined
template <typename F>
void B200TreeLearner::ForEachRank(F&& f) const {
  const int W = static_cast<int>(handles_.size());
  if (W == 1) { f(0); return; }
  std::vector<std::thread> th;
  std::vector<std::string> err(W);
  for (int r = 0; r < W; ++r) th.emplace_back([&, r]() {
    try { f(r); } catch (const std::exception& e) { err[r] = e.what(); } catch (...) { err[r] = "unknown error"; }
  });
  for (auto& t : th) t.join();
  for (int r = 0; r < W; ++r) if (!err[r].empty()) Log::Fatal("lgbm_b200 (GPU rank %d): %s", r, err[r].c_str());
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
  // feature-shard plan: contiguous runs of column groups per rank, in units of 64 columns (one histogram column-group
  // set) while there are enough of them; inner features are numbered group by group, so a rank's features are contiguous
  const int W = static_cast<int>(handles_.size());
  std::vector<int> col_begin(W + 1, 0);
  {
    const int unit = (C >= 64 * W) ? 64 : (C >= 32 * W ? 32 : 1);
    const int units = (C + unit - 1) / unit;
    for (int r = 0; r < W; ++r) col_begin[r + 1] = std::min(C, (units * (r + 1) / W) * unit);
    col_begin[W] = C;
  }
  feat_begin_.assign(W + 1, 0);
  for (int r = 0; r < W; ++r) {
    int f = feat_begin_[r];
    while (f < num_features_ && col[f] < col_begin[r + 1]) ++f;
    feat_begin_[r + 1] = f;
    if (W > 1 && (col_begin[r + 1] == col_begin[r] || f == feat_begin_[r])) Log::Fatal("lgbm_b200: num_gpu = %d is more than the %d feature groups can feed", W, C);
  }
  std::vector<std::vector<uint8_t>> bins(W);
  for (int r = 0; r < W; ++r) bins[r].resize(static_cast<size_t>(num_data_) * (col_begin[r + 1] - col_begin[r]));
  std::vector<int> owner(C);
  for (int r = 0; r < W; ++r) for (int g = col_begin[r]; g < col_begin[r + 1]; ++g) owner[g] = r;
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
          const int r = owner[g], Cr = col_begin[r + 1] - col_begin[r], gl = g - col_begin[r];
          uint8_t* dst = bins[r].data();
          for (int i = r0; i < r1; ++i) dst[static_cast<size_t>(i) * Cr + gl] = static_cast<uint8_t>(it->RawGet(i));
        }
      }
    }
  }
  if (iter_failed) Log::Fatal("lgbm_b200: cannot iterate a feature group");
  for (int r = 0; r < W; ++r) {
    const int f0 = feat_begin_[r], nf = feat_begin_[r + 1] - f0;
    std::vector<int32_t> lcol(nf);
    for (int f = 0; f < nf; ++f) lcol[f] = col[f0 + f] - col_begin[r];
    LGBMB200_Layout lay;
    lay.num_data = num_data_; lay.num_columns = col_begin[r + 1] - col_begin[r]; lay.num_features = nf;
    lay.feat_column = lcol.data(); lay.feat_lo = lo.data() + f0; lay.feat_num_bin = nbin.data() + f0;
    lay.feat_most_freq_bin = mfb.data() + f0; lay.feat_default_bin = dbin.data() + f0; lay.feat_missing_type = miss.data() + f0;
    lay.feat_real_index = real.data() + f0;
    Check(LGBMB200_LearnerInit(handles_[r], &lay, bins[r].data(), is_constant_hessian ? 1 : 0));
    std::vector<uint8_t>().swap(bins[r]);
  }
  if (W > 1) Check(LGBMB200_LearnersConnectLocal(handles_.data(), W, feat_begin_.data(), /*replicate_columns=*/1));
  col_sampler_.SetTrainingData(train_data);
}

void B200TreeLearner::ResetTrainingData(const Dataset* train_data, bool is_constant_hessian) {
  Init(train_data, is_constant_hessian);
}

void B200TreeLearner::ResetIsConstantHessian(bool is_constant_hessian) {
  for (LGBMB200_LearnerHandle h : handles_) Check(LGBMB200_LearnerSetConstantHessian(h, is_constant_hessian ? 1 : 0));
}

void B200TreeLearner::ResetConfig(const Config* config) {
  CheckSupported(config);
  config_ = config;
  for (size_t r = 0; r < handles_.size(); ++r) {
    LGBMB200_Config c = ToB200Config(config);
    if (handles_.size() > 1 || config->gpu_device_id >= 0) c.gpu_device_id = devices_[r];
    Check(LGBMB200_LearnerResetConfig(handles_[r], &c));
  }
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
    for (size_t r = 0; r < handles_.size(); ++r)
      Check(LGBMB200_LearnerSetFeatureMask(handles_[r], reinterpret_cast<const uint8_t*>(used.data()) + feat_begin_[r]));
    mask_set_ = true;
  } else if (mask_set_) {
    for (LGBMB200_LearnerHandle h : handles_) Check(LGBMB200_LearnerSetFeatureMask(h, nullptr));
    mask_set_ = false;
  }
  const int NL = config_->num_leaves;
  const int W = static_cast<int>(handles_.size());
  // every rank grows the identical tree (deterministic in-kernel exchange); rank 0's copy is replayed below
  std::vector<std::vector<LGBMB200_Split>> r_splits(W, std::vector<LGBMB200_Split>(NL));
  std::vector<std::vector<double>> r_value(W, std::vector<double>(NL)), r_weight(W, std::vector<double>(NL));
  std::vector<std::vector<int32_t>> r_count(W, std::vector<int32_t>(NL)), r_depth(W, std::vector<int32_t>(NL));
  std::vector<LGBMB200_Tree> r_tree(W);
  // host gradients (boosting_on_gpu_ == false: always so in a non-USE_CUDA build of GBDT, gbdt.cpp:110-135): every rank
  // copies them to its own GPU over its own PCIe link; device gradients (USE_CUDA build, CUDA objective) are read in place
  const int on_dev = on_device_ ? 1 : 0;
  ForEachRank([&](int r) {
    LGBMB200_Tree& tr = r_tree[r];
    tr.num_leaves = 0; tr.splits = r_splits[r].data(); tr.leaf_value = r_value[r].data(); tr.leaf_weight = r_weight[r].data();
    tr.leaf_count = r_count[r].data(); tr.leaf_depth = r_depth[r].data();
    if (LGBMB200_LearnerTrain(handles_[r], gradients, hessians, on_dev, &tr) != 0) throw std::runtime_error(LGBMB200_GetLastError());
  });
  const LGBMB200_Tree& t = r_tree[0];
  const std::vector<LGBMB200_Split>& splits = r_splits[0];
  const std::vector<double>& leaf_value = r_value[0];
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
#ifdef USE_CUDA
  const int idx_on_dev = 1;       // a -DUSE_CUDA GBDT hands the bag as a device array (bagging.hpp:110-111, goss.hpp:50-51)
#else
  const int idx_on_dev = 0;
#endif
  for (LGBMB200_LearnerHandle h : handles_) Check(LGBMB200_LearnerSetBaggingData(h, used_indices, num_data, idx_on_dev));
}

void B200TreeLearner::AddPredictionToScore(const Tree* tree, double* out_score) const {
  if (tree->num_leaves() <= 1) return;
  std::vector<double> lv(tree->num_leaves());
  for (int i = 0; i < tree->num_leaves(); ++i) lv[i] = tree->LeafOutput(i);
  Check(LGBMB200_LearnerAddPredictionToScore(handle_, lv.data(), tree->num_leaves(), out_score, on_device_ ? 1 : 0));
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
