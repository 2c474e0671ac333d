/*
 * lgbm_b200.h — C-ABI of the B200-native histogram tree learner (the drop-in boundary).
 *
 * This is the library a LightGBM maintainer binds under `device_type=cuda`: the entry points mirror,
 * one for one, the internal plug-in interface `class TreeLearner`
 * (reference include/LightGBM/tree_learner.h:27-114) as implemented today by
 * `CUDASingleGPUTreeLearner` (reference src/treelearner/cuda/cuda_single_gpu_tree_learner.{hpp,cpp}).
 * Plain C types only: pointers + sizes, no C++/torch types.  Every function returns 0 on success and
 * -1 on error with the message retrievable through LGBMB200_GetLastError() — the same convention as
 * the reference C API (reference include/LightGBM/c_api.h:1646-1664, src/c_api.cpp:41-55).
 *
 * Ownership: the learner owns its device copy of the bin matrix, the histogram pool, the row-index
 * partition and all scratch; the caller owns grad/hess/score and the output buffers it passes in.
 * INTEGRATION.md shows the ~150-line C++ adapter (`class B200TreeLearner : public TreeLearner`) and
 * the factory line that slot this library into the reference.
 */
#ifndef LGBM_B200_H_
#define LGBM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LGBMB200_EXPORT __attribute__((visibility("default")))
#else
#define LGBMB200_EXPORT
#endif

typedef void* LGBMB200_LearnerHandle;

/* MissingType — reference include/LightGBM/bin.h:28-32 */
enum { LGBMB200_MISSING_NONE = 0, LGBMB200_MISSING_ZERO = 1, LGBMB200_MISSING_NAN = 2 };

/*
 * The fields of `Config` the hot path reads (reference include/LightGBM/config.h:210-712; list in
 * SURVEY.md §5 "Config / flags").  Replaces the `const Config*` argument of
 * TreeLearner::CreateTreeLearner / ResetConfig (reference tree_learner.h:49,104-107).
 */
typedef struct {
  int32_t num_leaves;
  int32_t max_depth;                /* <= 0: no limit */
  int32_t min_data_in_leaf;
  int32_t gpu_device_id;            /* CUDA device ordinal; -1 = current device */
  double  min_sum_hessian_in_leaf;
  double  lambda_l1;
  double  lambda_l2;
  double  min_gain_to_split;
  double  max_delta_step;
  double  path_smooth;
  int32_t use_cuda_graph;           /* 1: replay the whole per-tree launch sequence as one CUDA graph */
  int32_t reserved;                 /* bit 0: do NOT keep the column-major copy of the bin matrix used by the partition
                                       kernels (saves num_data*num_columns bytes of HBM); other bits must be 0 (developer A/B
                                       switches live in the LGBMB200_DEBUG environment variable, see learner.cu) */
  /* ---- quantized-gradient training (reference config.h:626-651; gradient_discretizer.cpp) */
  int32_t use_quantized_grad;       /* 1: discretize (g,h) to int8 per tree, integer histograms, integer split scan */
  int32_t num_grad_quant_bins;      /* config.h:638, default 4 */
  int32_t quant_train_renew_leaf;   /* 1: leaf outputs re-derived from the original gradients after the tree */
  int32_t stochastic_rounding;      /* 1: g/scale + U[0,1) (own counter-based generator: NOT the reference's per-thread
                                       mt19937 streams); 0: round half away from zero, reproducing the reference exactly */
  int32_t seed;                     /* Config::seed, used by stochastic rounding */
  int32_t pad_;
} LGBMB200_Config;

/*
 * The layout contract handed over at Init — what `Dataset` exposes to a tree learner through
 * FeatureBinMapper(i), Feature2Group(i), feature_min_bin(i), RealFeatureIndex(i)
 * (reference include/LightGBM/dataset.h:638-647,806-810,985-1000).
 * `bins` is the row-major [num_data x num_columns] uint8 matrix of stored group values written by
 * FeatureGroup::PushData (reference include/LightGBM/feature_group.h:253-267):
 *   0               : every feature of that column is at its most-frequent bin
 *   feat_lo[f] + i  : feature f is at bin (i + (feat_most_freq_bin[f] == 0))
 */
typedef struct {
  int32_t num_data;
  int32_t num_columns;
  int32_t num_features;
  const int32_t* feat_column;
  const int32_t* feat_lo;
  const int32_t* feat_num_bin;
  const int32_t* feat_most_freq_bin;
  const int32_t* feat_default_bin;
  const int32_t* feat_missing_type;
  const int32_t* feat_real_index;
} LGBMB200_Layout;

/* One chosen split = the SplitInfo the reference applies in SerialTreeLearner::SplitInner
 * (reference src/treelearner/split_info.hpp:22-56, serial_tree_learner.cpp:769-925). */
typedef struct {
  int32_t leaf;                 /* leaf that was split: left child keeps the id, right child = index+1 */
  int32_t feature;              /* inner feature index                                                  */
  int32_t threshold;            /* threshold_in_bin                                                     */
  int32_t default_left;
  int32_t left_count;           /* true row counts after the partition                                  */
  int32_t right_count;
  double  gain;
  double  left_sum_gradient, left_sum_hessian, left_output;
  double  right_sum_gradient, right_sum_hessian, right_output;
} LGBMB200_Split;

/* Flat POD tree returned by Train: the fields Tree::Split fills (reference include/LightGBM/tree.h:543-585).
 * All arrays are caller-allocated with capacity num_leaves (config) / num_leaves-1 for `splits`. */
typedef struct {
  int32_t num_leaves;           /* out: leaves actually grown                                           */
  LGBMB200_Split* splits;       /* out [num_leaves-1], in split order (node i of the reference Tree)     */
  double*  leaf_value;          /* out [num_leaves]                                                      */
  double*  leaf_weight;         /* out [num_leaves] sum of hessians                                      */
  int32_t* leaf_count;          /* out [num_leaves]                                                      */
  int32_t* leaf_depth;          /* out [num_leaves]                                                      */
  double   root_sum_gradient;   /* out                                                                   */
  double   root_sum_hessian;    /* out                                                                   */
  double   grad_scale;          /* out: GradientDiscretizer::grad_scale() of this tree (0 without use_quantized_grad) */
  double   hess_scale;          /* out: GradientDiscretizer::hess_scale()                                */
} LGBMB200_Tree;

LGBMB200_EXPORT const char* LGBMB200_GetLastError(void);

/* TreeLearner::CreateTreeLearner("serial","cuda",config) — reference src/treelearner/tree_learner.cpp:17-55 */
LGBMB200_EXPORT int LGBMB200_LearnerCreate(const LGBMB200_Config* config, LGBMB200_LearnerHandle* out);

/* TreeLearner::Init(const Dataset*, bool is_constant_hessian) — reference tree_learner.h:38;
 * cuda_single_gpu_tree_learner.cpp:36-93.  Copies the bin matrix (host pointer, or a device pointer such as the output of
 * LGBMB200_BinnerTransform) into the learner's own padded HBM layout.
 * is_constant_hessian != 0 promises that every hessian passed to Train equals hessians[0] (the objective's
 * IsConstantHessian(), gbdt.cpp:92); Train then reads only hessians[0] from a host buffer. */
LGBMB200_EXPORT int LGBMB200_LearnerInit(LGBMB200_LearnerHandle h, const LGBMB200_Layout* layout,
                                         const uint8_t* bins_host, int32_t is_constant_hessian);

/* TreeLearner::ResetConfig(const Config*) — reference tree_learner.h:56 */
LGBMB200_EXPORT int LGBMB200_LearnerResetConfig(LGBMB200_LearnerHandle h, const LGBMB200_Config* config);

/* TreeLearner::ResetIsConstantHessian(bool) — reference tree_learner.h:49 (GBDT::ResetTrainingData may swap the
 * objective): switches between the count-and-scale histogram kernel and the general one. */
LGBMB200_EXPORT int LGBMB200_LearnerSetConstantHessian(LGBMB200_LearnerHandle h, int32_t is_constant_hessian);

/* ColSampler by-tree mask (reference src/treelearner/col_sampler.hpp; serial_tree_learner.cpp:297):
 * feature_used[num_features] bytes on the host, or NULL for "all features". */
LGBMB200_EXPORT int LGBMB200_LearnerSetFeatureMask(LGBMB200_LearnerHandle h, const uint8_t* feature_used);

/* TreeLearner::SetBaggingData(subset, used_indices, num_data) — reference tree_learner.h:95-96,
 * cuda_single_gpu_tree_learner.cpp:448-451.  used_indices == NULL restores "all rows". */
LGBMB200_EXPORT int LGBMB200_LearnerSetBaggingData(LGBMB200_LearnerHandle h, const int32_t* used_indices,
                                                   int32_t num_used, int32_t on_device);

/* GOSSStrategy::Bagging (reference src/boosting/goss.hpp:30-77, :118-167) on the device: grad/hess are DEVICE arrays
 * [num_data], modified in place (kept small-gradient rows are multiplied by (n - top_k) / other_k); the kept rows become
 * the learner's bagging set exactly as if SetBaggingData had been called with their ascending indices.  One exact global
 * threshold (the reference takes one per OpenMP chunk) and a counter-based uniform per (seed, iteration, row): which
 * small-gradient rows are drawn is statistically, not bitwise, the reference's.  The caller applies the reference's
 * "no sampling during the first 1/learning_rate iterations" rule (goss.hpp:33) and passes is_constant_hessian = 0 at
 * Init (IsHessianChange(), goss.hpp:105). */
LGBMB200_EXPORT int LGBMB200_LearnerGossSample(LGBMB200_LearnerHandle h, float* grad_dev, float* hess_dev, double top_rate,
                                               double other_rate, int32_t seed, int32_t iteration, int32_t* out_bag_count);
/* GBDT::UpdateScore with a bagging set (reference gbdt.cpp:505-530): the partition only holds the bagged rows, the
 * out-of-bag rows are scored by predicting the tree on them.  This routes EVERY row of the training matrix through the
 * last tree on the device (bin matrix + split records) and adds its leaf value to score_dev[row].  Single GPU or
 * row-shard (every column present). */
LGBMB200_EXPORT int LGBMB200_LearnerAddPredictionAllRows(LGBMB200_LearnerHandle h, const double* leaf_value, int32_t num_leaves,
                                                         double* score_dev);
/* the current bagging set (host copy; test / RenewTreeOutput hook) */
LGBMB200_EXPORT int LGBMB200_LearnerGetBaggingData(LGBMB200_LearnerHandle h, int32_t* indices_host, int32_t num_indices);

/* Tree* TreeLearner::Train(const score_t* gradients, const score_t* hessians, bool is_first_tree) —
 * reference tree_learner.h:71; serial_tree_learner.cpp:182-248.  grad/hess are device pointers iff
 * on_device (== boosting_on_cuda, cuda_single_gpu_tree_learner.cpp:101-106), else host pointers. */
LGBMB200_EXPORT int LGBMB200_LearnerTrain(LGBMB200_LearnerHandle h, const float* gradients, const float* hessians,
                                          int32_t on_device, LGBMB200_Tree* out_tree);

/* TreeLearner::AddPredictionToScore(const Tree*, double* out_score) — reference tree_learner.h:85,
 * serial_tree_learner.h:100-115: score[row] += leaf_value[leaf(row)] over the partition left by the last
 * Train.  leaf_value is a host array (already shrunk by the caller, gbdt.cpp:421). */
LGBMB200_EXPORT int LGBMB200_LearnerAddPredictionToScore(LGBMB200_LearnerHandle h, const double* leaf_value,
                                                         int32_t num_leaves, double* score, int32_t on_device);

/* DataPartition::GetIndexOnLeaf (reference src/treelearner/data_partition.hpp:85-91) for every leaf of
 * the last tree: leaf_begin/leaf_count [num_leaves] and the row ids [num_data] (host outputs). */
LGBMB200_EXPORT int LGBMB200_LearnerGetPartition(LGBMB200_LearnerHandle h, int32_t* leaf_begin, int32_t* leaf_count,
                                                 int32_t* indices);

/* Test hook: the histogram of one leaf of the last tree as fp64 (grad,hess) pairs, [num_columns*256*2],
 * i.e. what ConstructHistograms + FixHistogram/Subtract left in the HistogramPool slot of that leaf. */
LGBMB200_EXPORT int LGBMB200_LearnerGetLeafHistogram(LGBMB200_LearnerHandle h, int32_t leaf, double* out);

/* Stand-alone ConstructHistograms (reference dataset.cpp:1293, serial_tree_learner.cpp:411-478) on an
 * arbitrary row set (host indices, or NULL = all rows) — used by the kernel-level parity tests and
 * the roofline micro-benchmark.  hist_out: host [num_columns*256*2] doubles. */
LGBMB200_EXPORT int LGBMB200_LearnerConstructHistogram(LGBMB200_LearnerHandle h, const float* gradients,
                                                       const float* hessians, int32_t on_device,
                                                       const int32_t* indices_host, int32_t num_indices,
                                                       double* hist_out, float* elapsed_ms);

/* Device-resident boosting helpers ("next" row f-1 of SURVEY.md §8): L2 gradients
 * (reference src/objective/regression_objective.hpp:127-142) and the number of kernels launched so far. */
LGBMB200_EXPORT int LGBMB200_L2Gradients(LGBMB200_LearnerHandle h, const double* score_dev, const float* label_dev,
                                         float* grad_dev, float* hess_dev, int32_t n);
/* Binary logloss gradients (reference src/objective/binary_objective.hpp:105-121, unweighted); labels in {0,1}. */
LGBMB200_EXPORT int LGBMB200_BinaryGradients(LGBMB200_LearnerHandle h, const double* score_dev, const float* label_dev,
                                             float* grad_dev, float* hess_dev, int32_t n, double sigmoid);
LGBMB200_EXPORT int64_t LGBMB200_LearnerKernelLaunches(LGBMB200_LearnerHandle h);
/* Sum of CUDA-event time spent in histogram-construction launches since the last reset (ms) and the
 * algorithmic bytes they covered; used by bench.py for the roofline of the dominant kernel. */
LGBMB200_EXPORT int LGBMB200_LearnerHistStats(LGBMB200_LearnerHandle h, int32_t reset, double* hist_ms,
                                              double* hist_rows, int64_t* hist_launches);
LGBMB200_EXPORT int LGBMB200_LearnerSetProfiling(LGBMB200_LearnerHandle h, int32_t enable);
/* CUDA-event time (ms, accumulated since the last HistStats reset) per launch kind, profiling mode only:
 * [0] unused, [1] prep+root_init, [2] part_flags, [3] part_count, [4] part_scatter, [5] pool memset, [6] hist,
 * [7] scan, [8] select. */
LGBMB200_EXPORT int LGBMB200_LearnerProfileByKind(LGBMB200_LearnerHandle h, double* ms_out_9);

/* ---- Multi-GPU, feature-shard (semantic model: FeatureParallelTreeLearner, reference
 * src/treelearner/feature_parallel_tree_learner.cpp:37-78 + SyncUpGlobalBestSplit, parallel_tree_learner.h:207-232).
 * One learner per GPU (one process per GPU, or several learners in one process (LGBMB200_LearnersConnectLocal)); every learner is Init-ed with ALL
 * rows and ITS column slice (feat_real_index stays global).  Per split the ranks exchange their two per-leaf best
 * candidates and the split's owner pushes the go-left flags — both through NVLink peer memory inside the kernels
 * (no NCCL call, no host round trip).  Bootstrap: every rank exports a 64-byte CUDA-IPC handle of its exchange
 * block; the caller all-gathers the handles (any transport: torch.distributed, MPI, a file) and hands each rank the
 * full list plus feature_offsets[world+1] (first global inner-feature id of every rank's slice). */
LGBMB200_EXPORT int LGBMB200_LearnerCommExport(LGBMB200_LearnerHandle h, uint8_t* handle_out_64);
LGBMB200_EXPORT int LGBMB200_LearnerCommConnect(LGBMB200_LearnerHandle h, int32_t rank, int32_t world,
                                                const uint8_t* all_handles, const int32_t* feature_offsets);

/* The same feature-shard bootstrap for `world` learners inside ONE process, one per GPU (how the reference's own
 * multi-GPU mode is driven: one host thread per device, include/LightGBM/cuda/cuda_nccl_topology.hpp:177-188): peer
 * access replaces CUDA IPC.  handles[r] is rank r's learner, already Init-ed with its column slice; afterwards Train must
 * be called on ALL of them concurrently (one host thread each) with the same gradients.  replicate_columns = 1 also gives
 * every rank a copy of every rank's partition columns (see CommShareColumns). */
LGBMB200_EXPORT int LGBMB200_LearnersConnectLocal(LGBMB200_LearnerHandle* handles, int32_t world, const int32_t* feature_offsets,
                                                  int32_t replicate_columns);

/* Optional, after CommConnect: replicate every rank's column-major partition columns on every rank (costs
 * total_columns x num_data bytes of HBM per GPU, filled by NVLink peer copies — the 180 GB of a B200 hold a
 * 10M x 1024 matrix 17 times).  Every rank then computes the go-left flags of every split locally and the per-split
 * flag push + wait disappears; only the 88-byte candidate exchange stays on the per-split path.  The reference's
 * feature-parallel learner makes the same trade: every machine holds all the data so that Split() is local
 * (src/treelearner/feature_parallel_tree_learner.cpp:23-35, docs/Features.rst "Feature Parallel in LightGBM").
 * Bootstrap: CommExportColumns on every rank -> all-gather the handles -> CommShareColumns on every rank -> barrier. */
LGBMB200_EXPORT int LGBMB200_LearnerCommExportColumns(LGBMB200_LearnerHandle h, uint8_t* handle_out_64);
LGBMB200_EXPORT int LGBMB200_LearnerCommShareColumns(LGBMB200_LearnerHandle h, const uint8_t* all_column_handles);

/* ---- Multi-GPU, row-shard (semantic model: DataParallelTreeLearner, reference
 * src/treelearner/data_parallel_tree_learner.cpp, and the reference's own num_gpu>1 mode,
 * src/boosting/cuda/nccl_gbdt_component.hpp:30-57).  Every learner is Init-ed with ITS row slice and ALL columns;
 * Train takes the gradients of those rows.  Per split: local histogram of the smaller child -> the owning rank
 * of each feature slice sums the peers' histogram slots straight out of their HBM over NVLink inside the scan
 * kernel (exact int64) -> candidate exchange as in feature-shard mode -> local partition + an all-gather of the
 * left/right counts.  Root sums and counts are global.  Bootstrap: CommExport + CommExportPool handles of every
 * rank, then CommConnectRows. */
LGBMB200_EXPORT int LGBMB200_LearnerCommExportPool(LGBMB200_LearnerHandle h, uint8_t* handle_out_64);
LGBMB200_EXPORT int LGBMB200_LearnerCommConnectRows(LGBMB200_LearnerHandle h, int32_t rank, int32_t world,
                                                    const uint8_t* comm_handles, const uint8_t* pool_handles);

/* Leaf id of every row of [row_lo, row_hi) in the last tree, one byte per row (needs <= 255 leaves; 0xFF = the row is
 * outside the bagging set): the host-score path of AddPredictionToScore for a caller that keeps only a row slice of the
 * scores (N > 1 feature-sharded ranks each update 1/N of the host score). */
LGBMB200_EXPORT int LGBMB200_LearnerGetLeafIndexRange8(LGBMB200_LearnerHandle h, int32_t row_lo, int32_t row_hi,
                                                       uint8_t* leaf_index_host);

/* Per-row leaf id of the last tree (host output, -1 for rows outside the bag): what the CPU learner's
 * DataPartition encodes and the reference CUDA learner keeps in cuda_data_index_to_leaf_index_
 * (reference src/treelearner/cuda/cuda_data_partition.cu:113). */
LGBMB200_EXPORT int LGBMB200_LearnerGetLeafIndex(LGBMB200_LearnerHandle h, int32_t* leaf_index_host);

/* CUDA-event stopwatch on the learner's stream (the stream every kernel of this library is launched on). */
LGBMB200_EXPORT int LGBMB200_LearnerTimerStart(LGBMB200_LearnerHandle h);
LGBMB200_EXPORT int LGBMB200_LearnerTimerStop(LGBMB200_LearnerHandle h, float* elapsed_ms);

/* Pinned host memory for the host-buffer (e2e) path. */
LGBMB200_EXPORT int LGBMB200_HostAllocPinned(void** ptr, int64_t bytes);
LGBMB200_EXPORT int LGBMB200_HostFreePinned(void* ptr);

/* Raw device pointer helpers so a host program without torch can keep grad/hess/score in HBM. */
LGBMB200_EXPORT int LGBMB200_DeviceAlloc(void** ptr, int64_t bytes);
LGBMB200_EXPORT int LGBMB200_DeviceFree(void* ptr);
LGBMB200_EXPORT int LGBMB200_MemcpyH2D(void* dst_dev, const void* src_host, int64_t bytes);
LGBMB200_EXPORT int LGBMB200_MemcpyD2H(void* dst_host, const void* src_dev, int64_t bytes);

LGBMB200_EXPORT int LGBMB200_LearnerFree(LGBMB200_LearnerHandle h);

/* ---------------------------------------------------------------------------------------------------------------
 * Dataset construction: what LGBM_DatasetCreateFromMat does between the caller's dense matrix and
 * the binned Dataset a tree learner is Init-ed with — reference src/c_api.cpp:1296-1408 (row sample, :981-989),
 * DatasetLoader::ConstructFromSampleData (src/io/dataset_loader.cpp:600-761), BinMapper::FindBin (src/io/bin.cpp:315-512),
 * Dataset::Construct / FastFeatureBundling (src/io/dataset.cpp:66-440) under the rules of a `device_type=cuda` Dataset,
 * Dataset::PushOneRow -> FeatureGroup::PushData (include/LightGBM/feature_group.h:253-267).
 * Numerical features only; categorical_feature, forcedbins_filename, max_bin_by_feature and max_bin > 255 are not supported. */
typedef void* LGBMB200_BinnerHandle;

/* the Dataset parameters the construction reads (reference include/LightGBM/config.h "Dataset Parameters") */
typedef struct {
  int32_t max_bin;                    /* 2..255 */
  int32_t min_data_in_bin;
  int32_t min_data_in_leaf;           /* feature_pre_filter drops features that cannot be split under it */
  int32_t bin_construct_sample_cnt;
  int32_t data_random_seed;
  int32_t feature_pre_filter;
  int32_t use_missing;
  int32_t zero_as_missing;
  int32_t enable_bundle;
  int32_t gpu_device_id;              /* device of the value->bin pass; -1 = current device */
} LGBMB200_BinConfig;

LGBMB200_EXPORT int LGBMB200_BinnerCreate(const LGBMB200_BinConfig* config, LGBMB200_BinnerHandle* out);

/* Sample rows, find the bin mappers, bundle the features.  Host work only (no CUDA call).
 * data: dense host matrix, data_type 0 = float32 / 1 = float64 (C_API_DTYPE_FLOAT32/64, c_api.h:38-39). */
LGBMB200_EXPORT int LGBMB200_BinnerFit(LGBMB200_BinnerHandle h, const void* data, int32_t data_type, int32_t nrow, int32_t ncol,
                                       int32_t is_row_major);

/* The layout contract of the fitted Dataset; the arrays stay owned by the handle (valid until the next Fit / Free). */
LGBMB200_EXPORT int LGBMB200_BinnerGetLayout(LGBMB200_BinnerHandle h, LGBMB200_Layout* out);

/* BinMapper::bin_upper_bound_ of an inner feature (what Dataset::RealThreshold reads, dataset.h:853): num_bin doubles
 * (the last one +inf; a NaN-missing mapper ends {+inf, 2.0}: the reference stores the enumerator MissingType::NaN there).  upper_bounds_out may be NULL to query num_bin only. */
LGBMB200_EXPORT int LGBMB200_BinnerGetFeatureBounds(LGBMB200_BinnerHandle h, int32_t inner_feature, double* upper_bounds_out,
                                                    int32_t* num_bin_out);

/* LGBM_SampleIndices (c_api.h:96): the rows the mappers were found from.  indices_out may be NULL to query the count. */
LGBMB200_EXPORT int LGBMB200_BinnerGetSampleIndices(LGBMB200_BinnerHandle h, int32_t* indices_out, int32_t* num_out);

/* The value->bin pass over all rows on the device: row-major [nrow x ncol-of-Fit] matrix (host, streamed in chunks, or
 * already in HBM) -> [nrow x num_columns] stored bytes (host or device pointer).  A device `bins_out` can be handed to
 * LGBMB200_LearnerInit as is.  elapsed_ms (may be NULL): device time of the whole pass, copies included. */
LGBMB200_EXPORT int LGBMB200_BinnerTransform(LGBMB200_BinnerHandle h, const void* data, int32_t data_type, int32_t nrow,
                                             int32_t data_on_device, uint8_t* bins_out, int32_t out_on_device, float* elapsed_ms);

LGBMB200_EXPORT int LGBMB200_BinnerFree(LGBMB200_BinnerHandle h);

/* ---------------------------------------------------------------------------------------------------------------
 * Prediction with a trained model: LGBM_BoosterPredictForMat with C_API_PREDICT_RAW_SCORE
 * (reference include/LightGBM/c_api.h:1283-1330) -> GBDT::PredictRaw (src/boosting/gbdt_prediction.cpp:15-34) ->
 * Tree::Predict / NumericalDecision (include/LightGBM/tree.h:337-355, :587-620).  The model is handed over as the flat
 * arrays of the model text (Tree::ToString, src/io/tree.cpp:343-413), all trees back to back: per internal node
 * split_feature (real index), threshold, decision_type, left_child, right_child (>= 0 node, < 0 ~leaf); per leaf
 * leaf_value.  One tree per iteration (num_class = 1), numerical splits only. */
typedef void* LGBMB200_PredictorHandle;

LGBMB200_EXPORT int LGBMB200_PredictorCreate(int32_t gpu_device_id, int32_t num_trees, const int32_t* tree_num_leaves,
                                             const int32_t* split_feature, const double* threshold, const int8_t* decision_type,
                                             const int32_t* left_child, const int32_t* right_child, const double* leaf_value,
                                             int32_t max_feature_idx, LGBMB200_PredictorHandle* out);

/* Raw scores (sum of tree outputs in tree order, double) of a row-major [nrow x ncol] float32 (0) / float64 (1) matrix,
 * host (streamed in row chunks) or device resident; out_raw_score: nrow doubles, host or device. */
LGBMB200_EXPORT int LGBMB200_PredictorPredict(LGBMB200_PredictorHandle h, const void* data, int32_t data_type, int32_t nrow, int32_t ncol,
                                              int32_t data_on_device, double* out_raw_score, int32_t out_on_device, float* elapsed_ms);

LGBMB200_EXPORT int LGBMB200_PredictorFree(LGBMB200_PredictorHandle h);

#ifdef __cplusplus
}
#endif
#endif  /* LGBM_B200_H_ */
