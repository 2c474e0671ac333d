// types.cuh — device-resident state of one tree learner (sm_100a).
// Everything the per-tree launch sequence needs lives in HBM so that a whole tree is grown without
// a single host round-trip (the reference CUDA learner synchronises 7x per split).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr int kBinsPerColumn = 256;      // one histogram column = 256 slots of (grad, hess)
constexpr int kColGroup = 32;            // columns owned by one warp of the histogram kernel
// meta.h:54 — kEpsilon is a *float* literal widened to double at every use in the reference.
#define B200_KEPS (static_cast<double>(1e-15f))

// Per inner feature: the layout contract (reference feature_group.h:40-76, feature_histogram.hpp:1415-1454)
struct FeatMeta {
  int32_t col;          // column (feature group) in the bin matrix
  int32_t lo;           // stored value of slice entry 0  (bin_offsets_[sub_feature])
  int32_t nslice;       // num_bin - offset entries in the histogram slice
  int32_t num_bin;
  int32_t offset;       // most_freq_bin == 0
  int32_t mfb;
  int32_t default_bin;
  int32_t missing;      // 0 none, 1 zero, 2 nan
  int32_t real_index;   // tie-break key (split_info.hpp:138-164)
};

struct Params {
  int32_t num_leaves, max_depth, min_data_in_leaf, pad;
  double min_sum_hessian, l1, l2, min_gain_to_split, max_delta_step, path_smooth;
  // quantized-gradient training (reference gradient_discretizer.cpp, config.h:626-651)
  int32_t quant;               // use_quantized_grad
  int32_t quant_bins;          // num_grad_quant_bins
  int32_t quant_renew;         // quant_train_renew_leaf
  int32_t quant_stochastic;    // stochastic_rounding
  int32_t quant_const_hess;    // is_constant_hessian passed to Init
  int32_t quant_seed;
};

// The best split found for one (leaf[, feature]) — reference split_info.hpp:22-56
struct Cand {
  double gain;                 // -inf when none
  double lsg, lsh, lout;       // left sums / output
  double rsg, rsh, rout;
  int32_t feature;             // inner feature index, -1 when none
  int32_t threshold;
  int32_t default_left;
  int32_t left_count, right_count;   // estimated (RoundInt(hess*cnt_factor)) until the partition ran
  int32_t real;                // real (global) feature index: the cross-feature / cross-rank tie-break key
  int32_t owner;               // rank that owns `feature` (feature-shard mode); 0 on a single GPU
  int32_t pad;
  long long ilg, ilh;          // quantized training: exact integer left sums (SplitInfo::left_sum_gradient_and_hessian)
};

// LeafSplits + DataPartition entry + HistogramPool slot of one leaf
struct Leaf {
  int32_t begin, count;        // segment of the index buffer (data_partition.hpp:101-120)
  int32_t buf;                 // which of the two ping-pong index buffers holds the segment
  int32_t depth;
  int32_t slot;                // histogram pool slot
  int32_t lcount;              // rows of this leaf held by THIS rank (== count except in row-shard mode)
  double sum_g, sum_h, output; // leaf_splits.hpp: sum_gradients_, sum_hessians_, weight_
  long long isum_g, isum_h;    // quantized training: int_sum_gradients_and_hessians_ (leaf_splits.hpp:66-74), unpacked
  Cand best;                   // best_split_per_leaf_[leaf]
};

// Control block: what the next kernel in the sequence should work on.  Written by single threads of
// the `select` / `scatter` kernels, read by every block of the following launches.
struct Ctl {
  // current split to apply (snapshot taken by k_select so that k_scatter's bookkeeping cannot race)
  int32_t cur_valid;           // 0 => no further split (gain <= 0): every later kernel exits at once
  int32_t cur_leaf, cur_begin, cur_count, cur_buf;
  int32_t cur_feature, cur_threshold, cur_default_left;
  FeatMeta cur_meta;
  // leaves to histogram/scan next
  int32_t smaller, larger;     // larger = -1 for the root pass
  int32_t do_find;             // BeforeFindBestSplit() result (serial_tree_learner.cpp:343-370)
  int32_t num_leaves;          // leaves grown so far
  // fixed-point scales of the int64 histogram (power of two), set once per tree
  double g_scale, h_scale, g_inv, h_inv;
  double root_sum_g, root_sum_h;
  long long h_const_q;         // constant-hessian training: rint(hessians[0] * h_scale), the per-row fixed-point hessian
  // quantized training: GradientDiscretizer::grad_scale() / hess_scale() of this tree and their inverses
  double q_gscale, q_hscale, q_ginv, q_hinv;
  unsigned long long quant_iter;   // trees discretized so far (stochastic rounding stream id)
  int32_t root_count;          // rows in the root (bag size or num_data)
  int32_t root_identity;       // 1: root index list is 0..N-1 (no bagging) => histogram skips the index load
  // ---- multi-GPU (feature-shard) state; persistent across trees
  int32_t cur_owner;           // rank owning the feature of the split to apply
  int32_t error;               // set by a watchdog when a peer never showed up
  unsigned long long xchg_seq; // number of candidate exchanges done so far (mailbox sequence)
  unsigned long long flag_seq; // number of flag pushes done so far
  uint32_t part_blocks_done;   // last-block detection in k_part_flags
  uint32_t scan_done;          // last-block detection in k_scan (fused selection)
  unsigned long long hist_seq; // row-shard: number of "my local histogram is complete" signals sent
  unsigned long long misc_seq; // row-shard: number of small all-gathers (root sums, left counts) done
};

// One applied split, copied back to the host once per tree (mirrors LGBMB200_Split)
struct SplitRec {
  int32_t leaf, feature, threshold, default_left, left_count, right_count;
  int32_t owner, pad;
  double gain, lsg, lsh, lout, rsg, rsh, rout;
};

struct PartialSum { double g, h; float gmax, hmax; };

// ---- feature-shard exchange over NVLink peer memory (one CommBlock per rank, IPC-mapped by every peer)
constexpr int kMaxRanks = 8;
struct CommBlock {
  Cand mail[2][kMaxRanks][2];                 // [parity][source rank][smaller, larger] best candidates
  unsigned long long mail_seq[2][kMaxRanks];  // written by the source rank after its payload
  unsigned long long flags_seq[2];            // written by the split's owner after pushing the go-left flags
  int32_t num_features, num_columns;          // this rank's shard shape (read by the peers in CommShareColumns)
  unsigned long long pad[5];
  // row-shard mode
  unsigned long long hist_seq[kMaxRanks];     // hist_seq[r]: rank r's local histogram #seq is complete
  double misc[2][kMaxRanks][8];               // small all-gather payloads (root sums / left counts)
  unsigned long long misc_seq[2][kMaxRanks];
  int32_t blk_left[2][1024];                  // feature-shard: the owner's per-block left counts, pushed with the flags
  // followed by: uint32_t flag_words[2][ceil(num_data/32)] (bit-packed go-left flags),
  //              then FeatMeta[num_features] (this rank's layout contract, for replicated partition columns)
};
struct CommPeers {
  CommBlock* block[kMaxRanks];                // block[r] = rank r's CommBlock (own entry = local pointer)
  int32_t rank, world;
  int32_t mode;                               // 0 = feature-shard (all rows x column slice), 1 = row-shard (row slice x all columns),
                                              // 2 = feature-shard with every rank's partition columns replicated locally
  int32_t f_lo, f_cnt;                        // row-shard: the feature slice this rank reduces and scans
  int32_t pad;
  long long* pool[kMaxRanks];                 // row-shard: every rank's histogram pool (peer-mapped)
  int64_t flags_stride;                       // bytes between the two flag-word buffers
  const FeatMeta* gmeta;                      // mode 2: every rank's FeatMeta, columns re-based into the replicated column-major copy
  int32_t feat_off[kMaxRanks];                // mode 2: first gmeta entry of rank r
};
__host__ __device__ inline uint32_t* comm_flag_words(CommBlock* b, int parity, int64_t stride) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(b + 1) + parity * stride);
}
__host__ __device__ inline FeatMeta* comm_meta_tail(CommBlock* b, int64_t stride) {
  return reinterpret_cast<FeatMeta*>(reinterpret_cast<uint8_t*>(b + 1) + 2 * stride);
}

}  // namespace b200
