// predict.cuh — raw-score prediction of a boosted tree model over a dense float matrix on the device.
//
// Replaces: GBDT::PredictRaw -> Tree::Predict -> Tree::GetLeaf / NumericalDecision
// (reference src/boosting/gbdt_prediction.cpp:15-34, include/LightGBM/tree.h:337-355, :587-620, :701-713) as driven by
// LGBM_BoosterPredictForMat (src/c_api.cpp), numerical splits only.
//
// Layout: a CTA stages a tile of rows in shared memory with coalesced loads (X is read from HBM exactly once: 4 or 8 bytes
// per cell, the only traffic that scales with the data); the model's nodes (24 B each, all trees back to back, a few
// hundred KB) are read through L1/L2.  One warp scores one row at a time: lane l walks trees l, l + 32, ...; the 32 leaf
// values of a batch are then added IN TREE ORDER by every lane (shuffles), so the double-precision sum is the reference's
// sequential `output += tree->Predict(row)` bit for bit.
#pragma once
#include <cstdint>

namespace b200 {

struct PNode {
  double threshold;
  int32_t feature;        // real (column) index into the row
  int32_t left, right;    // >= 0: node, < 0: ~leaf
  int32_t decision;       // decision_type_: bit 1 = default left, bits 2..3 = missing type (tree.h:20-23, :258-270)
};
static_assert(sizeof(PNode) == 24, "PNode layout");

struct PredTable {
  const PNode* nodes;           // all trees, concatenated
  const double* leaf_value;     // all trees, concatenated
  const int32_t* node_first;    // [num_trees]
  const int32_t* leaf_first;    // [num_trees]
  const int32_t* num_leaves;    // [num_trees]
  int32_t num_trees;
};

constexpr int kPredThreads = 256;
constexpr int kPredTileBytes = 64 * 1024;      // rows staged per CTA; the rest of the SM's L1 serves the nodes

template <typename T>
__device__ __forceinline__ double tree_output(const PredTable& m, int t, const T* __restrict__ row) {
  if (m.num_leaves[t] <= 1) return m.leaf_value[m.leaf_first[t]];
  const PNode* nodes = m.nodes + m.node_first[t];
  int node = 0;
  while (node >= 0) {
    const PNode nd = nodes[node];
    double v = static_cast<double>(row[nd.feature]);
    const int missing = (nd.decision >> 2) & 3;
    const bool nan = v != v;
    if (nan && missing != LGBMB200_MISSING_NAN) v = 0.0;
    // kZeroThreshold = 1e-35f (meta.h:56), Tree::IsZero (tree.h:330-332)
    if ((missing == LGBMB200_MISSING_ZERO && v >= -static_cast<double>(1e-35f) && v <= static_cast<double>(1e-35f)) ||
        (missing == LGBMB200_MISSING_NAN && nan)) {
      node = (nd.decision & 2) ? nd.left : nd.right;
    } else {
      node = (v <= nd.threshold) ? nd.left : nd.right;
    }
  }
  return m.leaf_value[m.leaf_first[t] + ~node];
}

template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict(const T* __restrict__ x, int64_t ld, int64_t nrow, int32_t ncol, const PredTable m,
                                                          double* __restrict__ out, int32_t tile_rows) {
  extern __shared__ __align__(16) unsigned char psmem[];
  T* tile = reinterpret_cast<T*>(psmem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kPredThreads / 32;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * tile_rows; r0 < nrow; r0 += static_cast<int64_t>(gridDim.x) * tile_rows) {
    const int rows = static_cast<int>(min(static_cast<int64_t>(tile_rows), nrow - r0));
    if (tile_rows > 0 && ld == ncol) {
      // the tile is one contiguous run of the row-major matrix
      const T* src = x + r0 * ld;
      const int n = rows * ncol;
      for (int i = threadIdx.x; i < n; i += kPredThreads) tile[i] = src[i];
    } else {
      for (int r = warp; r < rows; r += kWarps)
        for (int c = lane; c < ncol; c += 32) tile[r * ncol + c] = x[(r0 + r) * ld + c];
    }
    __syncthreads();
    for (int r = warp; r < rows; r += kWarps) {
      const T* row = tile + r * ncol;
      double sum = 0.0;
      for (int t0 = 0; t0 < m.num_trees; t0 += 32) {
        const int t = t0 + lane;
        const double v = t < m.num_trees ? tree_output<T>(m, t, row) : 0.0;
        const int cnt = min(32, m.num_trees - t0);
        for (int j = 0; j < cnt; ++j) sum += __shfl_sync(0xffffffffu, v, j);      // tree order: the reference's summation
      }
      if (lane == 0) out[r0 + r] = sum;
    }
    __syncthreads();
  }
}

// rows too wide for the shared-memory tile: read the features straight from global memory
template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict_wide(const T* __restrict__ x, int64_t ld, int64_t nrow, const PredTable m,
                                                               double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * kPredThreads + threadIdx.x) >> 5;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (kPredThreads / 32);
  for (int64_t r = warp_global; r < nrow; r += warps) {
    const T* row = x + r * ld;
    double sum = 0.0;
    for (int t0 = 0; t0 < m.num_trees; t0 += 32) {
      const int t = t0 + lane;
      const double v = t < m.num_trees ? tree_output<T>(m, t, row) : 0.0;
      const int cnt = min(32, m.num_trees - t0);
      for (int j = 0; j < cnt; ++j) sum += __shfl_sync(0xffffffffu, v, j);
    }
    if (lane == 0) out[r] = sum;
  }
}

}  // namespace b200
