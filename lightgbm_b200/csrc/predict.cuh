// predict.cuh — raw-score prediction of a boosted tree model over a dense float matrix on the device.
//
// Replaces: GBDT::PredictRaw -> Tree::Predict -> Tree::GetLeaf / NumericalDecision
// (reference src/boosting/gbdt_prediction.cpp:15-34, include/LightGBM/tree.h:337-355, :587-620, :701-713) as driven by
// LGBM_BoosterPredictForMat (src/c_api.cpp), numerical splits only.
//
// Layout: a CTA stages a tile of rows in shared memory with coalesced loads — X is read from HBM exactly once, 4 or 8
// bytes per cell, the only traffic that scales with the data — with an odd row stride, so that 32 lanes reading one
// feature of 32 different rows hit 32 banks.  A THREAD owns a row and walks the trees in order, two at a time (two
// independent chains per lane), adding their outputs in tree order: the double-precision sum is the reference's sequential
// `output += tree->Predict(row)` bit for bit.  The 32 lanes of a warp are on the same tree, so near the root they read
// the same node (one L1 wavefront) and a tree's 126 nodes (3 KB) stay in L1 while the warp is on it.
// (The first version put the lanes on 32 different TREES: every node load touched 32 cache lines — 11.6 G tree visits/s,
// 1.8 % of the HBM roofline at 2M x 256 x 100 trees.)
#pragma once
#include <cstdint>

namespace b200 {

struct __align__(16) PNodeA { double threshold; int32_t feature; int32_t decision; };   // decision_type_: bit 1 default left, bits 2..3 missing type
struct __align__(8) PNodeB { int32_t left, right; };                                    // >= 0: node, < 0: ~leaf

struct PredTable {
  const PNodeA* node_a;         // all trees, concatenated
  const PNodeB* node_b;
  const double* leaf_value;     // all trees, concatenated
  const int32_t* node_first;    // [num_trees]
  const int32_t* leaf_first;    // [num_trees]
  const int32_t* num_leaves;    // [num_trees]
  int32_t num_trees;
};

constexpr int kPredThreads = 256;
constexpr int kPredTileBytes = 192 * 1024;     // rows staged per CTA

// one step of Tree::GetLeaf: NumericalDecision (tree.h:337-355)
template <typename T>
__device__ __forceinline__ int pred_step(const PNodeA* __restrict__ na, const PNodeB* __restrict__ nb, int node, const T* row) {
  const PNodeA a = na[node];
  const PNodeB b = nb[node];
  double v = static_cast<double>(row[a.feature]);
  const int missing = (a.decision >> 2) & 3;
  const bool nan = v != v;
  if (nan && missing != LGBMB200_MISSING_NAN) v = 0.0;
  // kZeroThreshold = 1e-35f (meta.h:56), Tree::IsZero (tree.h:330-332)
  if ((missing == LGBMB200_MISSING_ZERO && v >= -static_cast<double>(1e-35f) && v <= static_cast<double>(1e-35f)) ||
      (missing == LGBMB200_MISSING_NAN && nan)) {
    return (a.decision & 2) ? b.left : b.right;
  }
  return (v <= a.threshold) ? b.left : b.right;
}

template <typename T>
__device__ __forceinline__ double pred_row(const PredTable& m, const T* row) {
  double sum = 0.0;
  for (int t = 0; t < m.num_trees; t += 2) {
    const bool two = t + 1 < m.num_trees;
    const int f0 = m.node_first[t], f1 = two ? m.node_first[t + 1] : 0;
    int n0 = m.num_leaves[t] > 1 ? 0 : -1;                 // a single-leaf tree: ~0 = leaf 0
    int n1 = (two && m.num_leaves[t + 1] > 1) ? 0 : -1;
    while (n0 >= 0 || n1 >= 0) {
      if (n0 >= 0) n0 = pred_step<T>(m.node_a + f0, m.node_b + f0, n0, row);
      if (n1 >= 0) n1 = pred_step<T>(m.node_a + f1, m.node_b + f1, n1, row);
    }
    sum += m.leaf_value[m.leaf_first[t] + ~n0];             // tree order: the reference's summation
    if (two) sum += m.leaf_value[m.leaf_first[t + 1] + ~n1];
  }
  return sum;
}

template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict(const T* __restrict__ x, int64_t ld, int64_t nrow, int32_t ncol, const PredTable m,
                                                          double* __restrict__ out, int32_t tile_rows, int32_t stride) {
  extern __shared__ __align__(16) unsigned char psmem[];
  T* tile = reinterpret_cast<T*>(psmem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kPredThreads / 32;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * tile_rows; r0 < nrow; r0 += static_cast<int64_t>(gridDim.x) * tile_rows) {
    const int rows = static_cast<int>(min(static_cast<int64_t>(tile_rows), nrow - r0));
    for (int r = warp; r < rows; r += kWarps) {
      const T* src = x + (r0 + r) * ld;
      for (int c = lane; c < ncol; c += 32) tile[r * stride + c] = src[c];
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < rows) out[r0 + threadIdx.x] = pred_row<T>(m, tile + threadIdx.x * stride);
    __syncthreads();
  }
}

// rows too wide for a shared-memory tile of 32: a thread reads its row straight from global memory (L1 / L2 hold the
// sectors it has touched)
template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict_wide(const T* __restrict__ x, int64_t ld, int64_t nrow, const PredTable m,
                                                               double* __restrict__ out) {
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kPredThreads + threadIdx.x; r < nrow; r += static_cast<int64_t>(gridDim.x) * kPredThreads)
    out[r] = pred_row<T>(m, x + r * ld);
}

}  // namespace b200
