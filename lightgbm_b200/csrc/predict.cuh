// predict.cuh — raw-score prediction of a boosted tree model over a dense float matrix on the device.
//
// Replaces: GBDT::PredictRaw -> Tree::Predict -> Tree::GetLeaf / NumericalDecision
// (reference src/boosting/gbdt_prediction.cpp:15-34, include/LightGBM/tree.h:337-355, :587-620, :701-713) as driven by
// LGBM_BoosterPredictForMat (src/c_api.cpp), numerical splits only.
//
// Layout: a CTA stages a tile of rows in shared memory with coalesced loads — X is read from HBM exactly once, 4 or 8
// bytes per cell, the only traffic that scales with the data — with an odd row stride, so that 32 lanes reading one
// feature of 32 different rows hit 32 banks.  The lanes of a warp are 32 ROWS on the same tree: near the root they read
// the same node (one L1 wavefront) and a tree's 126 nodes (3 KB) stay in L1 while the warp is on it.  Tree outputs are
// added in tree order: the double-precision sum is the reference's sequential `output += tree->Predict(row)` bit for bit.
// (The first version put the lanes on 32 different TREES: every node load touched 32 cache lines — 11.6 G tree visits/s,
// 1.8 % of the HBM roofline at 2M x 256 x 100 trees.)
#pragma once
#include <cstdint>

#include "hist_common.cuh"

namespace b200 {

__device__ __forceinline__ void cp_async_elem(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_elem(double* smem_dst, const double* gsrc) { cp_async8(smem_dst, gsrc); }

struct __align__(16) PNodeA { double threshold; int32_t feature; int32_t decision; };   // decision_type_: bit 1 default left, bits 2..3 missing type
struct __align__(8) PNodeB { int32_t left, right; };                                    // >= 0: node, < 0: ~leaf

// float32 input: one 16-byte record per node and no FP64 on the walk.  (double)x <= t  <=>  x <= thr, thr = the largest float
// <= t, for every float x (NaN is handled before the compare); feature in the low 24 bits of `fd`, decision_type above.
struct __align__(16) PNodeF { float thr; uint32_t fd; int32_t left, right; };

struct PredTable {
  const PNodeA* node_a;         // all trees, concatenated
  const PNodeB* node_b;
  const PNodeF* node_f;
  const double* leaf_value;     // all trees, concatenated
  const int32_t* node_first;    // [num_trees]
  const int32_t* leaf_first;    // [num_trees]
  const int32_t* num_leaves;    // [num_trees]
  int32_t num_trees;
};

constexpr int kPredThreads = 256;
constexpr int kPredTileBytes = 192 * 1024;     // rows staged per CTA

// one step of Tree::GetLeaf: NumericalDecision (tree.h:337-355)
__device__ __forceinline__ int pred_step_f(const PNodeF* __restrict__ nf, int node, const float* row) {
  const PNodeF n = nf[node];
  float v = row[n.fd & 0xffffffu];
  const unsigned decision = n.fd >> 24;
  const unsigned missing = (decision >> 2) & 3u;
  const bool nan = v != v;
  if (nan && missing != LGBMB200_MISSING_NAN) v = 0.f;
  if ((missing == LGBMB200_MISSING_ZERO && fabsf(v) <= 1e-35f) || (missing == LGBMB200_MISSING_NAN && nan)) return (decision & 2u) ? n.left : n.right;
  return (v <= n.thr) ? n.left : n.right;
}

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

// leaf values of trees ta and tb (tb < 0: only ta) for one row, the two walks interleaved (two independent chains per lane)
template <typename T>
__device__ __forceinline__ void pred_pair(const PredTable& m, int ta, int tb, const T* row, double* va, double* vb) {
  const int fa = m.node_first[ta], fb = tb >= 0 ? m.node_first[tb] : 0;
  int na = m.num_leaves[ta] > 1 ? 0 : -1;                   // a single-leaf tree: ~0 = leaf 0
  int nb = (tb >= 0 && m.num_leaves[tb] > 1) ? 0 : -1;
  if (sizeof(T) == 4) {
    const float* frow = reinterpret_cast<const float*>(row);
    while (na >= 0 || nb >= 0) {
      if (na >= 0) na = pred_step_f(m.node_f + fa, na, frow);
      if (nb >= 0) nb = pred_step_f(m.node_f + fb, nb, frow);
    }
  } else {
    while (na >= 0 || nb >= 0) {
      if (na >= 0) na = pred_step<T>(m.node_a + fa, m.node_b + fa, na, row);
      if (nb >= 0) nb = pred_step<T>(m.node_a + fb, m.node_b + fb, nb, row);
    }
  }
  *va = m.leaf_value[m.leaf_first[ta] + ~na];
  if (tb >= 0) *vb = m.leaf_value[m.leaf_first[tb] + ~nb];
}

template <typename T>
__device__ __forceinline__ double pred_row(const PredTable& m, const T* row) {
  double sum = 0.0;
  for (int t = 0; t < m.num_trees; t += 2) {
    double va = 0.0, vb = 0.0;
    pred_pair<T>(m, t, t + 1 < m.num_trees ? t + 1 : -1, row, &va, &vb);
    sum += va;                                               // tree order: the reference's summation
    if (t + 1 < m.num_trees) sum += vb;
  }
  return sum;
}

// A CTA of 256 threads scores a tile of R = 64 (or 32) rows: thread = (row, tree group g of G = 256 / R).  The trees are
// taken `pass` (32) at a time: group g walks trees g, g + G, ... of the pass and parks the leaf values in shared memory,
// then the row's g = 0 thread adds the pass IN TREE ORDER to its running sum.  All 256 threads walk trees (the earlier
// one-thread-per-row form left a CTA with 160 busy threads and one CTA per SM: 40 ms for 2M x 256 x 100 trees; the walks
// are latency-bound, so what counts is the number of independent chains in flight per SM).
constexpr int kPredPassMax = 32;

template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict(const T* __restrict__ x, int64_t ld, int64_t nrow, int32_t ncol, const PredTable m,
                                                          double* __restrict__ out, int32_t R, int32_t stride, int32_t pass) {
  extern __shared__ __align__(16) unsigned char psmem[];
  double* vals = reinterpret_cast<double*>(psmem);                       // [pass][R]
  T* tile = reinterpret_cast<T*>(vals + pass * R);                       // [R][stride]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kPredThreads / 32;
  const int row = threadIdx.x & (R - 1), g = threadIdx.x / R, G = kPredThreads / R;
  for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * R; r0 < nrow; r0 += static_cast<int64_t>(gridDim.x) * R) {
    const int rows = static_cast<int>(min(static_cast<int64_t>(R), nrow - r0));
    // asynchronous element copies (LDGSTS): every thread has its whole share of the tile in flight at once
    for (int r = warp; r < R; r += kWarps) {
      const T* src = x + (r0 + min(r, rows - 1)) * ld;                   // ragged last tile: repeat its last row
      for (int c = lane; c < ncol; c += 32) cp_async_elem(tile + r * stride + c, src + c);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const T* my = tile + row * stride;
    double sum = 0.0;
    for (int t0 = 0; t0 < m.num_trees; t0 += pass) {
      const int cnt = min(pass, m.num_trees - t0);
      for (int k = g; k < cnt; k += 2 * G) {
        double va = 0.0, vb = 0.0;
        const int kb = k + G;
        pred_pair<T>(m, t0 + k, kb < cnt ? t0 + kb : -1, my, &va, &vb);
        vals[k * R + row] = va;
        if (kb < cnt) vals[kb * R + row] = vb;
      }
      __syncthreads();
      if (g == 0) {
#pragma unroll 8
        for (int k = 0; k < cnt; ++k) sum += vals[k * R + row];          // tree order: the reference's summation
      }
      __syncthreads();
    }
    if (g == 0 && row < rows) out[r0 + row] = sum;
  }
}

// (Two variants with the nodes in SHARED memory were built, checked bit-identical and measured on 2M x 256 x 100 trees:
// pass-major — one launch per 16 trees, running sums kept in `out` between passes — 16.9 ms (X is re-read once per pass and
// every (tile, pass) pays the 4-byte cp.async staging); tile-major with the current pass's nodes copied per (tile, pass) —
// 7.2 ms, the same as this kernel: ncu shows the long-scoreboard stalls gone (9.0 -> 1.7 per issue) and the issue slots
// 34 % busy with the same 110 M warp instructions — the leaf-wise trees are 30-60 levels deep on their main branch and the
// 32 rows of a warp diverge, so the walk is bound by instructions issued for the longest path.  Both removed.  A flat per-lane
// loop over the lane's trees of a pass (a chain moves on to its next tree the moment it reaches a leaf, so that the warp pays
// for sums of paths instead of the longest path once per tree) needs more instructions per step than it saves: 12.8 ms.)
// rows too wide for a shared-memory tile of 32: a thread reads its row straight from global memory (L1 / L2 hold the
// sectors it has touched)
template <typename T>
__global__ void __launch_bounds__(kPredThreads) k_predict_wide(const T* __restrict__ x, int64_t ld, int64_t nrow, const PredTable m,
                                                               double* __restrict__ out) {
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kPredThreads + threadIdx.x; r < nrow; r += static_cast<int64_t>(gridDim.x) * kPredThreads)
    out[r] = pred_row<T>(m, x + r * ld);
}

}  // namespace b200
