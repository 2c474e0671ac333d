// hist_kernel.cuh — per-leaf gradient/hessian histogram construction (sm_100a).
//
// Replaces: Dataset::ConstructHistograms -> MultiValDenseBin::ConstructHistogramInner
// (reference src/io/dataset.cpp:1293-1485, src/io/multi_val_dense_bin.hpp:58-102) and the reference
// CUDA kernel CUDAConstructHistogramDenseKernel (src/treelearner/cuda/cuda_histogram_constructor.cu:20),
// which does 2 shared-memory fp32 atomicAdds per cell — on sm_100a those compile to ATOMS.CAST.SPIN
// (a compare-and-swap loop), the dominant cost of the reference.
//
// Design (not a port): scatter-add WITHOUT atomics in the inner loop.
//   * A warp owns a group of 32 columns for a contiguous slice of the leaf's rows; lane L owns column L.
//     Two lanes of a warp therefore never touch the same (column, bin) cell, and no other warp shares
//     the warp's private histogram, so the update is a plain LDS.64 / FADD / STS.64.
//   * The warp-private histogram is laid out [bin][lane] with 8-byte (grad,hess) cells: the shared
//     memory bank of a cell depends on the lane only => every access is bank-conflict free whatever
//     the bin values are.  64 KB per warp, 3 warps per SM (one CTA of 96 threads per SM).
//   * The only hazard left is the same lane hitting the same bin in consecutive rows; rows are
//     processed four at a time with in-register forwarding, which keeps 4 independent LDS in flight
//     and preserves the exact sequential fp32 summation order (=> bitwise run-to-run determinism).
//   * Rows are gathered through the leaf's index list: 32-byte row segments and the (grad,hess) pairs
//     are staged into a per-warp 6-deep shared-memory ring with cp.async (LDGSTS) by a PARTNER PRODUCER
//     WARP; completion is tracked by mbarriers (cp.async.mbarrier.arrive), so the consumer warp never
//     touches global memory and never waits on a global-load scoreboard (the first version, which issued
//     its own loads, lost 37% of its cycles there — profiles/r01_hist_v1_stalls.txt).  No block barriers
//     in the main loop.
//   * Flush: fp32 partials -> int64 fixed point (power-of-two scale chosen per tree) added to the
//     leaf's slot of the histogram pool with RED.ADD.64.  Integer adds are associative, so the global
//     histogram does not depend on the order in which warps flush, and parent - child is exact.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "comm.cuh"
#include "types.cuh"

namespace b200 {

constexpr int kHistWarps = 3;                    // consumer warps (one per SMSP 0..2), each with a partner producer warp
constexpr int kHistThreads = 2 * kHistWarps * 32; // warps 0..2 consume, warps 3..5 produce (warp 3 owns SMSP 3)
constexpr int kStageRows = 32;
constexpr int kHistBatch = 4;                    // rows per in-register RMW batch (independent LDS.64 in flight per warp)
constexpr int kStages = 6;                       // ring depth; kStages-1 stages in flight
constexpr int kWarpHistBytes = kBinsPerColumn * 32 * 8;           // 65536
constexpr int kStageBinBytes = kStageRows * kColGroup;            // 1024
constexpr int kStageBytes = kStageBinBytes + kStageRows * 8;      // + (g,h) pairs = 1280
constexpr int kWarpSmemBytes = kWarpHistBytes + kStages * kStageBytes;
constexpr int kHistBarBytes = kHistWarps * kStages * 2 * 8;        // full[] + empty[] mbarriers per consumer warp
constexpr int kHistSmemBytes = kHistWarps * kWarpSmemBytes + kHistBarBytes;   // 219936 B <= 227 KB

struct HistArgs {
  const uint8_t* bins;            // [num_data x pitch] row-major stored values
  int64_t pitch;                  // bytes per row, multiple of 32
  const float2* gh;               // [num_data] (grad, hess)
  const int32_t* idx0;            // ping-pong row-index buffers
  const int32_t* idx1;
  const Leaf* leaves;
  const Ctl* ctl;
  unsigned long long* pool;       // int64 fixed-point histogram pool [slot][column][256][2]
  int64_t slot_stride;            // int64 elements per slot
  int32_t num_colgroups;          // ceil(num_columns / 32)
  int32_t min_rows_per_item;      // do not split a column group over more warps than n / this
  int32_t use_tma;                // 1: contiguous (root, un-bagged) stages are staged by TMA tile copies
  const int32_t* ghqo0;           // leaf-ordered packed quantized words (k_hist_q), or nullptr
  const int32_t* ghqo1;
  const float2* gho0;             // leaf-ordered (g,h) parallel to idx0 / idx1 (written by k_part_scatter), or nullptr
  const float2* gho1;
  int32_t l2_prefetch;            // > 0: gathered passes prefetch the bin sectors of the stage this many stages ahead into L2
  int32_t map_mode;               // 0: items dealt column-group-major; 1: one CTA = (column group, 3 row parts), adjacent CTAs = adjacent column groups
  // explicit mode (stand-alone ConstructHistogram hook): explicit_n >= 0
  int32_t explicit_n;
  int32_t explicit_slot;
  const int32_t* explicit_idx;    // nullptr = identity
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int kPfIdLead = 2;          // producer iterations between loading a future stage's row ids and prefetching its rows
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float2 lds64(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(unsigned addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y));
}

// Batches of K rows of the lane's column: all K cells are loaded BEFORE any store, so rows that hit the same bin see
// the same old value v; the new value of row i is v + (q_i + sum of the q_j, j < i, with b_j == b_i).  Those partial
// sums do not depend on the loads and are formed while the LDS are in flight, so the critical path per batch is
// LDS -> FADD -> STS with K independent LDS.64 in flight per warp.  Stores are issued in row order: the last store to
// a cell carries the complete sum.  Fixed evaluation order => bitwise run-to-run determinism.
// The same batch update split in two so that the caller can software-pipeline it: batch_prepare (pure ALU: cell
// addresses + the duplicate-combined increments) of batch k+1 is placed in the shadow of batch k's LDS latency.
template <bool QUANT> __device__ __forceinline__ float2 acc2(float2 a, float2 b);
template <int K, bool QUANT>
__device__ __forceinline__ void batch_prepare(unsigned hbase, const uint32_t (&b)[K], const float2 (&q)[K], unsigned (&addr)[K], float2 (&s)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    addr[i] = hbase + (b[i] << 8);
    s[i] = q[i];
#pragma unroll
    for (int j = 0; j < i; ++j) if (b[j] == b[i]) s[i] = acc2<QUANT>(s[i], q[j]);
  }
}

// ---- mbarrier helpers (shared::cta) -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
// all prior cp.async of this thread arrive on `bar` when they complete (count pre-accounted at init)
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 28)) __trap();          // watchdog: turn a protocol bug into a launch error, not a hang
  }
}

// Producer-side wait: the same protocol, but the try_wait carries a suspend-time hint so that a producer that has
// filled its ring parks in hardware instead of spinning (ncu on k_hist_q: the spin of the three producer warps was
// 40 % of all issued instructions and took issue slots from the consumer warps that share their SMSPs).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, unsigned parity) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

// ---- TMA (bulk async copies, completion by mbarrier complete_tx) ---------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
// 2-D tile [32 rows x 32 columns] of the row-major bin matrix -> shared memory (UTMALDG)
__device__ __forceinline__ void tma_load_tile_2d(void* smem_dst, const CUtensorMap* tmap, int col, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(tmap), "r"(col), "r"(row),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
// contiguous bytes -> shared memory (UBLKCP); size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc), "r"(bytes),
                 "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}

// Work decomposition shared by the producer and the consumer warp of a pair.
// Items are numbered column-group-major (item j -> column group j / splits, row part j % splits) and dealt to
// warp slots in order, so the three consumer warps of a CTA normally work on consecutive row parts of the SAME
// column group and can merge their private tables in shared memory before touching global memory.
struct HistWork {
  int n, begin, slot;
  const int32_t* idx;
  const float2* gh_ord;          // non-root leaf: (g,h) in leaf order (position-indexed), else nullptr => gather by row id
  const int32_t* ghq_ord;        // same for the packed quantized word (k_hist_q)
  int CG, splits, per, items, total_warps;
  int mode, rounds_total;        // mode 1: rounds_total = CG * ceil(splits/3) virtual CTAs
};

__device__ __forceinline__ bool hist_work_setup(const HistArgs& a, HistWork* w) {
  if (a.explicit_n >= 0) {
    w->n = a.explicit_n; w->begin = 0; w->slot = a.explicit_slot; w->idx = a.explicit_idx;
    w->gh_ord = nullptr; w->ghq_ord = nullptr;
  } else {
    const Ctl* c = a.ctl;
    if (!c->cur_valid || !c->do_find) return false;
    const Leaf& L = a.leaves[c->smaller];
    w->n = L.lcount; w->begin = L.begin; w->slot = L.slot;
    // the root of an un-bagged tree is the identity list: skip the index load altogether
    w->idx = (c->num_leaves == 1 && c->root_identity) ? nullptr : (L.buf ? a.idx1 : a.idx0);
    // every non-root segment was written by its parent's scatter, together with its (g,h) copy
    w->gh_ord = (c->num_leaves > 1 && a.gho0 != nullptr) ? (L.buf ? a.gho1 : a.gho0) + L.begin : nullptr;
    w->ghq_ord = (c->num_leaves > 1 && a.ghqo0 != nullptr) ? (L.buf ? a.ghqo1 : a.ghqo0) + L.begin : nullptr;
  }
  if (w->n <= 0) return false;
  w->total_warps = gridDim.x * kHistWarps;
  w->CG = a.num_colgroups;
  const int max_splits = max(1, w->total_warps / w->CG);
  // row parts per column group: balance the per-row work (~17 ns per row and warp) against the per-item fixed
  // cost (zeroing + the merged flush, ~16K/3 global atomics per item at ~0.2 G atomics/us):
  // t(s) = n/s * t_row + CG*s * t_item  =>  s* = sqrt(n * t_row / (CG * t_item)) ~ sqrt(0.7 n / CG)
  int sp = static_cast<int>(sqrtf(0.7f * static_cast<float>(w->n) / static_cast<float>(w->CG)));
  sp = max(1, min(sp, (w->n + a.min_rows_per_item - 1) / a.min_rows_per_item));
  // mapping choice: the CTA-per-(column group, row triple) mapping unless it would leave >5 % of the warps idle
  // on a leaf big enough to use them all (e.g. 32 column groups: 4 triples = 12 parts vs 13 parts column-group-major;
  // measured 8-13 % slower on 4M x 1024)
  const int max_triples0 = max(1, static_cast<int>(gridDim.x) / w->CG);
  w->mode = a.map_mode;
  if (w->mode == 1 && sp > max_triples0 * kHistWarps && max_triples0 * kHistWarps * 20 < max_splits * 19) w->mode = 0;
  if (w->mode == 1) {
    // one CTA = one column group x three consecutive row parts (always mergeable); CTAs b, b+1 work on adjacent
    // column groups of the SAME rows at the same time, so the two 32-byte sectors of a 64-byte DRAM atom are
    // fetched once (the second hits in L2).  One wave: at most floor(grid / CG) triples per column group.
    const int max_triples = max(1, static_cast<int>(gridDim.x) / w->CG);
    const int triples = min(max_triples, (sp + kHistWarps - 1) / kHistWarps);
    sp = min(sp, triples * kHistWarps);
    if (sp > kHistWarps) sp = triples * kHistWarps;          // full triples
    w->rounds_total = w->CG * triples;
  } else {
    sp = min(max_splits, sp);
    // whole CTAs per column group => every CTA can merge its tables (not when the warp count is the limit:
    // there the rows per warp matter more than the flush)
    if (sp >= kHistWarps && sp < max_splits) sp -= sp % kHistWarps;
    w->rounds_total = 0;
  }
  w->splits = sp;
  w->per = (((w->n + w->splits - 1) / w->splits) + 31) & ~31;
  w->items = w->CG * w->splits;
  return true;
}

// item of (round r, warp pair p): returns false if this warp has nothing to do in this round
struct HistItem { int cg, part; bool valid, merge; };
__device__ __forceinline__ bool hist_round_valid(const HistWork& w, int round) {
  return w.mode == 1 ? (static_cast<int>(blockIdx.x) + round * static_cast<int>(gridDim.x) < w.rounds_total)
                     : (static_cast<int>(blockIdx.x) * kHistWarps + round * w.total_warps < w.items);
}
__device__ __forceinline__ HistItem hist_item(const HistWork& w, int round, int pair) {
  HistItem it;
  if (w.mode == 1) {
    const int v = blockIdx.x + round * gridDim.x;
    it.cg = v % w.CG;
    const int t = v / w.CG;
    it.part = t * kHistWarps + pair;
    it.valid = it.part < w.splits;
    it.merge = (t * kHistWarps + kHistWarps - 1) < w.splits;
  } else {
    const int base = blockIdx.x * kHistWarps + round * w.total_warps;
    const int item = base + pair;
    it.valid = item < w.items;
    it.cg = item / w.splits; it.part = item % w.splits;
    it.merge = (base + kHistWarps - 1 < w.items) && (base / w.splits == (base + kHistWarps - 1) / w.splits);
  }
  return it;
}

__device__ __forceinline__ void consumer_bar_sync() {          // the 3 consumer warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(kHistWarps * 32) : "memory");
}

// (g,h) cell arithmetic: fp32 partial sums, or — quantized-gradient training — int32 partial sums whose bit
// patterns live in the same float2 slots (exact: a cell sees at most rows_per_part * 127 < 2^31).
template <bool QUANT>
__device__ __forceinline__ float2 acc2(float2 a, float2 b) {
  if (QUANT) return make_float2(__int_as_float(__float_as_int(a.x) + __float_as_int(b.x)), __int_as_float(__float_as_int(a.y) + __float_as_int(b.y)));
  return make_float2(a.x + b.x, a.y + b.y);
}

template <bool QUANT>
__global__ void __launch_bounds__(kHistThreads, 1) k_hist(const HistArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_producer = warp >= kHistWarps;
  const int pair = is_producer ? warp - kHistWarps : warp;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHistWarps * kWarpSmemBytes);
  uint64_t* full = bars + pair * (2 * kStages);       // producer -> consumer: stage landed
  uint64_t* empty = full + kStages;                   // consumer -> producer: stage consumed
  if (!is_producer && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 1); }
  }
  __syncthreads();
  pdl_enter();          // barrier init above overlaps the predecessor's tail; everything below reads its results

  HistWork w;
  if (!hist_work_setup(a, &w)) return;
  unsigned char* wbase = smem + pair * kWarpSmemBytes;
  unsigned char* ring = wbase + kWarpHistBytes;

  if (is_producer) {
    // ------------------------------------------------------------------ producer warp: stage rows
    int slot = 0; unsigned phase = 0;
    for (int round = 0; hist_round_valid(w, round); ++round) {
      const HistItem it = hist_item(w, round, pair);
      if (!it.valid) continue;
      const int cg = it.cg, part = it.part;
      const int r0 = part * w.per;
      const int r1 = min(w.n, r0 + w.per);
      if (r0 >= r1) continue;
      const uint8_t* colbase = a.bins + static_cast<int64_t>(cg) * kColGroup;
      const int half = (lane & 1) * 16;
      const int32_t* ip = w.idx ? w.idx + w.begin : nullptr;
      // Gathered passes: the ring holds only ~5 stages (6 KB) per warp, too little to cover the latency tail of 96
      // scattered requests per stage (ncu: consumers starved 28 % of the time, shared pipe 44 % busy vs 66 % at the
      // root).  The producer therefore also walks the index list `l2_prefetch` stages ahead and pulls every row's
      // 32-byte bin sector into L2 (fire-and-forget, no shared memory needed); the later cp.async hits L2.
      const int pf = (ip != nullptr) ? a.l2_prefetch : 0;
      int pfq[kPfIdLead];
#pragma unroll
      for (int d = 0; d < kPfIdLead; ++d) pfq[d] = -1;
      for (int p0 = r0; p0 < r1; p0 += kStageRows) {
        if (pf > 0) {
          if (pfq[0] >= 0) prefetch_l2(colbase + static_cast<int64_t>(pfq[0]) * a.pitch);
#pragma unroll
          for (int d = 0; d + 1 < kPfIdLead; ++d) pfq[d] = pfq[d + 1];
          const int pp = p0 + (pf + kPfIdLead) * kStageRows + lane;
          pfq[kPfIdLead - 1] = (pp < r1) ? __ldg(ip + pp) : -1;
        }
        if (a.use_tma && ip == nullptr && p0 + kStageRows <= r1) {
          // contiguous rows (root of an un-bagged tree): ONE 2-D TMA tile (32 rows x 32 columns of the row-major
          // matrix) + one bulk copy of the 32 (g,h) pairs per stage; both complete on the stage's mbarrier
          mbar_wait_parked(empty + slot, phase ^ 1);
          unsigned char* sb = ring + slot * kStageBytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(full + slot, kStageBytes);
            tma_load_tile_2d(sb, &tmap, cg * kColGroup, p0, full + slot);
            tma_load_1d(sb + kStageBinBytes, a.gh + p0, kStageRows * 8, full + slot);
          } else {
            mbar_arrive(full + slot);
          }
          if (++slot == kStages) { slot = 0; phase ^= 1; }
          continue;
        }
        // row ids: lane -> (g,h) of row p0+lane; lane pair -> 32-byte bin segment of rows p0+lane/2 and +16
        const int pa = p0 + (lane >> 1), pb = pa + 16, pg = p0 + lane;
        int ra = -1, rb = -1, rg = -1;
        if (pa < r1) ra = ip ? __ldg(ip + pa) : pa;
        if (pb < r1) rb = ip ? __ldg(ip + pb) : pb;
        if (pg < r1) rg = (w.gh_ord != nullptr) ? pg : (ip ? __ldg(ip + pg) : pg);       // leaf-ordered copy: the position IS the index
        mbar_wait_parked(empty + slot, phase ^ 1);           // the consumer released this ring slot
        unsigned char* sb = ring + slot * kStageBytes;
        if (ra >= 0) cp_async16(sb + (lane >> 1) * kColGroup + half, colbase + static_cast<int64_t>(ra) * a.pitch + half);
        if (rb >= 0) cp_async16(sb + (16 + (lane >> 1)) * kColGroup + half, colbase + static_cast<int64_t>(rb) * a.pitch + half);
        if (rg >= 0) cp_async8(sb + kStageBinBytes + lane * 8, (w.gh_ord != nullptr ? w.gh_ord : a.gh) + rg);
        mbar_arrive_on_cp_async(full + slot);
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warp: accumulate
  float2* H = reinterpret_cast<float2*>(wbase) + lane;            // lane's column of the [bin][lane] table
  const unsigned hbase = static_cast<unsigned>(__cvta_generic_to_shared(H));   // + bin*256 = the lane's cell
  const unsigned hbase0 = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float2*>(smem) + lane));
  const double gs = a.ctl->g_scale, hs = a.ctl->h_scale;
  int slot = 0; unsigned phase = 0;
  for (int round = 0; hist_round_valid(w, round); ++round) {
    const HistItem it = hist_item(w, round, pair);
    // merge path: all three warps of the CTA hold items of the same column group (CTA-uniform condition)
    const bool merge = it.merge;
    if (!it.valid) continue;                                       // only possible when !merge
    const int cg = it.cg, part = it.part;
    const int r0 = part * w.per;
    const int r1 = min(w.n, r0 + w.per);
    if (r0 >= r1 && !merge) continue;

    // zero the warp-private histogram (the producer is already filling the ring meanwhile)
    {
      float4* z = reinterpret_cast<float4*>(wbase);
#pragma unroll 8
      for (int i = lane; i < kWarpHistBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    for (int p0 = r0; p0 < r1; p0 += kStageRows) {
      mbar_wait(full + slot, phase);
      const int cnt = min(kStageRows, r1 - p0);
      const unsigned char* sb = ring + slot * kStageBytes;
      const unsigned char* sbin = sb + lane;
      const float2* sgh = reinterpret_cast<const float2*>(sb + kStageBinBytes);
      if (cnt == kStageRows) {
        constexpr int K = kHistBatch;
        // Software pipeline over the stage's 32/K batches (one warp per SMSP has no other warp to hide latency):
        //   LDS x K of batch k  |  [ALU: addresses + combined increments of batch k+1]  |  FADD + STS x K of batch k
        // and the raw bins / (g,h) of batch k+2 are fetched from the stage one iteration ahead.
        uint32_t nb[K]; float2 nq[K];            // raw inputs of the batch after next
        unsigned addrN[K]; float2 sN[K];         // prepared batch (next to be committed)
        auto fetch = [&](int r) {
#pragma unroll
          for (int i = 0; i < K; ++i) nb[i] = sbin[(r + i) * 32];
#pragma unroll
          for (int i = 0; i < K; i += 2) {
            const float4 t = *reinterpret_cast<const float4*>(sgh + r + i);
            nq[i] = make_float2(t.x, t.y); nq[i + 1] = make_float2(t.z, t.w);
          }
        };
        fetch(0);
        batch_prepare<K, QUANT>(hbase, nb, nq, addrN, sN);
        if (K < kStageRows) fetch(K);
#pragma unroll
        for (int r = 0; r < kStageRows; r += K) {
          unsigned addrC[K]; float2 sC[K], v[K];
#pragma unroll
          for (int i = 0; i < K; ++i) { addrC[i] = addrN[i]; sC[i] = sN[i]; }
#pragma unroll
          for (int i = 0; i < K; ++i) v[i] = lds64(addrC[i]);
          if (r + K < kStageRows) {
            batch_prepare<K, QUANT>(hbase, nb, nq, addrN, sN);     // batch r+K, in the shadow of the loads above
            if (r + 2 * K < kStageRows) fetch(r + 2 * K);
          }
#pragma unroll
          for (int i = 0; i < K; ++i) sts64(addrC[i], acc2<QUANT>(v[i], sC[i]));
        }
      } else {
        for (int r = 0; r < cnt; ++r) {
          const uint32_t b = sbin[r * 32];
          const float2 q = sgh[r];
          const unsigned addr = hbase + (b << 8);
          sts64(addr, acc2<QUANT>(lds64(addr), q));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
      if (++slot == kStages) { slot = 0; phase ^= 1; }
    }

    // flush: fp32 partials -> int64 fixed point, RED.ADD.64 into the leaf's pool slot.
    unsigned long long* dst = a.pool + static_cast<int64_t>(w.slot) * a.slot_stride +
                              (static_cast<int64_t>(cg) * kColGroup + lane) * (kBinsPerColumn * 2);
    if (merge) {
      // the CTA's three tables cover the same 32 columns: sum them in shared memory (fixed order 0,1,2 =>
      // deterministic) and let each warp flush a third of the bins => 3x fewer global atomics
      consumer_bar_sync();
      const int b_lo = pair * 86, b_hi = min(kBinsPerColumn, b_lo + 86);
#pragma unroll 2
      for (int b = b_lo; b < b_hi; ++b) {
        const float2 v0 = lds64(hbase0 + (b << 8));
        const float2 v1 = lds64(hbase0 + kWarpSmemBytes + (b << 8));
        const float2 v2 = lds64(hbase0 + 2 * kWarpSmemBytes + (b << 8));
        if (QUANT) {
          const int gi = __float_as_int(v0.x) + __float_as_int(v1.x) + __float_as_int(v2.x);
          const int hi = __float_as_int(v0.y) + __float_as_int(v1.y) + __float_as_int(v2.y);
          if (gi != 0 || hi != 0) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(static_cast<long long>(gi)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(static_cast<long long>(hi)));
          }
        } else {
          const float gx = (v0.x + v1.x) + v2.x, hx = (v0.y + v1.y) + v2.y;
          if (gx != 0.f || hx != 0.f) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(gx) * gs)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(hx) * hs)));
          }
        }
      }
      consumer_bar_sync();            // the tables are free to be zeroed for the next item
    } else {
#pragma unroll 4
      for (int b = 0; b < kBinsPerColumn; ++b) {
        const float2 v = lds64(hbase + (b << 8));
        if (QUANT) {
          const int gi = __float_as_int(v.x), hi = __float_as_int(v.y);
          if (gi != 0 || hi != 0) {
            atomicAdd(dst + 2 * b, static_cast<unsigned long long>(static_cast<long long>(gi)));
            atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(static_cast<long long>(hi)));
          }
        } else if (v.x != 0.f || v.y != 0.f) {
          atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v.x) * gs)));
          atomicAdd(dst + 2 * b + 1, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v.y) * hs)));
        }
      }
      __syncwarp();
    }
  }
}

// k_hist2: same staging, same work decomposition, but every column group is accumulated by TWO consumer warps that
// share the staged rows: a g-warp owning a [bin][lane] fp32 table of gradients and an h-warp owning the table of
// hessians (2 x 32 KB = the 64 KB of the float2 table).  Same shared-memory wavefronts per cell, but 6 consumer
// warps per SM instead of 3 (two per SMSP on half the SMSPs, all four SMSPs busy): a dependent-issue bubble of one
// warp is filled by the other, which is what the single-consumer version could not do.
constexpr int kHist2Threads = 3 * kHistWarps * 32;      // warps 0-2: g consumers, 3-5: h consumers, 6-8: producers
constexpr int kCompTableBytes = kBinsPerColumn * 32 * 4; // 32768

__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

template <int K>
__device__ __forceinline__ void batch_prepare1(unsigned tbase, const uint32_t (&b)[K], const float (&q)[K], unsigned (&addr)[K], float (&s)[K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    addr[i] = tbase + (b[i] << 7);
    s[i] = q[i];
#pragma unroll
    for (int j = 0; j < i; ++j) if (b[j] == b[i]) s[i] += q[j];
  }
}
__device__ __forceinline__ void consumer2_bar_sync() {          // the 6 consumer warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(2 * kHistWarps * 32) : "memory");
}

__global__ void __launch_bounds__(kHist2Threads, 1) k_hist2(const HistArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_producer = warp >= 2 * kHistWarps;
  const int comp = (warp >= kHistWarps && !is_producer) ? 1 : 0;      // 0: gradients, 1: hessians
  const int pair = warp % kHistWarps;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHistWarps * kWarpSmemBytes);
  uint64_t* full = bars + pair * (2 * kStages);       // producer -> consumer: stage landed
  uint64_t* empty = full + kStages;                   // consumer -> producer: stage consumed
  if (!is_producer && comp == 0 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 32); mbar_init(empty + i, 2); }   // two consumers release a slot
  }
  __syncthreads();
  pdl_enter();          // barrier init above overlaps the predecessor's tail; everything below reads its results

  HistWork w;
  if (!hist_work_setup(a, &w)) return;
  unsigned char* wbase = smem + pair * kWarpSmemBytes;
  unsigned char* ring = wbase + kWarpHistBytes;

  if (is_producer) {
    // ------------------------------------------------------------------ producer warp: stage rows
    int slot = 0; unsigned phase = 0;
    for (int round = 0; hist_round_valid(w, round); ++round) {
      const HistItem it = hist_item(w, round, pair);
      if (!it.valid) continue;
      const int cg = it.cg, part = it.part;
      const int r0 = part * w.per;
      const int r1 = min(w.n, r0 + w.per);
      if (r0 >= r1) continue;
      const uint8_t* colbase = a.bins + static_cast<int64_t>(cg) * kColGroup;
      const int half = (lane & 1) * 16;
      const int32_t* ip = w.idx ? w.idx + w.begin : nullptr;
      // Gathered passes: the ring holds only ~5 stages (6 KB) per warp, too little to cover the latency tail of 96
      // scattered requests per stage (ncu: consumers starved 28 % of the time, shared pipe 44 % busy vs 66 % at the
      // root).  The producer therefore also walks the index list `l2_prefetch` stages ahead and pulls every row's
      // 32-byte bin sector into L2 (fire-and-forget, no shared memory needed); the later cp.async hits L2.
      const int pf = (ip != nullptr) ? a.l2_prefetch : 0;
      int pfq[kPfIdLead];
#pragma unroll
      for (int d = 0; d < kPfIdLead; ++d) pfq[d] = -1;
      for (int p0 = r0; p0 < r1; p0 += kStageRows) {
        if (pf > 0) {
          if (pfq[0] >= 0) prefetch_l2(colbase + static_cast<int64_t>(pfq[0]) * a.pitch);
#pragma unroll
          for (int d = 0; d + 1 < kPfIdLead; ++d) pfq[d] = pfq[d + 1];
          const int pp = p0 + (pf + kPfIdLead) * kStageRows + lane;
          pfq[kPfIdLead - 1] = (pp < r1) ? __ldg(ip + pp) : -1;
        }
        if (a.use_tma && ip == nullptr && p0 + kStageRows <= r1) {
          // contiguous rows (root of an un-bagged tree): ONE 2-D TMA tile (32 rows x 32 columns of the row-major
          // matrix) + one bulk copy of the 32 (g,h) pairs per stage; both complete on the stage's mbarrier
          mbar_wait_parked(empty + slot, phase ^ 1);
          unsigned char* sb = ring + slot * kStageBytes;
          if (lane == 0) {
            mbar_arrive_expect_tx(full + slot, kStageBytes);
            tma_load_tile_2d(sb, &tmap, cg * kColGroup, p0, full + slot);
            tma_load_1d(sb + kStageBinBytes, a.gh + p0, kStageRows * 8, full + slot);
          } else {
            mbar_arrive(full + slot);
          }
          if (++slot == kStages) { slot = 0; phase ^= 1; }
          continue;
        }
        // row ids: lane -> (g,h) of row p0+lane; lane pair -> 32-byte bin segment of rows p0+lane/2 and +16
        const int pa = p0 + (lane >> 1), pb = pa + 16, pg = p0 + lane;
        int ra = -1, rb = -1, rg = -1;
        if (pa < r1) ra = ip ? __ldg(ip + pa) : pa;
        if (pb < r1) rb = ip ? __ldg(ip + pb) : pb;
        if (pg < r1) rg = (w.gh_ord != nullptr) ? pg : (ip ? __ldg(ip + pg) : pg);       // leaf-ordered copy: the position IS the index
        mbar_wait_parked(empty + slot, phase ^ 1);           // the consumer released this ring slot
        unsigned char* sb = ring + slot * kStageBytes;
        if (ra >= 0) cp_async16(sb + (lane >> 1) * kColGroup + half, colbase + static_cast<int64_t>(ra) * a.pitch + half);
        if (rb >= 0) cp_async16(sb + (16 + (lane >> 1)) * kColGroup + half, colbase + static_cast<int64_t>(rb) * a.pitch + half);
        if (rg >= 0) cp_async8(sb + kStageBinBytes + lane * 8, (w.gh_ord != nullptr ? w.gh_ord : a.gh) + rg);
        mbar_arrive_on_cp_async(full + slot);
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warp: accumulate ONE component
  unsigned char* tptr = wbase + comp * kCompTableBytes;             // this warp's [bin][lane] fp32 table
  const unsigned tbase = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float*>(tptr) + lane));
  const unsigned tbase0 = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float*>(smem + comp * kCompTableBytes) + lane));
  const double scale = comp == 0 ? a.ctl->g_scale : a.ctl->h_scale;
  int slot = 0; unsigned phase = 0;
  for (int round = 0; hist_round_valid(w, round); ++round) {
    const HistItem it = hist_item(w, round, pair);
    const bool merge = it.merge;
    if (!it.valid) continue;                                       // only possible when !merge
    const int cg = it.cg, part = it.part;
    const int r0 = part * w.per;
    const int r1 = min(w.n, r0 + w.per);
    if (r0 >= r1 && !merge) continue;
    {
      float4* z = reinterpret_cast<float4*>(tptr);
#pragma unroll 8
      for (int i = lane; i < kCompTableBytes / 16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    for (int p0 = r0; p0 < r1; p0 += kStageRows) {
      mbar_wait(full + slot, phase);
      const int cnt = min(kStageRows, r1 - p0);
      const unsigned char* sb = ring + slot * kStageBytes;
      const unsigned char* sbin = sb + lane;
      const float* sq = reinterpret_cast<const float*>(sb + kStageBinBytes) + comp;      // q of row r = sq[2 r]
      if (cnt == kStageRows) {
        constexpr int K = kHistBatch;
        uint32_t nb[K]; float nq[K];
        unsigned addrN[K]; float sN[K];
        auto fetch = [&](int r) {
#pragma unroll
          for (int i = 0; i < K; ++i) nb[i] = sbin[(r + i) * 32];
#pragma unroll
          for (int i = 0; i < K; i += 2) {
            const float4 t = *reinterpret_cast<const float4*>(sq - comp + 2 * (r + i));
            nq[i] = comp ? t.y : t.x; nq[i + 1] = comp ? t.w : t.z;
          }
        };
        fetch(0);
        batch_prepare1<K>(tbase, nb, nq, addrN, sN);
        if (K < kStageRows) fetch(K);
#pragma unroll
        for (int r = 0; r < kStageRows; r += K) {
          unsigned addrC[K]; float sC[K], v[K];
#pragma unroll
          for (int i = 0; i < K; ++i) { addrC[i] = addrN[i]; sC[i] = sN[i]; }
#pragma unroll
          for (int i = 0; i < K; ++i) v[i] = lds32(addrC[i]);
          if (r + K < kStageRows) {
            batch_prepare1<K>(tbase, nb, nq, addrN, sN);
            if (r + 2 * K < kStageRows) fetch(r + 2 * K);
          }
#pragma unroll
          for (int i = 0; i < K; ++i) sts32(addrC[i], v[i] + sC[i]);
        }
      } else {
        for (int r = 0; r < cnt; ++r) {
          const uint32_t b = sbin[r * 32];
          const unsigned addr = tbase + (b << 7);
          sts32(addr, lds32(addr) + sq[2 * r]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
      if (++slot == kStages) { slot = 0; phase ^= 1; }
    }

    // flush this component: fp32 partial -> int64 fixed point, RED.ADD.64 into the leaf's pool slot
    unsigned long long* dst = a.pool + static_cast<int64_t>(w.slot) * a.slot_stride +
                              (static_cast<int64_t>(cg) * kColGroup + lane) * (kBinsPerColumn * 2) + comp;
    if (merge) {
      consumer2_bar_sync();
      const int b_lo = pair * 86, b_hi = min(kBinsPerColumn, b_lo + 86);
#pragma unroll 2
      for (int b = b_lo; b < b_hi; ++b) {
        const float x = (lds32(tbase0 + (b << 7)) + lds32(tbase0 + kWarpSmemBytes + (b << 7))) + lds32(tbase0 + 2 * kWarpSmemBytes + (b << 7));
        if (x != 0.f) atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(x) * scale)));
      }
      consumer2_bar_sync();
    } else {
#pragma unroll 4
      for (int b = 0; b < kBinsPerColumn; ++b) {
        const float x = lds32(tbase + (b << 7));
        if (x != 0.f) atomicAdd(dst + 2 * b, static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(x) * scale)));
      }
      __syncwarp();
    }
  }
}

}  // namespace b200
